// MaskedSum: dest[col] = sum over rows of mask[row] * src[row, col].
//
// Replaces reference maskedsum.mako:38-68 (one thread per column walking all rows
// with a float32 fma chain).  Here a block owns 16 columns (one 128-byte line of
// complex64 per row) and its 64 row groups walk the rows interleaved, so a column
// sum is split 64 ways; partial sums are float64 and are combined in a fixed
// order, which makes the result the correctly rounded sum for all practical
// purposes and independent of the launch geometry.
#include "common.cuh"

namespace {

using namespace ksp;

constexpr int MS_COLS = 16;
constexpr int MS_GROUPS = 64;

template <bool AMPLITUDES>
__global__ void __launch_bounds__(MS_COLS * MS_GROUPS)
maskedsum_kernel(const float2 *__restrict__ src, const float *__restrict__ mask,
                 void *__restrict__ dest, int64_t rows, int64_t cols, int64_t stride, int abs_mode)
{
    __shared__ double part[MS_GROUPS][MS_COLS][2];
    const int cl = threadIdx.x & (MS_COLS - 1);
    const int grp = threadIdx.x / MS_COLS;
    const int64_t col_raw = (int64_t) blockIdx.x * MS_COLS + cl;
    const int64_t col = col_raw < cols ? col_raw : cols - 1;
    double re = 0.0, im = 0.0;
    int64_t r = grp;
    for (; r + 3 * MS_GROUPS < rows; r += 4 * MS_GROUPS) {
        float2 v[4];
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k] = ldg_stream_f2(src + (r + k * MS_GROUPS) * stride + col);
            m[k] = __ldg(mask + r + k * MS_GROUPS);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (AMPLITUDES) {
                re = fma((double) m[k], (double) abs_c64_rt(v[k].x, v[k].y, abs_mode), re);
            } else {
                re = fma((double) m[k], (double) v[k].x, re);
                im = fma((double) m[k], (double) v[k].y, im);
            }
        }
    }
    for (; r < rows; r += MS_GROUPS) {
        float2 v = ldg_stream_f2(src + r * stride + col);
        double m = (double) __ldg(mask + r);
        if (AMPLITUDES) {
            re = fma(m, (double) abs_c64_rt(v.x, v.y, abs_mode), re);
        } else {
            re = fma(m, (double) v.x, re);
            im = fma(m, (double) v.y, im);
        }
    }
    part[grp][cl][0] = re;
    part[grp][cl][1] = im;
    __syncthreads();
    if (grp == 0 && col_raw < cols) {
        double sre = 0.0, sim = 0.0;
        for (int g = 0; g < MS_GROUPS; g++) {
            sre += part[g][cl][0];
            sim += part[g][cl][1];
        }
        if (AMPLITUDES)
            reinterpret_cast<float *>(dest)[col_raw] = __double2float_rn(sre);
        else
            reinterpret_cast<float2 *>(dest)[col_raw] =
                make_float2(__double2float_rn(sre), __double2float_rn(sim));
    }
}

}  // namespace

extern "C" int ksp_maskedsum(void *stream, const void *src, const float *mask, void *dest,
                             int64_t rows, int64_t cols, int64_t src_stride, int use_amplitudes,
                             int abs_mode)
{
    if (rows < 0 || cols < 0 || src_stride < cols) return KSP_EINVAL;
    if (cols == 0) return 0;
    if (!src || !dest || (rows > 0 && !mask)) return KSP_EINVAL;
    if (abs_mode != KSP_ABS_NUMPY && abs_mode != KSP_ABS_HYPOT) return KSP_EINVAL;
    if ((uintptr_t) src % 8 || (uintptr_t) dest % (use_amplitudes ? 4 : 8)) return KSP_EALIGN;
    dim3 grid((unsigned) ksp_divup(cols, MS_COLS));
    cudaStream_t s = (cudaStream_t) stream;
    if (use_amplitudes)
        maskedsum_kernel<true><<<grid, MS_COLS * MS_GROUPS, 0, s>>>(
            (const float2 *) src, mask, dest, rows, cols, src_stride, abs_mode);
    else
        maskedsum_kernel<false><<<grid, MS_COLS * MS_GROUPS, 0, s>>>(
            (const float2 *) src, mask, dest, rows, cols, src_stride, abs_mode);
    KSP_CHECK_LAUNCH();
    return 0;
}
