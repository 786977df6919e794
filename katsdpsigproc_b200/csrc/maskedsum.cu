// MaskedSum: dest[col] = sum over rows of mask[row] * src[row, col].
//
// Replaces reference maskedsum.mako:38-68 (one thread per column walking all rows
// with a float32 fma chain).  Here a block owns a strip of 16 lanes x VEC columns
// (128 or 256 contiguous bytes of complex64 per row) and its 16 row groups walk the
// rows interleaved with 8 loads in flight per thread.  8320 columns are only 260
// strips, too few blocks to fill 148 SMs, so the ROWS are split as well: a thread
// block cluster of S blocks (S = 1, 2, 4 or 8, the largest that still fits the
// device in one wave) shares a strip, block k taking the row groups k, k + S, ...;
// each block reduces its groups in shared memory and block 0 of the cluster adds
// the S partial strips through distributed shared memory.  Partial sums are
// float64 and are combined in a fixed order, which makes the result the correctly
// rounded sum for all practical purposes and reproducible for a
// given shape and device.
#include <cooperative_groups.h>
#include "common.cuh"

namespace {

using namespace ksp;
namespace cg = cooperative_groups;

constexpr int MS_LANES = 16;    // column lanes of a block
constexpr int MS_GROUPS = 16;   // row groups of a block
constexpr int MS_UNROLL = 8;    // rows in flight per thread
constexpr int MS_THREADS = MS_LANES * MS_GROUPS;
constexpr int MS_BLOCKS_PER_SM = 4;

template <int VEC> struct MsLoad;
template <> struct MsLoad<1> {
    typedef float2 T;
    static __device__ __forceinline__ T ld(const float2 *p, bool) { return ldg_stream_f2(p); }
    static __device__ __forceinline__ float2 at(const T &v, int) { return v; }
};
template <> struct MsLoad<2> {
    typedef float4 T;
    static __device__ __forceinline__ T ld(const float2 *p, bool single)
    {
        float4 v;
        if (single) {                                     // the odd last column of the array
            const float2 h = ldg_stream_f2(p);
            return make_float4(h.x, h.y, 0.0f, 0.0f);
        }
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_evict_first()));
        return v;
    }
    static __device__ __forceinline__ float2 at(const T &v, int j)
    {
        return j ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
    }
};

// VEC = 2 needs an even stride and a 16-byte aligned source.  Lanes beyond the last column
// re-read the last one (their sums are not written).
template <bool AMPLITUDES, int VEC>
__global__ void __launch_bounds__(MS_THREADS, MS_BLOCKS_PER_SM)
maskedsum_kernel(const float2 *__restrict__ src, const float *__restrict__ mask,
                 void *__restrict__ dest, int64_t rows, int64_t cols, int64_t stride, int abs_mode)
{
    typedef MsLoad<VEC> L;
    __shared__ double part[MS_GROUPS][MS_LANES * VEC][2];
    cg::cluster_group cluster = cg::this_cluster();
    const int split = (int) cluster.num_blocks();
    const int rank = (int) cluster.block_rank();
    const int cl = threadIdx.x & (MS_LANES - 1);
    const int grp = threadIdx.x / MS_LANES;
    const int64_t col_raw = ((int64_t) blockIdx.x * MS_LANES + cl) * VEC;
    const int64_t col = col_raw < cols ? col_raw : (cols - 1) & ~(int64_t) (VEC - 1);
    const bool single = VEC == 2 && col + 1 == cols;
    double acc[VEC][2];
#pragma unroll
    for (int j = 0; j < VEC; j++) acc[j][0] = acc[j][1] = 0.0;
    const int64_t step = (int64_t) MS_GROUPS * split;       // rows between two of this thread's
    int64_t r = (int64_t) grp * split + rank;
    const float2 *p = src + r * stride + col;
    const int64_t pstep = step * stride;
    for (; r + (MS_UNROLL - 1) * step < rows; r += MS_UNROLL * step, p += MS_UNROLL * pstep) {
        typename L::T v[MS_UNROLL];
        float m[MS_UNROLL];
#pragma unroll
        for (int k = 0; k < MS_UNROLL; k++) {
            v[k] = L::ld(p + k * pstep, single);
            m[k] = __ldg(mask + r + k * step);
        }
#pragma unroll
        for (int k = 0; k < MS_UNROLL; k++) {
            const double mk = (double) m[k];
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                const float2 x = L::at(v[k], j);
                if (AMPLITUDES) {
                    acc[j][0] = fma(mk, (double) abs_c64_rt(x.x, x.y, abs_mode), acc[j][0]);
                } else {
                    acc[j][0] = fma(mk, (double) x.x, acc[j][0]);
                    acc[j][1] = fma(mk, (double) x.y, acc[j][1]);
                }
            }
        }
    }
    for (; r < rows; r += step, p += pstep) {
        const typename L::T v = L::ld(p, single);
        const double mk = (double) __ldg(mask + r);
#pragma unroll
        for (int j = 0; j < VEC; j++) {
            const float2 x = L::at(v, j);
            if (AMPLITUDES) {
                acc[j][0] = fma(mk, (double) abs_c64_rt(x.x, x.y, abs_mode), acc[j][0]);
            } else {
                acc[j][0] = fma(mk, (double) x.x, acc[j][0]);
                acc[j][1] = fma(mk, (double) x.y, acc[j][1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) {
        part[grp][cl * VEC + j][0] = acc[j][0];
        part[grp][cl * VEC + j][1] = acc[j][1];
    }
    __syncthreads();
    // the block's own strip: groups in order, left in part[0]
    constexpr int STRIP = MS_LANES * VEC;
    if (threadIdx.x < STRIP * 2) {
        const int c = threadIdx.x >> 1, ri = threadIdx.x & 1;
        double s = 0.0;
        for (int g = 0; g < MS_GROUPS; g++) s += part[g][c][ri];
        part[0][c][ri] = s;                                   // only this thread reads part[.][c][ri]
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x < STRIP) {
        const int c = threadIdx.x;
        double sre = 0.0, sim = 0.0;
        for (int k = 0; k < split; k++) {
            const double *remote = cluster.map_shared_rank(&part[0][0][0], k);
            sre += remote[c * 2 + 0];
            sim += remote[c * 2 + 1];
        }
        const int64_t out_col = (int64_t) blockIdx.x * STRIP + c;
        if (out_col < cols) {
            if (AMPLITUDES)
                reinterpret_cast<float *>(dest)[out_col] = __double2float_rn(sre);
            else
                reinterpret_cast<float2 *>(dest)[out_col] =
                    make_float2(__double2float_rn(sre), __double2float_rn(sim));
        }
    }
    cluster.sync();                                           // partial strips stay alive until read
}

template <bool AMPLITUDES, int VEC>
int launch(cudaStream_t s, const void *src, const float *mask, void *dest, int64_t rows,
           int64_t cols, int64_t stride, int abs_mode)
{
    const int64_t strips = ksp_divup(cols, MS_LANES * VEC);
    const int64_t slots = (int64_t) ksp_sm_count() * MS_BLOCKS_PER_SM;
    int split = 1;
    while (split < 8 && strips * split * 2 <= slots && rows >= (int64_t) MS_GROUPS * split * 2 * MS_UNROLL)
        split *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) strips, (unsigned) split);
    cfg.blockDim = dim3(MS_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned) split;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, maskedsum_kernel<AMPLITUDES, VEC>, (const float2 *) src,
                                       mask, dest, rows, cols, stride, abs_mode);
    ksp_count_launch();
    if (e != cudaSuccess) return (int) e;
    return 0;
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
    cudaStream_t s = (cudaStream_t) stream;
    const bool vec2 = cols >= 2 && src_stride % 2 == 0 && (uintptr_t) src % 16 == 0;
    if (use_amplitudes)
        return vec2 ? launch<true, 2>(s, src, mask, dest, rows, cols, src_stride, abs_mode)
                    : launch<true, 1>(s, src, mask, dest, rows, cols, src_stride, abs_mode);
    return vec2 ? launch<false, 2>(s, src, mask, dest, rows, cols, src_stride, abs_mode)
                : launch<false, 1>(s, src, mask, dest, rows, cols, src_stride, abs_mode);
}
