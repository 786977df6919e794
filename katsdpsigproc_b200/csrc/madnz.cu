// Median-of-absolute-deviations noise estimate and Percentile5.
//
// madnz_t  : replaces reference rfi/madnz_t.mako:72-87 (+ rank.mako:236-266).
// madnz    : replaces reference rfi/madnz.mako:105-123 (channel-major input).
// percentile5 : replaces reference percentile.mako:115-140.
//
// All three are exact selections.  madnz_t / percentile5 give
// one thread block per row: the row is read once from global memory, turned into
// order-preserving 32-bit keys in shared memory (rows up to 49 152 elements; longer
// rows are re-read from L2), and block_radix_select (select.cuh) finds the ranks.
// madnz (channel-major) gives one block per 32 baselines, lane == baseline, and
// makes its 4 radix passes over global memory with per-baseline histograms
// (hist[digit][lane]: conflict-free).
#include "common.cuh"
#include "select.cuh"

int ksp_row_mad(cudaStream_t s, const float *dev_t, float *noise, int64_t channels,
                int64_t baselines, int64_t dev_stride);

namespace {

using namespace ksp;

constexpr int SEL_THREADS = 1024;
constexpr int SMEM_KEY_CAP = 49152;  // 192 KB of keys + 32 KB histogram + misc < 227 KB

__device__ __forceinline__ uint32_t mad_key(float v)
{
    uint32_t b = __float_as_uint(v) & 0x7fffffffu;   // |v|
    return (b == 0u || b > 0x7f800000u) ? KEY_SKIP : b;  // zeros and NaN take no part
}

__device__ __forceinline__ float mad_finish(uint32_t lo_key, uint32_t hi_key)
{
    double lo = (double) __uint_as_float(lo_key), hi = (double) __uint_as_float(hi_key);
    double med = (lo_key == hi_key) ? lo : (lo + hi) * 0.5;
    return __double2float_rn(1.4826 * med);
}

// block-wide sum of a per-thread count; result broadcast
template <int THREADS>
__device__ uint32_t block_sum(uint32_t v, uint32_t *misc)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) misc[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < THREADS / 32; w++) t += misc[w];
    __syncthreads();
    return t;
}

// Lower/upper median keys of the non-skipped keys; n_valid > 0.
template <int THREADS, typename KeySource>
__device__ void block_median_keys(const KeySource &key_at, int n, uint32_t n_valid,
                                  const SelectScratch &sc, uint32_t &lo, uint32_t &hi)
{
    uint32_t k = (n_valid - 1) >> 1;
    lo = block_radix_select<THREADS>(key_at, n, k, sc);
    hi = lo;
    if (!(n_valid & 1u)) {
        uint32_t next, count_le;
        block_next_above<THREADS>(key_at, n, lo, next, count_le, sc);
        if (count_le < k + 2) hi = next;  // rank k+1 is the next distinct value
    }
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(SEL_THREADS, 1)
madnz_t_kernel(const float *__restrict__ dev_t, float *__restrict__ noise, int channels,
               int64_t stride)
{
    extern __shared__ __align__(16) uint32_t smem[];
    SelectScratch sc;
    sc.hist = smem;
    sc.misc = smem + SELECT_HIST_WORDS;
    uint32_t *keys = smem + SELECT_HIST_WORDS + 64;

    const float *row = dev_t + (int64_t) blockIdx.x * stride;
    const int tid = threadIdx.x;
    uint32_t valid = 0;
    const bool vec = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(dev_t) & 15) == 0);
    if (vec) {
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        for (int i = tid; i < (channels >> 2); i += SEL_THREADS) {
            float4 v = __ldg(row4 + i);
            uint4 k = make_uint4(mad_key(v.x), mad_key(v.y), mad_key(v.z), mad_key(v.w));
            valid += (k.x != KEY_SKIP) + (k.y != KEY_SKIP) + (k.z != KEY_SKIP) + (k.w != KEY_SKIP);
            if (IN_SMEM) reinterpret_cast<uint4 *>(keys)[i] = k;
        }
        for (int i = (channels & ~3) + tid; i < channels; i += SEL_THREADS) {
            uint32_t k = mad_key(row[i]);
            valid += (k != KEY_SKIP);
            if (IN_SMEM) keys[i] = k;
        }
    } else {
        for (int i = tid; i < channels; i += SEL_THREADS) {
            uint32_t k = mad_key(row[i]);
            valid += (k != KEY_SKIP);
            if (IN_SMEM) keys[i] = k;
        }
    }
    uint32_t n_valid = block_sum<SEL_THREADS>(valid, sc.misc);
    if (n_valid == 0) {
        if (tid == 0) noise[blockIdx.x] = __int_as_float(0x7fc00000);
        return;
    }
    uint32_t lo, hi;
    if (IN_SMEM) {
        auto src = [keys](int i) { return keys[i]; };
        block_median_keys<SEL_THREADS>(src, channels, n_valid, sc, lo, hi);
    } else {
        auto src = [row](int i) { return mad_key(row[i]); };
        block_median_keys<SEL_THREADS>(src, channels, n_valid, sc, lo, hi);
    }
    if (tid == 0) noise[blockIdx.x] = mad_finish(lo, hi);
}

// ------------------------------------------------------------------ channel-major MAD
// block = 32 warps x 32 lanes; lane == baseline, warps stride over channels.
__global__ void __launch_bounds__(1024, 1)
madnz_cm_kernel(const float *__restrict__ dev, float *__restrict__ noise, int channels,
                int baselines, int64_t stride)
{
    __shared__ uint32_t hist[256 * 32];   // hist[digit][lane]
    __shared__ uint32_t part[32][33];     // per-warp partials
    __shared__ uint32_t s_prefix[32], s_rank[32], s_nvalid[32], s_next[32], s_cle[32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b_raw = blockIdx.x * 32 + lane;
    const bool ok = b_raw < baselines;
    const int b = ok ? b_raw : baselines - 1;
    const float *col = dev + b;

    // count usable samples per baseline
    uint32_t cnt = 0;
    for (int c = warp; c < channels; c += 32) cnt += (mad_key(col[(int64_t) c * stride]) != KEY_SKIP);
    part[warp][lane] = cnt;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 32; w++) t += part[w][lane];
        s_nvalid[lane] = t;
        s_rank[lane] = t ? (t - 1) >> 1 : 0;
        s_prefix[lane] = 0;
    }
    __syncthreads();

    uint32_t prefix_mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256 * 32; i += 1024) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix[lane];
        for (int c = warp; c < channels; c += 32) {
            uint32_t k = mad_key(col[(int64_t) c * stride]);
            if (k != KEY_SKIP && (k & prefix_mask) == prefix)
                atomicAdd(&hist[((k >> shift) & 0xffu) * 32 + lane], 1u);
        }
        __syncthreads();
        // warp w resolves baseline w: lane j owns digits 8j .. 8j+7
        {
            const int bl = warp;
            uint32_t c8[8], tot = 0;
#pragma unroll
            for (int d = 0; d < 8; d++) {
                c8[d] = hist[(lane * 8 + d) * 32 + bl];
                tot += c8[d];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            uint32_t excl = incl - tot;
            const uint32_t rank = s_rank[bl];
            __syncwarp();
            if (s_nvalid[bl] != 0 && rank >= excl && rank < excl + tot) {
                uint32_t run = excl;
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    if (rank >= run && rank < run + c8[d]) {
                        s_prefix[bl] |= (uint32_t) (lane * 8 + d) << shift;
                        s_rank[bl] = rank - run;
                    }
                    run += c8[d];
                }
            }
        }
        prefix_mask |= 0xffu << shift;
        __syncthreads();
    }

    // upper median for even counts: smallest key above, and count of keys <= lower median
    const uint32_t lo = s_prefix[lane];
    uint32_t best = KEY_SKIP, cle = 0;
    for (int c = warp; c < channels; c += 32) {
        uint32_t k = mad_key(col[(int64_t) c * stride]);
        if (k != KEY_SKIP) {
            if (k <= lo) cle++;
            else best = min(best, k);
        }
    }
    part[warp][lane] = best;
    __syncthreads();
    if (warp == 0) {
        uint32_t m = KEY_SKIP;
        for (int w = 0; w < 32; w++) m = min(m, part[w][lane]);
        s_next[lane] = m;
    }
    __syncthreads();
    part[warp][lane] = cle;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 32; w++) t += part[w][lane];
        s_cle[lane] = t;
        if (ok) {
            uint32_t n = s_nvalid[lane];
            float out = __int_as_float(0x7fc00000);
            if (n) {
                uint32_t k = (n - 1) >> 1, hi = lo;
                if (!(n & 1u) && t < k + 2) hi = s_next[lane];
                out = mad_finish(lo, hi);
            }
            noise[b_raw] = out;
        }
    }
}

// ------------------------------------------------------------------ Percentile5
template <bool IN_SMEM>
__global__ void __launch_bounds__(SEL_THREADS, 1)
percentile5_kernel(const void *__restrict__ src, float *__restrict__ dest, int64_t src_stride,
                   int64_t dest_stride, int64_t first_col, int n_cols, int is_amplitude,
                   int abs_mode)
{
    extern __shared__ __align__(16) uint32_t smem[];
    SelectScratch sc;
    sc.hist = smem;
    sc.misc = smem + SELECT_HIST_WORDS;
    uint32_t *keys = smem + SELECT_HIST_WORDS + 64;

    const int64_t r = blockIdx.x;
    const int64_t off = r * src_stride + first_col;
    const float *fsrc = reinterpret_cast<const float *>(src);
    const float2 *csrc = reinterpret_cast<const float2 *>(src);
    auto value_key = [=](int i) -> uint32_t {
        float v;
        if (is_amplitude) {
            v = fabsf(fsrc[off + i]);
        } else {
            float2 z = csrc[off + i];
            v = abs_c64_rt(z.x, z.y, abs_mode);
        }
        return float_to_key(v);
    };
    const int tid = threadIdx.x;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i = tid; i < n_cols; i += SEL_THREADS) {
        uint32_t k = value_key(i);
        if (IN_SMEM) keys[i] = k;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
    }
    if ((tid & 31) == 0) {
        sc.misc[tid >> 5] = kmin;
        sc.hist[tid >> 5] = kmax;
    }
    __syncthreads();
    kmin = 0xffffffffu;
    kmax = 0u;
    for (int w = 0; w < SEL_THREADS / 32; w++) {
        kmin = min(kmin, sc.misc[w]);
        kmax = max(kmax, sc.hist[w]);
    }
    __syncthreads();

    const uint32_t n = (uint32_t) n_cols;
    const uint32_t ranks[3] = {(n - 1) / 4, ((n - 1) * 3) / 4, (n - 1) / 2};
    uint32_t found[3];
#pragma unroll 1
    for (int q = 0; q < 3; q++) {
        if (IN_SMEM) {
            auto srcf = [keys](int i) { return keys[i]; };
            found[q] = block_radix_select<SEL_THREADS>(srcf, n_cols, ranks[q], sc);
        } else {
            found[q] = block_radix_select<SEL_THREADS>(value_key, n_cols, ranks[q], sc);
        }
    }
    if (tid == 0) {
        dest[0 * dest_stride + r] = key_to_float(kmin);
        dest[1 * dest_stride + r] = key_to_float(kmax);
        dest[2 * dest_stride + r] = key_to_float(found[0]);
        dest[3 * dest_stride + r] = key_to_float(found[1]);
        dest[4 * dest_stride + r] = key_to_float(found[2]);
    }
}

size_t select_smem_bytes(int64_t n, bool in_smem)
{
    return (size_t) (SELECT_HIST_WORDS + 64 + (in_smem ? ((n + 3) & ~(int64_t) 3) : 0)) * 4;
}

}  // namespace

extern "C" int ksp_madnz_t(void *stream, const float *dev_t, float *noise, int64_t channels,
                           int64_t baselines, int64_t stride)
{
    if (channels < 0 || baselines < 0 || stride < channels) return KSP_EINVAL;
    if (baselines == 0) return 0;
    if (!dev_t || !noise) return KSP_EINVAL;
    if (channels > (int64_t) 1 << 30) return KSP_ETOOLARGE;
    cudaStream_t s = (cudaStream_t) stream;
    if (channels <= 32768)      // row fits one block of the row kernel (threshold.cu / mad.cuh)
        return ksp_row_mad(s, dev_t, noise, channels, baselines, stride);
    const bool in_smem = channels <= SMEM_KEY_CAP;
    const size_t smem = select_smem_bytes(channels, in_smem);
    if (in_smem) {
        KSP_CUDA(cudaFuncSetAttribute(madnz_t_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        madnz_t_kernel<true><<<(unsigned) baselines, SEL_THREADS, smem, s>>>(dev_t, noise,
                                                                             (int) channels, stride);
    } else {
        KSP_CUDA(cudaFuncSetAttribute(madnz_t_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        madnz_t_kernel<false><<<(unsigned) baselines, SEL_THREADS, smem, s>>>(
            dev_t, noise, (int) channels, stride);
    }
    KSP_CHECK_LAUNCH();
    return 0;
}

extern "C" int ksp_madnz(void *stream, const float *dev, float *noise, int64_t channels,
                         int64_t baselines, int64_t stride)
{
    if (channels < 0 || baselines < 0 || stride < baselines) return KSP_EINVAL;
    if (baselines == 0) return 0;
    if (!dev || !noise) return KSP_EINVAL;
    if (channels > (int64_t) 1 << 30 || baselines > (int64_t) 1 << 30) return KSP_ETOOLARGE;
    madnz_cm_kernel<<<(unsigned) ksp_divup(baselines, 32), 1024, 0, (cudaStream_t) stream>>>(
        dev, noise, (int) channels, (int) baselines, stride);
    KSP_CHECK_LAUNCH();
    return 0;
}

extern "C" int ksp_percentile5(void *stream, const void *src, float *dest, int64_t rows,
                               int64_t src_stride, int64_t dest_stride, int64_t first_col,
                               int64_t n_cols, int is_amplitude, int abs_mode)
{
    if (rows < 0 || n_cols <= 0 || first_col < 0 || src_stride < first_col + n_cols ||
        dest_stride < rows)
        return KSP_EINVAL;
    if (rows == 0) return 0;
    if (!src || !dest) return KSP_EINVAL;
    if (n_cols > (int64_t) 1 << 30) return KSP_ETOOLARGE;
    if (abs_mode != KSP_ABS_NUMPY && abs_mode != KSP_ABS_HYPOT) return KSP_EINVAL;
    cudaStream_t s = (cudaStream_t) stream;
    const bool in_smem = n_cols <= SMEM_KEY_CAP;
    const size_t smem = select_smem_bytes(n_cols, in_smem);
    if (in_smem) {
        KSP_CUDA(cudaFuncSetAttribute(percentile5_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        percentile5_kernel<true><<<(unsigned) rows, SEL_THREADS, smem, s>>>(
            src, dest, src_stride, dest_stride, first_col, (int) n_cols, is_amplitude, abs_mode);
    } else {
        KSP_CUDA(cudaFuncSetAttribute(percentile5_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        percentile5_kernel<false><<<(unsigned) rows, SEL_THREADS, smem, s>>>(
            src, dest, src_stride, dest_stride, first_col, (int) n_cols, is_amplitude, abs_mode);
    }
    KSP_CHECK_LAUNCH();
    return 0;
}
