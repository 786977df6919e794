// Thresholding: ThresholdSimple and the Offringa SumThreshold.
//
// threshold_sum replaces reference rfi/threshold_sum.mako:49-132 (42 barrier-
// separated Kogge-Stone passes over local memory per block for 7 windows).
// Results follow the HOST class (rfi/host.py:218-254) through
// the contract of oracle/contract.c:
//
//   F       = samples flagged by smaller windows,  u[j] = F[j] ? 0 : x[j]
//   D_0 = u,  D_{k+1}[i] = fl32(D_k[i] + D_k[i + 2^k])          (doubling tree)
//   window i of size 2^w fires  <=>  f64(D_w[i]) > f64(thr_w) * (2^w - #F in it)
//
// i.e. the flagged samples (which the reference replaces by thr_w) are moved to
// the right-hand side exactly, so the tree does not depend on w and is built
// ONCE for all window sizes as long as no new flags appear (the usual case
// after the single-sample pass): 1 add + at most 1 compare per sample and
// window size, instead of w adds + substitution.
//
// Mapping: one block per baseline row (rows up to 32 768 channels; longer rows
// in overlapping chunks), one thread per run of 32 consecutive channels.  The
// row is staged once in shared memory with 128-bit coalesced loads (run pitch
// 36 floats: conflict-free 128-bit reads by run).  Each thread then keeps its
// 32 partial sums in registers and its 32 flags in one word; a doubling step
// takes the missing right-hand operands from the next lane with shuffles (the
// next warp's through a 128-byte shared slot).  Per window size a thread first
// asks whether any window that touches its run can fire at all:
//   * no unflagged sample in reach exceeds thr_w            -> nothing can fire
//   * D_w[i] <= thr_w * (2^w - N_max), N_max = flags in reach -> window i cannot
// and only the survivors (rare) are evaluated exactly from shared memory.  When
// a window size > 1 does flag something, the block rebuilds its tree.
#include "common.cuh"
#include "mad.cuh"
#include <stdlib.h>

namespace {

using namespace ksp;

constexpr int RUN = 32;
constexpr int PITCH = 36;
constexpr int TS_MAX_THREADS = 1024;
constexpr int TS_MAX_WINDOWS = 7;   // windows up to 64 = two runs of reach
constexpr unsigned FULL = 0xffffffffu;

// MAD_MODE of the row kernel
constexpr int MAD_NONE = 0;   // thresholds only, noise is an input
constexpr int MAD_FUSED = 1;  // noise estimate (written to noise_out), then thresholds
constexpr int MAD_ONLY = 2;   // noise estimate only

struct TsArgs {
    const float *dev_t;
    const float *noise;
    float *noise_out;
    uint8_t *flags_t;      // byte output (or null)
    uint32_t *bits_t;      // bit-packed output (or null)
    int64_t channels, baselines;
    int64_t dev_stride, out_stride;   // out_stride: bytes per row, or words per row when packed
    int n_windows;
    int flag_value;
    int chunk_valid;       // channels produced per chunk (multiple of 32)
    int edge;              // halo on each side of a chunk (multiple of 32; 0 when 1 chunk)
    double n_sigma;
    double scales[TS_MAX_WINDOWS];
};

// bits j in [lo, hi) of a 32-bit word (any ints)
__device__ __forceinline__ uint32_t bit_range(int64_t lo, int64_t hi)
{
    if (lo < 0) lo = 0;
    if (hi > 32) hi = 32;
    if (hi <= lo) return 0u;
    uint32_t upto_hi = (hi == 32) ? FULL : ((1u << (int) hi) - 1u);
    return upto_hi & ~((1u << (int) lo) - 1u);
}

template <int K>
__device__ __forceinline__ void tree_step(float (&D)[RUN], float *Dex, int lane, int warp,
                                          int nwarps)
{
    constexpr int S = 1 << K;
    if (K == 5) {
        // shift by a whole run: every partial sum needs the same slot of the next thread
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < RUN; i++) Dex[warp * RUN + i] = D[i];
        }
        __syncthreads();
        const bool has_next = warp + 1 < nwarps;
        const float *nx = Dex + (warp + 1) * RUN;
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            float t = __shfl_down_sync(FULL, D[j], 1);
            if (lane == 31) t = has_next ? nx[j] : 0.0f;
            D[j] = D[j] + t;
        }
        __syncthreads();
    } else {
        float nb[S];
#pragma unroll
        for (int i = 0; i < S; i++) nb[i] = __shfl_down_sync(FULL, D[i], 1);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < S; i++) Dex[warp * RUN + i] = D[i];
        }
        __syncthreads();
        if (lane == 31) {
            const bool has_next = warp + 1 < nwarps;
            const float *nx = Dex + (warp + 1) * RUN;
#pragma unroll
            for (int i = 0; i < S; i++) nb[i] = has_next ? nx[i] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < RUN; j++) D[j] = D[j] + ((j + S < RUN) ? D[j + S] : nb[j + S - RUN]);
        __syncthreads();
    }
}

// Exact evaluation of the candidate windows of one thread (rare, divergent).
// Rebuilds D_w[i] in the same tree order from the staged row and the published flags.
__device__ __noinline__ uint32_t exact_windows(uint32_t cand, int run_index, int w, float thr_w,
                                               const float *rowbuf, const uint32_t *Fsm, int span)
{
    const int win = 1 << w;
    uint32_t fire = 0;
    float vals[64];
    while (cand) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1;
        const int p = run_index * RUN + j;
        int n_flagged = 0;
        for (int i = 0; i < win; i++) {
            const int q = p + i;
            float v = 0.0f;
            if (q < span) {
                const bool fl = (Fsm[q >> 5] >> (q & 31)) & 1u;
                n_flagged += fl;
                v = fl ? 0.0f : rowbuf[q + 4 * (q >> 5)];
            }
            vals[i] = v;
        }
        for (int h = 1; h < win; h <<= 1)
            for (int i = 0; i < win; i += 2 * h) vals[i] = vals[i] + vals[i + h];
        if ((double) vals[0] > (double) thr_w * (double) (win - n_flagged)) fire |= 1u << j;
    }
    return fire;
}

template <bool PACKED, int MAD_MODE>
__global__ void __launch_bounds__(TS_MAX_THREADS, 1)
threshold_sum_kernel(const TsArgs a)
{
    extern __shared__ __align__(16) float sm[];
    const int T = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int span = T * RUN;

    constexpr int AUX_WORDS = (MAD_MODE != MAD_NONE) ? 2 * MAD_BINS : 33 * RUN;
    float *rowbuf = sm;                                        // T * PITCH
    float *Dex = rowbuf + T * PITCH;                           // (32 + 1) * RUN; histograms of the MAD
    uint32_t *Fsm = reinterpret_cast<uint32_t *>(Dex + AUX_WORDS);  // T + 2
    float *umx = reinterpret_cast<float *>(Fsm + T + 2);       // T + 2
    uint32_t *car1 = reinterpret_cast<uint32_t *>(umx + T + 2);     // T
    uint32_t *car2 = car1 + T;                                 // T
    float *thr = reinterpret_cast<float *>(car2 + T);          // TS_MAX_WINDOWS (+1)
    uint32_t *mad_misc = reinterpret_cast<uint32_t *>(thr + 8);     // MAD_MISC_WORDS

    const int64_t row = blockIdx.x;
    const int C = (int) a.channels;
    const int base = (int) blockIdx.y * a.chunk_valid - a.edge;     // row channel of slot 0
    const float *src = a.dev_t + row * a.dev_stride;

    if (MAD_MODE == MAD_NONE) {
        if (tid < a.n_windows)
            thr[tid] = __double2float_rn((a.n_sigma * (double) a.noise[row]) * a.scales[tid]);
    } else {
        uint32_t *h = reinterpret_cast<uint32_t *>(Dex);
        for (int i = tid; i < 2 * MAD_BINS; i += T) h[i] = 0u;
        if (tid < 16) mad_misc[tid] = 0u;
    }
    if (tid < 2) {
        Fsm[T + tid] = 0u;
        umx[T + tid] = -__int_as_float(0x7f800000);
    }

    // ---- stage the chunk: coalesced 128-bit loads -> padded runs (zeros outside the band)
    const bool vec_ok = ((a.dev_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.dev_t) & 15) == 0) &&
                        ((base & 3) == 0);
    if (vec_ok && base >= 0 && base + span <= C) {
        // whole span inside the band: no tests, all loads of a thread in flight together
        const float4 *s4 = reinterpret_cast<const float4 *>(src + base) + tid;
        float *dst = rowbuf + (tid >> 3) * PITCH + 4 * (tid & 7);
        const int dstep = (T >> 3) * PITCH;
#pragma unroll
        for (int i = 0; i < RUN / 4; i++) {
            float4 v = __ldg(s4 + i * T);
            *reinterpret_cast<float4 *>(dst + i * dstep) = v;
        }
    } else {
        for (int q = tid; q < (span >> 2); q += T) {
            const int p = q << 2;
            const int g = base + p;
            float4 v;
            if (vec_ok && g >= 0 && g + 3 < C) {
                v = __ldg(reinterpret_cast<const float4 *>(src + g));
            } else {
                v.x = (g >= 0 && g < C) ? src[g] : 0.0f;
                v.y = (g + 1 >= 0 && g + 1 < C) ? src[g + 1] : 0.0f;
                v.z = (g + 2 >= 0 && g + 2 < C) ? src[g + 2] : 0.0f;
                v.w = (g + 3 >= 0 && g + 3 < C) ? src[g + 3] : 0.0f;
            }
            *reinterpret_cast<float4 *>(rowbuf + p + 4 * (p >> 5)) = v;
        }
    }
    __syncthreads();

    // ---- my run
    float D[RUN];
    const float *my = rowbuf + tid * PITCH;
#pragma unroll
    for (int i = 0; i < RUN / 4; i++) {
        float4 v = *reinterpret_cast<const float4 *>(my + 4 * i);
        D[4 * i] = v.x; D[4 * i + 1] = v.y; D[4 * i + 2] = v.z; D[4 * i + 3] = v.w;
    }
    const int64_t pos0 = (int64_t) base + (int64_t) tid * RUN;  // row channel of D[0]
    const uint32_t in_range = bit_range(-pos0, (int64_t) C - pos0);

    if (MAD_MODE != MAD_NONE) {
        // the block holds the whole row (single chunk): noise estimate first
        MadScratch sc;
        sc.hist_a = reinterpret_cast<uint32_t *>(Dex);
        sc.hist_b = sc.hist_a + MAD_BINS;
        sc.misc = mad_misc;
        const float sample = my[(tid * 5 + warp * 3) & 31];
        const float noise = block_mad_noise(D, sample, sc);
        if (tid == 0) a.noise_out[row] = noise;
        if (MAD_MODE == MAD_ONLY) return;
        __syncthreads();       // the histograms alias Dex
        if (tid < a.n_windows)
            thr[tid] = __double2float_rn((a.n_sigma * (double) noise) * a.scales[tid]);
        __syncthreads();
    }

    // ---- window size 1
    uint32_t F = 0;
    float um = -__int_as_float(0x7f800000);
    {
        const float t0 = thr[0];
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const bool f = D[j] > t0;
            F |= f ? (1u << j) : 0u;
            D[j] = f ? 0.0f : D[j];
            um = fmaxf(um, D[j]);
        }
        F &= in_range;
    }
    Fsm[tid] = F;
    umx[tid] = um;
    __syncthreads();

    int built = 0;   // D currently holds D_built
    for (int w = 1; w < a.n_windows; w++) {
        const int win = 1 << w;
        if (win > C) break;
        while (built < w) {
            switch (built) {
            case 0: tree_step<0>(D, Dex, lane, warp, nwarps); break;
            case 1: tree_step<1>(D, Dex, lane, warp, nwarps); break;
            case 2: tree_step<2>(D, Dex, lane, warp, nwarps); break;
            case 3: tree_step<3>(D, Dex, lane, warp, nwarps); break;
            case 4: tree_step<4>(D, Dex, lane, warp, nwarps); break;
            default: tree_step<5>(D, Dex, lane, warp, nwarps); break;
            }
            built++;
        }
        const float tw = thr[w];
        // windows that start in my run and lie inside the band
        const uint32_t valid = bit_range(-pos0, (int64_t) C - (int64_t) win - pos0 + 1);
        uint32_t fire = 0;
        const float reach_max = fmaxf(um, fmaxf(umx[tid + 1], umx[tid + 2]));
        if (valid != 0u && !(reach_max <= __fmul_rd(tw, 0.99999905f))) {
            const uint32_t F1 = Fsm[tid + 1], F2 = Fsm[tid + 2];
            const uint32_t m1 = (win > 32) ? FULL : ((1u << (win - 1)) - 1u);
            const uint32_t m2 = (win > 32) ? 0x7fffffffu : 0u;
            int nmax = __popc(F) + __popc(F1 & m1) + __popc(F2 & m2);
            nmax = min(nmax, win);
            // the bound needs thr_w >= 0; otherwise (negative or NaN) every window is a candidate
            const float tq = (tw >= 0.0f) ? __double2float_rd((double) tw * (double) (win - nmax))
                                          : -__int_as_float(0x7f800000);
            if (!(tw >= 0.0f)) nmax = win;   // force the exact path
            uint32_t cand = 0;
#pragma unroll
            for (int j = 0; j < RUN; j++) cand |= (D[j] > tq) ? (1u << j) : 0u;
            cand &= valid;
            if (nmax == 0)
                fire = cand;
            else if (cand != 0u)
                fire = exact_windows(cand, tid, w, tw, rowbuf, Fsm, span);
        }
        if (__syncthreads_or(fire != 0u)) {
            // spread every firing window over its 2^w samples (96-bit shift-or)
            uint32_t lo = fire, mid = 0, hi = 0;
            for (int k = 0; k < w; k++) {
                const int s = 1 << k;
                if (s < 32) {
                    hi |= __funnelshift_l(mid, hi, s);
                    mid |= __funnelshift_l(lo, mid, s);
                    lo |= lo << s;
                } else {
                    hi |= mid;
                    mid |= lo;
                }
            }
            car1[tid] = mid;
            car2[tid] = hi;
            __syncthreads();
            F |= lo | (tid >= 1 ? car1[tid - 1] : 0u) | (tid >= 2 ? car2[tid - 2] : 0u);
            // rebuild u from the staged row
            um = -__int_as_float(0x7f800000);
#pragma unroll
            for (int i = 0; i < RUN / 4; i++) {
                float4 v = *reinterpret_cast<const float4 *>(my + 4 * i);
                D[4 * i] = v.x; D[4 * i + 1] = v.y; D[4 * i + 2] = v.z; D[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                D[j] = ((F >> j) & 1u) ? 0.0f : D[j];
                um = fmaxf(um, D[j]);
            }
            Fsm[tid] = F;
            umx[tid] = um;
            built = 0;
            __syncthreads();
        }
    }

    // ---- write my 32 flags if my run belongs to this chunk's output range
    const int64_t out_lo = (int64_t) blockIdx.y * a.chunk_valid;
    const int64_t out_hi = min((int64_t) C, out_lo + (int64_t) a.chunk_valid);
    if (pos0 >= out_lo && pos0 < out_hi) {
        F &= in_range;
        if (PACKED) {
            a.bits_t[row * a.out_stride + (pos0 >> 5)] = F;
        } else {
            uint8_t *dst = a.flags_t + row * a.out_stride + pos0;
            const uint32_t fv = (uint32_t) a.flag_value & 0xffu;
            const bool vec = (pos0 + RUN <= C) && ((a.out_stride & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(a.flags_t) & 15) == 0);
            if (vec) {
                uint32_t wds[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t nib = (F >> (4 * i)) & 0xfu;
                    // spread 4 bits to 4 bytes: bit k -> byte k
                    wds[i] = ((nib * 0x00204081u) & 0x01010101u) * fv;
                }
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
            } else {
                for (int j = 0; j < RUN && pos0 + j < C; j++)
                    dst[j] = ((F >> j) & 1u) ? (uint8_t) fv : (uint8_t) 0;
            }
        }
    }
}

// ---------------------------------------------------------------- ThresholdSimple
// Replaces rfi/threshold_simple.mako:27-40 and threshold_simple_t.mako:28-42.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
threshold_simple_kernel(const float *__restrict__ dev, const float *__restrict__ noise,
                        uint8_t *__restrict__ flags, int64_t rows, int64_t cols,
                        int64_t dev_stride, int64_t flags_stride, double n_sigma, int flag_value)
{
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t) blockIdx.y * 64;
    if (c >= cols) return;
    const int64_t r1 = min(rows, r0 + 64);
    float thr = 0.0f;
    if (!TRANSPOSED) thr = __double2float_rn(n_sigma * (double) noise[c]);
    for (int64_t r = r0; r < r1; r++) {
        if (TRANSPOSED) thr = __double2float_rn(n_sigma * (double) noise[r]);
        flags[r * flags_stride + c] = dev[r * dev_stride + c] > thr ? (uint8_t) flag_value : (uint8_t) 0;
    }
}

// ---------------------------------------------------------------- packed flags -> bytes
// bits_t[b * wstride + c/32] (bit c%32) -> flags[c * fstride + b] = bit ? flag_value : 0.
// Tile: 128 baselines x 8 words (256 channels).  Replaces the uchar transpose of
// rfi/device.py:1161-1164 in the fused flagger.
__global__ void __launch_bounds__(256)
expand_flags_kernel(const uint32_t *__restrict__ bits_t, uint8_t *__restrict__ flags,
                    int64_t channels, int64_t baselines, int64_t wstride, int64_t fstride,
                    int flag_value)
{
    __shared__ __align__(16) uint32_t tile[8][132];
    const int64_t b0 = (int64_t) blockIdx.x * 128;
    const int64_t w0 = (int64_t) blockIdx.y * 8;
    const int64_t n_words = (channels + 31) >> 5;
    const int t = threadIdx.x;
    {
        const int w = t & 7;
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
            const int b = (t >> 3) + 32 * pass;
            uint32_t v = 0;
            if (b0 + b < baselines && w0 + w < n_words) v = bits_t[(b0 + b) * wstride + w0 + w];
            tile[w][b] = v;
        }
    }
    __syncthreads();
    const int lane = t & 31, w = t >> 5;   // warp <-> word
    const uint4 q = *reinterpret_cast<const uint4 *>(&tile[w][4 * lane]);
    const uint32_t fv = (uint32_t) flag_value & 0xffu;
    const int64_t c_base = (w0 + w) * 32;
    const int64_t b = b0 + 4 * lane;
    const bool vec = (b + 3 < baselines) && ((fstride & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(flags) & 3) == 0);
#pragma unroll 4
    for (int bit = 0; bit < 32; bit++) {
        const int64_t c = c_base + bit;
        if (c >= channels) break;
        const uint32_t x0 = (q.x >> bit) & 1u, x1 = (q.y >> bit) & 1u, x2 = (q.z >> bit) & 1u,
                       x3 = (q.w >> bit) & 1u;
        if (vec) {
            *reinterpret_cast<uint32_t *>(flags + c * fstride + b) =
                (x0 | (x1 << 8) | (x2 << 16) | (x3 << 24)) * fv;
        } else {
            if (b < baselines) flags[c * fstride + b] = (uint8_t) (x0 * fv);
            if (b + 1 < baselines) flags[c * fstride + b + 1] = (uint8_t) (x1 * fv);
            if (b + 2 < baselines) flags[c * fstride + b + 2] = (uint8_t) (x2 * fv);
            if (b + 3 < baselines) flags[c * fstride + b + 3] = (uint8_t) (x3 * fv);
        }
    }
}

size_t ts_smem_bytes(int threads, bool mad)
{
    const size_t aux = mad ? 2 * MAD_BINS : 33 * RUN;
    return sizeof(float) * ((size_t) threads * PITCH + aux + 4 * ((size_t) threads + 2) + 16 +
                            MAD_MISC_WORDS);
}

template <bool PACKED, int MAD_MODE>
int launch_row_kernel(cudaStream_t s, const TsArgs &a, dim3 grid, int threads)
{
    const bool mad = MAD_MODE != MAD_NONE;
    // (per device and cheap, so simply repeated on every launch)
    KSP_CUDA(cudaFuncSetAttribute(threshold_sum_kernel<PACKED, MAD_MODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int) ts_smem_bytes(TS_MAX_THREADS, mad)));
    threshold_sum_kernel<PACKED, MAD_MODE><<<grid, threads, ts_smem_bytes(threads, mad), s>>>(a);
    KSP_CHECK_LAUNCH();
    return 0;
}

// mad_mode MAD_NONE: thresholds with the given noise.  MAD_FUSED: noise estimate into noise_out,
// then thresholds.  MAD_ONLY: noise estimate only.  The MAD modes need the whole row in one
// block: channels <= TS_MAX_THREADS * RUN (KSP_ETOOLARGE otherwise; callers fall back).
int launch_threshold_sum(cudaStream_t s, const float *dev_t, const float *noise, float *noise_out,
                         uint8_t *flags_t, uint32_t *bits_t, int64_t channels, int64_t baselines,
                         int64_t dev_stride, int64_t out_stride, int n_windows, double n_sigma,
                         const double *scales, int flag_value, int mad_mode)
{
    if (channels < 0 || baselines < 0 || dev_stride < channels) return KSP_EINVAL;
    if (mad_mode != MAD_ONLY) {
        if (n_windows < 1 || !scales) return KSP_EINVAL;
        if (n_windows > TS_MAX_WINDOWS) return KSP_ETOOLARGE;
    }
    if (channels > 0x7fff0000) return KSP_ETOOLARGE;
    if (baselines == 0) return 0;
    if (channels == 0 && mad_mode == MAD_NONE) return 0;
    if (!dev_t) return KSP_EINVAL;
    if (mad_mode == MAD_NONE && !noise) return KSP_EINVAL;
    if (mad_mode != MAD_NONE && !noise_out) return KSP_EINVAL;
    if (mad_mode != MAD_ONLY && !flags_t && !bits_t) return KSP_EINVAL;
    if (baselines > 0x7fffffff) return KSP_ETOOLARGE;

    TsArgs a;
    a.dev_t = dev_t; a.noise = noise; a.noise_out = noise_out; a.flags_t = flags_t; a.bits_t = bits_t;
    a.channels = channels; a.baselines = baselines;
    a.dev_stride = dev_stride; a.out_stride = out_stride;
    a.n_windows = n_windows; a.flag_value = flag_value; a.n_sigma = n_sigma;
    for (int w = 0; w < TS_MAX_WINDOWS; w++) a.scales[w] = (scales && w < n_windows) ? scales[w] : 0.0;

    int threads, n_chunks;
    const int64_t runs = ksp_divup(channels, RUN);
    // Threads per block for long rows when only thresholding.  Blocks of fewer threads
    // (spans that overlap by the reach of the largest window) keep several blocks resident
    // per SM, which hides barrier and shared-memory latency.
    static const int chunk_threads = [] {
        const char *e = getenv("KSP_TS_THREADS");
        int v = e ? atoi(e) : 128;
        if (v < 32 || v > TS_MAX_THREADS || (v & 31)) v = 128;
        return v;
    }();
    const int max_threads = (mad_mode == MAD_NONE) ? chunk_threads : TS_MAX_THREADS;
    if (runs <= max_threads) {
        threads = (int) (ksp_divup(runs > 0 ? runs : 1, 32) * 32);
        a.edge = 0;
        a.chunk_valid = threads * RUN;
        n_chunks = 1;
    } else {
        if (mad_mode != MAD_NONE) return KSP_ETOOLARGE;
        threads = max_threads;
        const int reach = (1 << n_windows) - n_windows - 1;   // influence radius of a sample
        a.edge = (int) (ksp_divup(reach, RUN) * RUN);
        a.chunk_valid = threads * RUN - 2 * a.edge;
        n_chunks = (int) ksp_divup(channels, a.chunk_valid);
    }
    if (n_chunks > 65535) return KSP_ETOOLARGE;
    dim3 grid((unsigned) baselines, (unsigned) n_chunks);
    if (mad_mode == MAD_ONLY) return launch_row_kernel<false, MAD_ONLY>(s, a, grid, threads);
    if (mad_mode == MAD_FUSED)
        return bits_t ? launch_row_kernel<true, MAD_FUSED>(s, a, grid, threads)
                      : launch_row_kernel<false, MAD_FUSED>(s, a, grid, threads);
    return bits_t ? launch_row_kernel<true, MAD_NONE>(s, a, grid, threads)
                  : launch_row_kernel<false, MAD_NONE>(s, a, grid, threads);
}

}  // namespace

extern "C" int ksp_threshold_sum(void *stream, const float *dev_t, const float *noise,
                                 uint8_t *flags_t, int64_t channels, int64_t baselines,
                                 int64_t dev_stride, int64_t flags_stride, int n_windows,
                                 double n_sigma, const double *scales, int flag_value)
{
    if (flags_stride < channels) return KSP_EINVAL;
    return launch_threshold_sum((cudaStream_t) stream, dev_t, noise, nullptr, flags_t, nullptr,
                                channels, baselines, dev_stride, flags_stride, n_windows, n_sigma,
                                scales, flag_value, MAD_NONE);
}

// internal (fused flagger): bit-packed output, words_stride words per baseline row
int ksp_threshold_sum_packed(cudaStream_t s, const float *dev_t, const float *noise,
                             uint32_t *bits_t, int64_t channels, int64_t baselines,
                             int64_t dev_stride, int64_t words_stride, int n_windows,
                             double n_sigma, const double *scales)
{
    if (words_stride < ksp_divup(channels, 32)) return KSP_EINVAL;
    return launch_threshold_sum(s, dev_t, noise, nullptr, nullptr, bits_t, channels, baselines,
                                dev_stride, words_stride, n_windows, n_sigma, scales, 1, MAD_NONE);
}

// internal (fused flagger): noise estimate + thresholds in one launch, one block per row.
// KSP_ETOOLARGE if a row does not fit one block (the caller then uses the separate kernels).
int ksp_noise_threshold_packed(cudaStream_t s, const float *dev_t, float *noise, uint32_t *bits_t,
                               int64_t channels, int64_t baselines, int64_t dev_stride,
                               int64_t words_stride, int n_windows, double n_sigma,
                               const double *scales)
{
    if (words_stride < ksp_divup(channels, 32)) return KSP_EINVAL;
    return launch_threshold_sum(s, dev_t, nullptr, noise, nullptr, bits_t, channels, baselines,
                                dev_stride, words_stride, n_windows, n_sigma, scales, 1, MAD_FUSED);
}

// internal (ksp_madnz_t): noise estimate only, rows of up to TS_MAX_THREADS * RUN channels
int ksp_row_mad(cudaStream_t s, const float *dev_t, float *noise, int64_t channels,
                int64_t baselines, int64_t dev_stride)
{
    return launch_threshold_sum(s, dev_t, nullptr, noise, nullptr, nullptr, channels, baselines,
                                dev_stride, 0, 0, 0.0, nullptr, 1, MAD_ONLY);
}

int ksp_expand_flags(cudaStream_t s, const uint32_t *bits_t, uint8_t *flags, int64_t channels,
                     int64_t baselines, int64_t words_stride, int64_t flags_stride, int flag_value)
{
    if (channels == 0 || baselines == 0) return 0;
    dim3 grid((unsigned) ksp_divup(baselines, 128), (unsigned) ksp_divup(ksp_divup(channels, 32), 8));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    expand_flags_kernel<<<grid, 256, 0, s>>>(bits_t, flags, channels, baselines, words_stride,
                                             flags_stride, flag_value);
    KSP_CHECK_LAUNCH();
    return 0;
}

extern "C" int ksp_threshold_simple(void *stream, const float *dev, const float *noise,
                                    uint8_t *flags, int64_t rows, int64_t cols, int64_t dev_stride,
                                    int64_t flags_stride, double n_sigma, int flag_value,
                                    int transposed)
{
    if (rows < 0 || cols < 0 || dev_stride < cols || flags_stride < cols) return KSP_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!dev || !noise || !flags) return KSP_EINVAL;
    dim3 grid((unsigned) ksp_divup(cols, 256), (unsigned) ksp_divup(rows, 64));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    cudaStream_t s = (cudaStream_t) stream;
    if (transposed)
        threshold_simple_kernel<true><<<grid, 256, 0, s>>>(dev, noise, flags, rows, cols, dev_stride,
                                                           flags_stride, n_sigma, flag_value);
    else
        threshold_simple_kernel<false><<<grid, 256, 0, s>>>(dev, noise, flags, rows, cols,
                                                            dev_stride, flags_stride, n_sigma,
                                                            flag_value);
    KSP_CHECK_LAUNCH();
    return 0;
}
