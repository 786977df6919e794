// Median-of-absolute-deviations noise estimate and Percentile5.
//
// madnz_t  : replaces reference rfi/madnz_t.mako:72-87 (+ rank.mako:236-266).
// madnz    : replaces reference rfi/madnz.mako:105-123 (channel-major input).
// percentile5 : replaces reference percentile.mako:115-140.
//
// All three are exact selections.  madnz_t and percentile5 stream each
// row ONCE through a 256-thread block: a bracket around the wanted rank(s) from 1024 samples,
// one pass that counts the keys below the bracket and keeps the few inside it, and a small
// selection among those (madnz_stream_kernel, percentile5_stream_kernel below).  The
// block-wide radix select of select.cuh is the fallback for rows whose bracket misses and for
// very long Percentile5 rows.  madnz (channel-major) gives one block per 32 baselines,
// lane == baseline, and makes 4 radix passes over global memory with per-baseline histograms
// (hist[digit][lane]: conflict-free).
#include "common.cuh"
#include "select.cuh"
#include <math.h>
#include <stdlib.h>

#include "madnz_stream.cuh"

namespace {

// rows that fell back to the radix select (diagnostic, see ksp_selection_fallback_count)
__device__ unsigned long long g_selection_fallbacks;

__global__ void __launch_bounds__(MS_THREADS, 4)
madnz_stream_kernel(const float *__restrict__ dev_t, float *__restrict__ noise, int channels,
                    int64_t stride)
{
    __shared__ __align__(16) uint32_t smem[MS_SMEM_WORDS];
    const bool vec_ok = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(dev_t) & 15) == 0);
    madnz_stream_row<false>(dev_t + (int64_t) blockIdx.x * stride, noise + blockIdx.x, channels, vec_ok,
                            smem, &g_selection_fallbacks);
}

// ------------------------------------------------------------------ channel-major MAD
// block = 32 warps x 32 lanes; lane == baseline, warps stride over channels.
__global__ void __launch_bounds__(1024, 1)
madnz_cm_kernel(const float *__restrict__ dev, float *__restrict__ noise, int channels,
                int baselines, int64_t stride)
{
    __shared__ uint32_t hist[256 * 32];   // hist[digit][lane]
    __shared__ uint32_t part[32][33];     // per-warp partials
    __shared__ uint32_t s_prefix[32], s_rank[32], s_nvalid[32], s_next[32];

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

// Channel-major MAD in three passes (rows of at most 65535 channels): radix select with digits
// of 11, 11 and 9 bits instead of four of 8, the count of usable samples folded into the first
// pass and the "next key above" of even counts into the last.  The histogram
// hist[digit][baseline] holds 2048 x 32 counters as 16-bit halves of 1024 x 32 words
// (digit d -> word row d & 1023, half d >> 10; the column is skewed by the row so that both the
// atomics of a warp - lane == baseline - and the resolving warp's reads spread over the banks).
constexpr int CM3_ROWS = 1024;

// rank lookup for baseline `bl` by one warp: walks the digits 32 at a time (digit = 32 i + lane).
// Returns the digit holding `rank`, the rank inside it, its count and the next non-empty digit
// above it (n_digits if none).
struct Cm3Hit { uint32_t digit, r_in, count, next; };
__device__ __forceinline__ Cm3Hit cm3_locate(const uint32_t *hist, int bl, uint32_t rank, int n_digits,
                                             int lane)
{
    Cm3Hit h = {0u, 0u, 0u, (uint32_t) n_digits};
    bool found = false;
    uint32_t remaining = rank;
    for (int i = 0; i < n_digits / 32; i++) {
        const int d = 32 * i + lane;
        const int r = d & (CM3_ROWS - 1);
        const uint32_t word = hist[r * 32 + ((bl + r) & 31)];
        const uint32_t c = (d >> 10) ? (word >> 16) : (word & 0xffffu);
        if (!found) {
            const uint32_t incl = warp_scan_incl(c, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (remaining < total) {
                const int src = __ffs(__ballot_sync(0xffffffffu, remaining < incl)) - 1;
                h.digit = (uint32_t) (32 * i + src);
                h.count = __shfl_sync(0xffffffffu, c, src);
                h.r_in = remaining - (__shfl_sync(0xffffffffu, incl, src) - h.count);
                found = true;
                // next non-empty digit inside this group of 32
                const uint32_t above = __ballot_sync(0xffffffffu, c != 0u && lane > src);
                if (above) {
                    h.next = (uint32_t) (32 * i + __ffs(above) - 1);
                    break;
                }
            } else {
                remaining -= total;
            }
        } else {
            const uint32_t nz = __ballot_sync(0xffffffffu, c != 0u);
            if (nz) {
                h.next = (uint32_t) (32 * i + __ffs(nz) - 1);
                break;
            }
        }
    }
    return h;
}

__global__ void __launch_bounds__(1024, 1)
madnz_cm3_kernel(const float *__restrict__ dev, float *__restrict__ noise, int channels,
                 int baselines, int64_t stride)
{
    extern __shared__ uint32_t cm_hist[];              // CM3_ROWS x 32 words
    __shared__ uint32_t part[32][33];                  // per-warp partials
    __shared__ uint32_t s_prefix[32], s_rank[32], s_nvalid[32], s_rin[32], s_cnt[32], s_next[32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b_raw = blockIdx.x * 32 + lane;
    const bool ok = b_raw < baselines;
    const int b = ok ? b_raw : baselines - 1;
    const float *col = dev + b;

    auto clear = [&]() {
        for (int i = threadIdx.x; i < CM3_ROWS * 32; i += 1024) cm_hist[i] = 0u;
    };
    auto bump = [&](uint32_t d) {
        const uint32_t r = d & (CM3_ROWS - 1);
        atomicAdd(&cm_hist[r * 32 + ((lane + r) & 31)], (d >> 10) ? 0x10000u : 1u);
    };

    // ---- pass 1: usable count and the top 11 bits (keys are at most 0x7f800000: digit < 2048)
    clear();
    __syncthreads();
    {
        uint32_t cnt = 0;
        int c = warp;
        for (; c + 96 < channels; c += 128) {
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = mad_key(col[(int64_t) (c + 32 * u) * stride]);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k[u] != KEY_SKIP) { cnt++; bump(k[u] >> 20); }
        }
        for (; c < channels; c += 32) {
            const uint32_t k = mad_key(col[(int64_t) c * stride]);
            if (k != KEY_SKIP) { cnt++; bump(k >> 20); }
        }
        part[warp][lane] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 32; w++) t += part[w][lane];
        s_nvalid[lane] = t;
        s_rank[lane] = t ? (t - 1) >> 1 : 0;
    }
    __syncthreads();
    {
        const int bl = warp;                            // warp w resolves baseline w
        if (s_nvalid[bl] != 0) {
            const Cm3Hit h = cm3_locate(cm_hist, bl, s_rank[bl], 2048, lane);
            if (lane == 0) { s_prefix[bl] = h.digit << 20; s_rank[bl] = h.r_in; }
        } else if (lane == 0) {
            s_prefix[bl] = 0u;
        }
    }
    __syncthreads();

    // ---- pass 2: the next 11 bits among the keys that share the first digit
    clear();
    __syncthreads();
    {
        const uint32_t want = s_prefix[lane] >> 20;
        int c = warp;
        for (; c + 96 < channels; c += 128) {
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = mad_key(col[(int64_t) (c + 32 * u) * stride]);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k[u] != KEY_SKIP && (k[u] >> 20) == want) bump((k[u] >> 9) & 0x7ffu);
        }
        for (; c < channels; c += 32) {
            const uint32_t k = mad_key(col[(int64_t) c * stride]);
            if (k != KEY_SKIP && (k >> 20) == want) bump((k >> 9) & 0x7ffu);
        }
    }
    __syncthreads();
    {
        const int bl = warp;
        if (s_nvalid[bl] != 0) {
            const Cm3Hit h = cm3_locate(cm_hist, bl, s_rank[bl], 2048, lane);
            if (lane == 0) { s_prefix[bl] |= h.digit << 9; s_rank[bl] = h.r_in; }
        }
    }
    __syncthreads();

    // ---- pass 3: the last 9 bits; also the smallest key beyond the 22-bit prefix
    clear();
    __syncthreads();
    {
        const uint32_t want = s_prefix[lane] >> 9;
        uint32_t beyond = KEY_SKIP;
        int c = warp;
        for (; c + 96 < channels; c += 128) {
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = mad_key(col[(int64_t) (c + 32 * u) * stride]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k[u] == KEY_SKIP) continue;
                const uint32_t p = k[u] >> 9;
                if (p == want) bump(k[u] & 0x1ffu);
                else if (p > want) beyond = min(beyond, k[u]);
            }
        }
        for (; c < channels; c += 32) {
            const uint32_t k = mad_key(col[(int64_t) c * stride]);
            if (k == KEY_SKIP) continue;
            const uint32_t p = k >> 9;
            if (p == want) bump(k & 0x1ffu);
            else if (p > want) beyond = min(beyond, k);
        }
        part[warp][lane] = beyond;
    }
    __syncthreads();
    {
        const int bl = warp;
        if (s_nvalid[bl] != 0) {
            const Cm3Hit h = cm3_locate(cm_hist, bl, s_rank[bl], 512, lane);
            if (lane == 0) {
                const uint32_t p22 = s_prefix[bl];
                s_prefix[bl] = p22 | h.digit;
                s_rin[bl] = h.r_in;
                s_cnt[bl] = h.count;
                s_next[bl] = h.next < 512u ? (p22 | h.next) : KEY_SKIP;
            }
        }
    }
    __syncthreads();
    if (warp == 0 && ok) {
        const uint32_t n = s_nvalid[lane];
        float out = __int_as_float(0x7fc00000);
        if (n) {
            const uint32_t lo = s_prefix[lane];
            uint32_t hi = lo;
            // even count: the upper median is another copy of lo unless lo is the last of its value
            if (!(n & 1u) && s_rin[lane] + 1u == s_cnt[lane]) {
                hi = s_next[lane];
                if (hi == KEY_SKIP) {
                    for (int w = 0; w < 32; w++) hi = min(hi, part[w][lane]);
                }
            }
            out = mad_finish(lo, hi);
        }
        noise[b_raw] = out;
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

// ------------------------------------------------------------------ streaming Percentile5
// Same idea as madnz_stream_kernel, for three ranks at once: brackets around the 25 %, 50 % and
// 75 % ranks from 1024 samples of the row, ONE pass over the row (amplitude computed once per
// element) that tracks min / max, counts the keys below each bracket and keeps the keys inside
// any bracket (~36 %) in thread-private lists, then three small selections inside the lists.
// A rank whose bracket misses it, or a list that overflows, falls back to the radix select
// over global memory for that rank.  Replaces reference percentile.mako:115-140.
constexpr int P5_THREADS = 256;
constexpr int P5_MIN_SLOTS = 24;

struct P5Brackets {
    uint32_t lo[3], width[3];     // bracket k = keys in [lo, lo + width)
};

template <int MODE>   // 0: float32 amplitudes, 1: complex64 numpy rule, 2: complex64 hypot rule
__device__ __forceinline__ uint32_t p5_key(const void *src, int64_t index)
{
    float v;
    if (MODE == 0) {
        v = fabsf(reinterpret_cast<const float *>(src)[index]);
    } else {
        const float2 z = reinterpret_cast<const float2 *>(src)[index];
        v = abs_c64<MODE == 1 ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(z.x, z.y);
    }
    return float_to_key(v);
}

template <int MODE>
__global__ void __launch_bounds__(P5_THREADS)
percentile5_stream_kernel(const void *__restrict__ src, float *__restrict__ dest, int64_t src_stride,
                          int64_t dest_stride, int64_t first_col, int n, int slots)
{
    extern __shared__ __align__(16) uint32_t p5_smem[];
    uint32_t *lists = p5_smem;                               // slots * P5_THREADS (>= SELECT_HIST_WORDS)
    uint32_t *hist = lists + (size_t) slots * P5_THREADS;    // MS_BINS
    uint32_t *misc = hist + MS_BINS;                         // 320 words
    // misc: 0..2 below[k], 3 overflow, 4 kmin, 5 kmax, 6 bin, 7 before bin, 8 in bin, 9 small count,
    //       10..15 bounds (lo0 hi0 lo1 hi1 lo2 hi2), 16..63 scan / fallback scratch,
    //       64..255 group picks (6 x 32), 256..287 small list
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = P5_THREADS / 32;
    const int64_t off = (int64_t) blockIdx.x * src_stride + first_col;
    const uint32_t nn = (uint32_t) n;
    const uint32_t ranks[3] = {(nn - 1) / 4, ((nn - 1) * 3) / 4, (nn - 1) / 2};   // 25 %, 75 %, 50 %

    if (tid < 16) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;

    // ---- 1. brackets
    {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int i = tid * 4 + g;
            const int pos = (int) (((int64_t) i * n) >> 10);
            const uint32_t key = p5_key<MODE>(src, off + min(pos, n - 1));
            const uint32_t sorted = sort32(key, lane);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // group quantiles 6 % either side of the wanted rank (of the 32 samples of a group)
                const float q = (float) ranks[k] / (float) max(nn - 1u, 1u);
                const int i_lo = (int) floorf((q - 0.06f) * 32.0f);
                const int i_hi = (int) ceilf((q + 0.06f) * 32.0f);
                const uint32_t lo_g = __shfl_sync(0xffffffffu, sorted, max(i_lo, 0));
                const uint32_t hi_g = __shfl_sync(0xffffffffu, sorted, min(i_hi, 31));
                if (lane == 0) {
                    misc[64 + (2 * k) * 32 + warp * 4 + g] = (i_lo < 0) ? 0u : lo_g;
                    misc[64 + (2 * k + 1) * 32 + warp * 4 + g] = (i_hi > 31) ? 0xffffffffu : hi_g;
                }
            }
        }
        __syncthreads();
        if (warp < 6) {
            const uint32_t v = misc[64 + 32 * warp + lane];
            const uint32_t s = sort32(v, lane);
            // lower bounds take the lower median of the groups, upper bounds the upper one
            const uint32_t pick = __shfl_sync(0xffffffffu, s, (warp & 1) ? 16 : 15);
            if (lane == 0) misc[10 + warp] = pick;
        }
        __syncthreads();
    }
    P5Brackets br;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        br.lo[k] = misc[10 + 2 * k];
        const uint32_t hi = max(misc[11 + 2 * k], br.lo[k]);
        br.width[k] = (hi == 0xffffffffu && br.lo[k] == 0u) ? 0xffffffffu : hi - br.lo[k] + 1u;
    }

    // ---- 2. one pass over the row
    uint32_t n_mine = 0;
    {
        uint32_t below0 = 0, below1 = 0, below2 = 0, kmin = 0xffffffffu, kmax = 0u;
        uint32_t *slot = lists + tid;
        for (int i = tid; i < n; i += P5_THREADS) {
            const uint32_t key = p5_key<MODE>(src, off + i);
            kmin = min(kmin, key);
            kmax = max(kmax, key);
            below0 += (key < br.lo[0]) ? 1u : 0u;
            below1 += (key < br.lo[1]) ? 1u : 0u;
            below2 += (key < br.lo[2]) ? 1u : 0u;
            const bool in = (key - br.lo[0] < br.width[0]) || (key - br.lo[1] < br.width[1]) ||
                            (key - br.lo[2] < br.width[2]);
            if (in) {
                if (n_mine < (uint32_t) slots) *slot = key;
                slot += P5_THREADS;
                n_mine++;
            }
        }
        below0 = __reduce_add_sync(0xffffffffu, below0);
        below1 = __reduce_add_sync(0xffffffffu, below1);
        below2 = __reduce_add_sync(0xffffffffu, below2);
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        const bool over = __any_sync(0xffffffffu, n_mine > (uint32_t) slots);
        if (lane == 0) {
            atomicAdd(&misc[0], below0);
            atomicAdd(&misc[1], below1);
            atomicAdd(&misc[2], below2);
            atomicMin(&misc[4], kmin);
            atomicMax(&misc[5], kmax);
            if (over) misc[3] = 1u;
        }
    }
    __syncthreads();
    const bool overflow = misc[3] != 0u;
    const uint32_t *mine = lists + tid;

    // ---- 3. the three selections
    uint32_t found[3];
    bool redo[3];
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const uint32_t lo = br.lo[k], width = br.width[k];
        const uint32_t r_rel = ranks[k] - misc[k];               // wraps if the rank is below the bracket
        redo[k] = overflow;
        found[k] = 0;
        if (overflow) continue;                                  // block-uniform
        for (int i = tid; i < MS_BINS; i += P5_THREADS) hist[i] = 0u;
        if (tid == 0) misc[9] = 0u;
        __syncthreads();
        const int shift = (width <= (uint32_t) MS_BINS) ? 0 : (32 - __clz(width - 1u)) - 11;
        for (uint32_t j = 0; j < n_mine; j++) {
            const uint32_t d = mine[j * P5_THREADS] - lo;
            if (d < width) atomicAdd(&hist[d >> shift], 1u);
        }
        __syncthreads();
        {
            constexpr int BPT = MS_BINS / P5_THREADS;
            uint32_t h[BPT], own = 0;
#pragma unroll
            for (int b = 0; b < BPT; b++) {
                h[b] = hist[tid * BPT + b];
                own += h[b];
            }
            uint32_t incl = own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) misc[16 + warp] = incl;
            __syncthreads();
            const uint32_t w = (lane < NWARPS) ? misc[16 + lane] : 0u;
            const uint32_t total = __reduce_add_sync(0xffffffffu, w);
            uint32_t cum = __reduce_add_sync(0xffffffffu, lane < warp ? w : 0u) + incl - own;
            if (tid == 0) misc[8] = 0xffffffffu;                 // "not found" unless a thread finds it
            __syncthreads();
            if (r_rel < total && r_rel >= cum && r_rel < cum + own) {
#pragma unroll
                for (int b = 0; b < BPT; b++) {
                    if (r_rel >= cum && r_rel < cum + h[b]) {
                        misc[6] = (uint32_t) (tid * BPT + b);
                        misc[7] = cum;
                        misc[8] = h[b];
                    }
                    cum += h[b];
                }
            }
        }
        __syncthreads();
        const uint32_t bin = misc[6], in_bin = misc[8];
        if (in_bin == 0xffffffffu || in_bin > (uint32_t) MS_SMALL_CAP) {
            redo[k] = true;                                      // bracket missed or crowded bin
            continue;
        }
        const uint32_t r_bin = r_rel - misc[7];
        if (shift == 0) {
            found[k] = lo + bin;
            continue;
        }
        for (uint32_t j = 0; j < n_mine; j++) {
            const uint32_t key = mine[j * P5_THREADS];
            const uint32_t d = key - lo;
            if (d < width && (d >> shift) == bin) misc[256 + atomicAdd(&misc[9], 1u)] = key;
        }
        __syncthreads();
        if (warp == 0) {
            const uint32_t sorted = sort32(lane < (int) in_bin ? misc[256 + lane] : 0xffffffffu, lane);
            const uint32_t pick = __shfl_sync(0xffffffffu, sorted, (int) r_bin);
            if (lane == 0) misc[16] = pick;
        }
        __syncthreads();
        found[k] = misc[16];
        __syncthreads();
    }
    // ---- fallbacks: radix select over the row in global memory
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        if (!redo[k]) continue;                                  // block-uniform
        __syncthreads();
        SelectScratch sc;
        sc.hist = lists;
        sc.misc = misc + 16;
        auto key_at = [src, off](int i) { return p5_key<MODE>(src, off + i); };
        found[k] = block_radix_select<P5_THREADS>(key_at, n, ranks[k], sc);
        if (tid == 0) atomicAdd(&g_selection_fallbacks, 1ull);
    }
    if (tid == 0) {
        const int64_t r = blockIdx.x;
        dest[0 * dest_stride + r] = key_to_float(misc[4]);
        dest[1 * dest_stride + r] = key_to_float(misc[5]);
        dest[2 * dest_stride + r] = key_to_float(found[0]);
        dest[3 * dest_stride + r] = key_to_float(found[1]);
        dest[4 * dest_stride + r] = key_to_float(found[2]);
    }
}

// ------------------------------------------------------------------ Percentile5, fast path
// The techniques of madnz_stream_kernel applied to three ranks at once, for rows that can be
// read with aligned 16-byte loads (the usual case):
//   1. 2048 x 16 bytes of the row are sampled (8192 amplitudes, or 4096 from complex input)
//      into a two-level histogram over the float bit patterns; every warp locates the sample
//      quantiles h either side of the 25 %, 50 % and 75 % ranks (h = 4.5 standard errors)
//      and interpolates inside their bins: three disjoint brackets [lo_k, hi_k];
//   2. ONE pass over the row with 16-byte loads, branch-free per element: min, max, a NaN
//      detector, three predicated counters (a >= lo_k) and a predicated append of the
//      elements inside any bracket (~20 %) to thread-private lists;
//   3. one walk over the lists fills a two-level histogram per bracket, every warp locates the
//      three wanted bins, a second walk collects their keys, warps 0..2 sort them.
// Rows with NaN, overlapping brackets (heavy ties), a missed bracket, a crowded bin or a full
// list fall back to the radix select over global memory, rank by rank.
constexpr int PF_THREADS = 256;
constexpr int PF_FINE = 1024;                 // histogram bins per bracket: 32 coarse x 32
constexpr int PF_UNROLL = 4;                  // 16-byte loads in flight per thread
#ifndef PF_SAMPLE_VECS
#define PF_SAMPLE_VECS 8                      // 16-byte sample loads per thread (4 or 8)
#endif

// fine bin / rank lookup in a 32 x 32 two-level histogram (lane owns coarse bin `lane`)
__device__ __forceinline__ BinHit locate_rank32(const uint32_t *coarse, const uint32_t *fine,
                                                uint32_t r, int lane, uint32_t &total)
{
    const uint32_t c = coarse[lane];
    const uint32_t c_incl = warp_scan_incl(c, lane), c_excl = c_incl - c;
    total = __shfl_sync(0xffffffffu, c_incl, 31);
    BinHit h = {0u, 0u, 0u};
    if (r >= total) return h;                                  // warp-uniform: bracket missed
    const int src = __ffs(__ballot_sync(0xffffffffu, r >= c_excl && r < c_incl)) - 1;
    const uint32_t r_c = r - __shfl_sync(0xffffffffu, c_excl, src);
    const uint32_t f = fine[src * 32 + lane];
    const uint32_t f_incl = warp_scan_incl(f, lane), f_excl = f_incl - f;
    const int src2 = __ffs(__ballot_sync(0xffffffffu, r_c >= f_excl && r_c < f_incl)) - 1;
    h.bin = (uint32_t) (src * 32 + src2);
    h.count = __shfl_sync(0xffffffffu, f, src2);
    h.r_in = r_c - __shfl_sync(0xffffffffu, f_excl, src2);
    return h;
}

// one element of the pass (amplitude a >= 0, not NaN): three counters, append if inside a bracket
__device__ __forceinline__ void pf_step(float a, const float (&lo)[3], const float (&hi)[3],
                                        uint32_t &ge0, uint32_t &ge1, uint32_t &ge2, uint32_t &slot)
{
    asm volatile(
        "{\n\t"
        ".reg .pred g0, g1, g2, i0, i1, i2;\n\t"
        "setp.ge.f32 g0, %4, %5;\n\t"
        "setp.ge.f32 g1, %4, %7;\n\t"
        "setp.ge.f32 g2, %4, %9;\n\t"
        "setp.le.and.f32 i0, %4, %6, g0;\n\t"
        "setp.le.and.f32 i1, %4, %8, g1;\n\t"
        "setp.le.and.f32 i2, %4, %10, g2;\n\t"
        "or.pred i0, i0, i1;\n\t"
        "or.pred i0, i0, i2;\n\t"
        "@g0 add.u32 %0, %0, 1;\n\t"
        "@g1 add.u32 %1, %1, 1;\n\t"
        "@g2 add.u32 %2, %2, 1;\n\t"
        "@i0 st.shared.f32 [%3], %4;\n\t"
        "@i0 add.u32 %3, %3, %11;\n\t"
        "}"
        : "+r"(ge0), "+r"(ge1), "+r"(ge2), "+r"(slot)
        : "f"(a), "f"(lo[0]), "f"(hi[0]), "f"(lo[1]), "f"(hi[1]), "f"(lo[2]), "f"(hi[2]),
          "n"(PF_THREADS * 4));
}

__device__ __forceinline__ void pf_step_checked(float a, const float (&lo)[3], const float (&hi)[3],
                                                uint32_t &ge0, uint32_t &ge1, uint32_t &ge2,
                                                uint32_t &slot, uint32_t slot_end, bool &over)
{
    ge0 += (a >= lo[0]) ? 1u : 0u;
    ge1 += (a >= lo[1]) ? 1u : 0u;
    ge2 += (a >= lo[2]) ? 1u : 0u;
    if ((a >= lo[0] && a <= hi[0]) || (a >= lo[1] && a <= hi[1]) || (a >= lo[2] && a <= hi[2])) {
        if (slot < slot_end) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot), "f"(a));
            slot += PF_THREADS * 4;
        } else {
            over = true;
        }
    }
}

template <int MODE>   // 0: float32 amplitudes, 1: complex64 numpy rule, 2: complex64 hypot rule
__global__ void __launch_bounds__(PF_THREADS, 3)
percentile5_fast_kernel(const void *__restrict__ src, float *__restrict__ dest, int64_t src_stride,
                        int64_t dest_stride, int64_t first_col, int n, int slots, float half_width)
{
    constexpr int EPV = (MODE == 0) ? 4 : 2;                  // elements per 16-byte load
    constexpr int ESZ = (MODE == 0) ? 4 : 8;
    extern __shared__ __align__(16) uint32_t pf_smem[];
    uint32_t *lists = pf_smem;                                // slots * 256 words
    uint32_t *hist = lists + (size_t) slots * PF_THREADS;     // 3 * PF_FINE
    uint32_t *coarse = hist + 3 * PF_FINE;                    // 3 * 32
    uint32_t *small = coarse + 3 * 32;                        // 3 * 32
    uint32_t *misc = small + 3 * 32;                          // 128 words
    // misc: 0..2 counts of a >= lo_k, 3 list overflow, 4 min bits, 5 max bits, 6 NaN seen,
    //       7..9 small-list counts, 10..12 results, 64..127 fallback scratch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = (int64_t) blockIdx.x * src_stride + first_col;
    const char *row = reinterpret_cast<const char *>(src) + off * ESZ;
    const uint32_t nn = (uint32_t) n;
    const uint32_t ranks[3] = {(nn - 1) / 4, (nn - 1) / 2, ((nn - 1) * 3) / 4};   // ascending

    auto amplitudes = [](const float4 v, float (&a)[4]) {
        if (MODE == 0) {
            a[0] = fabsf(v.x); a[1] = fabsf(v.y); a[2] = fabsf(v.z); a[3] = fabsf(v.w);
        } else {
            a[0] = abs_c64<MODE == 1 ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(v.x, v.y);
            a[1] = abs_c64<MODE == 1 ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(v.z, v.w);
            a[2] = a[3] = 0.0f;
        }
    };
    auto scalar_amplitude = [row](int i) -> float {
        if (MODE == 0) return fabsf(reinterpret_cast<const float *>(row)[i]);
        const float2 z = reinterpret_cast<const float2 *>(row)[i];
        return abs_c64<MODE == 1 ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(z.x, z.y);
    };

    // ---- 1. brackets from a sample histogram (borrows the list memory)
    uint32_t *s_fine = lists, *s_coarse = lists + MS_BINS;
    for (int i = tid; i < MS_BINS; i += PF_THREADS) s_fine[i] = 0u;
    for (int i = tid; i < 3 * PF_FINE; i += PF_THREADS) hist[i] = 0u;
    if (tid < MS_BINS / 32) s_coarse[tid] = 0u;
    if (tid < 3 * 32) coarse[tid] = 0u;
    if (tid < 16) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;
    const int n_vec = n / EPV;
    const float4 *row4 = reinterpret_cast<const float4 *>(row);
    float4 sv[PF_SAMPLE_VECS];
    {
        constexpr int LOG_N = 8 + (PF_SAMPLE_VECS == 8 ? 3 : 2);             // log2(sample vectors)
        const uint32_t step = (uint32_t) n_vec >> LOG_N;
#pragma unroll
        for (int g = 0; g < PF_SAMPLE_VECS; g++) {
            const uint32_t i = (uint32_t) tid * PF_SAMPLE_VECS + g;
            uint32_t pos = (uint32_t) (((uint64_t) i * (uint32_t) n_vec) >> LOG_N);
            pos += (((i * 2654435761u) >> 16) * step) >> 16;
            sv[g] = __ldg(row4 + min((int) pos, n_vec - 1));
        }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < PF_SAMPLE_VECS; g++) {
        float a[4];
        amplitudes(sv[g], a);
#pragma unroll
        for (int e = 0; e < EPV; e++) {
            const uint32_t k = __float_as_uint(a[e]);
            if (k <= KEY_INF) {                                // not NaN
                atomicAdd(&s_fine[k >> 20], 1u);
                atomicAdd(&s_coarse[k >> 25], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t lo_k[3], hi_k[3];
    bool degenerate;
    {
        const uint32_t c0 = s_coarse[2 * lane];
        const uint32_t c01 = c0 + s_coarse[2 * lane + 1];
        const uint32_t c_incl = warp_scan_incl(c01, lane), c_excl = c_incl - c01;
        const uint32_t m = __shfl_sync(0xffffffffu, c_incl, 31);
        degenerate = m == 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo_k[k] = 0u;
            hi_k[k] = KEY_INF;
            if (m > 0) {
                const float q = (float) ranks[k] / (float) max(nn - 1u, 1u);
                const float f_lo = (q - half_width) * (float) m, f_hi = (q + half_width) * (float) m;
                if (f_lo >= 0.0f) {
                    const BinHit b = locate_rank(s_coarse, s_fine, min((uint32_t) f_lo, m - 1u), c0, c_excl, lane);
                    lo_k[k] = (b.bin << 20) + (uint32_t) (__fdividef((float) b.r_in, (float) b.count) * 1048576.0f);
                }
                if (f_hi < (float) (m - 1u)) {
                    const BinHit b = locate_rank(s_coarse, s_fine, (uint32_t) ceilf(f_hi), c0, c_excl, lane);
                    hi_k[k] = min((b.bin << 20) + (uint32_t) (__fdividef((float) (b.r_in + 1u), (float) b.count) * 1048576.0f),
                                  KEY_INF);
                }
                hi_k[k] = max(hi_k[k], lo_k[k]);
            }
        }
        degenerate |= !(hi_k[0] < lo_k[1] && hi_k[1] < lo_k[2]);
    }
    __syncthreads();                                           // the lists may overwrite the sample histogram

    // ---- 2. one pass over the row
    uint32_t n_mine = 0;
    bool redo[3] = {degenerate, degenerate, degenerate};
    if (!degenerate) {                                         // block-uniform
        const float lo[3] = {__uint_as_float(lo_k[0]), __uint_as_float(lo_k[1]), __uint_as_float(lo_k[2])};
        const float hi[3] = {__uint_as_float(hi_k[0]), __uint_as_float(hi_k[1]), __uint_as_float(hi_k[2])};
        const uint32_t slot0 = (uint32_t) __cvta_generic_to_shared(lists + tid);
        const uint32_t slot_end = slot0 + (uint32_t) slots * PF_THREADS * 4;
        constexpr uint32_t BATCH_BYTES = PF_UNROLL * EPV * PF_THREADS * 4;
        uint32_t ge0 = 0, ge1 = 0, ge2 = 0, slot = slot0;
        float mn = __int_as_float(0x7f800000), mx = 0.0f, nan_acc = 0.0f;
        bool over = false;
        int i = tid;
        for (; i + (PF_UNROLL - 1) * PF_THREADS < n_vec; i += PF_UNROLL * PF_THREADS) {
            float4 v[PF_UNROLL];
#pragma unroll
            for (int u = 0; u < PF_UNROLL; u++) v[u] = __ldg(row4 + i + u * PF_THREADS);
            const bool room = slot + BATCH_BYTES <= slot_end;
#pragma unroll
            for (int u = 0; u < PF_UNROLL; u++) {
                float a[4];
                amplitudes(v[u], a);
#pragma unroll
                for (int e = 0; e < EPV; e++) {
                    mn = fminf(mn, a[e]);
                    mx = fmaxf(mx, a[e]);
                    nan_acc += a[e];
                    if (room) pf_step(a[e], lo, hi, ge0, ge1, ge2, slot);
                    else pf_step_checked(a[e], lo, hi, ge0, ge1, ge2, slot, slot_end, over);
                }
            }
        }
        for (; i < n_vec; i += PF_THREADS) {
            float a[4];
            amplitudes(__ldg(row4 + i), a);
#pragma unroll
            for (int e = 0; e < EPV; e++) {
                mn = fminf(mn, a[e]);
                mx = fmaxf(mx, a[e]);
                nan_acc += a[e];
                pf_step_checked(a[e], lo, hi, ge0, ge1, ge2, slot, slot_end, over);
            }
        }
        for (int j = n_vec * EPV + tid; j < n; j += PF_THREADS) {
            const float a = scalar_amplitude(j);
            mn = fminf(mn, a);
            mx = fmaxf(mx, a);
            nan_acc += a;
            pf_step_checked(a, lo, hi, ge0, ge1, ge2, slot, slot_end, over);
        }
        n_mine = (slot - slot0) / (PF_THREADS * 4);
        ge0 = __reduce_add_sync(0xffffffffu, ge0);
        ge1 = __reduce_add_sync(0xffffffffu, ge1);
        ge2 = __reduce_add_sync(0xffffffffu, ge2);
        const uint32_t mn_w = __reduce_min_sync(0xffffffffu, __float_as_uint(mn));
        const uint32_t mx_w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        const bool over_w = __any_sync(0xffffffffu, over);
        const bool nan_w = __any_sync(0xffffffffu, nan_acc != nan_acc);
        if (lane == 0) {
            atomicAdd(&misc[0], ge0);
            atomicAdd(&misc[1], ge1);
            atomicAdd(&misc[2], ge2);
            atomicMin(&misc[4], mn_w);
            atomicMax(&misc[5], mx_w);
            if (over_w) misc[3] = 1u;
            if (nan_w) misc[6] = 1u;
        }

        // ---- 3. histograms of the three brackets in one walk over the lists
        const uint32_t *mine = lists + tid;
        // one bin width for the three brackets (that of the widest), so that a key only has to
        // find its bracket's origin: key - origin_k >> shift, histogram k
        int shift;
        {
            const uint32_t widest = max(max(hi_k[0] - lo_k[0], hi_k[1] - lo_k[1]), hi_k[2] - lo_k[2]) + 1u;
            shift = (widest <= (uint32_t) PF_FINE) ? 0 : (32 - __clz(widest - 1u)) - 10;
        }
        const uint32_t d01 = lo_k[1] - lo_k[0], d12 = lo_k[2] - lo_k[1];
        for (uint32_t j = 0; j < n_mine; j++) {
            const uint32_t b = mine[j * PF_THREADS];
            const uint32_t in1 = b >= lo_k[1] ? 1u : 0u, in2 = b >= lo_k[2] ? 1u : 0u;
            const uint32_t k = in1 + in2;
            const uint32_t bin = (b - lo_k[0] - in1 * d01 - in2 * d12) >> shift;
            atomicAdd(&hist[k * PF_FINE + bin], 1u);
            atomicAdd(&coarse[k * 32 + (bin >> 5)], 1u);
        }
        __syncthreads();
        const bool bad_row = (misc[3] | misc[6]) != 0u;        // full list or NaN: redo all three
        BinHit hit[3];
        bool collect[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t below = nn - misc[k];               // elements smaller than lo_k
            const uint32_t r_rel = ranks[k] - below;           // wraps if the rank is below the bracket
            uint32_t kept;
            hit[k] = locate_rank32(coarse + k * 32, hist + k * PF_FINE, r_rel, lane, kept);
            const bool missed = r_rel >= kept;
            const bool crowded = shift != 0 && hit[k].count > 32u;
            redo[k] = bad_row || missed || crowded;
            collect[k] = !redo[k] && shift != 0;
        }
        // second walk: the keys of the three wanted bins
        // second walk: the keys of the three wanted bins.  A bracket that is not collected gets
        // an impossible bin number; (histogram, bin) is compared as one word.
        const uint32_t want0 = collect[0] ? hit[0].bin : 0xffffu;
        const uint32_t want1 = (1u << 16) | (collect[1] ? hit[1].bin : 0xffffu);
        const uint32_t want2 = (2u << 16) | (collect[2] ? hit[2].bin : 0xffffu);
        for (uint32_t j = 0; j < n_mine; j++) {
            const uint32_t b = mine[j * PF_THREADS];
            const uint32_t in1 = b >= lo_k[1] ? 1u : 0u, in2 = b >= lo_k[2] ? 1u : 0u;
            const uint32_t k = in1 + in2;
            const uint32_t tag = (k << 16) | ((b - lo_k[0] - in1 * d01 - in2 * d12) >> shift);
            if (tag == want0 || tag == want1 || tag == want2)
                small[k * 32 + atomicAdd(&misc[7 + k], 1u)] = b;
        }
        __syncthreads();
        if (warp < 3) {
            const int k = warp;
            const BinHit h = warp == 0 ? hit[0] : (warp == 1 ? hit[1] : hit[2]);
            const bool rd = warp == 0 ? redo[0] : (warp == 1 ? redo[1] : redo[2]);
            const int sh = shift;
            const uint32_t base = warp == 0 ? lo_k[0] : (warp == 1 ? lo_k[1] : lo_k[2]);
            if (!rd) {
                uint32_t v;
                if (sh == 0) {
                    v = base + h.bin;
                } else {
                    const uint32_t srt = sort32(lane < (int) h.count ? small[k * 32 + lane] : 0xffffffffu, lane);
                    v = __shfl_sync(0xffffffffu, srt, (int) h.r_in);
                }
                if (lane == 0) misc[10 + k] = v;
            }
        }
        __syncthreads();
    }

    // ---- fallbacks: radix select over the row in global memory (keys in float_to_key order)
    uint32_t found[3];
    uint32_t kmin = float_to_key(__uint_as_float(misc[4])), kmax = float_to_key(__uint_as_float(misc[5]));
#pragma unroll
    for (int k = 0; k < 3; k++) found[k] = float_to_key(__uint_as_float(misc[10 + k]));
    if (redo[0] || redo[1] || redo[2]) {                       // block-uniform
        __syncthreads();
        SelectScratch sc;
        sc.hist = lists;
        sc.misc = misc + 64;
        auto key_at = [&](int i) { return float_to_key(scalar_amplitude(i)); };
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            if (!redo[k]) continue;
            found[k] = block_radix_select<PF_THREADS>(key_at, n, ranks[k], sc);
            if (tid == 0) atomicAdd(&g_selection_fallbacks, 1ull);
        }
        if (degenerate || misc[6] != 0u) {                     // min / max in key order (NaN sorts last)
            uint32_t lo_key = 0xffffffffu, hi_key = 0u;
            for (int i = tid; i < n; i += PF_THREADS) {
                const uint32_t key = key_at(i);
                lo_key = min(lo_key, key);
                hi_key = max(hi_key, key);
            }
            lo_key = __reduce_min_sync(0xffffffffu, lo_key);
            hi_key = __reduce_max_sync(0xffffffffu, hi_key);
            __syncthreads();
            if (tid == 0) { misc[64] = 0xffffffffu; misc[65] = 0u; }
            __syncthreads();
            if (lane == 0) { atomicMin(&misc[64], lo_key); atomicMax(&misc[65], hi_key); }
            __syncthreads();
            kmin = misc[64];
            kmax = misc[65];
        }
    }
    if (tid == 0) {
        const int64_t r = blockIdx.x;
        dest[0 * dest_stride + r] = key_to_float(kmin);
        dest[1 * dest_stride + r] = key_to_float(kmax);
        dest[2 * dest_stride + r] = key_to_float(found[0]);    // 25 %
        dest[3 * dest_stride + r] = key_to_float(found[2]);    // 75 %
        dest[4 * dest_stride + r] = key_to_float(found[1]);    // 50 %
    }
}

int baselines_limit_ok(int64_t rows) { return rows > 0x7fffffff ? 1 : 0; }

size_t select_smem_bytes(int64_t n, bool in_smem)
{
    return (size_t) (SELECT_HIST_WORDS + 64 + (in_smem ? ((n + 3) & ~(int64_t) 3) : 0)) * 4;
}

}  // namespace

extern "C" int ksp_selection_fallback_count(void *stream, unsigned long long *count, int reset)
{
    if (!count) return KSP_EINVAL;
    KSP_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    KSP_CUDA(cudaMemcpyFromSymbol(count, g_selection_fallbacks, sizeof(*count)));
    if (reset) {
        const unsigned long long zero = 0;
        KSP_CUDA(cudaMemcpyToSymbol(g_selection_fallbacks, &zero, sizeof(zero)));
    }
    return 0;
}

extern "C" int ksp_madnz_t(void *stream, const float *dev_t, float *noise, int64_t channels,
                           int64_t baselines, int64_t stride)
{
    if (channels < 0 || baselines < 0 || stride < channels) return KSP_EINVAL;
    if (baselines == 0) return 0;
    if (!dev_t || !noise) return KSP_EINVAL;
    if (channels > (int64_t) 1 << 30 || baselines > 0x7fffffff) return KSP_ETOOLARGE;
    madnz_stream_kernel<<<(unsigned) baselines, MS_THREADS, 0, (cudaStream_t) stream>>>(
        dev_t, noise, (int) channels, stride);
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
    if (channels <= 65535) {                                 // 16-bit counters suffice
        const size_t smem = (size_t) CM3_ROWS * 32 * sizeof(uint32_t);
        KSP_CUDA(cudaFuncSetAttribute(madnz_cm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) smem));
        madnz_cm3_kernel<<<(unsigned) ksp_divup(baselines, 32), 1024, smem, (cudaStream_t) stream>>>(
            dev, noise, (int) channels, (int) baselines, stride);
        KSP_CHECK_LAUNCH();
        return 0;
    }
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
    if (baselines_limit_ok(rows) != 0) return KSP_ETOOLARGE;
    // fast path: rows that can be read with aligned 16-byte loads and are long enough to sample
    {
        const size_t esz = is_amplitude ? 4 : 8;
        const int epv = is_amplitude ? 4 : 2;
        const bool aligned = ((uintptr_t) src % 16) == 0 && ((size_t) src_stride * esz) % 16 == 0 &&
                             ((size_t) first_col * esz) % 16 == 0;
        static const bool fast_allowed = [] {
            const char *e = getenv("KSP_P5_FAST");
            return !(e && atoi(e) == 0);
        }();
        if (fast_allowed && aligned && n_cols >= 8192) {
            const double m = 256.0 * PF_SAMPLE_VECS * epv;          // samples
            const double h = 4.5 * sqrt(0.25 / m);                  // bracket half-width (rank fraction)
            const double frac = 6.0 * h + 3.0 / 1024.0;             // kept: three brackets + bin interpolation slack
            const double per_t = (double) ksp_divup(n_cols, PF_THREADS);
            int64_t fslots = (int64_t) (per_t * frac + 5.0 * sqrt(per_t * frac * (1.0 - frac)) + 4.0);
            if (fslots * PF_THREADS < SELECT_HIST_WORDS) fslots = SELECT_HIST_WORDS / PF_THREADS;
            const size_t fsmem = ((size_t) fslots * PF_THREADS + 3 * PF_FINE + 3 * 32 + 3 * 32 + 128) * sizeof(uint32_t);
            if (fsmem <= 110 * 1024) {                              // 2-3 blocks per SM
                const int mode = is_amplitude ? 0 : (abs_mode == KSP_ABS_NUMPY ? 1 : 2);
#define KSP_PF_CASE(M)                                                                         \
                if (mode == M) {                                                               \
                    KSP_CUDA(cudaFuncSetAttribute(percentile5_fast_kernel<M>,                  \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fsmem)); \
                    percentile5_fast_kernel<M><<<(unsigned) rows, PF_THREADS, fsmem, s>>>(     \
                        src, dest, src_stride, dest_stride, first_col, (int) n_cols, (int) fslots, (float) h); \
                }
                KSP_PF_CASE(0)
                KSP_PF_CASE(1)
                KSP_PF_CASE(2)
#undef KSP_PF_CASE
                KSP_CHECK_LAUNCH();
                return 0;
            }
        }
    }
    // list slots per thread: the three brackets keep ~36 % of a thread's n / 256 keys
    const int64_t per_thread = ksp_divup(n_cols, P5_THREADS);
    int64_t slots = (per_thread * 36 + 99) / 100 + 5 * (int64_t) ceil(sqrt(0.23 * (double) per_thread)) + 4;
    if (slots < P5_MIN_SLOTS) slots = P5_MIN_SLOTS;
    if (slots * P5_THREADS < SELECT_HIST_WORDS) slots = SELECT_HIST_WORDS / P5_THREADS;
    const size_t smem = ((size_t) slots * P5_THREADS + MS_BINS + 320) * sizeof(uint32_t);
    if (smem <= 200 * 1024) {
        const int mode = is_amplitude ? 0 : (abs_mode == KSP_ABS_NUMPY ? 1 : 2);
#define KSP_P5_CASE(M)                                                                         \
        if (mode == M) {                                                                       \
            KSP_CUDA(cudaFuncSetAttribute(percentile5_stream_kernel<M>,                        \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
            percentile5_stream_kernel<M><<<(unsigned) rows, P5_THREADS, smem, s>>>(            \
                src, dest, src_stride, dest_stride, first_col, (int) n_cols, (int) slots);     \
        }
        KSP_P5_CASE(0)
        KSP_P5_CASE(1)
        KSP_P5_CASE(2)
#undef KSP_P5_CASE
        KSP_CHECK_LAUNCH();
        return 0;
    }
    // very long rows: block-wide radix select re-reading global memory
    const size_t smem_old = select_smem_bytes(n_cols, false);
    KSP_CUDA(cudaFuncSetAttribute(percentile5_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_old));
    percentile5_kernel<false><<<(unsigned) rows, SEL_THREADS, smem_old, s>>>(
        src, dest, src_stride, dest_stride, first_col, (int) n_cols, is_amplitude, abs_mode);
    KSP_CHECK_LAUNCH();
    return 0;
}
