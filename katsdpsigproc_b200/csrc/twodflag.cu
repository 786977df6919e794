// 2-D SumThreshold flagger (time x frequency per baseline).
//
// Replaces reference rfi/twodflag.py:67-890 (numba, CPU only): _average_freq, _time_median,
// _get_background2d (masked box-Gaussian filters, per-chunk MAD rejection, NaN interpolation),
// _sum_threshold along both axes, _combine_flags, _unaverage_freq, _get_flags_impl.
//
// The results of that code depend on the ORDER of its float64 running sums (box filters,
// cumulative sums of the SumThreshold) and on which intermediate is float32, so this file keeps
// every such recurrence serial along its axis, in the reference's order and precision, and takes
// its parallelism from everything the reference loops over independently: baselines (one
// 128-thread block per baseline, 8 blocks per SM, persistent over the batch), and inside a
// baseline the columns of a time-axis recurrence, the rows (x frequency chunks) of a
// frequency-axis one.  What changed against a line-by-line restatement is WHERE the recurrences
// keep their state:
//
//   * box filters (twodflag.py:255-309): the reference runs four in-place passes over a padded
//     copy of the line.  Pass p at position i reads only what pass p - 1 left at i .. i + 2 r, so
//     the four passes are chained here as four stages of ONE sweep over the line, each stage
//     with its float64 running sum in a register and its last 2 r inputs in a shared-memory
//     ring; the line is read once and written once (frequency axis: through a transposing tile,
//     128-byte rows of 32 lines at a time; time axis: directly, neighbouring threads hold
//     neighbouring columns).  Same additions, same order, same float32 roundings between passes.
//   * SumThreshold (twodflag.py:493-560): one sweep per window size, the last w + 1 cumulative
//     sums in a shared-memory ring, flags as bit masks; "every hit flags the w samples of its
//     window" becomes "sample k is flagged if any of the last w hits fired", a shift register.
//   * medians (thresholds per row and chunk, MAD per chunk): exact radix select by ONE warp per
//     set, the warps of a block working on different sets.
//
// oracle/twodflag_numpy.py states the reference's arithmetic in numpy and is pinned bit for bit
// against the reference; the tests compare this file's flags with both.
//
// Three launches per batch of baselines:
//   twod_average_kernel   (time, freq, baseline) input -> baseline-major averaged magnitudes + flags
//   twod_baseline_kernel  everything of rfi/twodflag.py:768-881 for one baseline per block
//   twod_output_kernel    baseline-major flags -> (time, freq, baseline), OR isnan(input)
#include "common.cuh"
#include "select.cuh"
#include <math.h>
#include <string.h>

namespace {

using namespace ksp;

constexpr int TD_THREADS = 128;
constexpr int TD_WARPS = TD_THREADS / 32;
constexpr int TD_BLOCKS_PER_SM = 8;
constexpr int TD_SMEM_WORDS = 5120;          // 20 KB of rings / tiles per block
constexpr int TD_PASSES = 4;                 // box filters per Gaussian (reference default, twodflag.py:313)
constexpr double TD_MAD_NORMAL = 1.4826;     // rfi/__init__.py:31
constexpr unsigned FULL = 0xffffffffu;

struct TdArgs {
    ksp_twodflag_params p;
    const void *data;            // (time, freq, baseline)
    const uint8_t *in_flags;     // same shape, non-zero = flagged
    uint8_t *out_flags;          // same shape
    int64_t bl0, nb;             // batch of baselines
    int a_freq;                  // averaged channels
    int max_rt, max_rf;          // largest box radii
    int max_wt, max_wf;          // largest windows
    int max_cl;                  // longest frequency chunk
    int big_radius;              // a box radius too large for the shared-memory rings: in-place fallback
    char *scratch;
    size_t per_bl;               // scratch bytes per baseline
};

// Cycles block 0 spent in each phase of the launches since the last reset (ksp_twodflag_phases):
// 0 spectrum median, 1 spectrum Gaussians, 2 spectrum MAD, 3 spectrum interpolation,
// 4 spectrum thresholds, 5 spectrum SumThreshold, 6 2-D Gaussians, 7 2-D MAD, 8 2-D interpolation,
// 9 SumThreshold in time, 10 thresholds per row and chunk, 11 SumThreshold in frequency,
// 12 combination, 13 un-averaging and fill rules, 15 elementwise steps in between; inside the
// Gaussians: 16 time-axis box passes, 17 frequency-axis box passes, 18 masking and normalisation.
__device__ unsigned long long td_phase[KSP_TWOD_PHASES];
// ---- per-baseline scratch layout (A = n_time * a_freq elements)
struct TdBuffers {
    float *data, *bg, *weight, *vals;
    float *pad;                  // in-place work areas of the large-radius fallback (two of pad_n floats)
    int64_t pad_n;
    float *spec_data, *spec_bg, *spec_weight;
    uint8_t *flags, *work, *tfl, *ffl, *spec_flags, *spec_work, *spec_out, *comb;
    uint32_t *posw, *negw;       // SumThreshold bit masks, [word][line]
    uint8_t *outb;               // (n_time, n_freq) flags of this baseline at the original resolution
    uint8_t *row_full, *col_full; // n_time, n_freq
    float *thr;                  // thresholds per (row, chunk)
};

__host__ __device__ inline size_t td_align(size_t x) { return (x + 255) / 256 * 256; }

// words of one SumThreshold bit mask: the larger of the frequency-axis lines (every (row, chunk)
// has its slice of chunk + 2 (largest window - 1) samples) and the time-axis ones (one per column)
__host__ __device__ inline size_t td_mask_words(const TdArgs &a)
{
    const size_t T = (size_t) a.p.n_time, F = (size_t) a.a_freq;
    const size_t wf = ((size_t) a.max_cl + 2 * (size_t) a.max_wf + 31) / 32 * T * (size_t) a.p.n_chunks;
    const size_t wt = (T + 31) / 32 * F;
    return wf > wt ? wf : wt;
}

__host__ __device__ inline size_t td_layout(const TdArgs &a, char *base, TdBuffers *b)
{
    const size_t T = (size_t) a.p.n_time, F = (size_t) a.a_freq, A = T * F;
    size_t pad_n = 0;
    if (a.big_radius) {
        const size_t pad_t = (T + (size_t) a.max_rt * TD_PASSES) * F, pad_f = T * (F + (size_t) a.max_rf * TD_PASSES);
        pad_n = pad_t > pad_f ? pad_t : pad_f;
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += td_align(bytes); return base ? base + o : (char *) nullptr; };
    float *f_data = (float *) take(A * 4), *f_bg = (float *) take(A * 4), *f_w = (float *) take(A * 4);
    float *f_vals = (float *) take(A * 4), *f_pad = (float *) take(2 * pad_n * 4);
    float *s_data = (float *) take(F * 4), *s_bg = (float *) take(F * 4), *s_w = (float *) take(F * 4);
    uint8_t *u[8];
    for (int i = 0; i < 4; i++) u[i] = (uint8_t *) take(A);
    for (int i = 4; i < 7; i++) u[i] = (uint8_t *) take(F);
    u[7] = (uint8_t *) take(A);
    const size_t mw = td_mask_words(a);
    uint32_t *posw = (uint32_t *) take(mw * 4), *negw = (uint32_t *) take(mw * 4);
    uint8_t *outb = (uint8_t *) take(T * (size_t) a.p.n_freq);
    uint8_t *row_full = (uint8_t *) take(T), *col_full = (uint8_t *) take((size_t) a.p.n_freq);
    float *thr = (float *) take((T + 1) * (size_t) (a.p.n_chunks > 0 ? a.p.n_chunks : 1) * 4);
    if (b) {
        b->data = f_data; b->bg = f_bg; b->weight = f_w; b->vals = f_vals; b->pad = f_pad;
        b->pad_n = (int64_t) pad_n;
        b->spec_data = s_data; b->spec_bg = s_bg; b->spec_weight = s_w;
        b->flags = u[0]; b->work = u[1]; b->tfl = u[2]; b->ffl = u[3];
        b->spec_flags = u[4]; b->spec_work = u[5]; b->spec_out = u[6]; b->comb = u[7];
        b->posw = posw; b->negw = negw;
        b->outb = outb; b->thr = thr; b->row_full = row_full; b->col_full = col_full;
    }
    return off;
}

__device__ __forceinline__ void td_buffers(const TdArgs &a, int64_t blr, TdBuffers *b)
{
    td_layout(a, a.scratch + (size_t) blr * a.per_bl, b);
}

// ------------------------------------------------------------------ averaging (twodflag.py:68-116)
// Block = 32 baselines x 32 averaged channels of one dump: the input is read with the baseline
// fastest (its layout), the per-baseline arrays are written with the channel fastest (theirs).
template <bool COMPLEX>
__global__ void __launch_bounds__(256)
twod_average_kernel(const TdArgs a)
{
    __shared__ float s_val[32][33];
    __shared__ uint8_t s_flag[32][33];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t bl_tile = (int64_t) blockIdx.x * 32;
    const int f_tile = (int) blockIdx.y * 32;
    const int t = (int) blockIdx.z;
    const int factor = a.p.average_freq;
    const int64_t blr = bl_tile + lane;
    for (int fo = wrp; fo < 32; fo += 8) {
        const int jo = f_tile + fo;
        float sum = 0.0f;
        int count = 0;
        if (blr < a.nb && jo < a.a_freq) {
            for (int j = jo * factor; j < (jo + 1) * factor && j < a.p.n_freq; j++) {
                const int64_t idx = ((int64_t) t * a.p.n_freq + j) * a.p.n_bl + a.bl0 + blr;
                float mag;
                if (COMPLEX) {
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(a.data) + idx);
                    // numba's abs(complex64): the correctly rounded hypot
                    mag = (isnan(v.x) || isnan(v.y)) ? __int_as_float(0x7fc00000)
                                                     : abs_slow(fabsf(v.x), fabsf(v.y), KSP_ABS_HYPOT);
                } else {
                    mag = fabsf(__ldg(reinterpret_cast<const float *>(a.data) + idx));
                }
                if (!__ldg(a.in_flags + idx) && !isnan(mag)) {
                    sum = __fadd_rn(sum, mag);
                    count++;
                }
            }
        }
        s_val[fo][lane] = count ? __fdiv_rn(sum, (float) count) : 0.0f;
        s_flag[fo][lane] = count == 0;
    }
    __syncthreads();
    const int jo = f_tile + lane;
    for (int bo = wrp; bo < 32; bo += 8) {
        const int64_t bl = bl_tile + bo;
        if (bl < a.nb && jo < a.a_freq) {
            TdBuffers b;
            td_buffers(a, bl, &b);
            const size_t o = (size_t) t * a.a_freq + jo;
            b.data[o] = s_val[lane][bo];
            b.flags[o] = s_flag[lane][bo];
        }
    }
}

// ------------------------------------------------------------------ shared state of twod_baseline_kernel
// The phases below are separate (non-inlined) functions so that each gets its own register
// allocation; what they share lives here instead of in arguments.
__shared__ TdArgs s_a;                                   // the launch arguments
__shared__ TdBuffers s_b;                                // scratch arrays of the block's current baseline
__shared__ __align__(16) float s_sm[TD_SMEM_WORDS];      // rings and tiles
__shared__ uint32_t s_whist[TD_WARPS * 256];             // one 256-bin histogram per warp
__shared__ long long s_last;                             // td_mark's clock

__device__ __forceinline__ void td_mark(int k)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long now = clock64();
        td_phase[k] += (unsigned long long) (now - s_last);
        s_last = now;
    }
}

// Thread number with the warps rotated by a per-block amount.  The phases that cannot use every
// warp (a box filter along frequency has 2 T lines) hand their work to the first warps IN THIS
// NUMBERING: warp w of a block always runs on scheduler w mod 4 of its SM, and with the plain
// numbering the working warps of all the blocks of an SM would queue for the conversion pipe of
// one scheduler while the other three sat idle.
__device__ __forceinline__ int td_vtid()
{
    const unsigned rot = (blockIdx.x * 0x9E3779B1u) >> 30;
    return (int) (((threadIdx.x >> 5) + rot) & (TD_WARPS - 1)) * 32 + (int) (threadIdx.x & 31);
}
static_assert(TD_WARPS == 4, "td_vtid rotates four warps");

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ float td_nan() { return __int_as_float(0x7fc00000); }

__device__ __forceinline__ float median_sorted_f32(const float *s, int64_t stride, int n)
{
    if (n & 1) return s[(int64_t) (n / 2) * stride];
    return __fmul_rn(__fadd_rn(s[(int64_t) (n / 2 - 1) * stride], s[(int64_t) (n / 2) * stride]), 0.5f);
}

// insertion sort of a thread-private strided segment (the time axis of one column; neighbouring
// threads hold neighbouring columns, so the accesses of a warp coalesce)
__device__ void sort_small(float *v, int64_t stride, int n)
{
    for (int i = 1; i < n; i++) {
        const float x = v[(int64_t) i * stride];
        int j = i - 1;
        while (j >= 0 && v[(int64_t) j * stride] > x) {
            v[(int64_t) (j + 1) * stride] = v[(int64_t) j * stride];
            j--;
        }
        v[(int64_t) (j + 1) * stride] = x;
    }
}

// Median of the unflagged samples of one column (values at x[t * stride], flags at fl[t * stride],
// t < T <= N) in registers: flagged samples become +inf, Batcher's odd-even merge sort of N values
// with every index known at compile time, then the middle of the n unflagged ones.  ABS: of |x|.
// n == 0: returns NaN and *count = 0.
template <int N, bool ABS>
__device__ __forceinline__ float column_median(const float *x, const uint8_t *fl, int64_t stride, int T, int *count)
{
    float v[N];
    int n = 0;
#pragma unroll
    for (int t = 0; t < N; t++) {
        float val = __int_as_float(0x7f800000);
        if (t < T && !fl[(int64_t) t * stride]) {
            val = x[(int64_t) t * stride];
            if (ABS) val = fabsf(val);
            n++;
        }
        v[t] = val;
    }
    *count = n;
    if (n == 0) return td_nan();
#pragma unroll
    for (int p = 1; p < N; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k)
#pragma unroll
                for (int i = 0; i <= (k - 1 < N - j - k - 1 ? k - 1 : N - j - k - 1); i++)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const float lo = fminf(v[i + j], v[i + j + k]), hi = fmaxf(v[i + j], v[i + j + k]);
                        v[i + j] = lo;
                        v[i + j + k] = hi;
                    }
    const int hi_i = n >> 1, lo_i = (n - 1) >> 1;
    float a = 0.0f, b = 0.0f;
#pragma unroll
    for (int t = 0; t < N; t++) {
        if (t == lo_i) a = v[t];
        if (t == hi_i) b = v[t];
    }
    return (n & 1) ? b : __fmul_rn(__fadd_rn(a, b), 0.5f);
}

// The same for any T: the samples are gathered into `work` (stride as the inputs) and sorted there.
template <bool ABS>
__device__ float column_median_any(const float *x, const uint8_t *fl, float *work, int64_t stride, int T, int *count)
{
    if (T <= 16) return column_median<16, ABS>(x, fl, stride, T, count);
    if (T <= 32) return column_median<32, ABS>(x, fl, stride, T, count);
    int n = 0;
    for (int t = 0; t < T; t++)
        if (!fl[(int64_t) t * stride]) {
            const float val = x[(int64_t) t * stride];
            work[(int64_t) (n++) * stride] = ABS ? fabsf(val) : val;
        }
    *count = n;
    if (n == 0) return td_nan();
    sort_small(work, stride, n);
    return median_sorted_f32(work, stride, n);
}

__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Median of the non-skipped keys of a set, by ONE warp: exact radix select on the bit patterns of
// non-negative floats, 4 passes of 8 bits with a 256-bin histogram of the warp's own; float32 mean
// of the two middle values as numba's np.median of float32.  NaN if the set is empty.  All lanes
// return the same value.
//
// The end of a pass: which bin holds `rank`, and the rank inside it.  c: the lane's 8 bins.
struct BinChoice {
    uint32_t bin, r_in, count, total;
};
__device__ __forceinline__ BinChoice warp_choose_bin(const uint32_t *hist, uint32_t rank, bool first, int lane,
                                                     uint32_t (&c)[8])
{
    uint32_t tot = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        c[j] = hist[8 * lane + j];
        tot += c[j];
    }
    const uint32_t incl = warp_scan_incl(tot, lane), excl = incl - tot;
    BinChoice ch;
    ch.total = __shfl_sync(FULL, incl, 31);
    if (first) rank = ch.total ? (ch.total - 1) >> 1 : 0;          // the lower median
    const bool mine = rank >= excl && rank < incl;
    const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
    uint32_t bin = 0, r_in = 0, cnt = 0;
    if (mine) {
        uint32_t e = excl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (rank >= e && rank < e + c[j]) {
                bin = 8u * (uint32_t) lane + j;
                r_in = rank - e;
                cnt = c[j];
            }
            e += c[j];
        }
    }
    ch.bin = __shfl_sync(FULL, bin, max(src, 0));
    ch.r_in = __shfl_sync(FULL, r_in, max(src, 0));
    ch.count = __shfl_sync(FULL, cnt, max(src, 0));
    return ch;
}

// Upper median once the lower one (lo: the key `prefix24 | bin` with `count` copies, of which
// number r_in is the lower median) is known: lo itself if another copy follows, else the next key
// - the next occupied bin of the last pass's histogram (same upper 24 bits), else `beyond`, the
// smallest key above all keys that share lo's upper 24 bits (collected during the last pass).
__device__ __forceinline__ uint32_t warp_upper_median(uint32_t lo, const BinChoice &ch, const uint32_t (&c)[8],
                                                      uint32_t beyond, int lane)
{
    if (ch.r_in + 1 < ch.count) return lo;
    uint32_t next_bin = 0xffffffffu;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        const uint32_t b = 8u * (uint32_t) lane + j;
        if (c[j] != 0u && b > ch.bin) next_bin = b;
    }
    next_bin = __reduce_min_sync(FULL, next_bin);
    if (next_bin != 0xffffffffu) return (lo & 0xffffff00u) | next_bin;
    return __reduce_min_sync(FULL, beyond);
}

__device__ __forceinline__ float warp_median_finish(uint32_t lo, uint32_t hi)
{
    const float a = __uint_as_float(lo), b = __uint_as_float(hi);
    return lo == hi ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
}

// The set is rows x cols in memory: key_first(r, c) is called for every element in the first
// pass (it may compute and store what key_at(r, c) then re-reads in the other three); both return
// KEY_SKIP for elements that take no part.  Eight loads in flight per lane.
template <typename KeyFirst, typename KeyAt>
__device__ float warp_median(const KeyFirst &key_first, const KeyAt &key_at, int rows, int cols, uint32_t *hist,
                             int lane)
{
    uint32_t prefix = 0, prefix_mask = 0, rank = 0, n_valid = 0, beyond = KEY_SKIP;
    uint32_t c[8];
    BinChoice ch;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int bin = lane; bin < 256; bin += 32) hist[bin] = 0u;
        __syncwarp();
        for (int r = 0; r < rows; r += 2)
            for (int col = lane; col < cols; col += 128) {
                uint32_t k[8];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool in0 = col + 32 * u < cols, in1 = in0 && r + 1 < rows;
                    if (shift == 24) {
                        k[u] = in0 ? key_first(r, col + 32 * u) : KEY_SKIP;
                        k[4 + u] = in1 ? key_first(r + 1, col + 32 * u) : KEY_SKIP;
                    } else {
                        k[u] = in0 ? key_at(r, col + 32 * u) : KEY_SKIP;
                        k[4 + u] = in1 ? key_at(r + 1, col + 32 * u) : KEY_SKIP;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (k[u] == KEY_SKIP) continue;
                    if ((k[u] & prefix_mask) == prefix) atomicAdd(&hist[(k[u] >> shift) & 0xffu], 1u);
                    else if (shift == 0 && k[u] > prefix) beyond = min(beyond, k[u]);
                }
            }
        __syncwarp();
        ch = warp_choose_bin(hist, rank, shift == 24, lane, c);
        if (shift == 24) {
            n_valid = ch.total;
            if (n_valid == 0) return td_nan();                      // warp-uniform
        }
        rank = ch.r_in;
        prefix |= ch.bin << shift;
        prefix_mask |= 0xffu << shift;
        __syncwarp();
    }
    const uint32_t lo = prefix;
    const uint32_t hi = (n_valid & 1u) ? lo : warp_upper_median(lo, ch, c, beyond, lane);
    return warp_median_finish(lo, hi);
}

// The same for a set small enough to sit in registers (n <= 32 * PER_LANE): read once.
template <int PER_LANE, typename KeyAt>
__device__ float warp_median_small(const KeyAt &key_at, int n, uint32_t *hist, int lane)
{
    uint32_t k[PER_LANE];
#pragma unroll
    for (int u = 0; u < PER_LANE; u++) k[u] = (lane + 32 * u < n) ? key_at(lane + 32 * u) : KEY_SKIP;
    uint32_t prefix = 0, prefix_mask = 0, rank = 0, n_valid = 0, beyond = KEY_SKIP;
    uint32_t c[8];
    BinChoice ch;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int bin = lane; bin < 256; bin += 32) hist[bin] = 0u;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < PER_LANE; u++) {
            if (k[u] == KEY_SKIP) continue;
            if ((k[u] & prefix_mask) == prefix) atomicAdd(&hist[(k[u] >> shift) & 0xffu], 1u);
            else if (shift == 0 && k[u] > prefix) beyond = min(beyond, k[u]);
        }
        __syncwarp();
        ch = warp_choose_bin(hist, rank, shift == 24, lane, c);
        if (shift == 24) {
            n_valid = ch.total;
            if (n_valid == 0) return td_nan();                      // warp-uniform
        }
        rank = ch.r_in;
        prefix |= ch.bin << shift;
        prefix_mask |= 0xffu << shift;
        __syncwarp();
    }
    const uint32_t lo = prefix;
    const uint32_t hi = (n_valid & 1u) ? lo : warp_upper_median(lo, ch, c, beyond, lane);
    return warp_median_finish(lo, hi);
}

// ------------------------------------------------------------------ box filter (twodflag.py:255-309)
// The reference pads the line with 4 r zeros in front (padded = [0] * 4 r + line, len = n + 4 r;
// beyond the end counts as zero too) and runs pass p = 1 .. 4 in place:
//     s = sum(padded[0 : 2 r]);  for i < len:  s += padded[i + 2 r];  out = float32(s);
//                                              s -= padded[i];  padded[i] = out
// (its start / stop bookkeeping only skips positions whose inputs are known zeros), then returns
// padded[0 : n] / float32((2 r + 1)^4).  Pass p at position i needs pass p - 1 at i .. i + 2 r
// only, so the passes run here as four stages of one sweep: at step k (k = 0 .. len - 1) stage 1
// takes line[k] (its index 4 r + k), stage 2 the value stage 1 has just produced (index 2 r + k),
// stage 3 what stage 2 produced (index k), stage 4 what stage 3 produced (index k - 2 r), and
// the result for line position k - 4 r leaves stage 4.  Each stage keeps its float64 sum and a
// ring of its last 2 r inputs; all four rings turn together, so one slot number serves them.
// The rings of stages 1 and 2 start as zeros (the padding), those of 3 and 4 are filled while
// these stages sum their first 2 r inputs.
//
// The stages are SKEWED by one iteration each - iteration `it` runs stage 1 at step it, stage 2
// at step it - 1 (on what stage 1 produced in the previous iteration), stage 3 at step it - 2,
// stage 4 at step it - 3 - so that nothing in an iteration waits for anything else in it.  Two
// layouts: along frequency (2 T long lines) the four stages of a line are four neighbouring LANES
// (lane = 4 * line + stage; a shuffle hands a stage's output to the next lane at the start of the
// next iteration), so that 32 lines keep all four warps of a block - all four schedulers of the
// SM, each with its own float64 and conversion pipes - busy; along time (2 F short lines) one
// thread runs all four stages of a line as four independent chains in one branch-free block.
//
// Two arithmetics.  BoxF64: the reference's - float64 sums, float32 between the stages.
// BoxInt: the same sums as 32-bit integers, for lines whose samples are 0 or 1 (the weights of a
// masked filter): every intermediate is an integer below (2 r + 1)^4, so for r <= 31 it is below
// 2^24 and both the float64 sums and their float32 roundings are exact - integer adds give the
// same bits without touching the float64 and conversion pipes.
struct BoxF64 {
    typedef double Sum;
    typedef float Val;
    static __device__ __forceinline__ double add(double s, float v) { return __dadd_rn(s, (double) v); }
    static __device__ __forceinline__ double sub(double s, float v) { return __dsub_rn(s, (double) v); }
    static __device__ __forceinline__ float round(double s) { return (float) s; }
    static __device__ __forceinline__ float load(const float *p) { return *p; }
    static __device__ __forceinline__ void store(float *p, float v) { *p = v; }
    static __device__ __forceinline__ float from_input(float v) { return v; }
    static __device__ __forceinline__ float to_float(float v) { return v; }
};
struct BoxInt {
    typedef int Sum;
    typedef int Val;
    static __device__ __forceinline__ int add(int s, int v) { return s + v; }
    static __device__ __forceinline__ int sub(int s, int v) { return s - v; }
    static __device__ __forceinline__ int round(int s) { return s; }
    static __device__ __forceinline__ int load(const float *p) { return __float_as_int(*p); }
    static __device__ __forceinline__ void store(float *p, int v) { *p = __int_as_float(v); }
    static __device__ __forceinline__ int from_input(float v) { return v != 0.0f; }
    static __device__ __forceinline__ float to_float(int v) { return (float) v; }
};
constexpr int TD_INT_RADIUS = 31;                 // (2 * 31 + 1)^4 < 2^24

// One stage of one line.  Stage p (0-based) runs its step it - p in iteration `it`, so with
// `it` as the common clock: it adds its input while it < add_limit (stage 1: the n samples; stage
// 2: n + 2 r values of stage 1; stages 3, 4: always), and it produces / pops from iteration
// pop_start on (stages 1, 2: at once, their rings start as the zero padding; stage 3 after its
// first 2 r inputs; stage 4 likewise) - before that `e` stays 0, which is a no-op downstream.
template <typename A>
struct BoxStage {
    typename A::Sum s;
    typename A::Val e;            // what this stage produced in the previous iteration
    int add_limit, pop_start;
    __device__ __forceinline__ void init(int p, int n, int r2)
    {
        s = 0;
        e = 0;
        add_limit = p == 0 ? n : p == 1 ? n + r2 + 1 : 0x7fffffff;
        pop_start = p == 2 ? r2 + 2 : p == 3 ? 2 * r2 + 3 : 0;
    }
    // `in`: the sample (stage 1) or the previous lane's e; rp: this thread's ring slot it mod 2 r.
    // Returns what the stage produced; it counts (and e takes it) from pop_start on.  Branch-free.
    __device__ __forceinline__ typename A::Val step(int it, typename A::Val in, float *rp)
    {
        const bool add_ok = it < add_limit, pop_ok = it >= pop_start;
        const typename A::Sum sa0 = A::add(s, in);
        const typename A::Sum sa = add_ok ? sa0 : s;
        const typename A::Val en = A::round(sa);
        const typename A::Sum sb = A::sub(sa, A::load(rp));
        s = pop_ok ? sb : sa;
        e = pop_ok ? en : e;
        A::store(rp, in);
        return en;
    }
    // The same when the stage both adds and pops (add_limit > it >= pop_start): no conditions.
    __device__ __forceinline__ typename A::Val step_steady(typename A::Val in, float *rp)
    {
        const typename A::Sum sa = A::add(s, in);
        e = A::round(sa);
        s = A::sub(sa, A::load(rp));
        A::store(rp, in);
        return e;
    }
};

// float32(2 r + 1) ** K as numba evaluates it: binary exponentiation in float32 (K = 4: the
// square of the square)
__device__ __forceinline__ float box_divisor(int r)
{
    static_assert(sizeof(TdArgs) % 4 == 0, "copied by words");
    static_assert(TD_PASSES == 4, "square of the square");
    const float d = (float) (2 * r + 1);
    const float d2 = __fmul_rn(d, d);
    return __fmul_rn(d2, d2);
}

// Lines of a ring-based pass that fit the shared memory: a multiple of 32 when at least 32 fit.
__device__ __forceinline__ int td_lines_that_fit(int words_available, int words_per_line, int most)
{
    int L = words_available / words_per_line;
    if (L > most) L = most;
    if (L >= 32) L &= ~31;
    return L;
}

__device__ __forceinline__ float td_shfl_up1(float v) { return __shfl_up_sync(FULL, v, 1); }

// Time axis of a masked filter.  Lines 0 .. F - 1 are the weights of the columns (1 where
// unflagged), lines F .. 2 F - 1 the data with flagged samples zeroed, both made on the fly from
// `data` and `flags`; results into weight / out.  The samples of the next four iterations (and of
// the thread's next line) are requested before the current four are worked on: the arrays live
// in DRAM, and nothing else hides that latency.
struct TimeLine {
    const float *data;
    const uint8_t *flags;
    int T, F;
    // Four raw samples of a line, t0 .. t0 + 3: the loads are only ISSUED here - masking them now
    // would wait for them.
    struct Raw {
        float d[4];
        uint8_t f[4];
    };
    __device__ __forceinline__ void load4(int line, int t0, Raw &r) const
    {
        const bool is_data = line >= F;
        const int f = is_data ? line - F : line;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            r.d[k] = 1.0f;
            r.f[k] = 1;
            if (t0 + k < T) {
                const int64_t o = (int64_t) (t0 + k) * F + f;
                if (is_data) r.d[k] = data[o];
                r.f[k] = flags[o];
            }
        }
    }
    static __device__ __forceinline__ float masked(const Raw &r, int k) { return r.f[k] ? 0.0f : r.d[k]; }
};

// The time-axis passes have thousands of short lines and use the same four stages inside ONE
// thread per line (less work per line and iteration than four lanes and a shuffle; with that many
// lines every warp is busy anyway).  Iteration `it` again runs stage 1 at step it, stage 2 at step
// it - 1, ... so that nothing in an iteration waits for anything else in it.
template <typename A>
struct BoxState {
    typename A::Sum s1, s2, s3, s4;
    typename A::Val e1, e2, e3;   // what stages 1 - 3 produced in the previous iteration
};

// rp: this line's slot (it mod 2 r) of the stage-1 ring; the other stages' rings follow at
// multiples of `pitch` (all four must be readable: zero the lot before the first iteration).
// Returns true when `out` holds the unnormalised result for line position it - 3 - 2 * r2.  Call
// for it = 0 .. n + 2 * r2 + 2.
// The body is branch-free (selects and one predicated result): the scheduler issues in order, and
// only inside one basic block can the four chains overlap.  Stages that have not started yet see
// zeros (e1 .. e3 start as 0, the rings of stages 1 and 2 too), for which every operation below
// is a no-op, so only four conditions remain.
template <typename A>
__device__ __forceinline__ bool box_iter(BoxState<A> &st, float *rp, int pitch, int it, int n, int r2, float v_in,
                                         float &out)
{
    typedef typename A::Sum Sum;
    typedef typename A::Val Val;
    const Val v = A::from_input(v_in);
    const bool c1 = it < n;                 // stage 1 still has input
    const bool c2 = it - 1 < n + r2;        // stage 2 still has input
    const bool d3 = it - 2 >= r2;           // stage 3 has summed its first 2 r inputs: it produces and pops
    const bool d4 = it - 3 >= 2 * r2;       // stage 4 likewise
    const Val old1 = A::load(rp), old2 = A::load(rp + pitch), old3 = A::load(rp + 2 * pitch),
              old4 = A::load(rp + 3 * pitch);
    // stage 4 at step it - 3, on what stage 3 produced in the previous iteration
    const Sum s4a = A::add(st.s4, st.e3);
    out = A::to_float(A::round(s4a));
    const Sum s4b = A::sub(s4a, old4);
    st.s4 = d4 ? s4b : s4a;
    A::store(rp + 3 * pitch, st.e3);
    // stage 3 at step it - 2
    const Sum s3a = A::add(st.s3, st.e2);
    const Val e3n = A::round(s3a);
    const Sum s3b = A::sub(s3a, old3);
    st.e3 = d3 ? e3n : st.e3;
    st.s3 = d3 ? s3b : s3a;
    A::store(rp + 2 * pitch, st.e2);
    // stage 2 at step it - 1
    const Sum s2a0 = A::add(st.s2, st.e1);
    const Sum s2a = c2 ? s2a0 : st.s2;
    st.e2 = A::round(s2a);
    st.s2 = A::sub(s2a, old2);
    A::store(rp + pitch, st.e1);
    // stage 1 at step it
    const Sum s1a0 = A::add(st.s1, v);
    const Sum s1a = c1 ? s1a0 : st.s1;
    st.e1 = A::round(s1a);
    st.s1 = A::sub(s1a, old1);
    A::store(rp, v);
    return d4;
}

// One line by one thread.  cur: its first four samples.  ring: slot s of stage q at
// ring[(q * r2 + s) * L].
template <typename A>
__device__ __forceinline__ void box_time_line(const TimeLine &in, int line, TimeLine::Raw &cur, float *dst,
                                              float *ring, int L, int r2, float div)
{
    const int T = in.T, F = in.F, total = T + 2 * r2 + 3, out_start = 2 * r2 + 3, pitch = r2 * L;
    for (int s = 0; s < 4 * r2; s++) ring[s * L] = 0.0f;                  // all four rings (0 in either arithmetic)
    BoxState<A> st = {0, 0, 0, 0, 0, 0, 0};
    int slot = 0;
    for (int it0 = 0; it0 < total; it0 += 4) {
        TimeLine::Raw next;
#pragma unroll
        for (int k = 0; k < 4; k++) next.f[k] = 1;
        if (it0 + 4 < T) in.load4(line, it0 + 4, next);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int it = it0 + k;
            if (it < total) {
                float o;
                box_iter<A>(st, ring + slot * L, pitch, it, T, r2, TimeLine::masked(cur, k), o);
                if (it >= out_start) dst[(int64_t) (it - out_start) * F] = __fdiv_rn(o, div);
                slot = (slot + 1 == r2) ? 0 : slot + 1;
            }
        }
        cur = next;
    }
}

__device__ __noinline__ void box_time_pass(const float *__restrict__ data, const uint8_t *__restrict__ flags,
                                           float *__restrict__ weight, float *__restrict__ out, int T, int F, int r)
{
    const int tid = td_vtid(), r2 = 2 * r;
    const int L = td_lines_that_fit(TD_SMEM_WORDS, 4 * r2, TD_THREADS);
    const float div = box_divisor(r);
    const int lines = 2 * F;
    const TimeLine in = {data, flags, T, F};
    TimeLine::Raw nxt;
#pragma unroll
    for (int k = 0; k < 4; k++) nxt.f[k] = 1;
    if (tid < L && tid < lines) in.load4(tid, 0, nxt);
    for (int base = 0; base < lines; base += L) {
        const int line = base + tid;
        if (tid < L && line < lines) {
            TimeLine::Raw cur = nxt;
            if (line + L < lines) in.load4(line + L, 0, nxt);             // the next round's first samples
            float *ring = s_sm + tid;
            if (line < F) {
                if (r <= TD_INT_RADIUS) box_time_line<BoxInt>(in, line, cur, weight + line, ring, L, r2, div);
                else box_time_line<BoxF64>(in, line, cur, weight + line, ring, L, r2, div);
            } else {
                box_time_line<BoxF64>(in, line, cur, out + (line - F), ring, L, r2, div);
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *src, bool valid)
{
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;                              // 0: nothing is read, the word is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// Frequency axis of a masked filter.  Lines are pairs (weights of row t, data of row t); a block
// takes L <= 32 of them at a time, 8 per warp, and moves them through transposing tiles: rows of
// 128 bytes in (cp.async, the next tile in flight while this one is worked on), one column per
// iteration inside.  FROM_MASK: the lines are made on the fly from `data` and `flags` (no
// time-axis pass before this one), else they are weight / out as the time-axis pass left them.
// The results leave through the tile they came in by, 4 r + 3 positions behind the reads, already
// normalised: out = filtered data / filtered weight, NaN where that weight is 0
// (twodflag.py:392-399); `weight` is not written.
template <bool FROM_MASK>
__device__ __noinline__ void box_freq_pass(const float *__restrict__ data, const uint8_t *__restrict__ flags,
                                           const float *weight, float *out, int T, int F, int r)
{
    const int tid = td_vtid(), lane = tid & 31, warp = tid >> 5, r2 = 2 * r;
    const int lines = 2 * T;
    // per line: its four rings and a row in each of the two tiles
    int L = td_lines_that_fit(TD_SMEM_WORDS, 4 * r2 + 66, TD_THREADS / 4) & ~1;
    const float div = box_divisor(r);
    const int total = F + 2 * r2 + 3;
    const int ll = tid >> 2, p = tid & 3;                 // this thread's line of the group, and its stage
    float *tiles = s_sm, *ring = s_sm + 2 * L * 33 + tid;
    const int rstride = 4 * L;
    const int ll0 = warp * 8;                             // the warp's lines: ll0 .. ll0 + 7
    for (int first = 0; first < lines; first += L) {
        const int nl = min(L, lines - first);             // lines of this group
        const int mine = max(0, min(8, nl - ll0));        // of which this warp's
        if (mine == 0) continue;                          // warp-uniform
        const bool active = ll < nl;
        // request the tile of iterations step0 .. step0 + 31 (one commit group per tile)
        auto request = [&](float *tile, int step0) {
            const int idx = step0 + lane;
            const bool inside = idx < F;
            for (int j = 0; j < mine; j++) {
                const int line = first + ll0 + j;
                const int64_t row = (int64_t) (line >> 1) * F;
                float *dstw = tile + (ll0 + j) * 33 + lane;
                if (FROM_MASK) {
                    float x = 0.0f;
                    if (inside) {
                        x = (line & 1) ? data[row + idx] : 1.0f;
                        if (flags[row + idx]) x = 0.0f;
                    }
                    *dstw = x;
                } else {
                    const float *src = ((line & 1) ? out : weight) + row;
                    cp_async4(dstw, src + (inside ? idx : 0), inside);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // threads without a line run the same instructions on a word of their own (no branch in
        // the loops: a divergence check around every shuffle costs as much as the stage itself)
        float *const my_ring = active ? ring : reinterpret_cast<float *>(s_whist) + tid;
        const int my_stride = active ? rstride : 0;
        if (active)
            for (int s = 0; s < r2; s++) ring[s * rstride] = 0.0f;
        BoxStage<BoxF64> st;
        st.init(p, F, r2);
        const bool head = active && p == 0, tail = active && p == 3;
        float *rp = my_ring;
        int slot = 0, tb = 0;
        request(tiles, 0);
        for (int step0 = 0; step0 < total; step0 += 32, tb ^= 1) {
            float *tile = tiles + tb * L * 33;
            if (step0 + 32 < total) {
                request(tiles + (tb ^ 1) * L * 33, step0 + 32);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            float *cell = tile + ll * 33;
            const int kmax = min(32, total - step0);
            if (step0 >= 2 * r2 + 3 && step0 + 32 <= F) {
                // the bulk of the line: every stage adds and pops in all 32 iterations of the tile
#pragma unroll 4
                for (int k = 0; k < 32; k++) {
                    const float from_prev = td_shfl_up1(st.e);
                    const float en = st.step_steady(head ? cell[k] : from_prev, rp);
                    if (tail) cell[k] = en;
                    slot++;
                    rp += my_stride;
                    if (slot == r2) {
                        slot = 0;
                        rp = my_ring;
                    }
                }
            } else
            for (int k = 0; k < kmax; k++) {
                const float from_prev = td_shfl_up1(st.e);
                const float x = head ? cell[k] : from_prev;
                const float en = st.step(step0 + k, x, rp);
                if (tail) cell[k] = en;                                // unnormalised; only positions >= 0 are used
                slot++;
                rp += my_stride;
                if (slot == r2) {
                    slot = 0;
                    rp = my_ring;
                }
            }
            __syncwarp();
            const int oi = step0 + lane - 3 - 2 * r2;            // line position of what sits in column `lane`
            if (oi >= 0 && oi < F) {
                for (int j = 0; j < mine; j += 2) {
                    const float w = __fdiv_rn(tile[(ll0 + j) * 33 + lane], div);
                    const float d = __fdiv_rn(tile[(ll0 + j + 1) * 33 + lane], div);
                    out[(int64_t) ((first + ll0 + j) >> 1) * F + oi] = (w == 0.0f) ? td_nan() : __fdiv_rn(d, w);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

// ---- fallback for radii whose rings do not fit the shared memory: the reference's four passes
// in place on a padded copy in global memory.  One line: n samples at data[k * stride], work
// array padded[k * pstride] of n + 4 r floats.
__device__ __noinline__ void box_line_in_place(const float *data, int64_t stride, float *padded, int64_t pstride, int n, int r,
                                  float *out, int64_t ostride, float divisor)
{
    const int K = TD_PASSES;
    const int padding = r * K, len = n + padding;
    for (int i = 0; i < padding; i++) padded[(int64_t) i * pstride] = 0.0f;
    for (int i = 0; i < n; i++) padded[(int64_t) (padding + i) * pstride] = data[(int64_t) i * stride];
    for (int p = 1; p <= K; p++) {
        double s = 0.0;
        for (int i = 0; i < min(2 * r, len); i++) s = __dadd_rn(s, (double) padded[(int64_t) i * pstride]);
        for (int i = 0; i < len; i++) {
            if (i + 2 * r < len) s = __dadd_rn(s, (double) padded[(int64_t) (i + 2 * r) * pstride]);
            const float prev = padded[(int64_t) i * pstride];
            padded[(int64_t) i * pstride] = (float) s;
            s = __dsub_rn(s, (double) prev);
        }
    }
    for (int i = 0; i < n; i++) out[(int64_t) i * ostride] = __fdiv_rn(padded[(int64_t) i * pstride], divisor);
}

// twodflag.py:360-400: out = filtered (data with flagged samples zeroed) / filtered (weights),
// NaN where the filtered weight is exactly 0.  Time axis first, then frequency
// (twodflag.py:313-356).  The ring-based passes mask on the way in and normalise on the way out;
// radii too large for the rings go through separate elementwise steps and the in-place lines.
__device__ __noinline__ void masked_gaussian(const float *data, const uint8_t *flags, int T, int F, int r_t, int r_f,
                                             float *out, float *weight)
{
    const int tid = threadIdx.x, A = T * F;
    const bool time_ring = r_t > 0 && 8 * r_t <= TD_SMEM_WORDS;                 // a line fits
    const bool freq_ring = r_f > 0 && 2 * (8 * r_f + 66) <= TD_SMEM_WORDS;     // 2 lines fit
    const bool masked_first = time_ring || (r_t == 0 && freq_ring);
    if (!masked_first) {
        for (int i = tid; i < A; i += TD_THREADS) {
            const bool fl = flags[i] != 0;
            weight[i] = fl ? 0.0f : 1.0f;
            out[i] = fl ? 0.0f : data[i];
        }
        __syncthreads();
        td_mark(18);
    }
    float *pad = s_b.pad;
    const int64_t pad_n = s_b.pad_n;
    if (time_ring) {
        box_time_pass(data, flags, weight, out, T, F, r_t);
        td_mark(16);
    } else if (r_t > 0) {
        const float div = box_divisor(r_t);
        for (int i = tid; i < 2 * F; i += TD_THREADS) {
            const int which = i >= F, f = which ? i - F : i;
            float *arr = which ? out : weight;
            box_line_in_place(arr + f, F, pad + which * pad_n + f, F, T, r_t, arr + f, F, div);
        }
        __syncthreads();
    }
    if (freq_ring) {
        if (r_t == 0) box_freq_pass<true>(data, flags, weight, out, T, F, r_f);
        else box_freq_pass<false>(data, flags, weight, out, T, F, r_f);
        td_mark(17);
        return;
    }
    if (r_f > 0) {
        const float div = box_divisor(r_f);
        const int64_t plen = F + (int64_t) r_f * TD_PASSES;
        for (int i = tid; i < 2 * T; i += TD_THREADS) {
            const int which = i >= T, t = which ? i - T : i;
            float *arr = which ? out : weight;
            box_line_in_place(arr + (int64_t) t * F, 1, pad + which * pad_n + t * plen, 1, F, r_f,
                              arr + (int64_t) t * F, 1, div);
        }
        __syncthreads();
    }
    for (int i = tid; i < A; i += TD_THREADS) {
        const float w = weight[i];
        out[i] = (w == 0.0f) ? td_nan() : __fdiv_rn(out[i], w);
    }
    __syncthreads();
    td_mark(18);
}

// twodflag.py:200-251, one row by one warp.  NaN runs are filled from the values on both sides
// (linearly, in float64), or with the nearest value at the ends of the row, or with zeros if the
// row holds nothing else.  `next_valid` (one int per element) is scratch.
__device__ __noinline__ void interpolate_row(float *row, int *next_valid, int n, int lane)
{
    // backward sweep: for every element the position of the first valid one at or after it
    int carry = n;
    bool any_nan = false;
    for (int base = ((n - 1) / 32) * 32; base >= 0; base -= 32) {
        const int i = base + lane;
        const bool valid = i < n && !isnan(row[i]);
        const uint32_t m = __ballot_sync(FULL, valid);
        if (__popc(m) != min(32, n - base)) any_nan = true;
        const uint32_t at_or_after = m & (FULL << lane);
        if (i < n) next_valid[i] = at_or_after ? base + __ffs(at_or_after) - 1 : carry;
        if (m) carry = base + __ffs(m) - 1;
    }
    if (!any_nan) return;                                          // warp-uniform
    __syncwarp();
    // forward sweep: the last valid position before each NaN, then the fill (valid elements never change)
    int prev_i = -1;
    float prev_v = 0.0f;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const float v = i < n ? row[i] : 0.0f;
        const bool valid = i < n && !isnan(v);
        const uint32_t m = __ballot_sync(FULL, valid);
        const uint32_t before = m & ((1u << lane) - 1u);
        const int src = before ? 31 - __clz(before) : 0;
        const float tv = __shfl_sync(FULL, v, src);
        const int p = before ? base + src : prev_i;
        const float pv = before ? tv : prev_v;
        const int last_src = m ? 31 - __clz(m) : 0;
        const float lv = __shfl_sync(FULL, v, last_src);
        if (i < n && !valid) {
            const int q = next_valid[i];
            float fill;
            if (p < 0 && q >= n) {
                fill = 0.0f;
            } else if (p < 0) {
                fill = row[q];
            } else if (q >= n) {
                fill = pv;
            } else {
                const double grad = __ddiv_rn((double) __fsub_rn(row[q], pv), (double) (q - p));
                fill = (float) __dadd_rn((double) pv, __dmul_rn((double) (i - p), grad));
            }
            row[i] = fill;
        }
        if (m) {
            prev_i = base + last_src;
            prev_v = lv;
        }
    }
}

// twodflag.py:404-463.  flags_in is not modified; `work` receives the growing mask.
// spectrum: the 1 x F median spectrum (no time axis) instead of the T x F array.
__device__ __noinline__ void background2d(bool spectrum, int ph)
{
    const TdArgs &a = s_a;
    const int tid = td_vtid(), lane = tid & 31, warp = tid >> 5;
    const int T = spectrum ? 1 : (int) a.p.n_time, F = a.a_freq, A = T * F;
    const float *data = spectrum ? s_b.spec_data : s_b.data;
    const uint8_t *flags_in = spectrum ? s_b.spec_flags : s_b.flags;
    float *bg = spectrum ? s_b.spec_bg : s_b.bg;
    float *weight = spectrum ? s_b.spec_weight : s_b.weight;
    uint8_t *work = spectrum ? s_b.spec_work : s_b.work;
    for (int i = tid; i < A; i += TD_THREADS) work[i] = flags_in[i] != 0;
    __syncthreads();
    for (int ef = a.p.background_iterations; ef >= 1; ef--) {
        masked_gaussian(data, work, T, F, spectrum ? 0 : a.p.r_time[ef], a.p.r_freq[ef], bg, weight);
        td_mark(ph);
        // per chunk (they are disjoint sets of columns): |residual|, its median over the samples
        // still unflagged, rejection above background_reject sigma.  One warp per chunk.
        for (int c = warp; c < a.p.n_chunks; c += TD_WARPS) {
            const int c0 = (int) a.p.chunk_ends[c], cl = (int) a.p.chunk_ends[c + 1] - c0;
            if (cl <= 0) continue;
            // the first pass of the select makes the residuals (and leaves them in bg)
            auto key_first = [=](int t, int j) -> uint32_t {
                const int o = t * F + c0 + j;
                const float res = fabsf(__fsub_rn(data[o], bg[o]));
                bg[o] = res;
                return work[o] ? KEY_SKIP : __float_as_uint(res);     // residuals are >= 0
            };
            auto key_at = [=](int t, int j) -> uint32_t {
                const int o = t * F + c0 + j;
                return work[o] ? KEY_SKIP : __float_as_uint(bg[o]);
            };
            const float med = warp_median(key_first, key_at, T, cl, s_whist + warp * 256, lane);
            const double threshold = __dmul_rn((double) med, __dmul_rn(TD_MAD_NORMAL, a.p.background_reject));
            for (int t = 0; t < T; t++)
                for (int j = lane; j < cl; j += 128) {
                    float res[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) res[u] = (j + 32 * u < cl) ? bg[t * F + c0 + j + 32 * u] : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (j + 32 * u < cl && (double) res[u] > threshold) work[t * F + c0 + j + 32 * u] = 1;
                }
        }
        __syncthreads();
        td_mark(ph + 1);
    }
    masked_gaussian(data, work, T, F, spectrum ? 0 : a.p.r_time[1], a.p.r_freq[1], bg, weight);
    td_mark(ph);
    for (int t = warp; t < T; t += TD_WARPS)
        interpolate_row(bg + (int64_t) t * F, reinterpret_cast<int *>(weight) + (int64_t) t * F, F, lane);
    __syncthreads();
    td_mark(ph + 2);
}

// ------------------------------------------------------------------ SumThreshold, one line (twodflag.py:493-560)
// The reference, per window size w (in order): lim = thr / tf; samples already flagged (positive
// or negative side) are clamped to +-lim; cum = running float64 sum; window k .. k + w - 1 is a
// hit if (cum[k + w] - cum[k]) * float32(1 / w) > lim (positive side; the negative side with
// -scale); after the sweep every hit flags the w samples of its window.  Here: one sweep per
// window size with the last w + 1 cumulative sums in a ring (ring[slot * L]); sample k is flagged
// exactly if one of the windows k - w + 1 .. k fired, i.e. if the last hit is fewer than w windows
// back when window k has been tested - so the flag of sample k is final w - 1 steps after it was read, and
// reads of the masks (32 samples at a time) always see the state before this window size.
// x(i), i < len: the line.  posw / negw: the line's bit masks, word j at [j * wpitch].
template <typename LineAt>
__device__ void sum_threshold_line(const LineAt &x_at, int len, float thr32, const int *windows, int n_windows,
                                   const double *tf, uint32_t *posw, uint32_t *negw, int64_t wpitch, double *ring,
                                   int L)
{
    const int nw = (len + 31) >> 5;
    for (int j = 0; j < nw; j++) posw[j * wpitch] = negw[j * wpitch] = 0u;
    for (int wi = 0; wi < n_windows; wi++) {
        const int w = windows[wi];
        const float lim = (float) __ddiv_rn((double) thr32, tf[wi]);
        const float nlim = -lim;
        const double scale = (double) (float) __ddiv_rn(1.0, (double) w);       // np.float32(1.0 / window)
        const double dlim = (double) lim, ndlim = -dlim;
        // "any of the last w hits": windows since the last hit (0 = this one) < w
        int since_p = w, since_n = w;
        double cum = 0.0;
        double *const ring_end = ring + (int64_t) (w + 1) * L;
        double *head = ring;                                  // the newest cumulative sum
        ring[0] = 0.0;
        uint32_t rp = 0u, rn = 0u, wp = 0u, wn = 0u;
        const int total = len + w - 1;
        float xs[4], xn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) xs[u] = (u < len) ? x_at(u) : 0.0f;
        for (int i0 = 0; i0 < total; i0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) xn[u] = (i0 + 4 + u < len) ? x_at(i0 + 4 + u) : 0.0f;   // one group ahead
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u;
                if (i >= total) break;
                const bool inside = i < len;
                if (inside) {
                    if ((i & 31) == 0) {
                        rp = posw[(i >> 5) * wpitch];
                        rn = negw[(i >> 5) * wpitch];
                    }
                    const uint32_t bit = 1u << (i & 31);
                    float x = xs[u];
                    if ((rp & bit) && x > lim) x = lim;
                    else if ((rn & bit) && x < nlim) x = nlim;
                    cum = __dadd_rn(cum, (double) x);
                    head += L;
                    if (head == ring_end) head = ring;
                    *head = cum;
                }
                const int k = i + 1 - w;
                if (k >= 0) {
                    bool hp = false, hn = false;
                    if (inside) {
                        double *oldest = head + L;
                        if (oldest == ring_end) oldest = ring;
                        // sum * -scale > lim  <=>  sum * scale < -lim: negation is exact
                        const double avg = __dmul_rn(__dsub_rn(cum, *oldest), scale);
                        hp = avg > dlim;
                        hn = avg < ndlim;
                    }
                    since_p = hp ? 0 : since_p + 1;
                    since_n = hn ? 0 : since_n + 1;
                    wp |= (since_p < w ? 1u : 0u) << (k & 31);
                    wn |= (since_n < w ? 1u : 0u) << (k & 31);
                    if ((k & 31) == 31 || k == len - 1) {
                        if (wp) posw[(k >> 5) * wpitch] |= wp;
                        if (wn) negw[(k >> 5) * wpitch] |= wn;
                        wp = wn = 0u;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) xs[u] = xn[u];
        }
    }
}

__device__ __forceinline__ float scaled_threshold(float med, double outlier_nsigma)
{
    if (isnan(med)) return __int_as_float(0x7f800000);
    return (float) __dmul_rn((double) med, __dmul_rn(outlier_nsigma, TD_MAD_NORMAL));
}

// SumThreshold along frequency with per-chunk thresholds (axis 1): T rows x n_chunks chunks.
// Thresholds: one warp per (row, chunk).  Lines: one thread per (row, chunk) over the chunk
// extended by the largest window - 1 on both sides (clipped to the row); the flags of the chunk
// itself are then written out by one warp per line.
__device__ __noinline__ void sum_threshold_freq(bool spectrum, int ph)
{
    const TdArgs &a = s_a;
    const TdBuffers &b = s_b;
    const int tid = td_vtid(), lane = tid & 31, warp = tid >> 5;
    const int T = spectrum ? 1 : (int) a.p.n_time, F = a.a_freq;
    const float *data = spectrum ? b.spec_data : b.data;
    const uint8_t *flags = spectrum ? b.spec_flags : b.flags;
    uint8_t *out = spectrum ? b.spec_out : b.ffl;
    const int nc = a.p.n_chunks, max_w = a.max_wf, lines = T * nc;
    for (int i = warp; i < lines; i += TD_WARPS) {
        const int t = i / nc, c = i - t * nc;
        const int c0 = (int) a.p.chunk_ends[c], cl = (int) a.p.chunk_ends[c + 1] - c0;
        const float *row = data + (int64_t) t * F + c0;
        const uint8_t *frow = flags + (int64_t) t * F + c0;
        auto key_at = [=](int, int j) -> uint32_t {
            return frow[j] ? KEY_SKIP : (__float_as_uint(row[j]) & 0x7fffffffu);
        };
        float med;
        if (cl <= 512) {
            auto key1 = [=](int j) -> uint32_t { return key_at(0, j); };
            med = warp_median_small<16>(key1, cl, s_whist + warp * 256, lane);
        } else {
            med = warp_median(key_at, key_at, 1, cl, s_whist + warp * 256, lane);
        }
        if (lane == 0) b.thr[i] = scaled_threshold(med, a.p.outlier_nsigma);
    }
    __syncthreads();
    td_mark(ph);
    double *ring = reinterpret_cast<double *>(s_sm);
    const int L = td_lines_that_fit(TD_SMEM_WORDS / 2, max_w + 1, TD_THREADS);
    for (int base = 0; base < lines; base += L) {
        const int i = base + tid;
        if (tid < L && i < lines) {
            const int t = i / nc, c = i - t * nc;
            const int c0 = (int) a.p.chunk_ends[c], c1 = (int) a.p.chunk_ends[c + 1];
            if (c1 > c0) {
                const int p0 = max(c0 - max_w + 1, 0), p1 = min(c1 + max_w - 1, F);
                const float *line = data + (int64_t) t * F + p0;
                auto x_at = [=](int j) -> float { return line[j]; };
                sum_threshold_line(x_at, p1 - p0, b.thr[i], a.p.windows_freq, a.p.n_windows_freq, a.p.tf_freq,
                                   b.posw + i, b.negw + i, lines, ring + tid, L);
            }
        }
    }
    __syncthreads();
    for (int i = warp; i < lines; i += TD_WARPS) {
        const int t = i / nc, c = i - t * nc;
        const int c0 = (int) a.p.chunk_ends[c], c1 = (int) a.p.chunk_ends[c + 1];
        const int coff = c0 - max(c0 - max_w + 1, 0);
        for (int j = lane; j < c1 - c0; j += 32) {
            const int bitpos = coff + j;
            const uint32_t m = b.posw[(int64_t) (bitpos >> 5) * lines + i] | b.negw[(int64_t) (bitpos >> 5) * lines + i];
            out[(int64_t) t * F + c0 + j] = (m >> (bitpos & 31)) & 1u;
        }
    }
    __syncthreads();
    td_mark(ph + 1);
}

// Median over time of every column's unflagged samples (twodflag.py:120-158)
__device__ __noinline__ void median_spectrum()
{
    const TdBuffers &b = s_b;
    const int T = (int) s_a.p.n_time, F = s_a.a_freq;
    for (int f = threadIdx.x; f < F; f += TD_THREADS) {
        int n;
        const float med = column_median_any<false>(b.data + f, b.flags + f, b.vals + f, F, T, &n);
        b.spec_data[f] = n ? med : 0.0f;
        b.spec_flags[f] = n == 0;
    }
    __syncthreads();
}

// SumThreshold along time: one column per thread, one chunk [0, T)
__device__ __noinline__ void sum_threshold_time()
{
    const TdArgs &a = s_a;
    const TdBuffers &b = s_b;
    const int tid = td_vtid();
    const int T = (int) a.p.n_time, F = a.a_freq;
    double *ring = reinterpret_cast<double *>(s_sm);
    const int L = td_lines_that_fit(TD_SMEM_WORDS / 2, a.max_wt + 1, TD_THREADS);
    const int words = (T + 31) >> 5;
    for (int base = 0; base < F; base += L) {
        const int f = base + tid;
        if (tid < L && f < F) {
            int n;
            const float med = column_median_any<true>(b.data + f, b.flags + f, b.vals + f, F, T, &n);
            const float *col = b.data + f;
            const int64_t stride = F;
            auto x_at = [=](int j) -> float { return col[(int64_t) j * stride]; };
            sum_threshold_line(x_at, T, scaled_threshold(med, a.p.outlier_nsigma), a.p.windows_time,
                               a.p.n_windows_time, a.p.tf_time, b.posw + f, b.negw + f, F, ring + tid, L);
            for (int j = 0; j < words; j++) {
                const uint32_t m = b.posw[(int64_t) j * F + f] | b.negw[(int64_t) j * F + f];
                for (int t = j * 32; t < min(T, j * 32 + 32); t++) {
                    const uint8_t fl = (m >> (t & 31)) & 1u;
                    b.tfl[(int64_t) t * F + f] = fl;
                    b.flags[(int64_t) t * F + f] |= fl;
                }
            }
        }
    }
    __syncthreads();
}

// Combination of the three flag sets, smearing in time; back to the original channels, smearing
// in frequency, fill rules (twodflag.py:691-764)
__device__ __noinline__ void combine_and_unaverage()
{
    const TdArgs &a = s_a;
    const TdBuffers &b = s_b;
    const int tid = td_vtid(), lane = tid & 31, warp = tid >> 5;
    const int T = (int) a.p.n_time, F = a.a_freq;
    {
        const int lo = -(a.p.time_extend / 2), hi = lo + a.p.time_extend;
        for (int f = tid; f < F; f += TD_THREADS) {
            // any flag in rows [t + lo, t + hi) clipped to the array
            const uint8_t spec = b.spec_out[f];
            for (int t = 0; t < T; t++) {
                const int t0 = max(t + lo, 0), t1 = min(t + hi, T);
                uint8_t any = 0;
                for (int k = t0; k < t1; k++)
                    any |= spec | b.tfl[(int64_t) k * F + f] | b.ffl[(int64_t) k * F + f];
                b.comb[(int64_t) t * F + f] = any;
            }
        }
    }
    __syncthreads();
    td_mark(12);
    const int OF = (int) a.p.n_freq, avg = a.p.average_freq;
    const int lo = -(a.p.freq_extend / 2), hi = lo + a.p.freq_extend;
    for (int t = 0; t < T; t++) {
        const uint8_t *crow = b.comb + (int64_t) t * F;
        for (int f = tid; f < OF; f += TD_THREADS) {
            const int f0 = max(f + lo, 0), f1 = min(f + hi, OF);
            uint8_t any = 0;
            if (avg == 1)
                for (int k = f0; k < f1; k++) any |= crow[k];
            else
                for (int k = f0; k < f1; k++) any |= crow[k / avg];
            b.outb[(int64_t) t * OF + f] = any;
        }
    }
    __syncthreads();
    // rows with too many flags (counted before any filling), then columns likewise
    uint8_t *row_full = b.row_full, *col_full = b.col_full;
    for (int t = warp; t < T; t += TD_WARPS) {
        int tot = 0;
        for (int f = lane; f < OF; f += 32) tot += b.outb[(int64_t) t * OF + f];
        tot = __reduce_add_sync(FULL, tot);
        if (lane == 0) row_full[t] = (double) tot > a.p.flag_all_freq_frac * (double) OF;
    }
    for (int f = tid; f < OF; f += TD_THREADS) {
        int tot = 0;
        for (int t = 0; t < T; t++) tot += b.outb[(int64_t) t * OF + f];
        col_full[f] = (double) tot > (double) T * a.p.flag_all_time_frac;
    }
    __syncthreads();
    for (int t = 0; t < T; t++) {
        const bool rf = row_full[t] != 0;
        for (int f = tid; f < OF; f += TD_THREADS)
            if (rf || col_full[f]) b.outb[(int64_t) t * OF + f] = 1;
    }
    __syncthreads();
    td_mark(13);
}

// ------------------------------------------------------------------ one baseline (twodflag.py:768-881)
__global__ void __launch_bounds__(TD_THREADS, TD_BLOCKS_PER_SM)
twod_baseline_kernel(const TdArgs a)
{
    const int tid = threadIdx.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&a);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_a);
        for (int i = tid; i < (int) (sizeof(TdArgs) / 4); i += TD_THREADS) dst[i] = src[i];
    }
    const int T = (int) a.p.n_time, F = a.a_freq, A = T * F;
    for (int64_t blr = blockIdx.x; blr < a.nb; blr += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            td_buffers(a, blr, &s_b);
            s_last = clock64();
        }
        __syncthreads();
        const TdBuffers &b = s_b;
        median_spectrum();
        td_mark(0);
        // ---- background and SumThreshold of the spectrum
        background2d(true, 1);
        for (int f = tid; f < F; f += TD_THREADS) b.spec_data[f] = __fsub_rn(b.spec_data[f], b.spec_bg[f]);
        __syncthreads();
        sum_threshold_freq(true, 4);
        for (int t = 0; t < T; t++)
            for (int f = tid; f < F; f += TD_THREADS) b.flags[(int64_t) t * F + f] |= b.spec_out[f];
        __syncthreads();
        td_mark(15);
        // ---- 2-D background
        background2d(false, 6);
        for (int i = tid; i < A; i += TD_THREADS) b.data[i] = __fsub_rn(b.data[i], b.bg[i]);
        __syncthreads();
        td_mark(15);
        sum_threshold_time();
        td_mark(9);
        sum_threshold_freq(false, 10);
        combine_and_unaverage();
    }
}

// ------------------------------------------------------------------ output (twodflag.py:680-688)
// Block = 32 baselines x 32 channels of one dump, transposed through shared memory.
template <bool COMPLEX>
__global__ void __launch_bounds__(256)
twod_output_kernel(const TdArgs a)
{
    __shared__ uint8_t s_flag[32][33];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t bl_tile = (int64_t) blockIdx.x * 32;
    const int f_tile = (int) blockIdx.y * 32;
    const int t = (int) blockIdx.z;
    const int OF = (int) a.p.n_freq;
    {
        const int f = f_tile + lane;
        for (int bo = wrp; bo < 32; bo += 8) {
            const int64_t bl = bl_tile + bo;
            uint8_t v = 0;
            if (bl < a.nb && f < OF) {
                TdBuffers b;
                td_buffers(a, bl, &b);
                v = b.outb[(int64_t) t * OF + f];
            }
            s_flag[bo][lane] = v;
        }
    }
    __syncthreads();
    const int64_t blr = bl_tile + lane;
    for (int fo = wrp; fo < 32; fo += 8) {
        const int f = f_tile + fo;
        if (blr < a.nb && f < OF) {
            const int64_t idx = ((int64_t) t * OF + f) * a.p.n_bl + a.bl0 + blr;
            bool nan_in;
            if (COMPLEX) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(a.data) + idx);
                nan_in = isnan(v.x) || isnan(v.y);
            } else {
                nan_in = isnan(__ldg(reinterpret_cast<const float *>(a.data) + idx));
            }
            a.out_flags[idx] = (s_flag[lane][fo] || nan_in) ? 1 : 0;
        }
    }
}

void fill_derived(TdArgs &a)
{
    a.a_freq = (int) ((a.p.n_freq + a.p.average_freq - 1) / a.p.average_freq);
    a.max_rt = a.max_rf = 0;
    for (int e = 1; e <= a.p.background_iterations; e++) {
        if (a.p.r_time[e] > a.max_rt) a.max_rt = a.p.r_time[e];
        if (a.p.r_freq[e] > a.max_rf) a.max_rf = a.p.r_freq[e];
    }
    a.max_wf = a.max_wt = 1;
    for (int i = 0; i < a.p.n_windows_freq; i++)
        if (a.p.windows_freq[i] > a.max_wf) a.max_wf = a.p.windows_freq[i];
    for (int i = 0; i < a.p.n_windows_time; i++)
        if (a.p.windows_time[i] > a.max_wt) a.max_wt = a.p.windows_time[i];
    a.max_cl = 0;
    for (int c = 0; c < a.p.n_chunks; c++) {
        const int cl = (int) (a.p.chunk_ends[c + 1] - a.p.chunk_ends[c]);
        if (cl > a.max_cl) a.max_cl = cl;
    }
    a.big_radius = 8 * a.max_rt > TD_SMEM_WORDS || 2 * (8 * a.max_rf + 66) > TD_SMEM_WORDS;
    a.per_bl = td_layout(a, nullptr, nullptr);
}

int check_params(const ksp_twodflag_params *p)
{
    if (!p) return KSP_EINVAL;
    if (p->n_time < 1 || p->n_freq < 1 || p->n_bl < 0) return KSP_EINVAL;
    if (p->n_time > 4096 || p->n_freq > (1 << 21)) return KSP_ETOOLARGE;
    if (p->average_freq < 1 || p->background_iterations < 1 || p->background_iterations > 63) return KSP_EINVAL;
    if (p->n_windows_time < 0 || p->n_windows_time > KSP_TWOD_MAX_WINDOWS) return KSP_EINVAL;
    if (p->n_windows_freq < 0 || p->n_windows_freq > KSP_TWOD_MAX_WINDOWS) return KSP_EINVAL;
    for (int i = 0; i < p->n_windows_time; i++)
        if (p->windows_time[i] < 1 || p->windows_time[i] > KSP_TWOD_MAX_WINDOW) return KSP_ETOOLARGE;
    for (int i = 0; i < p->n_windows_freq; i++)
        if (p->windows_freq[i] < 1 || p->windows_freq[i] > KSP_TWOD_MAX_WINDOW) return KSP_ETOOLARGE;
    if (p->n_chunks < 1 || p->n_chunks > KSP_TWOD_MAX_CHUNKS) return KSP_EINVAL;
    const int64_t a_freq = (p->n_freq + p->average_freq - 1) / p->average_freq;
    if (p->chunk_ends[0] != 0 || p->chunk_ends[p->n_chunks] != a_freq) return KSP_EINVAL;
    for (int c = 0; c < p->n_chunks; c++)
        if (p->chunk_ends[c + 1] < p->chunk_ends[c]) return KSP_EINVAL;
    if (p->time_extend < 1 || p->freq_extend < 1) return KSP_EINVAL;
    for (int e = 1; e <= p->background_iterations; e++)
        if (p->r_time[e] < 0 || p->r_freq[e] < 0 || p->r_time[e] > 4096 || p->r_freq[e] > 4096) return KSP_EINVAL;
    if (p->n_time * a_freq > 0x3fffffff) return KSP_ETOOLARGE;
    return 0;
}

}  // namespace

extern "C" size_t ksp_twodflag_scratch_bytes(const ksp_twodflag_params *p, int64_t batch_baselines)
{
    if (check_params(p) || batch_baselines < 1) return 0;
    TdArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    fill_derived(a);
    return a.per_bl * (size_t) batch_baselines;
}

extern "C" int ksp_twodflag_resident_baselines(void)
{
    return TD_BLOCKS_PER_SM * ksp_sm_count();
}

extern "C" int ksp_twodflag_phases(unsigned long long *out, int n, int reset)
{
    if (n < 0 || n > KSP_TWOD_PHASES || (n > 0 && !out)) return KSP_EINVAL;
    if (n > 0) KSP_CUDA(cudaMemcpyFromSymbol(out, td_phase, sizeof(unsigned long long) * (size_t) n));
    if (reset) {
        const unsigned long long zero[KSP_TWOD_PHASES] = {0};
        KSP_CUDA(cudaMemcpyToSymbol(td_phase, zero, sizeof(zero)));
    }
    return 0;
}

extern "C" int ksp_twodflag(void *stream, const ksp_twodflag_params *p, const void *data,
                            const uint8_t *in_flags, uint8_t *out_flags, void *scratch, size_t scratch_bytes,
                            int64_t batch_baselines)
{
    int rc = check_params(p);
    if (rc) return rc;
    if (p->n_bl == 0) return 0;
    if (!data || !in_flags || !out_flags || !scratch || batch_baselines < 1) return KSP_EINVAL;
    if ((uintptr_t) scratch % 256) return KSP_EALIGN;
    cudaStream_t s = (cudaStream_t) stream;
    TdArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.data = data; a.in_flags = in_flags; a.out_flags = out_flags;
    fill_derived(a);
    if (scratch_bytes < a.per_bl * (size_t) batch_baselines) return KSP_ESCRATCH;
    a.scratch = (char *) scratch;
    const int resident = TD_BLOCKS_PER_SM * ksp_sm_count();
    for (int64_t bl0 = 0; bl0 < p->n_bl; bl0 += batch_baselines) {
        a.bl0 = bl0;
        a.nb = p->n_bl - bl0 < batch_baselines ? p->n_bl - bl0 : batch_baselines;
        if (ksp_divup(a.nb, 32) > 0x7fffffff) return KSP_ETOOLARGE;
        const dim3 g_avg((unsigned) ksp_divup(a.nb, 32), (unsigned) ksp_divup(a.a_freq, 32), (unsigned) p->n_time);
        const dim3 g_out((unsigned) ksp_divup(a.nb, 32), (unsigned) ksp_divup(p->n_freq, 32), (unsigned) p->n_time);
        if (p->is_complex) twod_average_kernel<true><<<g_avg, 256, 0, s>>>(a);
        else twod_average_kernel<false><<<g_avg, 256, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
        const int blocks = (int) (a.nb < resident ? a.nb : resident);
        twod_baseline_kernel<<<blocks, TD_THREADS, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
        if (p->is_complex) twod_output_kernel<true><<<g_out, 256, 0, s>>>(a);
        else twod_output_kernel<false><<<g_out, 256, 0, s>>>(a);
        KSP_CHECK_LAUNCH();
    }
    return 0;
}
