// Median of the non-zero magnitudes of one row held in registers (32 values per thread,
// one thread block per row): the selection behind NoiseEstMADT in the fused row kernel.
//
// Replaces the reference's 31-step bisection with ~33 block reductions per row
// (rank.mako:186-266, rfi/madnz_t.mako:72-87).  Exact, as a pure selection on the
// positive-float bit patterns ("keys"; ordering of keys == ordering of magnitudes):
//
//   1. Bracket.  Every warp bitonic-sorts 32 samples of the row with shuffles and takes
//      its 40 % and 60 % quantiles; the medians over the warps of those (one more warp
//      sort each) bracket the row median, [LO, HI], with ~20 % of the row inside.
//   2. Histogram select.  One pass over the registers counts the keys below LO and
//      histograms the keys inside [LO, HI] into 4096 linear-in-key bins (shared-memory
//      atomics, ~2 hits per bin, so next to no contention); a block scan finds the bin
//      holding the wanted rank; the next pass histograms that bin alone, now one key
//      per bin.  Two passes instead of the four of a plain 8-bit radix select, and
//      the first one touches the atomics for a fifth of the samples only.
//      If the rank falls outside the bracket (unrepresentative samples) the search
//      restarts on the full key range and simply takes one or two more passes.
//   3. Even count: the upper median is the same key if enough copies exist, else the
//      smallest key above it (one min-reduction pass).
//
// All threads of the block must call; T = blockDim.x is a multiple of 32, at most 1024.
#pragma once
#include "common.cuh"

namespace ksp {

constexpr int MAD_BINS = 4096;
constexpr int MAD_MISC_WORDS = 96;
constexpr uint32_t MAD_KEY_INF = 0x7f800000u;

struct MadScratch {
    uint32_t *hist_a;   // MAD_BINS words, zeroed by the caller's block before the call
    uint32_t *hist_b;   // MAD_BINS words, zeroed likewise
    uint32_t *misc;     // MAD_MISC_WORDS words, words 0..15 zeroed likewise
};

__device__ __forceinline__ uint32_t warp_sort_asc(uint32_t v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, v, j);
            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
            v = keep_min ? min(v, other) : max(v, other);
        }
    }
    return v;
}

// Exclusive prefix over the threads of the block of `own`, plus the block total.
// Uses misc[16..47]; contains two barriers.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t own, uint32_t *misc, int tid,
                                                         int nwarps, uint32_t &total)
{
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t incl = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) misc[16 + warp] = incl;
    __syncthreads();
    const uint32_t w = (lane < nwarps) ? misc[16 + lane] : 0u;
    const uint32_t before = __reduce_add_sync(0xffffffffu, lane < warp ? w : 0u);
    total = __reduce_add_sync(0xffffffffu, w);
    __syncthreads();
    return before + incl - own;
}

// D: this thread's 32 row values (zero where the row has no sample).  sample: one row value
// per thread, any position (used for the bracket only).  Returns noise = float32(1.4826 *
// median of the non-zero, non-NaN magnitudes), NaN if there is none; the same value in
// every thread.
__device__ float block_mad_noise(const float (&D)[32], float sample, const MadScratch &sc)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    uint32_t *misc = sc.misc;
    // misc words: 0 n_valid, 1 below, 2 bin, 3 count before bin, 4 count in bin,
    //             5 min key above, 6 LO, 7 HI, 16..47 per-warp slots, 48..79 per-warp slots

    // ---- 1. bracket from per-warp sample quantiles
    {
        uint32_t key = __float_as_uint(sample) & 0x7fffffffu;
        const bool valid = (key - 1u) < MAD_KEY_INF;          // 0 < |x| <= inf
        key = valid ? key : 0xffffffffu;
        const uint32_t sorted = warp_sort_asc(key, lane);
        const int m = __popc(__ballot_sync(0xffffffffu, valid));
        const int i_lo = (m * 13) >> 5;                        // ~40 % quantile
        const int i_hi = min(max(m - 1, 0), (m * 19 + 31) >> 5);  // ~60 % quantile
        const uint32_t lo_w = __shfl_sync(0xffffffffu, sorted, i_lo);
        const uint32_t hi_w = __shfl_sync(0xffffffffu, sorted, i_hi);
        if (lane == 0) {
            misc[16 + warp] = m ? lo_w : 0xffffffffu;
            misc[48 + warp] = m ? hi_w : 0xffffffffu;
        }
        if (tid == 0) misc[5] = 0xffffffffu;
        __syncthreads();
        if (warp < 2) {
            // warp 0 -> LO from the lows, warp 1 (warp 0 again if it is alone) -> HI
            for (int which = warp; which < 2; which += nwarps) {
                const uint32_t v = (lane < nwarps) ? misc[16 + 32 * which + lane] : 0xffffffffu;
                const uint32_t s = warp_sort_asc(v, lane);
                const int cnt = __popc(__ballot_sync(0xffffffffu, v != 0xffffffffu));
                const uint32_t pick = __shfl_sync(0xffffffffu, s, which ? cnt >> 1 : (max(cnt, 1) - 1) >> 1);
                if (lane == 0) misc[6 + which] = cnt ? min(pick, MAD_KEY_INF) : (which ? MAD_KEY_INF : 1u);
            }
        }
        __syncthreads();
    }
    uint32_t lo = misc[6];
    uint32_t hi = max(misc[7], lo);
    uint32_t width = hi - lo + 1u;

    // ---- 2. first pass: count valid keys, keys below LO, histogram of [LO, HI]
    int shift = (width <= (uint32_t) MAD_BINS) ? 0 : (32 - __clz(width - 1u)) - 12;
    {
        uint32_t nv = 0, below = 0;
        const uint32_t lo_m1 = lo - 1u;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t key = __float_as_uint(D[j]) & 0x7fffffffu;
            const uint32_t km1 = key - 1u;                     // zero wraps to the top: never counted
            nv += (km1 < MAD_KEY_INF) ? 1u : 0u;
            below += (km1 < lo_m1) ? 1u : 0u;
            const uint32_t d = key - lo;
            if (d < width) atomicAdd(&sc.hist_a[d >> shift], 1u);
        }
        nv = __reduce_add_sync(0xffffffffu, nv);
        below = __reduce_add_sync(0xffffffffu, below);
        if (lane == 0) {
            atomicAdd(&misc[0], nv);
            atomicAdd(&misc[1], below);
        }
    }
    __syncthreads();
    const uint32_t n_valid = misc[0];
    if (n_valid == 0) return __int_as_float(0x7fc00000);       // block-uniform
    const uint32_t rank = (n_valid - 1u) >> 1;                 // lower median, 0-based
    uint32_t below_total = misc[1];                            // valid keys < lo

    // ---- levels: locate the bin of `rank`, then refine inside it
    uint32_t *hist = sc.hist_a;
    uint32_t v1 = 0, count_le = 0;
    int level = 0;
    for (;;) {
        // bins of this thread: [tid * bpt, tid * bpt + bpt)
        const int bpt = (MAD_BINS + T - 1) / T;
        const int b_first = tid * bpt;
        uint32_t own = 0;
        for (int i = 0; i < bpt; i++) {
            const int b = b_first + i;
            own += (b < MAD_BINS) ? hist[b] : 0u;
        }
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(own, misc, tid, nwarps, total);
        const uint32_t r_rel = rank - below_total;             // wraps if rank < below_total
        const bool outside = r_rel >= total;                   // block-uniform
        if (!outside && r_rel >= excl && r_rel < excl + own) {
            uint32_t cum = excl;
            for (int i = 0; i < bpt; i++) {
                const int b = b_first + i;
                const uint32_t h = (b < MAD_BINS) ? hist[b] : 0u;
                if (r_rel >= cum && r_rel < cum + h) {
                    misc[2] = (uint32_t) b;
                    misc[3] = cum;
                    misc[4] = h;
                }
                cum += h;
            }
        }
        __syncthreads();
        if (outside) {
            // the bracket missed the median: search the whole key range instead
            lo = 1u;
            width = MAD_KEY_INF;                               // keys 1 .. 0x7f800000
            below_total = 0;
        } else {
            const uint32_t bin = misc[2];
            if (shift == 0) {
                v1 = lo + bin;
                count_le = below_total + misc[3] + misc[4];
                break;
            }
            below_total += misc[3];
            lo += bin << shift;
            width = min(1u << shift, hi - lo + 1u);
        }
        hi = lo + width - 1u;
        shift = (width <= (uint32_t) MAD_BINS) ? 0 : (32 - __clz(width - 1u)) - 12;
        // next histogram: hist_b is still clean on the first refinement, later ones re-clear
        level++;
        hist = (level & 1) ? sc.hist_b : sc.hist_a;
        if (level > 1) {
            for (int i = tid; i < MAD_BINS; i += T) hist[i] = 0u;
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t key = __float_as_uint(D[j]) & 0x7fffffffu;
            const uint32_t d = key - lo;
            if (d < width) atomicAdd(&hist[d >> shift], 1u);
        }
        __syncthreads();
    }

    // ---- 3. upper median for even counts
    uint32_t v2 = v1;
    if (!(n_valid & 1u) && count_le < rank + 2u) {             // block-uniform
        uint32_t m = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t key = __float_as_uint(D[j]) & 0x7fffffffu;
            m = min(m, key > v1 ? key : 0xffffffffu);
        }
        m = __reduce_min_sync(0xffffffffu, m);
        if (lane == 0) atomicMin(&misc[5], m);
        __syncthreads();
        v2 = misc[5];
    }
    const double a = (double) __uint_as_float(v1), b = (double) __uint_as_float(v2);
    const double med = (v1 == v2) ? a : (a + b) * 0.5;
    return __double2float_rn(1.4826 * med);
}

}  // namespace ksp
