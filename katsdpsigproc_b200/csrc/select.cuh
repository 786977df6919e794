// Exact rank selection for one row per thread block (shared by madnz_t and
// percentile5).
//
// Replaces the reference's rank library (rank.mako:186-266: 31-step bisection on
// the float bit pattern, each step a block-wide count with 2-3 barriers) by a
// most-significant-digit radix select: 4 passes of 8 bits over 32-bit
// order-preserving keys.  Each pass histograms the digit of the keys that still
// match the prefix found so far.  The histogram is replicated per lane
// (hist[bin][lane], bank == lane), so the 32 shared-memory atomics of a warp
// never collide, however clustered the float exponents are.  A pass ends with a
// 256-wide scan done by 8 warps with shuffles.
//
// Keys are unsigned 32-bit, larger key == larger value; KEY_SKIP (all ones)
// marks samples that take no part (zeros / NaN for MAD, padding).
#pragma once
#include "common.cuh"

namespace ksp {

constexpr uint32_t KEY_SKIP = 0xffffffffu;
constexpr int SELECT_HIST_WORDS = 256 * 32;

// order-preserving map float -> uint32 (handles negative values too)
__device__ __forceinline__ uint32_t float_to_key(float v)
{
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

struct SelectScratch {
    uint32_t *hist;     // SELECT_HIST_WORDS words
    uint32_t *misc;     // >= 48 words: [0..31] warp totals, [32] bin, [33] rank in bin, [34..] free
};

// KeySource: functor  uint32_t operator()(int i) const  for i in [0, n).
// Returns the key of rank `rank` (0-based, ascending) among keys != KEY_SKIP.
// Requires rank < number of such keys.  All threads of the block must call it;
// all return the same value.  THREADS must be a multiple of 256.
template <int THREADS, typename KeySource>
__device__ uint32_t block_radix_select(const KeySource &key_at, int n, uint32_t rank,
                                       const SelectScratch &sc)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint32_t prefix = 0;       // bits found so far (in place)
    uint32_t prefix_mask = 0;  // which bits of the key are already decided
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < SELECT_HIST_WORDS; i += THREADS) sc.hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) {
            uint32_t k = key_at(i);
            if (k != KEY_SKIP && (k & prefix_mask) == prefix)
                atomicAdd(&sc.hist[((k >> shift) & 0xffu) * 32 + lane], 1u);
        }
        __syncthreads();
        // 256 bins: thread t < 256 totals bin t over the 32 lane replicas, then the
        // first 8 warps scan the totals with shuffles
        uint32_t count = 0, excl_in_warp = 0;
        if (tid < 256) {
#pragma unroll 8
            for (int j = 0; j < 32; j++) count += sc.hist[tid * 32 + ((j + tid) & 31)];
            uint32_t incl = count;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) sc.misc[tid >> 5] = incl;
            excl_in_warp = incl - count;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t base = 0;
            for (int w = 0; w < (tid >> 5); w++) base += sc.misc[w];
            uint32_t excl = base + excl_in_warp;
            if (count != 0 && rank >= excl && rank < excl + count) {
                sc.misc[32] = (uint32_t) tid;
                sc.misc[33] = rank - excl;
            }
        }
        __syncthreads();
        uint32_t bin = sc.misc[32];
        rank = sc.misc[33];
        prefix |= bin << shift;
        prefix_mask |= 0xffu << shift;
        __syncthreads();
    }
    return prefix;
}

// Smallest key strictly greater than `v` (KEY_SKIP if none), and the number of
// keys <= v, over the non-skipped keys.  Block-wide; result broadcast.
template <int THREADS, typename KeySource>
__device__ void block_next_above(const KeySource &key_at, int n, uint32_t v, uint32_t &next,
                                 uint32_t &count_le, const SelectScratch &sc)
{
    const int tid = threadIdx.x;
    uint32_t best = KEY_SKIP, cnt = 0;
    for (int i = tid; i < n; i += THREADS) {
        uint32_t k = key_at(i);
        if (k != KEY_SKIP) {
            if (k <= v) cnt++;
            else best = min(best, k);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((tid & 31) == 0) {
        sc.misc[tid >> 5] = best;
        sc.hist[tid >> 5] = cnt;
    }
    __syncthreads();
    best = KEY_SKIP;
    cnt = 0;
    for (int w = 0; w < THREADS / 32; w++) {
        best = min(best, sc.misc[w]);
        cnt += sc.hist[w];
    }
    __syncthreads();
    next = best;
    count_le = cnt;
}

}  // namespace ksp
