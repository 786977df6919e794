// Streaming median-of-absolute-deviations of one baseline-major row: the device code shared by
// madnz_stream_kernel (madnz.cu) and the dataflow flagger (dataflow.cu).  See madnz.cu for the
// reference citations.
#pragma once
#include "common.cuh"
#include "select.cuh"
#include <math.h>

namespace {

using namespace ksp;

constexpr int SEL_THREADS = 1024;

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

// ------------------------------------------------------------------ streaming MAD (baseline-major)
// One block of 256 threads per row, any row length, ~49 KB of shared memory so that four
// blocks share an SM and hide each other's latency.  The row is streamed from global memory
// ONCE:
//   1. bracket [LO, HI] around the median from 1024 samples of the row (32 warp-sorted groups
//      of 32; LO / HI = medians over the groups of their 44 % / 56 % quantiles);
//   2. one pass over the row: every thread counts its usable keys and those below LO, and
//      keeps the ~12 % of its keys that fall inside the bracket in a private list in shared
//      memory (slot n of thread t at word n * 256 + t: conflict-free, no atomics);
//   3. select inside the lists: histogram (2048 bins over the bracket), block scan, the 2-3
//      keys of the wanted bin are sorted by one warp.
// If the bracket misses the median (~0.2 % of rows), a private list overflows or the wanted
// bin is crowded (heavy ties), the row is redone with the 4-pass radix select of select.cuh.
constexpr int MS_THREADS = 256;
#ifndef MS_SLOTS_N
#define MS_SLOTS_N 38
#endif
constexpr int MS_SLOTS = MS_SLOTS_N;              // list slots per thread (mean use ~16 of 128 keys)
constexpr int MS_BINS = 2048;
#ifndef MS_QLO
#define MS_QLO 448u                       // bracket = these sample quantiles, in 1/1024
#define MS_QHI 576u
#endif
#ifndef MS_QLO_WIDE
#define MS_QLO_WIDE 478u                  // the same with 4096 samples: +-3.4 % (4.3 standard errors)
#define MS_QHI_WIDE 546u
#endif
#ifndef MS_UNROLL
#define MS_UNROLL 8                       // float4 loads in flight per thread (multiple of 4)
#endif
constexpr int MS_SMALL_CAP = 32;
constexpr uint32_t KEY_INF = 0x7f800000u;
static_assert(MS_SLOTS * MS_THREADS >= SELECT_HIST_WORDS, "the fallback's histogram aliases the lists");

__device__ __forceinline__ uint32_t sort32(uint32_t v, int lane)
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

// N independent bitonic sorts across the warp, interleaved so that the shuffle latencies overlap
template <int N>
__device__ __forceinline__ void sort32xN(uint32_t (&v)[N], int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
#pragma unroll
            for (int g = 0; g < N; g++) {
                const uint32_t other = __shfl_xor_sync(0xffffffffu, v[g], j);
                v[g] = keep_min ? min(v[g], other) : max(v[g], other);
            }
        }
    }
}

// warp-inclusive prefix sum
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Two-level histogram lookup by one warp: `coarse[j]` = sum of fine[32 j .. 32 j + 31], 64
// coarse bins.  Finds the fine bin holding rank r (0-based, r < total), the rank inside that
// bin and the bin's count.  Every lane returns the same values.
struct BinHit { uint32_t bin, r_in, count; };

__device__ __forceinline__ BinHit locate_rank(const uint32_t *coarse, const uint32_t *fine,
                                              uint32_t r, uint32_t c0, uint32_t c_excl, int lane)
{
    // c0 = coarse[2 lane], c_excl = exclusive prefix of this lane's two coarse bins
    const uint32_t c1 = coarse[2 * lane + 1];
    const int src = __ffs(__ballot_sync(0xffffffffu, r >= c_excl && r < c_excl + c0 + c1)) - 1;
    const uint32_t base = __shfl_sync(0xffffffffu, c_excl, src);
    const uint32_t first = __shfl_sync(0xffffffffu, c0, src);
    const bool second = r >= base + first;
    const uint32_t cbin = 2u * (uint32_t) src + (second ? 1u : 0u);
    const uint32_t r_c = r - base - (second ? first : 0u);            // rank inside the coarse bin
    const uint32_t f = fine[cbin * 32u + lane];
    const uint32_t f_incl = warp_scan_incl(f, lane), f_excl = f_incl - f;
    const int src2 = __ffs(__ballot_sync(0xffffffffu, r_c >= f_excl && r_c < f_incl)) - 1;
    BinHit h;
    h.bin = cbin * 32u + (uint32_t) src2;
    h.count = __shfl_sync(0xffffffffu, f, src2);
    h.r_in = r_c - __shfl_sync(0xffffffffu, f_excl, src2);
    return h;
}

// One element of the pass, branch-free (7 instructions): three float compares with |x| as an
// operand modifier, two predicated counters and a predicated append of the RAW bits (readers
// of the lists mask the sign).  In positive-float order these are exactly the integer key
// tests: usable <=> |x| > 0 (false for zero and NaN), and the bracket is LO <= |x| <= HI.
// `slot` is the shared-space byte address of the thread's next free list slot.
__device__ __forceinline__ void stream_key_fast(float x, float lo_f, float hi_f, uint32_t &nv,
                                                uint32_t &ge, uint32_t &slot)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p2, p3;\n\t"
        ".reg .f32 ax;\n\t"
        "abs.f32 ax, %3;\n\t"
        "setp.gt.f32 p0, ax, 0f00000000;\n\t"
        "setp.ge.f32 p2, ax, %4;\n\t"
        "setp.le.and.f32 p3, ax, %5, p2;\n\t"
        "@p0 add.u32 %0, %0, 1;\n\t"
        "@p2 add.u32 %1, %1, 1;\n\t"
        "@p3 st.shared.f32 [%2], %3;\n\t"
        "@p3 add.u32 %2, %2, %6;\n\t"
        "}"
        : "+r"(nv), "+r"(ge), "+r"(slot)
        : "f"(x), "f"(lo_f), "f"(hi_f), "n"(MS_THREADS * 4));
}

// The same with a capacity check, for batches that might fill the list and for the row tail.
__device__ __forceinline__ void stream_key_checked(float x, float lo_f, float hi_f, uint32_t &nv,
                                                   uint32_t &ge, uint32_t &slot, uint32_t slot_end,
                                                   bool &over)
{
    const float ax = fabsf(x);
    nv += (ax > 0.0f) ? 1u : 0u;
    ge += (ax >= lo_f) ? 1u : 0u;
    if (ax >= lo_f && ax <= hi_f) {
        if (slot < slot_end) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot), "f"(x));
            slot += MS_THREADS * 4;
        } else {
            over = true;
        }
    }
}

// Developer aid (tools/micro/madnz_phases.cu): cycles of thread 0 of every block between the ticks
#ifdef MS_PHASE_CLOCK
__device__ unsigned long long ms_phase[8];
#define MS_TICK(k)                                                                              \
    do {                                                                                        \
        if (threadIdx.x == 0) {                                                                 \
            const long long now__ = clock64();                                                  \
            atomicAdd(&ms_phase[k], (unsigned long long) (now__ - tick__));                     \
            tick__ = now__;                                                                     \
        }                                                                                       \
    } while (0)
#define MS_TICK_START() long long tick__ = clock64()
#else
#define MS_TICK(k) do { } while (0)
#define MS_TICK_START() do { } while (0)
#endif

// Shared memory of one row: lists, hist, coarse, misc in this order.
constexpr int MS_SMEM_WORDS = MS_SLOTS * MS_THREADS + MS_BINS + MS_BINS / 32 + 160;

// loads of the row: read-only path, or L2-coherent (ld.global.cg) when the row lives in a buffer
// that other blocks of the same launch rewrite (dataflow flagger)
template <bool COHERENT> __device__ __forceinline__ float4 ms_load4(const float4 *p)
{
    return COHERENT ? __ldcg(p) : __ldg(p);
}
template <bool COHERENT> __device__ __forceinline__ float ms_load(const float *p)
{
    return COHERENT ? __ldcg(p) : __ldg(p);
}

// One row by one block of MS_THREADS threads: *noise_out = noise of row[0 .. channels).
// vec_ok: the row starts on a 16-byte boundary.  smem: MS_SMEM_WORDS words, 16-byte aligned.
// Returns through block-uniform paths only; no barrier after the last shared-memory access.
template <bool COHERENT>
__device__ __forceinline__ void madnz_stream_row(const float *__restrict__ row, float *__restrict__ noise_out,
                                                 const int channels, const bool vec_ok, uint32_t *smem,
                                                 unsigned long long *fallbacks)
{
    uint32_t *lists = smem;                                   // MS_SLOTS * MS_THREADS
    uint32_t *hist = lists + MS_SLOTS * MS_THREADS;           // MS_BINS
    uint32_t *coarse = hist + MS_BINS;                        // MS_BINS / 32: sums of 32 consecutive bins
    uint32_t *misc = coarse + MS_BINS / 32;                   // 160
    // misc: 0 nv, 1 below, 2 overflow flag, 8 small count, 10 min key beyond the wanted bin,
    //       12 kept total, 16..63 fallback scratch, 64..95 group lows, 96..127 group highs,
    //       128..159 small list
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (channels <= 0) {                                      // empty row: no median
        if (tid == 0) *noise_out = __int_as_float(0x7fc00000);
        return;
    }
    MS_TICK_START();
    // the sample histogram of step 1 borrows the (still unused) list memory
    uint32_t *s_fine = lists, *s_coarse = lists + MS_BINS;
    for (int i = tid; i < MS_BINS; i += MS_THREADS) {
        hist[i] = 0u;
        s_fine[i] = 0u;
    }
    if (tid < MS_BINS / 32) {
        coarse[tid] = 0u;
        s_coarse[tid] = 0u;
    }
    if (tid < 16) misc[tid] = (tid == 10) ? 0xffffffffu : 0u;

    // ---- 1. bracket.  1024 samples of the row go into a histogram over the float bit
    //         patterns (bin = key >> 20: exponent and 3 mantissa bits; coarse = bin >> 5); every
    //         warp then finds the 44 % and 56 % sample quantiles with two warp scans and
    //         interpolates linearly inside their bins.  The bracket only has to CONTAIN the
    //         median - that is verified after the pass - so the estimate need not be exact.
    uint32_t lo, hi;
    {
        // 1024 sample positions; rows that allow 16-byte loads take the whole aligned float4 at
        // each position (4096 samples for the same number of memory sectors), which halves the
        // width of the bracket
        const bool wide = vec_ok && channels >= 4096;
        uint32_t key[16];
        if (wide) {
            const float4 *row4s = reinterpret_cast<const float4 *>(row);
            const uint32_t nv4 = (uint32_t) channels >> 2, step = nv4 >> 10;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t i = (uint32_t) tid * 4u + g;              // sample number, 0..1023
                uint32_t pos = (uint32_t) (((uint64_t) i * nv4) >> 10);
                pos += (((i * 2654435761u) >> 16) * step) >> 16;         // jitter inside the stride
                const float4 v = ms_load4<COHERENT>(row4s + min(pos, nv4 - 1u));
                key[4 * g] = __float_as_uint(v.x);
                key[4 * g + 1] = __float_as_uint(v.y);
                key[4 * g + 2] = __float_as_uint(v.z);
                key[4 * g + 3] = __float_as_uint(v.w);
            }
        } else {
            const uint32_t step = (uint32_t) channels >> 10;
            const int last = channels - 1;
#pragma unroll
            for (int g = 0; g < 16; g++) key[g] = 0u;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t i = (uint32_t) tid * 4u + g;
                uint32_t pos = (uint32_t) (((uint64_t) i * (uint32_t) channels) >> 10);
                pos += (((i * 2654435761u) >> 16) * step) >> 16;
                key[g] = __float_as_uint(ms_load<COHERENT>(row + min((int) pos, last)));
            }
        }
        __syncthreads();                                                 // histograms are zeroed
#pragma unroll
        for (int g = 0; g < 16; g++) {
            const uint32_t k = key[g] & 0x7fffffffu;
            if ((k - 1u) < KEY_INF) {
                atomicAdd(&s_fine[k >> 20], 1u);
                atomicAdd(&s_coarse[k >> 25], 1u);
            }
        }
        const uint32_t q_lo = wide ? MS_QLO_WIDE : MS_QLO, q_hi = wide ? MS_QHI_WIDE : MS_QHI;
        __syncthreads();
        MS_TICK(0);
        const uint32_t c0 = s_coarse[2 * lane];
        const uint32_t c01 = c0 + s_coarse[2 * lane + 1];
        const uint32_t c_incl = warp_scan_incl(c01, lane), c_excl = c_incl - c01;
        const uint32_t m = __shfl_sync(0xffffffffu, c_incl, 31);          // usable samples
        lo = 1u;
        hi = KEY_INF;
        if (m > 0) {                                                      // block-uniform
            const uint32_t r_lo = (m * q_lo) >> 10;
            const uint32_t r_hi = min(m - 1u, (m * q_hi + 1023u) >> 10);
            const BinHit a = locate_rank(s_coarse, s_fine, r_lo, c0, c_excl, lane);
            const BinHit b = locate_rank(s_coarse, s_fine, r_hi, c0, c_excl, lane);
            const float scale = 1048576.0f;                               // keys per bin
            lo = (a.bin << 20) + (uint32_t) (__fdividef((float) a.r_in, (float) a.count) * scale);
            hi = (b.bin << 20) + (uint32_t) (__fdividef((float) (b.r_in + 1u), (float) b.count) * scale);
            lo = max(lo, 1u);
            hi = min(hi, KEY_INF);
            hi = max(hi, lo);
        }
        __syncthreads();                              // the lists may now overwrite the sample histogram
        MS_TICK(1);
    }
    const uint32_t width = hi - lo + 1u;

    // ---- 2. the one pass over the row
    uint32_t n_mine;
    {
        const float lo_f = __uint_as_float(lo), hi_f = __uint_as_float(hi);
        const uint32_t slot0 = (uint32_t) __cvta_generic_to_shared(lists + tid);
        const uint32_t slot_end = slot0 + MS_SLOTS * MS_THREADS * 4;
        constexpr uint32_t BATCH_BYTES = 16 * MS_THREADS * 4;   // 16 appends
        uint32_t nv = 0, ge = 0, slot = slot0;
        bool over = false;
        if (vec_ok) {
            const float4 *row4 = reinterpret_cast<const float4 *>(row);
            const int n4 = channels >> 2;
            int i = tid;
            for (; i + (MS_UNROLL - 1) * MS_THREADS < n4; i += MS_UNROLL * MS_THREADS) {
                float4 v[MS_UNROLL];
#pragma unroll
                for (int u = 0; u < MS_UNROLL; u++) v[u] = ms_load4<COHERENT>(row4 + i + u * MS_THREADS);
#pragma unroll
                for (int h = 0; h < MS_UNROLL; h += 4) {
                    if (slot + BATCH_BYTES <= slot_end) {        // room for all 16: no checks
#pragma unroll
                        for (int u = h; u < h + 4; u++) {
                            stream_key_fast(v[u].x, lo_f, hi_f, nv, ge, slot);
                            stream_key_fast(v[u].y, lo_f, hi_f, nv, ge, slot);
                            stream_key_fast(v[u].z, lo_f, hi_f, nv, ge, slot);
                            stream_key_fast(v[u].w, lo_f, hi_f, nv, ge, slot);
                        }
                    } else {
#pragma unroll
                        for (int u = h; u < h + 4; u++) {
                            stream_key_checked(v[u].x, lo_f, hi_f, nv, ge, slot, slot_end, over);
                            stream_key_checked(v[u].y, lo_f, hi_f, nv, ge, slot, slot_end, over);
                            stream_key_checked(v[u].z, lo_f, hi_f, nv, ge, slot, slot_end, over);
                            stream_key_checked(v[u].w, lo_f, hi_f, nv, ge, slot, slot_end, over);
                        }
                    }
                }
            }
            for (; i < n4; i += MS_THREADS) {
                const float4 v = ms_load4<COHERENT>(row4 + i);
                stream_key_checked(v.x, lo_f, hi_f, nv, ge, slot, slot_end, over);
                stream_key_checked(v.y, lo_f, hi_f, nv, ge, slot, slot_end, over);
                stream_key_checked(v.z, lo_f, hi_f, nv, ge, slot, slot_end, over);
                stream_key_checked(v.w, lo_f, hi_f, nv, ge, slot, slot_end, over);
            }
            for (int j = (n4 << 2) + tid; j < channels; j += MS_THREADS)
                stream_key_checked(ms_load<COHERENT>(row + j), lo_f, hi_f, nv, ge, slot, slot_end, over);
        } else {
            for (int j = tid; j < channels; j += MS_THREADS)
                stream_key_checked(ms_load<COHERENT>(row + j), lo_f, hi_f, nv, ge, slot, slot_end, over);
        }
        n_mine = (slot - slot0) / (MS_THREADS * 4);
        const uint32_t nv_w = __reduce_add_sync(0xffffffffu, nv);
        const uint32_t below_w = __reduce_add_sync(0xffffffffu, nv - ge);
        const uint32_t kept_w = __reduce_add_sync(0xffffffffu, n_mine);
        const bool over_w = __any_sync(0xffffffffu, over);
        if (lane == 0) {
            atomicAdd(&misc[0], nv_w);
            atomicAdd(&misc[1], below_w);
            atomicAdd(&misc[12], kept_w);
            if (over_w) misc[2] = 1u;
        }
        MS_TICK(2);
    }

    // ---- 3. select inside the lists.  Histogram of the kept keys at two levels (2048 bins and
    //         their sums in groups of 32), so that after ONE barrier every warp can locate the
    //         wanted bin by itself with two warp scans.
    const uint32_t *mine = lists + tid;
    const int shift = (width <= (uint32_t) MS_BINS) ? 0 : (32 - __clz(width - 1u)) - 11;
    for (uint32_t k = 0; k < n_mine; k++) {
        const uint32_t b = ((mine[k * MS_THREADS] & 0x7fffffffu) - lo) >> shift;
        atomicAdd(&hist[b], 1u);
        atomicAdd(&coarse[b >> 5], 1u);
    }
    __syncthreads();
    MS_TICK(3);
    const uint32_t n_valid = misc[0];
    if (n_valid == 0) {                                       // block-uniform
        if (tid == 0) *noise_out = __int_as_float(0x7fc00000);
        return;
    }
    const bool even = !(n_valid & 1u);
    const uint32_t rank = (n_valid - 1u) >> 1;                // lower median, 0-based
    const uint32_t r_rel = rank - misc[1];                    // rank inside the lists (wraps if below)
    const uint32_t kept = misc[12];
    // the bracket must hold the lower median and, for an even count, the key after it
    bool fallback = (misc[2] != 0u) || (r_rel >= kept) || (even && r_rel + 1u >= kept);

    if (!fallback) {                                          // block-uniform
        const uint32_t c0 = coarse[2 * lane];
        const uint32_t c01 = c0 + coarse[2 * lane + 1];
        const uint32_t c_excl = warp_scan_incl(c01, lane) - c01;
        const BinHit h = locate_rank(coarse, hist, r_rel, c0, c_excl, lane);
        const bool exact_bins = shift == 0;                   // a bin is one key value
        const bool crowded = !exact_bins && h.count > (uint32_t) MS_SMALL_CAP;
        const bool collect = !exact_bins && !crowded;

        // second walk over the lists: the keys of the wanted bin, and the smallest key beyond
        // it.  d = key - first key of the bin; keys beyond the bin have d >= span, and
        // d - span wraps to >= 2^31 for all others.
        const uint32_t bin_first = lo + (h.bin << shift), span = 1u << shift;
        uint32_t beyond = 0xffffffffu;
        for (uint32_t k = 0; k < n_mine; k++) {
            const uint32_t key = mine[k * MS_THREADS] & 0x7fffffffu;
            const uint32_t d = key - bin_first;
            if (d < span && collect) misc[128 + atomicAdd(&misc[8], 1u)] = key;
            beyond = min(beyond, d - span);
        }
        beyond = __reduce_min_sync(0xffffffffu, beyond);
        if (lane == 0 && beyond < 0x80000000u) atomicMin(&misc[10], beyond);
        __syncthreads();
        MS_TICK(4);
        if (crowded) {
            fallback = true;                                  // heavy ties: block-uniform
        } else {
            if (warp == 0) {
                uint32_t v1, nxt;                             // ranks r_in and r_in + 1 inside the bin
                if (exact_bins) {
                    v1 = bin_first;
                    nxt = v1;
                } else {
                    const uint32_t srt = sort32(lane < (int) h.count ? misc[128 + lane] : 0xffffffffu, lane);
                    v1 = __shfl_sync(0xffffffffu, srt, (int) h.r_in);
                    nxt = __shfl_sync(0xffffffffu, srt, (int) min(h.r_in + 1u, 31u));
                }
                const uint32_t v2 = !even ? v1 : (h.r_in + 1u < h.count ? nxt : bin_first + span + misc[10]);
                if (lane == 0) *noise_out = mad_finish(v1, v2);
            }
            MS_TICK(5);
            return;
        }
    }
    // ---- plain radix select over the row in global memory (rare)
    {
        __syncthreads();
        uint32_t v1, v2;
        SelectScratch sc;
        sc.hist = lists;
        sc.misc = misc + 16;
        auto src = [row](int i) { return mad_key(ms_load<COHERENT>(row + i)); };
        block_median_keys<MS_THREADS>(src, channels, n_valid, sc, v1, v2);
        if (tid == 0) {
            *noise_out = mad_finish(v1, v2);
            atomicAdd(fallbacks, 1ull);
        }
    }
}

}  // namespace
