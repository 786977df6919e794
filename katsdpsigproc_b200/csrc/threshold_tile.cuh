// SumThreshold: the per-tile device code shared by threshold_sum_kernel (threshold.cu) and the
// dataflow flagger (dataflow.cu).  See threshold.cu for the algorithm and the reference citations.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace {

using namespace ksp;

constexpr int RUN = 32;
// The staged span is an array of 128-byte runs in the layout TMA's 128-byte swizzle produces:
// 16-byte chunk c of run r lives at chunk (c ^ (r & 7)).  Eight consecutive threads reading
// "their" chunk i thus touch eight different bank groups, with no padding.
constexpr int PITCH = 32;
constexpr int TS_MAX_THREADS = 256;
constexpr int TS_MAX_WINDOWS = 7;    // windows up to 64 = two runs of reach
constexpr unsigned FULL = 0xffffffffu;
constexpr float FILTER_ERR = 2.5e-4f;

struct TsArgs {
    const float *dev_t;
    const float *noise;
    uint8_t *flags_t;      // byte output (or null)
    uint32_t *bits_t;      // bit-packed output (or null)
    int64_t channels, baselines;
    int64_t dev_stride, out_stride;   // out_stride: bytes per row, or words per row when packed
    int n_windows;
    int flag_value;
    int use_tma;           // stage the span with one TMA tile load (needs tmap)
    int two_buffers;       // with TMA: second span buffer, the next tile loads while this one is worked on
    int n_chunks;          // spans per row
    uint32_t *work;        // two-pass mode: work[0] = number of listed tiles, work[1..] = their ids
    int chunk_valid;       // channels produced per block (multiple of 32)
    int edge;              // halo on each side of a span (multiple of 32; 0 when one block per row)
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

// float offset of element q of the staged span / of chunk i of run r (swizzled layout)
__device__ __forceinline__ int span_offset(int q)
{
    const int r = q >> 5;
    return (r << 5) + ((((q >> 2) & 7) ^ (r & 7)) << 2) + (q & 3);
}
__device__ __forceinline__ int chunk_offset(int r, int i) { return (r << 5) + ((i ^ (r & 7)) << 2); }

// Exact evaluation of the candidate windows of one thread (rare, divergent): D_w[i] in tree
// order from the staged row and the published flags.
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
                v = fl ? 0.0f : rowbuf[span_offset(q)];
            }
            vals[i] = v;
        }
        for (int h = 1; h < win; h <<= 1)
            for (int i = 0; i < win; i += 2 * h) vals[i] = vals[i] + vals[i + h];
        if ((double) vals[0] > (double) thr_w * (double) (win - n_flagged)) fire |= 1u << j;
    }
    return fire;
}

// Running sums p[0..32] of staged run r with its flagged samples zeroed (rare path).
__device__ __forceinline__ void run_sums(const float *rowbuf, int r, uint32_t F, float (&p)[RUN + 1])
{
    p[0] = 0.0f;
#pragma unroll
    for (int k = 0; k < RUN / 4; k++) {
        const float4 q = *reinterpret_cast<const float4 *>(rowbuf + chunk_offset(r, k));
        const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = 4 * k + i;
            p[j + 1] = p[j] + (((F >> j) & 1u) ? 0.0f : v[i]);
        }
    }
}

// Candidate windows of size W (2..64) that start in this thread's run (rare path).
//   rowbuf, r the staged span and the own run's index; runs r + 1, r + 2 exist (zeros
//             beyond the span)
//   F, F1, F2 flag words of the three runs;  tw = thr_w;  err = bound on |S~ - D_w|
// The running sums are formed exactly as the kernel's rebuild() forms them.
template <int W>
__device__ __noinline__ uint32_t window_candidates(const float *rowbuf, int r, uint32_t F,
                                                   uint32_t F1, uint32_t F2, float tw, float err)
{
    float p[RUN + 1], p1[RUN + 1], p2[RUN + 1];
    run_sums(rowbuf, r, F, p);
    run_sums(rowbuf, r + 1, F1, p1);
    if (W > RUN) run_sums(rowbuf, r + 2, F2, p2);
    else p2[0] = 0.0f;
    const bool any_flag = (F | F1 | F2) != 0u;
    const float t_full = tw * (float) W;                   // exact: W is a power of two
    const float slack = err + 1.2e-7f * t_full;            // rounding of tw * (W - n) below
    uint32_t cand = 0;
#pragma unroll
    for (int j = 0; j < RUN; j++) {
        float s;
        if (W <= RUN) {
            if (j + W <= RUN) s = p[j + W] - p[j];
            else s = (p[RUN] - p[j]) + p1[j + W - RUN];
        } else {
            s = ((p[RUN] - p[j]) + p1[RUN]) + p2[j];
        }
        float t = t_full;
        if (any_flag) {
            int n = __popc(__funnelshift_r(F, F1, j) & ((W >= 32) ? FULL : ((1u << (W & 31)) - 1u)));
            if (W > RUN) n += __popc(__funnelshift_r(F1, F2, j));
            t = tw * (float) (W - n);
        }
        cand |= (!(s <= t - slack)) ? (1u << j) : 0u;
    }
    return cand;
}

// Per-window thresholds of one row (noise `noise`) and the thread-invariant parts of the vote,
// computed by ONE WARP (all 32 lanes call it; lane = its lane number) into thr[0 .. TS_THR_WORDS):
//   thr[w], w < n_windows   thr_w = f32((n_sigma * f64(noise)) * scales[w])
//   thr[8]                  strictest limit of the sizes 2..8 that exist (+inf if none)
//   thr[9] (bits)           bit 0: some existing size has a negative threshold (everything is
//                           hot); bits 4..6: sizes 16 / 32 / 64 exist with a threshold >= 0
// A size "exists" if it is below n_windows, not larger than the band and its threshold is not NaN.
constexpr int TS_THR_WORDS = 16;
__device__ __forceinline__ void ts_thresholds(float *thr, int lane, int n_windows, double n_sigma,
                                              float noise, const double *scales, int C)
{
    const float inf = __int_as_float(0x7f800000);
    float tw = 0.0f;
    if (lane < n_windows) tw = __double2float_rn((n_sigma * (double) noise) * scales[lane]);
    const bool exists = lane >= 1 && lane < n_windows && (1 << lane) <= C && tw == tw;
    const bool neg = exists && !(tw >= 0.0f);
    const unsigned some_neg = __ballot_sync(0xffffffffu, neg);
    const unsigned plain = __ballot_sync(0xffffffffu, exists && !neg);
    float lim = (exists && !neg && lane <= 3) ? __fmul_rd(tw, 0.99999905f) : inf;
    lim = fminf(lim, __shfl_xor_sync(0xffffffffu, lim, 1));
    lim = fminf(lim, __shfl_xor_sync(0xffffffffu, lim, 2));
    if (lane < n_windows) thr[lane] = tw;
    if (lane == 0) {
        thr[8] = lim;
        thr[9] = __uint_as_float((some_neg ? 1u : 0u) | (plain & 0x70u));
    }
}

// One staged span, processed by the T threads that staged it (thread <-> run of 32 channels).
struct TsTile {
    const float *rowbuf;   // the span, swizzled, T + 2 runs (two runs of zeros at the end)
    float4 *stat;          // T + 2 run statistics (entries T, T + 1 preset: -inf, 0, 0, 0)
    uint32_t *Fsm;         // T + 2 flag words (entries T, T + 1 preset to 0)
    uint32_t *car1, *car2; // T words each
    const float *thr;      // per-window thresholds and vote constants (ts_thresholds)
    int T, span, C, n_windows;   // T: staged runs; threads tid >= T only take part in the barriers
    int64_t pos0;          // row channel of this thread's element 0
};

// Window size 1, the filters and the block-wide vote; unless LEAN, the remaining window sizes
// where the vote says that one might fire.  F returns the thread's 32 flags (not yet masked to
// the band).  Returns true iff LEAN and the tile still needs the full algorithm (F is then
// incomplete).  Block-wide barriers inside: every thread of the block must call it.
template <bool LEAN>
__device__ __forceinline__ bool ts_process_tile(const TsTile &t, uint32_t &F)
{
    const int tid = threadIdx.x, span = t.span, C = t.C;
    const float *rowbuf = t.rowbuf;
    float4 *stat = t.stat;
    uint32_t *Fsm = t.Fsm, *car1 = t.car1, *car2 = t.car2;
    const float *thr = t.thr;
    const int64_t pos0 = t.pos0;
    const float neg_inf = -__int_as_float(0x7f800000);
    // ---- my run: window size 1, then running sums of what is left
    const int my_sw = tid & 7;                                 // swizzle of my run
    const float *my = rowbuf + tid * PITCH;
    const uint32_t in_range = bit_range(-pos0, (int64_t) C - pos0);

    F = 0;
    const bool active = tid < t.T;
    float m8[4] = {neg_inf, neg_inf, neg_inf, neg_inf};   // maxima of u over the four groups of 8 samples
    float ppos = 0.0f, sabs = 0.0f;                       // sum of the positive u, sum of |u|
    // (re)build u = F ? 0 : x, its running sums and the run statistics; publish them
    auto rebuild = [&](bool first) {
        if (!active) return;
        float x[RUN];
#pragma unroll
        for (int i = 0; i < RUN / 4; i++) {
            const float4 v = *reinterpret_cast<const float4 *>(my + ((i ^ my_sw) << 2));
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        if (first) {
            // window size 1: flag and zero in one go (samples outside the band are zeros and
            // stay zeros whether or not their bit survives the mask)
            // (three instructions per sample: compare, predicated OR of the flag bit, predicated zero)
            const float t0 = thr[0];
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                asm("{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.gt.f32 p, %1, %2;\n\t"
                    "@p or.b32 %0, %0, %3;\n\t"
                    "@p mov.f32 %1, 0f00000000;\n\t"
                    "}"
                    : "+r"(F), "+f"(x[j])
                    : "f"(t0), "r"(1u << j));
            }
            F &= in_range;
        } else {
#pragma unroll
            for (int j = 0; j < RUN; j++) x[j] = ((F >> j) & 1u) ? 0.0f : x[j];
        }
        sabs = 0.0f;
        float total = 0.0f;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            float m = neg_inf;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = 8 * g + k;
                const float u = x[j];
                m = fmaxf(m, u);
                sabs += fabsf(u);
                total += u;
            }
            m8[g] = m;
        }
        // >= sum of max(u, 0): the two serial sums are each off by at most 31 * 2^-24 * sabs, so
        // the slack has to scale with sabs (deep negative dips), not with the result alone
        ppos = fmaf(sabs, 4e-6f, 0.5f * (total + sabs) * 1.00002f);
        Fsm[tid] = F;
        stat[tid] = make_float4(m8[0], ppos, sabs, 0.0f);
    };
    rebuild(true);
    __syncthreads();

    // runs whose windows of every size lie inside the band need no masks
    const bool interior = (pos0 >= 0) && (pos0 + RUN + 64 <= (int64_t) C);
    // neighbours' statistics and flags: reloaded only after a rebuild
    float4 st1 = stat[tid + 1], st2 = stat[tid + 2];
    uint32_t F1 = Fsm[tid + 1], F2 = Fsm[tid + 2];

    // window starts of size 2^w in my run that the cheap tests cannot rule out
    auto hot_starts = [&](int w, float tw) -> uint32_t {
        const int win = 1 << w;
        const bool two = win > RUN;                            // reach covers two more runs
        // windows that start in my run and lie inside the band
        const uint32_t valid = interior ? FULL
                                        : bit_range(-pos0, (int64_t) C - (int64_t) win - pos0 + 1);
        uint32_t hot = FULL;
        if (tw >= 0.0f) {
            if (win <= 8) {
                // a window of unflagged samples that are all <= thr_w cannot fire
                const float lim = __fmul_rd(tw, 0.99999905f);
                hot = 0u;
                hot |= !(fmaxf(m8[0], m8[1]) <= lim) ? 0x000000ffu : 0u;
                hot |= !(fmaxf(m8[1], m8[2]) <= lim) ? 0x0000ff00u : 0u;
                hot |= !(fmaxf(m8[2], m8[3]) <= lim) ? 0x00ff0000u : 0u;
                hot |= !(fmaxf(m8[3], st1.x) <= lim) ? 0xff000000u : 0u;
            } else {
                // no window sum exceeds the sum of the positive samples within reach
                const float bound = ppos + st1.y + (two ? st2.y : 0.0f);
                const int nf = __popc(F) + __popc(F1) + (two ? __popc(F2) : 0);
                const float t_min = tw * (float) max(win - nf, 0);
                if (bound <= __fmul_rd(t_min, 0.99999f)) hot = 0u;
            }
        }
        return hot & valid;
    };

    // Almost always no thread of the block has anything left to look at: one vote then
    // replaces the whole window-size loop (and its barrier per size).  The vote uses a cheaper,
    // slightly more conservative form of hot_starts: sizes 2..8 share the strictest of their
    // limits (one comparison against the largest sample within reach), sizes 16..64 keep the
    // positive-sum bound, and band edges are ignored (a false "maybe" only costs the loop).
    {
        const float lim_small = thr[8];                        // strictest limit of the sizes <= 8
        const uint32_t kinds = __float_as_uint(thr[9]);
        bool any = (kinds & 1u) != 0u;                         // a negative threshold: everything is hot
        const int nf1 = __popc(F) + __popc(F1), nf2 = nf1 + __popc(F2);
        const float bound1 = ppos + st1.y, bound2 = bound1 + st2.y;
        if (kinds & 0x10u) any |= !(bound1 <= __fmul_rd(thr[4] * (float) max(16 - nf1, 0), 0.99999f));
        if (kinds & 0x20u) any |= !(bound1 <= __fmul_rd(thr[5] * (float) max(32 - nf1, 0), 0.99999f));
        if (kinds & 0x40u) any |= !(bound2 <= __fmul_rd(thr[6] * (float) max(64 - nf2, 0), 0.99999f));
        const float reach_max = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), st1.x);
        any |= !(reach_max <= lim_small);
        if (!__syncthreads_or(any && active)) return false;
    }
    if (!LEAN)
    for (int w = 1; w < t.n_windows; w++) {
        const int win = 1 << w;
        if (win > C) break;
        const float tw = thr[w];
        if (tw != tw) continue;                                // NaN threshold: nothing can fire
        uint32_t fire = 0;
        {
            const bool two = win > RUN;
            uint32_t hot = active ? hot_starts(w, tw) : 0u;
            if (hot != 0u) {
                const float err = FILTER_ERR * ((sabs + st1.z) + (two ? st2.z : 0.0f));
                const uint32_t G2 = two ? F2 : 0u;
                uint32_t cand;
                switch (w) {
                case 1: cand = window_candidates<2>(rowbuf, tid, F, F1, G2, tw, err); break;
                case 2: cand = window_candidates<4>(rowbuf, tid, F, F1, G2, tw, err); break;
                case 3: cand = window_candidates<8>(rowbuf, tid, F, F1, G2, tw, err); break;
                case 4: cand = window_candidates<16>(rowbuf, tid, F, F1, G2, tw, err); break;
                case 5: cand = window_candidates<32>(rowbuf, tid, F, F1, G2, tw, err); break;
                default: cand = window_candidates<64>(rowbuf, tid, F, F1, G2, tw, err); break;
                }
                cand &= hot;
                if (cand != 0u) fire = exact_windows(cand, tid, w, tw, rowbuf, Fsm, span);
            }
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
            const uint32_t Fnew = F | lo | (tid >= 1 ? car1[tid - 1] : 0u) | (tid >= 2 ? car2[tid - 2] : 0u);
            if (Fnew != F) {
                F = Fnew;
                rebuild(false);
            }
            __syncthreads();
            st1 = stat[tid + 1];
            st2 = stat[tid + 2];
            F1 = Fsm[tid + 1];
            F2 = Fsm[tid + 2];
        }
    }

    return LEAN;
}

// ---------------------------------------------------------------- packed flags -> bytes
// bits_t[b * wstride + c/32] (bit c%32) -> flags[c * fstride + b] = bit ? flag_value : 0.
// Tile: 128 baselines x 8 words (256 channels).  Replaces the uchar transpose of
// rfi/device.py:1161-1164 in the fused flagger.
// One tile by 256 threads: baselines b0 .. b0 + 127, words w0 .. w0 + 7; the bits of baseline b are
// in row b + row_off of bits_t (ring buffers; COHERENT: read them with ld.global.cg).
constexpr int EXPAND_SMEM_WORDS = 8 * 132;
template <bool COHERENT>
__device__ __forceinline__ void expand_flags_tile(const uint32_t *__restrict__ bits_t, uint8_t *__restrict__ flags,
                                                  int64_t channels, int64_t baselines, int64_t wstride,
                                                  int64_t fstride, int flag_value, const int64_t b0,
                                                  const int64_t w0, const int64_t row_off,
                                                  uint32_t (*tile)[132])
{
    const int64_t n_words = (channels + 31) >> 5;
    const int t = threadIdx.x;
    {
        const int w = t & 7;
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
            const int b = (t >> 3) + 32 * pass;
            uint32_t v = 0;
            if (b0 + b < baselines && w0 + w < n_words) {
                const uint32_t *p = bits_t + (b0 + b + row_off) * wstride + w0 + w;
                v = COHERENT ? __ldcg(p) : __ldg(p);
            }
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
    if (vec && c_base + 32 <= channels) {
        // 4 baselines x 4 channels at a time: nibble -> 4 bytes (one multiply), then a 4 x 4
        // byte transpose with byte permutes gives, per channel, the 4 baselines' flag bytes
        uint8_t *out = flags + c_base * fstride + b;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const uint32_t r0 = (((q.x >> (4 * g)) & 0xfu) * 0x00204081u & 0x01010101u) * fv;
            const uint32_t r1 = (((q.y >> (4 * g)) & 0xfu) * 0x00204081u & 0x01010101u) * fv;
            const uint32_t r2 = (((q.z >> (4 * g)) & 0xfu) * 0x00204081u & 0x01010101u) * fv;
            const uint32_t r3 = (((q.w >> (4 * g)) & 0xfu) * 0x00204081u & 0x01010101u) * fv;
            // r_i: byte k = channel 4g+k of baseline i.  Want t_k: byte i = baseline i of channel 4g+k.
            const uint32_t a01 = __byte_perm(r0, r1, 0x5140);   // r0.b0 r1.b0 r0.b1 r1.b1
            const uint32_t b01 = __byte_perm(r0, r1, 0x7362);   // r0.b2 r1.b2 r0.b3 r1.b3
            const uint32_t a23 = __byte_perm(r2, r3, 0x5140);
            const uint32_t b23 = __byte_perm(r2, r3, 0x7362);
            const uint32_t t0 = __byte_perm(a01, a23, 0x5410);  // channel 4g
            const uint32_t t1 = __byte_perm(a01, a23, 0x7632);  // channel 4g + 1
            const uint32_t t2 = __byte_perm(b01, b23, 0x5410);  // channel 4g + 2
            const uint32_t t3 = __byte_perm(b01, b23, 0x7632);  // channel 4g + 3
            uint8_t *o = out + (int64_t) (4 * g) * fstride;
            // write-once output: first to leave the L2
            stg_stream_u32(reinterpret_cast<uint32_t *>(o), t0);
            stg_stream_u32(reinterpret_cast<uint32_t *>(o + fstride), t1);
            stg_stream_u32(reinterpret_cast<uint32_t *>(o + 2 * fstride), t2);
            stg_stream_u32(reinterpret_cast<uint32_t *>(o + 3 * fstride), t3);
        }
        return;
    }
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


}  // namespace
