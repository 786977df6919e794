// Width-13 sliding-median background filter: the per-tile device code shared by
// bg13_kernel (background.cu) and the dataflow flagger (dataflow.cu).  See background.cu for
// the reference citations and the semantics.
#pragma once
#include "common.cuh"
#include "median13.cuh"

namespace {

using namespace ksp;

constexpr int IN_NUMPY = 0;   // complex64, numpy AVX-512 amplitude rule
constexpr int IN_HYPOT = 1;   // complex64, correctly rounded hypot
constexpr int IN_AMP = 2;     // float32 amplitudes

constexpr int BG_THREADS = 256;
#ifndef BG_TC_VALUE
#define BG_TC_VALUE 256
#endif
constexpr int BG_TC = BG_TC_VALUE;          // channels per tile of the width-13 kernel
#ifndef BG_MIN_BLOCKS
#define BG_MIN_BLOCKS 4
#endif

struct BgArgs {
    const void *vis;
    float *out;
    const uint8_t *flags;
    int64_t channels, baselines;
    int64_t vis_stride, out_stride, flags_stride;
    int seg;  // channels per segment (multiple of 32)
};

template <int IN_MODE>
__device__ __forceinline__ float load_amp(const void *vis, int64_t idx)
{
    if (IN_MODE == IN_AMP) {
        return ldg_stream_f(reinterpret_cast<const float *>(vis) + idx);
    } else {
        float2 v = ldg_stream_f2(reinterpret_cast<const float2 *>(vis) + idx);
        return abs_c64<IN_MODE == IN_NUMPY ? KSP_ABS_NUMPY : KSP_ABS_HYPOT>(v.x, v.y);
    }
}

// Masked medians for the 4 outputs of one step.  e is in logical order.
__device__ __noinline__ float4 slow_step(const float *e, unsigned bad)
{
    float r[4];
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        float out = 0.0f;
        if (!((bad >> (j + 6)) & 1u)) {
            float w[13];
#pragma unroll
            for (int k = 0; k < 13; k++) w[k] = e[j + k];
            float lo, hi;
            unsigned valid = ~(bad >> j) & 0x1fffu;
            if (median_masked13(w, valid, lo, hi)) {
                double med = (lo == hi) ? (double) lo : ((double) lo + (double) hi) * 0.5;
                out = __double2float_rn((double) e[j + 6] - med);
            }
        }
        r[j] = out;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

// ---------------------------------------------------------------- width 13, tiled
// One block = one tile of 32 baselines x TC channels.
//
// Phase 1 (lane <-> baseline): every warp reads groups of 4 consecutive channel rows, 256
// coalesced bytes of complex64 per row, turns them into amplitudes and stores them
// baseline-major in shared memory (row pitch P floats, P/4 odd: the 128-bit stores of
// a warp, one row apart, fall into distinct bank groups).  Unusable samples (outside
// the band, flagged, NaN amplitude) are stored as NaN.  The tile carries 8 halo
// channels on the left and 12 on the right so that groups stay 16-byte aligned.
// Phase 2 (lane <-> run of 4 channels): every thread reads the 20 amplitudes around its 4
// outputs with 5 conflict-free 128-bit loads, runs the shared selection network
// (median13.cuh) and stores 4 deviations -- one 128-bit store per thread, 512 contiguous
// bytes per warp, in the baseline-major (_t) variant; 4 row-strided coalesced stores in
// the channel-major one.  If the tile holds any unusable sample the threads test their 16
// window samples first and take the masked slow path where needed.
constexpr int TILE_B = 32;
constexpr int HALO_L = 8;            // smem index s = c - c0 + HALO_L
constexpr int HALO_R = 12;

template <int TC>
struct TileGeom {
    static constexpr int P = TC + HALO_L + HALO_R;      // floats per smem row
    static_assert(((P / 4) & 1) == 1, "P/4 must be odd for conflict-free 128-bit row-strided access");
    static constexpr int SMEM_BYTES = TILE_B * P * 4;
};

// Phase-1 work of one thread: U groups of 4 consecutive channels (the groups are GS channels
// apart) of one baseline -> amplitudes (NaN where unusable) -> one 128-bit shared store per
// group.  `p` addresses the sample (channel c, this thread's baseline), `fp` its flag;
// row_bytes / flag_row are the distances to the next channel.  INTERIOR tiles skip every
// bounds test.  Returns true if some sample is unusable.
template <int IN_MODE, int FLAG_MODE, bool INTERIOR, int U, int GS = 32>
__device__ __forceinline__ bool amplitudes_to_smem(const char *p, int64_t row_bytes,
                                                   const uint8_t *fp, int64_t flag_row, int c,
                                                   int C, float *dst)
{
    const float nan = __int_as_float(0x7fc00000);
    float v[U][4];
    unsigned usable = 0;                                 // bit 4u+k: sample may be used
    float2 raw[U][4];
    // all loads first (predicated at most, no control flow in between), then the arithmetic
#pragma unroll
    for (int u = 0; u < U; u++) {
        const char *q = p + (int64_t) (GS * u) * row_bytes;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ck = c + GS * u + k;
            const bool in = INTERIOR || (ck >= 0 && ck < C);
            raw[u][k] = make_float2(1.0f, 0.0f);
            if (IN_MODE == IN_AMP) {
                if (in) raw[u][k].x = ldg_stream_f(reinterpret_cast<const float *>(q));
            } else {
                if (in) raw[u][k] = ldg_stream_f2(reinterpret_cast<const float2 *>(q));
            }
            usable |= (in ? 1u : 0u) << (4 * u + k);
            q += row_bytes;
        }
    }
    if (FLAG_MODE != KSP_FLAGS_NONE) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint8_t *fq = fp + (int64_t) (GS * u) * flag_row;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if ((usable >> (4 * u + k)) & 1u) {
                    if (*fq) usable &= ~(1u << (4 * u + k));
                }
                fq += flag_row;
            }
        }
    }
    if (IN_MODE == IN_AMP) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < 4; k++) v[u][k] = raw[u][k].x;
    } else if (IN_MODE == IN_NUMPY) {
        unsigned redo = 0;
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                bool ok;
                v[u][k] = abs_numpy_try(raw[u][k].x, raw[u][k].y, ok);
                redo |= (ok ? 0u : 1u) << (4 * u + k);
            }
        if (redo) {                                      // zeros, denormals, huge, inf, NaN
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((redo >> (4 * u + k)) & 1u)
                        v[u][k] = abs_slow_call(raw[u][k].x, raw[u][k].y, KSP_ABS_NUMPY);
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < 4; k++)
                v[u][k] = abs_slow_call(raw[u][k].x, raw[u][k].y, KSP_ABS_HYPOT);
    }
    bool any_bad = false;
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (!(INTERIOR && FLAG_MODE == KSP_FLAGS_NONE)) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[u][k] = ((usable >> (4 * u + k)) & 1u) ? v[u][k] : nan;
        }
        const float probe = (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]);
        any_bad |= (probe != probe);                     // NaN iff some NaN (or inf - inf)
        *reinterpret_cast<float4 *>(dst + GS * u) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
    }
    return any_bad;
}

// NW warps (numbered `warp`) share the tile's groups: warp w takes groups w, w + NW, ...
template <int IN_MODE, int FLAG_MODE, bool INTERIOR, int TC, int NW = BG_THREADS / 32>
__device__ __forceinline__ bool tile_phase1(const BgArgs &a, int64_t b0, int c0, float *amp_sm,
                                            const int warp, const int lane)
{
    using G = TileGeom<TC>;
    constexpr int GS = 4 * NW;                           // channels covered by one group per warp
    constexpr int NEEDED = (TC + HALO_L + 6 + 3) / 4;    // groups that phase 2 reads
    constexpr int FULL_ITERS = NEEDED / (2 * NW);        // 2 NW groups per iteration (U = 2)
    constexpr int REST = NEEDED - 2 * NW * FULL_ITERS;   // < 2 NW, handled with U = 1
    const int C = (int) a.channels;
    const int64_t b = min(b0 + lane, a.baselines - 1);   // clamp: duplicates are never stored
    constexpr int ESZ = (IN_MODE == IN_AMP) ? 4 : 8;
    const int64_t row_bytes = a.vis_stride * ESZ;
    int c = c0 - HALO_L + 4 * warp;                       // first channel of this warp's group
    const char *p = reinterpret_cast<const char *>(a.vis) + ((int64_t) c * a.vis_stride + b) * ESZ;
    const uint8_t *fp = nullptr;
    int64_t flag_row = 0;
    if (FLAG_MODE == KSP_FLAGS_CHANNEL) { fp = a.flags + c; flag_row = 1; }
    if (FLAG_MODE == KSP_FLAGS_FULL) { fp = a.flags + (int64_t) c * a.flags_stride + b; flag_row = a.flags_stride; }
    float *dst = amp_sm + lane * G::P + 4 * warp;
    bool any_bad = false;
#pragma unroll 1
    for (int i = 0; i < FULL_ITERS; i++) {
        any_bad |= amplitudes_to_smem<IN_MODE, FLAG_MODE, INTERIOR, 2, GS>(p, row_bytes, fp, flag_row,
                                                                           c, C, dst);
        p += 2 * GS * row_bytes;
        fp += 2 * GS * flag_row;
        c += 2 * GS;
        dst += 2 * GS;
    }
    if (REST > NW) {
        if (warp < REST - NW)
            any_bad |= amplitudes_to_smem<IN_MODE, FLAG_MODE, INTERIOR, 2, GS>(p, row_bytes, fp,
                                                                               flag_row, c, C, dst);
        else
            any_bad |= amplitudes_to_smem<IN_MODE, FLAG_MODE, INTERIOR, 1, GS>(p, row_bytes, fp,
                                                                               flag_row, c, C, dst);
    } else if (REST > 0) {
        if (warp < REST)
            any_bad |= amplitudes_to_smem<IN_MODE, FLAG_MODE, INTERIOR, 1, GS>(p, row_bytes, fp,
                                                                               flag_row, c, C, dst);
    }
    return any_bad;
}

// Interior tiles without flags, complex input: the same work with the loads software-pipelined
// BG_PF groups ahead in registers, so that every thread keeps loads in flight while it does the
// amplitude arithmetic of an earlier group (phase 1 is otherwise latency-bound: issue 8 loads,
// wait, compute, repeat).
#ifndef BG_PF
#define BG_PF 3
#endif
#ifndef BG_X8
#define BG_X8 1                      // phase 2 computes 8 medians per step (median13x8)
#endif
template <int IN_MODE, int TC, int PF, int NW = BG_THREADS / 32>
__device__ __forceinline__ bool tile_phase1_pipelined(const BgArgs &a, int64_t b0, int c0, float *amp_sm,
                                                      const int warp, const int lane)
{
    using G = TileGeom<TC>;
    constexpr int NWARPS = NW;
    constexpr int NEEDED = (TC + HALO_L + 6 + 3) / 4;    // groups that phase 2 reads
    constexpr int MAX_IT = (NEEDED + NWARPS - 1) / NWARPS;
    const int64_t b = min(b0 + lane, a.baselines - 1);
    const uint32_t row_bytes = (uint32_t) a.vis_stride * 8u;         // (the caller checks that it fits)
    const int n_it = (NEEDED - warp + NWARPS - 1) / NWARPS;          // groups warp, warp + 8, ...
    const char *p = reinterpret_cast<const char *>(a.vis) +
                    ((int64_t) (c0 - HALO_L + 4 * warp) * a.vis_stride + b) * 8;
    float *dst = amp_sm + lane * G::P + 4 * warp;
    float2 buf[PF + 1][4];
    // row `row` of this thread's column: one multiply-add on the FMA pipe per address
    // (IMAD.WIDE with an immediate row number) instead of a chain of 64-bit additions on the
    // ALU pipe, which the selection network of phase 2 keeps busy
    auto load = [&](float2 (&r)[4], int row) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t q;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(q) : "r"(row_bytes), "r"((uint32_t) (row + k)),
                "l"(reinterpret_cast<uint64_t>(p)));
            r[k] = ldg_stream_f2(reinterpret_cast<const float2 *>(q));
        }
    };
#pragma unroll
    for (int d = 0; d < PF; d++) {
        if (d < n_it) load(buf[d], 4 * NWARPS * d);
    }
    bool any_bad = false;
#pragma unroll
    for (int it = 0; it < MAX_IT; it++) {
        if (it < n_it) {
            if (it + PF < n_it) load(buf[(it + PF) % (PF + 1)], 4 * NWARPS * (it + PF));
            const float2 (&r)[4] = buf[it % (PF + 1)];
            float v[4];
            unsigned redo = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                bool ok;
                v[k] = abs_numpy_try(r[k].x, r[k].y, ok);
                redo |= (ok ? 0u : 1u) << k;
            }
            if (redo) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((redo >> k) & 1u)
                        v[k] = abs_slow_call(r[k].x, r[k].y, IN_MODE == IN_NUMPY ? KSP_ABS_NUMPY : KSP_ABS_HYPOT);
            }
            const float probe = (v[0] + v[1]) + (v[2] + v[3]);
            any_bad |= (probe != probe);
            *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            dst += 4 * NWARPS;
        }
    }
    return any_bad;
}

// One tile: baselines b0 .. b0 + 31, channels c0 .. c0 + TC - 1; amp_sm = TileGeom<TC>::SMEM_BYTES of
// shared memory.  Every thread of the 256-thread block must call it; block-wide barriers inside.
// `at` says where the tile is: at.b0(), at.c0(), and at.row_off(), the distance from baseline b to
// its row of a baseline-major output (0, or a ring-buffer offset).  It is asked again in each
// phase, so that a caller whose coordinates live in shared memory (dataflow.cu) holds no
// register for them while the other phase runs - this kernel has none to spare.
struct BlockTile {               // grid (strips, channel tiles)
    template <int TC> __device__ __forceinline__ int c0() const { return (int) blockIdx.y * TC; }
    __device__ __forceinline__ int64_t b0() const { return (int64_t) blockIdx.x * 32; }
    __device__ __forceinline__ int64_t row_off() const { return 0; }
    static constexpr bool KEEP_IN_L2 = false;
};

// Phase 1 of a tile by NW warps (this thread: warp `warp`, lane `lane`): amplitudes into amp_sm.
// Returns true if this thread stored an unusable sample (the tile then needs the checked phase 2).
template <int IN_MODE, int FLAG_MODE, int TC, typename Where, int PF, int NW>
__device__ __forceinline__ bool bg13_phase1(const BgArgs &a, const Where &at, float *amp_sm,
                                            const int warp, const int lane)
{
    const int C = (int) a.channels;
    const int64_t b0 = at.b0();
    const int c0 = at.template c0<TC>();
    const bool interior = (c0 - HALO_L >= 0) && (c0 + TC + HALO_R <= C);   // block-uniform
    if (interior && IN_MODE == IN_NUMPY && FLAG_MODE == KSP_FLAGS_NONE && PF > 0 &&
        a.vis_stride < ((int64_t) 1 << 28))
        return tile_phase1_pipelined<IN_MODE, TC, PF, NW>(a, b0, c0, amp_sm, warp, lane);
    if (interior)
        return tile_phase1<IN_MODE, FLAG_MODE, true, TC, NW>(a, b0, c0, amp_sm, warp, lane);
    return tile_phase1<IN_MODE, FLAG_MODE, false, TC, NW>(a, b0, c0, amp_sm, warp, lane);
}

// Phase 2 of a tile by BG_THREADS threads (this one: `tid`): medians, deviations, stores.
template <int IN_MODE, int FLAG_MODE, bool TRANSPOSED, int TC, typename Where>
__device__ __forceinline__ void bg13_phase2(const BgArgs &a, const Where &at, const float *amp_sm,
                                            const int tile_bad, const int tid)
{
    using G = TileGeom<TC>;
    const int C = (int) a.channels;
    const int64_t b0 = at.b0();
    const int c0 = at.template c0<TC>();
    const int64_t row_off = at.row_off();

#if BG_X8
    // ---- phase 2, fast: no unusable sample anywhere in the tile.  Every thread produces runs of
    // 8 outputs: the 6 samples common to all 8 windows are sorted once (median13x8).
    if (!tile_bad) {
        constexpr int RUNS8 = TC / 8;
        // whole tile inside the array and 16-byte aligned rows: no test per store (block-uniform)
        const bool whole = TRANSPOSED && (b0 + TILE_B <= a.baselines) && (c0 + TC <= C) &&
                           ((a.out_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
        if (TRANSPOSED && whole && (TILE_B * RUNS8) % BG_THREADS == 0 && BG_THREADS % (2 * RUNS8) == 0) {
            // Whole tile, baseline-major output: the thread's runs are BG_THREADS / RUNS8 baseline
            // rows apart with the same channel offset, so everything that depends on the thread is
            // computed once and each further run costs two additions (this part of the kernel is
            // bound by the ALU pipe, which the selection network needs for itself).
            constexpr int STEP_B = BG_THREADS / RUNS8;            // baseline rows between two runs of a thread
            const int t = tid;
            const int bl = 2 * (t / (2 * RUNS8)) + (t & 1);       // (see the general loop below)
            const int j = (t % (2 * RUNS8)) >> 1;
            const float *src = amp_sm + bl * G::P + 8 * j;
            float *o = a.out + (b0 + bl + row_off) * a.out_stride + c0 + 8 * j;
            const int64_t o_step = (int64_t) STEP_B * a.out_stride;
#pragma unroll
            for (int k = 0; k < (TILE_B * RUNS8) / BG_THREADS; k++) {
                float w[24];
#pragma unroll
                for (int q4 = 0; q4 < 6; q4++) {
                    float4 q = *reinterpret_cast<const float4 *>(src + k * STEP_B * G::P + 4 * q4);
                    w[4 * q4] = q.x; w[4 * q4 + 1] = q.y; w[4 * q4 + 2] = q.z; w[4 * q4 + 3] = q.w;
                }
                float e[20];
#pragma unroll
                for (int i = 0; i < 20; i++) e[i] = w[i + 2];    // channels c-6 .. c+13
                float m[8], o8[8];
                median13x8(e, m);
#pragma unroll
                for (int i = 0; i < 8; i++) o8[i] = e[6 + i] - m[i];
                if (Where::KEEP_IN_L2) {
                    stg_keep_f4(o, make_float4(o8[0], o8[1], o8[2], o8[3]));
                    stg_keep_f4(o + 4, make_float4(o8[4], o8[5], o8[6], o8[7]));
                } else {
                    reinterpret_cast<float4 *>(o)[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
                    reinterpret_cast<float4 *>(o)[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
                }
                o += o_step;
            }
            return;
        }
        for (int t = tid; t < TILE_B * RUNS8; t += BG_THREADS) {
            int bl, j;
            if (TRANSPOSED) {
                // lanes walk along channels, alternating between two baseline rows: 8 consecutive
                // lanes then read 128-bit words from 8 different bank groups (rows are P = 276
                // floats apart, 20 banks), and each row still gets 512 contiguous bytes per warp
                bl = 2 * (t / (2 * RUNS8)) + (t & 1);
                j = (t % (2 * RUNS8)) >> 1;
            } else {                                              // lanes walk along baselines
                j = t / TILE_B;
                bl = t % TILE_B;
            }
            const float *src = amp_sm + bl * G::P + 8 * j;        // channel c - 8
            float w[24];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                float4 q = *reinterpret_cast<const float4 *>(src + 4 * k);
                w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
            }
            float e[20];
#pragma unroll
            for (int k = 0; k < 20; k++) e[k] = w[k + 2];        // channels c-6 .. c+13
            float m[8], o8[8];
            median13x8(e, m);
#pragma unroll
            for (int k = 0; k < 8; k++) o8[k] = e[6 + k] - m[k];
            const int c = c0 + 8 * j;
            const int64_t b = b0 + bl;
            if (!whole && (b >= a.baselines || c >= C)) continue;
            if (TRANSPOSED) {
                float *o = a.out + (b + row_off) * a.out_stride + c;
                if (whole || (c + 7 < C && ((reinterpret_cast<uintptr_t>(o) & 15) == 0))) {
                    if (Where::KEEP_IN_L2) {
                        stg_keep_f4(o, make_float4(o8[0], o8[1], o8[2], o8[3]));
                        stg_keep_f4(o + 4, make_float4(o8[4], o8[5], o8[6], o8[7]));
                    } else {
                        reinterpret_cast<float4 *>(o)[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
                        reinterpret_cast<float4 *>(o)[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (c + k < C) o[k] = o8[k];
                }
            } else {
                float *o = a.out + c * a.out_stride + b;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (c + k < C) o[k * a.out_stride] = o8[k];
            }
        }
        return;
    }
#endif
    // ---- phase 2, fast: no unusable sample anywhere in the tile
    if (!tile_bad) {
        constexpr int RUNS = TC / 4;                 // runs of 4 outputs per baseline row
        for (int t = tid; t < TILE_B * RUNS; t += BG_THREADS) {
            int bl, j;
            if (TRANSPOSED) { bl = t / RUNS; j = t % RUNS; }      // lanes walk along channels
            else            { j = t / TILE_B; bl = t % TILE_B; }  // lanes walk along baselines
            const float *src = amp_sm + bl * G::P + 4 * j;
            float w[20];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                float4 q = *reinterpret_cast<const float4 *>(src + 4 * k);
                w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
            }
            float e[16];
#pragma unroll
            for (int k = 0; k < 16; k++) e[k] = w[k + 2];        // channels c-6 .. c+9
            float m0, m1, m2, m3;
            median13x4(e, 0, m0, m1, m2, m3);
            const float o0 = e[6] - m0, o1 = e[7] - m1, o2 = e[8] - m2, o3 = e[9] - m3;
            const int c = c0 + 4 * j;
            const int64_t b = b0 + bl;
            if (b >= a.baselines || c >= C) continue;
            if (TRANSPOSED) {
                float *o = a.out + (b + row_off) * a.out_stride + c;
                if (c + 3 < C && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    *reinterpret_cast<float4 *>(o) = make_float4(o0, o1, o2, o3);
                } else {
                    o[0] = o0;
                    if (c + 1 < C) o[1] = o1;
                    if (c + 2 < C) o[2] = o2;
                    if (c + 3 < C) o[3] = o3;
                }
            } else {
                float *o = a.out + c * a.out_stride + b;
                o[0] = o0;
                if (c + 1 < C) o[a.out_stride] = o1;
                if (c + 2 < C) o[2 * a.out_stride] = o2;
                if (c + 3 < C) o[3 * a.out_stride] = o3;
            }
        }
        return;
    }

    // ---- phase 2, checked: per-thread test of the 16 window samples
    {
        constexpr int RUNS = TC / 4;
        for (int t = tid; t < TILE_B * RUNS; t += BG_THREADS) {
            int bl, j;
            if (TRANSPOSED) { bl = t / RUNS; j = t % RUNS; }
            else            { j = t / TILE_B; bl = t % TILE_B; }
            const int c = c0 + 4 * j;
            const int64_t b = b0 + bl;
            if (b >= a.baselines || c >= C) continue;
            const float *src = amp_sm + bl * G::P + 4 * j;
            float e[16];
            unsigned bad = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                e[k] = src[k + 2];
                bad |= (e[k] != e[k] ? 1u : 0u) << k;
            }
            float o0, o1, o2, o3;
            if (bad == 0) {
                float m0, m1, m2, m3;
                median13x4(e, 0, m0, m1, m2, m3);
                o0 = e[6] - m0; o1 = e[7] - m1; o2 = e[8] - m2; o3 = e[9] - m3;
            } else {
                float4 r = slow_step(e, bad);
                o0 = r.x; o1 = r.y; o2 = r.z; o3 = r.w;
            }
            if (TRANSPOSED) {
                float *o = a.out + (b + row_off) * a.out_stride + c;
                o[0] = o0;
                if (c + 1 < C) o[1] = o1;
                if (c + 2 < C) o[2] = o2;
                if (c + 3 < C) o[3] = o3;
            } else {
                float *o = a.out + c * a.out_stride + b;
                o[0] = o0;
                if (c + 1 < C) o[a.out_stride] = o1;
                if (c + 2 < C) o[2 * a.out_stride] = o2;
                if (c + 3 < C) o[3 * a.out_stride] = o3;
            }
        }
    }
}

// One tile by one block of BG_THREADS threads: phase 1, a block-wide vote, phase 2.
template <int IN_MODE, int FLAG_MODE, bool TRANSPOSED, int TC, typename Where, int PF = BG_PF>
__device__ __forceinline__ void bg13_tile(const BgArgs &a, const Where &at, float *amp_sm)
{
    const bool any_bad = bg13_phase1<IN_MODE, FLAG_MODE, TC, Where, PF, BG_THREADS / 32>(
        a, at, amp_sm, (int) (threadIdx.x >> 5), (int) (threadIdx.x & 31));
    const int tile_bad = __syncthreads_or(any_bad);       // does the tile need the checked path?
    bg13_phase2<IN_MODE, FLAG_MODE, TRANSPOSED, TC, Where>(a, at, amp_sm, tile_bad, (int) threadIdx.x);
}

}  // namespace
