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
// the right-hand side exactly.
//
// Almost no window ever fires, so the kernel does not build the trees.  It FILTERS with
// cheap window sums and evaluates the tree only for the survivors:
//
//   * mapping: one thread per run of 32 consecutive channels, one tile = a span of T * 32
//     channels of one baseline (spans overlap by the reach of the largest window, 128
//     channels, so tiles are independent); a persistent grid (4 blocks of 128 threads per SM)
//     walks the tiles;
//   * staging: ONE TMA tile load per tile (cp.async.bulk.tensor over dev_t viewed as
//     [baseline][run][32 floats], 128-byte swizzle, hardware zero fill outside the band,
//     completion on an mbarrier) puts the span in shared memory, double-buffered: the load of
//     a block's next tile is issued before it starts on the current one, so the memory system
//     always has 2 x 16 KB per block in flight (the kernel was latency-bound without it);
//     plain vector loads into the same layout when the shape does not allow a tensor map;
//   * every thread publishes its 32 flags and three statistics of its run (maximum of the
//     first 8 samples, sum of the positive samples, sum of |u|) in shared memory;
//   * window size 2^w, thread by thread, cheapest test first:
//       - sizes 2..8, per group of 8 window starts: no sample of the 16 the group can reach
//         exceeds thr_w                                  -> none of those windows can fire;
//       - sizes 16..64: the positive samples within reach add up to less than
//         thr_w * (2^w - #F within reach)                -> none of the thread's windows can;
//       - else S~[j] = differences of running sums p[k] = u[0] + .. + u[k-1] of the own and
//         the next one or two runs; |S~ - D_w| <= E with E = 2.5e-4 * sum|u| within reach
//         (3 serial sums of 32 terms, the tree's own 6 roundings, with slack), so
//         S~[j] <= thr_w * (2^w - #F) - E rules the window out;
//       - survivors (real interference, or decisions within ~1e-4 of the threshold) are
//         evaluated exactly in tree order from the staged row (exact_windows);
//   * if a window size flags anything the block dilates the hits over their windows,
//     rebuilds u and the running sums, and goes on.
// Non-finite samples make the sums non-finite: every window within reach is then a survivor.
#include "common.cuh"
#include "tma.cuh"
#include <stdlib.h>
#include <string.h>

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

// TFIX: block size known at compile time (0 = use blockDim.x); lets the staging loop use
// immediate offsets.
#ifndef TS_MIN_BLOCKS
#define TS_MIN_BLOCKS 4
#endif
// MODE 0: everything in one kernel.  Two-pass mode (fused flagger): MODE 1 does window size 1 and
// the vote for every tile - the whole job for almost all of them - and lists the few tiles where
// some larger window might fire; MODE 2 then runs the full algorithm over the listed tiles.  The
// first pass carries none of the rare paths, needs half the registers and runs with more blocks
// per SM.
#ifndef TS_LEAN_BLOCKS
#define TS_LEAN_BLOCKS 7
#endif
template <bool PACKED, int TFIX, int MODE>
__global__ void __launch_bounds__(TFIX ? TFIX : TS_MAX_THREADS,
                                  MODE == 1 ? TS_LEAN_BLOCKS : (TFIX ? TS_MIN_BLOCKS : 4))
threshold_sum_kernel(const TsArgs a, const __grid_constant__ CUtensorMap tmap)
{
    if (MODE == 2 && blockIdx.x >= a.work[0]) return;           // nothing listed for this block
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    // the swizzle pattern is a function of the shared-memory address: align the span to 1 KB
    float *sm = reinterpret_cast<float *>(sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u));
    const int T = TFIX ? TFIX : (int) blockDim.x;
    const int tid = threadIdx.x;
    const int span = T * RUN;
    const float neg_inf = -__int_as_float(0x7f800000);

    // two span buffers of (T + 2) * PITCH floats (the second only with two_buffers), then the
    // per-run state
    // (each buffer starts on a 1 KB boundary: the swizzle is a function of the address)
    const int buf_floats = (((T + 2) * PITCH + 255) / 256) * 256;
    float *rowbuf0 = sm;
    float *rowbuf1 = rowbuf0 + (a.two_buffers ? buf_floats : 0);
    float4 *stat = reinterpret_cast<float4 *>(rowbuf1 + buf_floats);        // T + 2: run statistics
    uint32_t *Fsm = reinterpret_cast<uint32_t *>(stat + T + 2);             // T + 2
    uint32_t *car1 = Fsm + T + 2;                              // T
    uint32_t *car2 = car1 + T;                                 // T
    float *thr = reinterpret_cast<float *>(car2 + T);          // TS_MAX_WINDOWS (+1)
    uint64_t *mbar = reinterpret_cast<uint64_t *>(thr + 8);    // TMA completion barriers, one per buffer

    const int C = (int) a.channels;
    const int64_t total = a.baselines * (int64_t) a.n_chunks;  // tiles = (row, span) pairs

    // ---- once per block
    if (tid < 2) {
        Fsm[T + tid] = 0u;
        stat[T + tid] = make_float4(neg_inf, 0.0f, 0.0f, 0.0f);
    }
    for (int i = tid; i < 2 * PITCH; i += T) {                 // two runs of zeros past each span
        rowbuf0[T * PITCH + i] = 0.0f;
        rowbuf1[T * PITCH + i] = 0.0f;
    }
    if (a.use_tma && tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
    }
    __syncthreads();
    // one TMA tile load stages a span: T runs of 128 bytes, swizzled, out-of-band runs zero-filled
    // by the hardware, completion on the buffer's mbarrier
    auto issue_tile = [&](int64_t r, int y, int buf) {
        mbar_expect_tx(&mbar[buf], (uint32_t) span * 4u);
        tma_load_3d(buf ? rowbuf1 : rowbuf0, &tmap, 0, (y * a.chunk_valid - a.edge) >> 5, (int) r, &mbar[buf]);
    };
    // tiles of this block: MODE 2 takes them from the list of the first pass, the others walk all
    // (row, span) pairs without a division per tile, stepping both by the grid size
    const int64_t n_iter = (MODE == 2) ? (int64_t) a.work[0] : total;
    auto listed = [&](int64_t i, int64_t &r, int &y) {
        const uint32_t t = a.work[1 + i];
        r = t / (uint32_t) a.n_chunks;
        y = (int) (t - (uint32_t) r * (uint32_t) a.n_chunks);
    };
    int64_t row = (int64_t) blockIdx.x / a.n_chunks;
    int span_y = (int) ((int64_t) blockIdx.x - row * a.n_chunks);
    if (MODE == 2 && (int64_t) blockIdx.x < n_iter) listed(blockIdx.x, row, span_y);
    const int64_t step_rows = (int64_t) gridDim.x / a.n_chunks;
    const int step_y = (int) ((int64_t) gridDim.x - step_rows * a.n_chunks);
    // the noise of a tile's row is fetched one tile ahead, like its samples
    float noise_now = 0.0f;
    if (tid < a.n_windows && (int64_t) blockIdx.x < n_iter) noise_now = a.noise[row];
    if (a.use_tma && a.two_buffers && tid == 0 && (int64_t) blockIdx.x < n_iter) issue_tile(row, span_y, 0);

    // ---- persistent loop over tiles; with two buffers the next tile's load is in flight while
    //      this one is processed
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_iter; tile += gridDim.x, it++) {
    const int buf = a.two_buffers ? (it & 1) : 0;
    float *rowbuf = buf ? rowbuf1 : rowbuf0;
    const int base = span_y * a.chunk_valid - a.edge;             // row channel of slot 0
    const float *src = a.dev_t + row * a.dev_stride;
    const bool has_next = tile + gridDim.x < n_iter;
    int64_t next_row = row + step_rows;
    int next_y = span_y + step_y;
    if (MODE == 2) {
        if (has_next) listed(tile + gridDim.x, next_row, next_y);
    } else if (next_y >= a.n_chunks) {
        next_y -= a.n_chunks;
        next_row++;
    }

    // (everybody has left the previous tile's last barrier: thr, Fsm, stat and the other buffer
    // are free)
    if (tid < a.n_windows) {
        thr[tid] = __double2float_rn((a.n_sigma * (double) noise_now) * a.scales[tid]);
        if (has_next) noise_now = a.noise[next_row];
    }

    // ---- stage the span (zeros outside the band)
    if (a.use_tma) {
        if (a.two_buffers) {
            if (tid == 0 && has_next) issue_tile(next_row, next_y, buf ^ 1);
            mbar_wait(&mbar[buf], (uint32_t) (it >> 1) & 1u);
        } else {
            if (tid == 0) issue_tile(row, span_y, 0);
            mbar_wait(&mbar[0], (uint32_t) it & 1u);
        }
    } else {
        const bool vec_ok = ((a.dev_stride & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(a.dev_t) & 15) == 0) && ((base & 3) == 0) &&
                            ((C & 3) == 0);
        if (vec_ok) {
            // all loads of a thread in flight together, then the swizzled stores
            float4 v[RUN / 4];
#pragma unroll
            for (int i = 0; i < RUN / 4; i++) {
                const int g = base + ((tid + i * T) << 2);
                v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (g >= 0 && g < C) v[i] = __ldg(reinterpret_cast<const float4 *>(src + g));
            }
#pragma unroll
            for (int i = 0; i < RUN / 4; i++) {
                const int q = tid + i * T;                       // 16-byte chunk of the span
                *reinterpret_cast<float4 *>(rowbuf + chunk_offset(q >> 3, q & 7)) = v[i];
            }
        } else {
            for (int q = tid; q < (span >> 2); q += T) {
                const int g = base + (q << 2);
                float4 v;
                v.x = (g >= 0 && g < C) ? src[g] : 0.0f;
                v.y = (g + 1 >= 0 && g + 1 < C) ? src[g + 1] : 0.0f;
                v.z = (g + 2 >= 0 && g + 2 < C) ? src[g + 2] : 0.0f;
                v.w = (g + 3 >= 0 && g + 3 < C) ? src[g + 3] : 0.0f;
                *reinterpret_cast<float4 *>(rowbuf + chunk_offset(q >> 3, q & 7)) = v;
            }
        }
    }
    __syncthreads();                                           // thr[] (and the plain staging) visible

    // ---- my run: window size 1, then running sums of what is left
    const int my_sw = tid & 7;                                 // swizzle of my run
    const float *my = rowbuf + tid * PITCH;
    const int64_t pos0 = (int64_t) base + (int64_t) tid * RUN;  // row channel of element 0
    const uint32_t in_range = bit_range(-pos0, (int64_t) C - pos0);

    uint32_t F = 0;
    float m8[4];        // maxima of u over the four groups of 8 samples
    float ppos, sabs;   // sum of the positive u, sum of |u|
    // (re)build u = F ? 0 : x, its running sums and the run statistics; publish them
    auto rebuild = [&](bool first) {
        float x[RUN];
#pragma unroll
        for (int i = 0; i < RUN / 4; i++) {
            const float4 v = *reinterpret_cast<const float4 *>(my + ((i ^ my_sw) << 2));
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        if (first) {
            // window size 1: flag and zero in one go (samples outside the band are zeros and
            // stay zeros whether or not their bit survives the mask)
            const float t0 = thr[0];
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                const bool f = x[j] > t0;
                F |= f ? (1u << j) : 0u;
                x[j] = f ? 0.0f : x[j];
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
        ppos = 0.5f * (total + sabs) * 1.00002f;               // >= sum of max(u, 0), with slack
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
        float lim_small = __int_as_float(0x7f800000);          // strictest limit of the sizes <= 8
        bool any = false;
        const int nf1 = __popc(F) + __popc(F1), nf2 = nf1 + __popc(F2);
        const float bound1 = ppos + st1.y, bound2 = bound1 + st2.y;
        for (int w = 1; w < a.n_windows; w++) {
            const int win = 1 << w;
            if (win > C) break;
            const float tw = thr[w];
            if (tw != tw) continue;
            if (!(tw >= 0.0f)) {
                any = true;                                    // negative threshold: everything is hot
            } else if (win <= 8) {
                lim_small = fminf(lim_small, __fmul_rd(tw, 0.99999905f));
            } else {
                const bool two = win > RUN;
                const float t_min = tw * (float) max(win - (two ? nf2 : nf1), 0);
                any |= !((two ? bound2 : bound1) <= __fmul_rd(t_min, 0.99999f));
            }
        }
        const float reach_max = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), st1.x);
        any |= !(reach_max <= lim_small);
        if (!__syncthreads_or(any)) goto write_out;
    }
    if (MODE == 1) {
        // first pass: leave the tile to the second pass (its flags are not written here)
        if (tid == 0) a.work[1 + atomicAdd(&a.work[0], 1u)] = (uint32_t) (row * a.n_chunks + span_y);
        goto next_tile;
    }

    if (MODE != 1)
    for (int w = 1; w < a.n_windows; w++) {
        const int win = 1 << w;
        if (win > C) break;
        const float tw = thr[w];
        if (tw != tw) continue;                                // NaN threshold: nothing can fire
        uint32_t fire = 0;
        {
            const bool two = win > RUN;
            uint32_t hot = hot_starts(w, tw);
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

write_out:
    {
    // ---- write my 32 flags if my run belongs to this block's output range
    const int64_t out_lo = (int64_t) span_y * a.chunk_valid;
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
next_tile:
    row = next_row;
    span_y = next_y;
    }   // tiles
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
            *reinterpret_cast<uint32_t *>(o) = t0;
            *reinterpret_cast<uint32_t *>(o + fstride) = t1;
            *reinterpret_cast<uint32_t *>(o + 2 * fstride) = t2;
            *reinterpret_cast<uint32_t *>(o + 3 * fstride) = t3;
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

// ---------------------------------------------------------------- more than 7 window sizes
// The kernel above is built around a reach of two runs (windows up to 64).  Window sizes 128 ..
// 1024 (n_windows 8 .. 11) take this plain statement of the contract instead: a block stages a
// chunk of 8192 channels of one row (chunks overlap by the reach of all the sizes together, so
// they are independent), keeps u = flagged ? 0 : x and the flag counts as doubling trees in
// shared memory (one more level per window size while nothing new is flagged, a rebuild after a
// size that flagged something) and dilates firing windows with a doubling OR.
constexpr int TG_CH = 8192;
constexpr int TG_THREADS = 1024;
constexpr int TG_MAX_WINDOWS = KSP_MAX_WINDOWS;

struct TgArgs {
    const float *dev_t;
    const float *noise;
    uint8_t *flags_t;
    uint32_t *bits_t;
    int64_t channels, baselines, dev_stride, out_stride;
    int n_windows, flag_value, edge, valid;
    double n_sigma;
    double scales[TG_MAX_WINDOWS];
};

__global__ void __launch_bounds__(TG_THREADS, 1)
threshold_sum_general_kernel(const TgArgs a)
{
    extern __shared__ __align__(16) uint8_t tg_raw[];
    float *x = reinterpret_cast<float *>(tg_raw);                 // TG_CH samples
    float *ta = x + TG_CH, *tb = ta + TG_CH;                      // tree level, ping-pong
    uint16_t *ca = reinterpret_cast<uint16_t *>(tb + TG_CH);      // flagged-sample counts, ping-pong
    uint16_t *cb = ca + TG_CH;
    uint8_t *fl = reinterpret_cast<uint8_t *>(cb + TG_CH);        // flags so far
    uint8_t *ma = fl + TG_CH, *mb = ma + TG_CH;                   // firing starts -> covered samples
    __shared__ float thr[TG_MAX_WINDOWS];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x;
    const int C = (int) a.channels;
    const int base = (int) blockIdx.y * a.valid - a.edge;        // row channel of slot 0
    const float *src = a.dev_t + row * a.dev_stride;
    if (tid < a.n_windows)
        thr[tid] = __double2float_rn((a.n_sigma * (double) a.noise[row]) * a.scales[tid]);
    for (int i = tid; i < TG_CH; i += TG_THREADS) {
        const int g = base + i;
        x[i] = (g >= 0 && g < C) ? src[g] : 0.0f;
        fl[i] = 0;
    }
    __syncthreads();

    bool tree_ok = false;                                         // ta / ca hold level w - 1 of the current flags
    for (int w = 0; w < a.n_windows; w++) {
        const int win = 1 << w;
        if (win > C) break;
        if (!tree_ok || w == 0) {
            for (int i = tid; i < TG_CH; i += TG_THREADS) {
                ta[i] = fl[i] ? 0.0f : x[i];
                ca[i] = fl[i];
            }
            __syncthreads();
            for (int step = 1; step < win; step <<= 1) {
                for (int i = tid; i < TG_CH; i += TG_THREADS) {
                    const bool in = i + step < TG_CH;
                    tb[i] = ta[i] + (in ? ta[i + step] : 0.0f);
                    cb[i] = (uint16_t) (ca[i] + (in ? ca[i + step] : 0));
                }
                __syncthreads();
                float *tf = ta; ta = tb; tb = tf;
                uint16_t *cf = ca; ca = cb; cb = cf;
            }
        } else {
            const int step = win >> 1;
            for (int i = tid; i < TG_CH; i += TG_THREADS) {
                const bool in = i + step < TG_CH;
                tb[i] = ta[i] + (in ? ta[i + step] : 0.0f);
                cb[i] = (uint16_t) (ca[i] + (in ? ca[i + step] : 0));
            }
            __syncthreads();
            float *tf = ta; ta = tb; tb = tf;
            uint16_t *cf = ca; ca = cb; cb = cf;
        }
        // windows that lie inside the band and inside the chunk
        const double tw = (double) thr[w];
        bool any = false;
        for (int i = tid; i < TG_CH; i += TG_THREADS) {
            const int g = base + i;
            bool fire = false;
            if (g >= 0 && g + win <= C && i + win <= TG_CH)
                fire = (double) ta[i] > tw * (double) (win - (int) ca[i]);
            ma[i] = fire ? 1 : 0;
            any |= fire;
        }
        tree_ok = !__syncthreads_or(any);
        if (!tree_ok) {
            // covered[j] = OR of fire[j - k], k < win: doubling OR towards higher indices
            for (int step = 1; step < win; step <<= 1) {
                for (int i = tid; i < TG_CH; i += TG_THREADS) mb[i] = ma[i] | (i >= step ? ma[i - step] : 0);
                __syncthreads();
                uint8_t *mf = ma; ma = mb; mb = mf;
            }
            for (int i = tid; i < TG_CH; i += TG_THREADS) fl[i] |= ma[i];
            __syncthreads();
        }
    }

    // ---- the chunk's own part of the row
    const int out_lo = (int) blockIdx.y * a.valid;
    const int out_hi = min(C, out_lo + a.valid);
    if (a.bits_t) {
        for (int wd = tid; wd < a.valid / 32; wd += TG_THREADS) {
            const int g0 = out_lo + 32 * wd;
            if (g0 >= out_hi) break;
            uint32_t bits = 0;
            for (int j = 0; j < 32; j++)
                if (g0 + j < out_hi && fl[g0 + j - base]) bits |= 1u << j;
            a.bits_t[row * a.out_stride + (g0 >> 5)] = bits;
        }
    } else {
        for (int g = out_lo + tid; g < out_hi; g += TG_THREADS)
            a.flags_t[row * a.out_stride + g] = fl[g - base] ? (uint8_t) a.flag_value : (uint8_t) 0;
    }
}

int launch_threshold_sum_general(cudaStream_t s, const float *dev_t, const float *noise,
                                 uint8_t *flags_t, uint32_t *bits_t, int64_t channels,
                                 int64_t baselines, int64_t dev_stride, int64_t out_stride,
                                 int n_windows, double n_sigma, const double *scales, int flag_value)
{
    TgArgs a;
    a.dev_t = dev_t; a.noise = noise; a.flags_t = flags_t; a.bits_t = bits_t;
    a.channels = channels; a.baselines = baselines;
    a.dev_stride = dev_stride; a.out_stride = out_stride;
    a.n_windows = n_windows; a.flag_value = flag_value; a.n_sigma = n_sigma;
    for (int w = 0; w < TG_MAX_WINDOWS; w++) a.scales[w] = w < n_windows ? scales[w] : 0.0;
    const int reach = (1 << n_windows) - n_windows - 1;            // influence radius of a sample
    a.edge = (int) (ksp_divup(reach, RUN) * RUN);
    a.valid = TG_CH - 2 * a.edge;
    if (a.valid < RUN) return KSP_ETOOLARGE;
    const int64_t n_chunks = ksp_divup(channels, a.valid);
    if (n_chunks > 65535) return KSP_ETOOLARGE;
    const size_t smem = (size_t) TG_CH * (3 * sizeof(float) + 2 * sizeof(uint16_t) + 3);
    KSP_CUDA(cudaFuncSetAttribute(threshold_sum_general_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    dim3 grid((unsigned) baselines, (unsigned) n_chunks);
    threshold_sum_general_kernel<<<grid, TG_THREADS, smem, s>>>(a);
    KSP_CHECK_LAUNCH();
    return 0;
}

size_t ts_smem_bytes(int threads, int buffers)
{
    // span buffer(s), 7 words of state per run, thresholds and mbarriers, + 1 KB so that the
    // spans can be aligned for the swizzle
    const size_t buf_floats = ((((size_t) threads + 2) * PITCH + 255) / 256) * 256;   // whole KB
    return sizeof(float) * ((size_t) buffers * buf_floats + 7 * ((size_t) threads + 2) + 16) + 1024 + 16;
}

// Blocks of the persistent grid: as many as fit on the device at once.
template <typename Kernel>
int ts_blocks(Kernel kernel, const TsArgs &a, int threads, size_t smem, int64_t *blocks)
{
    int per_sm = 0;
    KSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total = a.baselines * (int64_t) a.n_chunks;
    *blocks = (int64_t) per_sm * ksp_sm_count();
    if (*blocks > total) *blocks = total;
    return 0;
}

template <typename Kernel>
int ts_launch(Kernel kernel, cudaStream_t s, const TsArgs &a, const CUtensorMap &tmap, int threads,
              size_t smem)
{
    int per_sm = 0;
    KSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total = a.baselines * (int64_t) a.n_chunks;
    int64_t blocks = (int64_t) per_sm * ksp_sm_count();
    if (blocks > total) blocks = total;
    kernel<<<(unsigned) blocks, threads, smem, s>>>(a, tmap);
    KSP_CHECK_LAUNCH();
    return 0;
}

int launch_threshold_sum(cudaStream_t s, const float *dev_t, const float *noise, uint8_t *flags_t,
                         uint32_t *bits_t, int64_t channels, int64_t baselines, int64_t dev_stride,
                         int64_t out_stride, int n_windows, double n_sigma, const double *scales,
                         int flag_value, uint32_t *work = nullptr)
{
    if (channels < 0 || baselines < 0 || dev_stride < channels) return KSP_EINVAL;
    if (n_windows < 1 || !scales) return KSP_EINVAL;
    if (n_windows > TG_MAX_WINDOWS) return KSP_ETOOLARGE;
    if (channels > 0x7fff0000) return KSP_ETOOLARGE;
    if (channels == 0 || baselines == 0) return 0;
    if (!dev_t || !noise || (!flags_t && !bits_t)) return KSP_EINVAL;
    if (baselines > 0x7fffffff) return KSP_ETOOLARGE;
    if (n_windows > TS_MAX_WINDOWS)
        return launch_threshold_sum_general(s, dev_t, noise, flags_t, bits_t, channels, baselines,
                                            dev_stride, out_stride, n_windows, n_sigma, scales,
                                            flag_value);

    TsArgs a;
    a.dev_t = dev_t; a.noise = noise; a.flags_t = flags_t; a.bits_t = bits_t;
    a.channels = channels; a.baselines = baselines;
    a.dev_stride = dev_stride; a.out_stride = out_stride;
    a.n_windows = n_windows; a.flag_value = flag_value; a.n_sigma = n_sigma;
    for (int w = 0; w < TS_MAX_WINDOWS; w++) a.scales[w] = w < n_windows ? scales[w] : 0.0;

    // Threads per block for long rows (KSP_TS_THREADS, default 128): spans of 4096 channels
    // that overlap by the reach of the largest window keep many blocks resident per SM.
    static const int chunk_threads = [] {
        const char *e = getenv("KSP_TS_THREADS");
        int v = e ? atoi(e) : 128;
        if (v < 32 || v > TS_MAX_THREADS || (v & 31)) v = 128;
        return v;
    }();
    int threads, n_chunks;
    const int64_t runs = ksp_divup(channels, RUN);
    if (runs <= chunk_threads) {
        threads = (int) (ksp_divup(runs, 32) * 32);
        a.edge = 0;
        a.chunk_valid = threads * RUN;
        n_chunks = 1;
    } else {
        const int reach = (1 << n_windows) - n_windows - 1;   // influence radius of a sample
        a.edge = (int) (ksp_divup(reach, RUN) * RUN);
        // balance the spans: as few blocks as chunk_threads allows, equal valid parts
        const int max_valid = chunk_threads * RUN - 2 * a.edge;
        if (max_valid < RUN) return KSP_EINVAL;
        n_chunks = (int) ksp_divup(channels, max_valid);
        a.chunk_valid = (int) (ksp_divup(ksp_divup(channels, n_chunks), RUN) * RUN);
        threads = (int) (ksp_divup((a.chunk_valid + 2 * a.edge) / RUN, 32) * 32);
        a.chunk_valid = threads * RUN - 2 * a.edge;
        n_chunks = (int) ksp_divup(channels, a.chunk_valid);
    }
    a.n_chunks = n_chunks;
    // TMA staging: dev_t viewed as [baselines][channels / 32][32] floats, box = [1][threads][32],
    // 128-byte swizzle, zero fill outside the tensor (needs whole runs and 16-byte alignment)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    a.use_tma = 0;
    static const bool tma_allowed = [] {
        const char *e = getenv("KSP_TS_TMA");
        return !(e && atoi(e) == 0);
    }();
    if (tma_allowed && (channels % RUN) == 0 && (dev_stride % 4) == 0 &&
        ((uintptr_t) dev_t % 16) == 0 && threads <= 256 && tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t) RUN, (cuuint64_t) (channels / RUN), (cuuint64_t) baselines};
        const cuuint64_t strides[2] = {RUN * sizeof(float), (cuuint64_t) dev_stride * sizeof(float)};
        const cuuint32_t box[3] = {RUN, (cuuint32_t) threads, 1};
        const cuuint32_t elem[3] = {1, 1, 1};
        CUresult rc = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *) dev_t, dims,
                                           strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        a.use_tma = (rc == CUDA_SUCCESS) ? 1 : 0;
    }
    // second span buffer while it fits in the default 48 KB of dynamic shared memory
    a.two_buffers = (a.use_tma && ts_smem_bytes(threads, 2) <= 48 * 1024 &&
                     baselines * (int64_t) n_chunks > 1) ? 1 : 0;
    const size_t smem = ts_smem_bytes(threads, a.two_buffers ? 2 : 1);
    a.work = nullptr;
    if (bits_t) {
        if (threads == 128 && work && a.two_buffers && baselines * (int64_t) n_chunks < 0x7fffffff) {
            // two passes (see the kernel): lean first pass over every tile, full algorithm over
            // the tiles it lists
            static const bool two_pass = [] {
                const char *e = getenv("KSP_TS_TWO_PASS");
                return !(e && atoi(e) == 0);
            }();
            if (two_pass) {
                a.work = work;
                KSP_CUDA(cudaMemsetAsync(work, 0, sizeof(uint32_t), s));
                // the first pass runs single-buffered: with its small footprint (56 registers,
                // 21 KB) nine blocks share an SM and hide each other's tile loads
                TsArgs a1 = a;
                a1.two_buffers = 0;
                int rc = ts_launch(threshold_sum_kernel<true, 128, 1>, s, a1, tmap, threads,
                                   ts_smem_bytes(threads, 1));
                if (rc) return rc;
                return ts_launch(threshold_sum_kernel<true, 128, 2>, s, a, tmap, threads, smem);
            }
        }
        if (threads == 128) return ts_launch(threshold_sum_kernel<true, 128, 0>, s, a, tmap, threads, smem);
        return ts_launch(threshold_sum_kernel<true, 0, 0>, s, a, tmap, threads, smem);
    }
    if (threads == 128) return ts_launch(threshold_sum_kernel<false, 128, 0>, s, a, tmap, threads, smem);
    return ts_launch(threshold_sum_kernel<false, 0, 0>, s, a, tmap, threads, smem);
}

}  // namespace

extern "C" int ksp_threshold_sum(void *stream, const float *dev_t, const float *noise,
                                 uint8_t *flags_t, int64_t channels, int64_t baselines,
                                 int64_t dev_stride, int64_t flags_stride, int n_windows,
                                 double n_sigma, const double *scales, int flag_value)
{
    if (flags_stride < channels) return KSP_EINVAL;
    return launch_threshold_sum((cudaStream_t) stream, dev_t, noise, flags_t, nullptr, channels,
                                baselines, dev_stride, flags_stride, n_windows, n_sigma, scales,
                                flag_value);
}

// internal (fused flagger): bit-packed output, words_stride words per baseline row
// work: optional scratch of ksp_threshold_work_bytes() bytes for the two-pass mode
int ksp_threshold_sum_packed(cudaStream_t s, const float *dev_t, const float *noise,
                             uint32_t *bits_t, int64_t channels, int64_t baselines,
                             int64_t dev_stride, int64_t words_stride, int n_windows,
                             double n_sigma, const double *scales, uint32_t *work)
{
    if (words_stride < ksp_divup(channels, 32)) return KSP_EINVAL;
    return launch_threshold_sum(s, dev_t, noise, nullptr, bits_t, channels, baselines, dev_stride,
                                words_stride, n_windows, n_sigma, scales, 1, work);
}

// upper bound on the tile list of the two-pass mode: one word per (baseline, span) + the count
size_t ksp_threshold_work_bytes(int64_t channels, int64_t baselines)
{
    const int64_t spans = channels / 2048 + 2;              // spans hold at least 2048 own channels
    return (size_t) (baselines * spans + 1) * sizeof(uint32_t);
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
