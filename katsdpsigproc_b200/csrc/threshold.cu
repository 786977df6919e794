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

#include "threshold_tile.cuh"

namespace {

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
    float *thr = reinterpret_cast<float *>(car2 + T);          // TS_THR_WORDS
    uint64_t *mbar = reinterpret_cast<uint64_t *>(thr + TS_THR_WORDS);   // TMA completion barriers, one per buffer

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
    if (tid < 32 && (int64_t) blockIdx.x < n_iter) noise_now = a.noise[row];
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
    if (tid < 32) {
        ts_thresholds(thr, tid, a.n_windows, a.n_sigma, noise_now, a.scales, C);
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

    // ---- my run: window size 1, the filters and the vote, then (MODE != 1) the window sizes
    const int64_t pos0 = (int64_t) base + (int64_t) tid * RUN;  // row channel of element 0
    const uint32_t in_range = bit_range(-pos0, (int64_t) C - pos0);
    uint32_t F = 0;
    {
        TsTile tl;
        tl.rowbuf = rowbuf; tl.stat = stat; tl.Fsm = Fsm; tl.car1 = car1; tl.car2 = car2; tl.thr = thr;
        tl.T = T; tl.span = span; tl.C = C; tl.n_windows = a.n_windows; tl.pos0 = pos0;
        if (ts_process_tile<MODE == 1>(tl, F)) {
            // first pass: leave the tile to the second pass (its flags are not written here)
            if (tid == 0) a.work[1 + atomicAdd(&a.work[0], 1u)] = (uint32_t) (row * a.n_chunks + span_y);
            goto next_tile;
        }
    }

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

// ---------------------------------------------------------------- packed flags -> bytes (threshold_tile.cuh)

// Tile: 256 baselines x 8 words (256 channels) by 128 threads.  A thread owns 16 consecutive
// baselines of one word (32 channels): 16 words in, 32 rows of 16 flag bytes out, each row ONE
// 128-bit streaming store - the 16 lanes that share a word fill 256 contiguous bytes of a flags row.
// Bits become bytes through a 256-entry table in shared memory (8 flags per 64-bit load; flags
// are sparse, so almost every lookup is the broadcast of entry 0) and the 4 x 4 byte blocks are
// transposed with byte permutes: about one instruction per flag byte (the 32-bit stores of
// expand_flags_tile, which the dataflow kernel keeps, cost 4.2).
constexpr int EX_THREADS = 128;
constexpr int EX_BL = 256;               // baselines per tile
constexpr int EX_WORDS = 8;              // words per tile
constexpr int EX_PITCH = 20;             // words per group of 16 baselines: conflict-free 128-bit reads

__device__ __forceinline__ void stg_stream_u4(uint8_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(a), "r"(b), "r"(c),
                 "r"(d), "l"(l2_evict_first())
                 : "memory");
}

__global__ void __launch_bounds__(EX_THREADS)
expand_flags_kernel(const uint32_t *__restrict__ bits_t, uint8_t *__restrict__ flags,
                    int64_t channels, int64_t baselines, int64_t wstride, int64_t fstride,
                    int flag_value)
{
    __shared__ __align__(16) uint32_t tile[EX_WORDS][EX_BL / 16][EX_PITCH];
    __shared__ __align__(8) uint2 lut[256];
    const int t = threadIdx.x;
    const int64_t b0 = (int64_t) blockIdx.x * EX_BL, w0 = (int64_t) blockIdx.y * EX_WORDS;
    const int64_t n_words = (channels + 31) >> 5;
    const uint32_t fv = (uint32_t) flag_value & 0xffu;
    for (int v = t; v < 256; v += EX_THREADS)           // byte of 8 flag bits -> 8 flag bytes
        lut[v] = make_uint2((((uint32_t) v & 0xfu) * 0x00204081u & 0x01010101u) * fv,
                            (((uint32_t) v >> 4) * 0x00204081u & 0x01010101u) * fv);
#pragma unroll
    for (int k = 0; k < EX_BL * EX_WORDS / EX_THREADS; k++) {
        const int i = t + EX_THREADS * k, w = i & (EX_WORDS - 1), b = i / EX_WORDS;
        uint32_t v = 0;
        if (b0 + b < baselines && w0 + w < n_words) v = __ldg(bits_t + (b0 + b) * wstride + w0 + w);
        tile[w][b >> 4][b & 15] = v;
    }
    __syncthreads();
    const int bg = t & 15, w = t >> 4;
    if (w0 + w >= n_words) return;
    uint32_t wv[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint4 q = *reinterpret_cast<const uint4 *>(&tile[w][bg][4 * k]);
        wv[4 * k] = q.x; wv[4 * k + 1] = q.y; wv[4 * k + 2] = q.z; wv[4 * k + 3] = q.w;
    }
    const int64_t c_base = (w0 + w) * 32;
    const int64_t b = b0 + 16 * bg;
    if (b >= baselines) return;
    const bool vec = (b + 16 <= baselines) && (c_base + 32 <= channels) && ((fstride & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(flags) & 15) == 0);
    if (vec) {
        uint8_t *out = flags + c_base * fstride + b;
#pragma unroll
        for (int h = 0; h < 4; h++) {                   // channels 8 h .. 8 h + 7
            uint2 e[16];
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = lut[(wv[i] >> (8 * h)) & 0xffu];
#pragma unroll
            for (int half = 0; half < 2; half++) {      // channels 8 h + 4 half .. + 3
                uint32_t o[4][4];                       // o[k][j]: channel k, baselines 4 j .. 4 j + 3
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t r0 = half ? e[4 * j].y : e[4 * j].x, r1 = half ? e[4 * j + 1].y : e[4 * j + 1].x;
                    const uint32_t r2 = half ? e[4 * j + 2].y : e[4 * j + 2].x, r3 = half ? e[4 * j + 3].y : e[4 * j + 3].x;
                    // r_i: byte k = channel k of baseline 4 j + i; o[k][j]: byte i = baseline 4 j + i
                    const uint32_t a01 = __byte_perm(r0, r1, 0x5140), b01 = __byte_perm(r0, r1, 0x7362);
                    const uint32_t a23 = __byte_perm(r2, r3, 0x5140), b23 = __byte_perm(r2, r3, 0x7362);
                    o[0][j] = __byte_perm(a01, a23, 0x5410);
                    o[1][j] = __byte_perm(a01, a23, 0x7632);
                    o[2][j] = __byte_perm(b01, b23, 0x5410);
                    o[3][j] = __byte_perm(b01, b23, 0x7632);
                }
                uint8_t *row = out + (int64_t) (8 * h + 4 * half) * fstride;
#pragma unroll
                for (int k = 0; k < 4; k++)             // write-once output: first to leave the L2
                    stg_stream_u4(row + k * fstride, o[k][0], o[k][1], o[k][2], o[k][3]);
            }
        }
        return;
    }
    // ragged edges, unaligned output: flag by flag
    for (int bit = 0; bit < 32 && c_base + bit < channels; bit++)
#pragma unroll 4
        for (int i = 0; i < 16; i++)
            if (b + i < baselines) flags[(c_base + bit) * fstride + b + i] = (uint8_t) (((wv[i] >> bit) & 1u) * fv);
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
    return sizeof(float) * ((size_t) buffers * buf_floats + 7 * ((size_t) threads + 2) + 32) + 1024 + 16;
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
    dim3 grid((unsigned) ksp_divup(baselines, EX_BL), (unsigned) ksp_divup(ksp_divup(channels, 32), EX_WORDS));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    expand_flags_kernel<<<grid, EX_THREADS, 0, s>>>(bits_t, flags, channels, baselines, words_stride,
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
