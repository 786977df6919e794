// Dataflow flagger: the whole median + MAD + SumThreshold pipeline of reference
// rfi/device.py:1111-1166 (background -> transpose -> noise -> threshold -> transpose) as ONE
// persistent kernel per dump.
//
// The four stages of flagger.cu (bg13 tiles, noise rows, threshold spans, bit -> byte expansion)
// become four kinds of work ITEM.  A fixed schedule interleaves them strip by strip (a strip =
// 32 baselines x all channels):
//
//   step t:  background tiles of strip t,
//            noise rows of strip t - L1,
//            threshold rows of strip t - L2,
//            a quarter of the expansion tiles of the group (4 strips) that finished L3 steps ago
//
// Every block (256 threads, 4 per SM) takes the next item with one atomic ticket, waits until
// the items it depends on have been completed (per-strip counters in global memory; they were
// handed out L1 / L2 - L1 / ... steps earlier, so the wait is almost always over before it
// starts), runs the same device code as the stand-alone kernels and bumps its strip's counter.
// The deviations live in a RING of R strips (R x 32 x channels floats, 48 MB at 32768
// channels): they are written, read twice and overwritten while still in the 126 MB L2, so
// HBM sees the visibilities once and the flag bytes once (9 B/vis instead of 20).  The packed
// flags go through a second, small ring.
//
// Deadlock freedom: every wait is on items with a SMALLER ticket, tickets are handed out in
// order to running blocks, and a block holds one ticket at a time - the block with the smallest
// unfinished ticket never waits.  A wait that lasts longer than DF_TIMEOUT cycles sets the error
// word and every block leaves (ksp_flagger_stats reports it).
#include "common.cuh"
#include "tma.cuh"
#include "bg13.cuh"
#include "madnz_stream.cuh"
#include "threshold_tile.cuh"
#include "dataflow.h"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int DF_THREADS = 256;
constexpr int DF_BLOCKS_PER_SM = 4;
#ifndef DF_BG_PF_VALUE
#define DF_BG_PF_VALUE 2
#endif
constexpr int DF_BG_PF = DF_BG_PF_VALUE;    // load groups in flight per thread in the background tile's phase 1
constexpr long long DF_TIMEOUT = 20000000000ll;     // cycles (~10 s): a broken schedule, not a slow one
static_assert(DF_THREADS == BG_THREADS && DF_THREADS == MS_THREADS, "one block shape for every item");

struct DfArgs {
    BgArgs bg;                 // vis, ring (out), input flags, shapes and strides
    float *noise;
    uint8_t *flags;
    int64_t flags_stride;
    uint32_t *bits;            // ring of RB strips x 32 rows x words_stride words
    int64_t words_stride;
    int flag_value;
    int n_windows;
    double n_sigma;
    double scales[TS_MAX_WINDOWS];
    int T;                     // staged runs per threshold span (multiple of 32, <= 256)
    int n_chunks, chunk_valid, edge;
    int S, G;                  // strips, groups of 4 strips
    int n_bg, n_exp, n_expq, K, steps;
    uint32_t total;            // steps * K items
    int L1, L2, L3, R, RB;
    uint32_t *ctl;             // [0] ticket, [1] error
    uint32_t *bg_done, *noise_done, *thr_done, *exp_done;
    unsigned long long *stats; // KSP_DF_STATS words
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Thread 0 of the block: wait until *ctr >= target.  false: the launch is being abandoned.
__device__ __noinline__ bool df_wait(const uint32_t *ctr, uint32_t target, uint32_t *err,
                                     long long &waited)
{
    if (ld_acquire(ctr) >= target) return true;
    const long long t0 = clock64();
    unsigned ns = 64;
    for (;;) {
        __nanosleep(ns);
        if (ns < 2048) ns *= 2;
        if (ld_acquire(ctr) >= target) break;
        if (*reinterpret_cast<volatile uint32_t *>(err) != 0u) return false;
        if (clock64() - t0 > DF_TIMEOUT) {
            atomicExch(err, 1u);
            return false;
        }
    }
    waited += clock64() - t0;
    return true;
}

// rows (baselines) of strip s
__device__ __forceinline__ int strip_rows(const DfArgs &a, int s)
{
    const int64_t left = a.bg.baselines - (int64_t) s * 32;
    return left >= 32 ? 32 : (int) left;
}

// Block-wide bookkeeping in shared memory: nothing of it occupies registers across the items.
struct DfShared {
    long long cycles[6];       // per kind of item, [4] waiting, [5] start of the current item
    uint64_t mbar[2];          // TMA completion barriers of the two threshold span buffers
    uint32_t ticket, parity, items;
    uint32_t *counter;         // what the item that just ran has to bump (null: nothing)
    int kind;                  // its kind, for the time accounting
    int go;
    int strip, tile, row_off;  // the current background tile (see WhereInRing)
};

// The launch's arguments and the bookkeeping live at fixed shared-memory addresses: the item
// functions below are real calls (not inlined, so each kind of item gets the block's whole
// register budget - the background tile needs all 64) and reach everything through these
// without holding a pointer.
__shared__ DfArgs s_args;
__shared__ DfShared s_sh;
extern __shared__ __align__(1024) uint8_t df_sm_raw[];

// the threshold span's swizzle pattern is a function of the shared-memory address: 1 KB aligned
__device__ __forceinline__ uint8_t *df_smem()
{
    return df_sm_raw + ((1024u - (smem_u32(df_sm_raw) & 1023u)) & 1023u);
}

// Coordinates of a background tile of the dataflow kernel, re-read from shared memory whenever
// bg13_tile asks (volatile: never kept in a register from one phase to the next).
struct WhereInRing {
    const volatile DfShared *sh;
    template <int TC> __device__ __forceinline__ int c0() const { return sh->tile * TC; }
    __device__ __forceinline__ int64_t b0() const { return (int64_t) sh->strip * 32; }
    __device__ __forceinline__ int64_t row_off() const { return (int64_t) sh->row_off; }
    static constexpr bool KEEP_IN_L2 = true;     // the ring is read twice and overwritten while in L2
};

// Item done: say which counter to bump.  The scheduling loop does it after its next barrier (all
// of the item's writes are then complete), on another warp than the one that fetches the next
// ticket, so that the fence and the ticket's round trip overlap.
__device__ __forceinline__ void df_complete(uint32_t *counter, DfShared *sh, int kind)
{
    if (threadIdx.x == 0) {
        sh->counter = counter;
        sh->kind = kind;
    }
}

// The four kinds of item; they return false when the launch is being abandoned.

// ---- background tile `tile` of strip s
template <int IN_MODE, int FLAG_MODE>
__device__ __noinline__ bool df_background(int s, int tile)
{
    const DfArgs &a = s_args;
    DfShared *sh = &s_sh;
    uint8_t *sm = df_smem();
    if (threadIdx.x == 0) {
        int go = 1;
        if (s >= a.R)                      // the ring slot's previous strip has been thresholded
            go = df_wait(&a.thr_done[s - a.R], (uint32_t) strip_rows(a, s - a.R),
                         &a.ctl[1], sh->cycles[4]);
        sh->go = go;
        sh->strip = s;
        sh->tile = tile;
        sh->row_off = 32 * (s % a.R) - 32 * s;
    }
    __syncthreads();
    if (!sh->go) return false;
    WhereInRing at;
    at.sh = sh;
    bg13_tile<IN_MODE, FLAG_MODE, true, BG_TC, WhereInRing, DF_BG_PF>(a.bg, at, reinterpret_cast<float *>(sm));
    if (threadIdx.x == 0) {
        sh->counter = &a.bg_done[sh->strip];
        sh->kind = 0;
    }
    return true;
}

// ---- noise of row r of strip s
__device__ __noinline__ bool df_noise(int s, int r)
{
    const DfArgs &a = s_args;
    DfShared *sh = &s_sh;
    uint8_t *sm = df_smem();
    if (threadIdx.x == 0) sh->go = df_wait(&a.bg_done[s], (uint32_t) a.n_bg, &a.ctl[1], sh->cycles[4]);
    __syncthreads();
    if (!sh->go) return false;
    const float *row = a.bg.out + ((int64_t) 32 * (s % a.R) + r) * a.bg.out_stride;
    const bool vec_ok = ((a.bg.out_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.bg.out) & 15) == 0);
    madnz_stream_row<true>(row, a.noise + (int64_t) s * 32 + r, (int) a.bg.channels, vec_ok,
                           reinterpret_cast<uint32_t *>(sm), &a.stats[KSP_DF_STAT_FALLBACKS]);
    df_complete(&a.noise_done[s], sh, 1);
    return true;
}

// The full SumThreshold of a span whose vote said that some larger window might fire (rare; a
// call of its own so that its registers do not weigh on the common path).
__device__ __noinline__ uint32_t df_threshold_full(const TsTile &tl)
{
    uint32_t F = 0;
    __syncthreads();                       // the vote's readers are done with stat / Fsm
    ts_process_tile<false>(tl, F);
    return F;
}

// ---- threshold of row r of strip s: its spans one after the other, the next span's TMA load in
//      flight while this one is worked on
__device__ __noinline__ bool df_threshold(const CUtensorMap *tmap, int s, int r)
{
    const DfArgs &a = s_args;
    DfShared *sh = &s_sh;
    uint8_t *sm = df_smem();
    const int tid = threadIdx.x;
    const int T = a.T;
    // layout: two spans of (T + 2) runs, statistics, flag words, carries, thresholds
    const int buf_floats = (((T + 2) * PITCH + 255) / 256) * 256;
    float *rowbuf0 = reinterpret_cast<float *>(sm);
    float4 *stat = reinterpret_cast<float4 *>(rowbuf0 + 2 * buf_floats);         // T + 2 (<= DF_THREADS + 2)
    uint32_t *Fsm = reinterpret_cast<uint32_t *>(stat + DF_THREADS + 2);         // T + 2
    uint32_t *car1 = Fsm + DF_THREADS + 2;                                       // DF_THREADS
    uint32_t *car2 = car1 + DF_THREADS;                                          // DF_THREADS
    float *thr = reinterpret_cast<float *>(car2 + DF_THREADS);                   // TS_THR_WORDS
    const int64_t b = (int64_t) s * 32 + r;
    const int ring_row = 32 * (s % a.R) + r;
    // this shared memory was last written with ordinary stores (previous item): order them before
    // the TMA writes of the spans
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        int go = df_wait(&a.noise_done[s], (uint32_t) strip_rows(a, s), &a.ctl[1], sh->cycles[4]);
        if (go && s >= a.RB)               // the bit ring's previous strip has been expanded
            go = df_wait(&a.exp_done[(s - a.RB) >> 2], (uint32_t) a.n_exp, &a.ctl[1], sh->cycles[4]);
        if (go) {
            // the deviations were written with ordinary stores by other blocks: order them before
            // this block's reads through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
            for (int y = 0; y < 2 && y < a.n_chunks; y++) {
                mbar_expect_tx(&sh->mbar[y], (uint32_t) T * RUN * 4u);
                tma_load_3d(rowbuf0 + y * buf_floats, tmap, 0, (y * a.chunk_valid - a.edge) >> 5, ring_row,
                            &sh->mbar[y]);
            }
        }
        sh->go = go;
    }
    if (tid >= 32 && tid < 34) {
        Fsm[T + tid - 32] = 0u;
        stat[T + tid - 32] = make_float4(-__int_as_float(0x7f800000), 0.0f, 0.0f, 0.0f);
    }
    if (tid >= 64 && tid < 64 + 2 * PITCH) {               // two runs of zeros past each span
        rowbuf0[T * PITCH + tid - 64] = 0.0f;
        rowbuf0[buf_floats + T * PITCH + tid - 64] = 0.0f;
    }
    __syncthreads();
    if (!sh->go) return false;
    if (tid < 32)
        ts_thresholds(thr, tid, a.n_windows, a.n_sigma, __ldcg(a.noise + b), a.scales, (int) a.bg.channels);
    const int C = (int) a.bg.channels;
    uint32_t *bits_row = a.bits + ((int64_t) 32 * (s % a.RB) + r) * a.words_stride;
    for (int y = 0; y < a.n_chunks; y++) {
        float *rowbuf = rowbuf0 + (y & 1) * buf_floats;
        mbar_wait(&sh->mbar[y & 1], (sh->parity >> (y & 1)) & 1u);
        __syncthreads();                   // thr[] visible; the previous span's statistics are free
        const int base = y * a.chunk_valid - a.edge;               // row channel of slot 0
        const int64_t pos0 = (int64_t) base + (int64_t) tid * RUN;
        uint32_t F = 0;
        TsTile tl;
        tl.rowbuf = rowbuf; tl.stat = stat; tl.Fsm = Fsm; tl.car1 = car1; tl.car2 = car2; tl.thr = thr;
        tl.T = T; tl.span = T * RUN; tl.C = C; tl.n_windows = a.n_windows; tl.pos0 = pos0;
        if (ts_process_tile<true>(tl, F)) F = df_threshold_full(tl);   // block-uniform
        const int64_t out_lo = (int64_t) y * a.chunk_valid;
        const int64_t out_hi = min((int64_t) C, out_lo + (int64_t) a.chunk_valid);
        if (tid < T && pos0 >= out_lo && pos0 < out_hi)
            bits_row[pos0 >> 5] = F & bit_range(-pos0, (int64_t) C - pos0);
        __syncthreads();                   // everybody is done with this buffer
        if (tid == 0) {
            sh->parity ^= 1u << (y & 1);
            if (y + 2 < a.n_chunks) {
                mbar_expect_tx(&sh->mbar[y & 1], (uint32_t) T * RUN * 4u);
                tma_load_3d(rowbuf, tmap, 0, ((y + 2) * a.chunk_valid - a.edge) >> 5, ring_row, &sh->mbar[y & 1]);
            }
        }
    }
    df_complete(&a.thr_done[s], sh, 2);
    return true;
}

// ---- expansion tile `tile` of group g
__device__ __noinline__ bool df_expand(int g, int tile)
{
    const DfArgs &a = s_args;
    DfShared *sh = &s_sh;
    uint8_t *sm = df_smem();
    if (threadIdx.x == 0) {
        int go = 1;
        for (int s = 4 * g; s < 4 * g + 4 && s < a.S && go; s++)
            go = df_wait(&a.thr_done[s], (uint32_t) strip_rows(a, s), &a.ctl[1], sh->cycles[4]);
        sh->go = go;
    }
    __syncthreads();
    if (!sh->go) return false;
    const int64_t row_off = (int64_t) 32 * ((4 * g) % a.RB) - (int64_t) 128 * g;
    expand_flags_tile<true>(a.bits, a.flags, a.bg.channels, a.bg.baselines, a.words_stride,
                            a.flags_stride, a.flag_value, (int64_t) 128 * g, (int64_t) 8 * tile, row_off,
                            reinterpret_cast<uint32_t (*)[132]>(sm));
    df_complete(&a.exp_done[g], sh, 3);
    return true;
}

template <int IN_MODE, int FLAG_MODE>
__global__ void __launch_bounds__(DF_THREADS, DF_BLOCKS_PER_SM)
dataflow_kernel(const __grid_constant__ DfArgs a_param, const __grid_constant__ CUtensorMap tmap)
{
    const int tid = threadIdx.x;
    {
        static_assert(sizeof(DfArgs) % 4 == 0, "copied word by word");
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&a_param);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_args);
        for (int i = tid; i < (int) (sizeof(DfArgs) / 4); i += DF_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        mbar_init(&s_sh.mbar[0], 1);
        mbar_init(&s_sh.mbar[1], 1);
        for (int i = 0; i < 6; i++) s_sh.cycles[i] = 0;
        s_sh.parity = 0;
        s_sh.items = 0;
        s_sh.counter = nullptr;
        s_sh.kind = -1;
        s_sh.cycles[5] = clock64();
    }
    for (;;) {
        __syncthreads();                  // the previous item's writes are done, its shared memory is free
        const DfArgs &a = s_args;
        if (threadIdx.x == 32 && s_sh.counter) {       // publish the previous item
            __threadfence();
            atomicAdd(s_sh.counter, 1u);
        }
        if (threadIdx.x == 0) {
            uint32_t t = atomicAdd(&a.ctl[0], 1u);
            const long long now = clock64();
            if (s_sh.kind >= 0) {
                s_sh.cycles[s_sh.kind] += now - s_sh.cycles[5];
                s_sh.items++;
            }
            s_sh.cycles[5] = now;
            if (*reinterpret_cast<volatile uint32_t *>(&a.ctl[1]) != 0u) t = 0xffffffffu;
            s_sh.ticket = t;
        }
        __syncthreads();
        if (threadIdx.x == 0) {           // (thread 32 has read them)
            s_sh.counter = nullptr;
            s_sh.kind = -1;
        }
        const uint32_t ticket = s_sh.ticket;
        if (ticket >= a.total) break;
        const int step = (int) (ticket / (uint32_t) a.K);
        int slot = (int) (ticket - (uint32_t) step * (uint32_t) a.K);
        bool ok = true;
        if (slot < a.n_bg) {
            if (step < a.S) ok = df_background<IN_MODE, FLAG_MODE>(step, slot);
        } else if ((slot -= a.n_bg) < 32) {
            const int s = step - a.L1;
            if (s >= 0 && s < a.S && slot < strip_rows(a, s)) ok = df_noise(s, slot);
        } else if ((slot -= 32) < 32) {
            const int s = step - a.L2;
            if (s >= 0 && s < a.S && slot < strip_rows(a, s)) ok = df_threshold(&tmap, s, slot);
        } else {
            slot -= 32;
            const int e = step - a.L3;
            const int g = e >> 2, tile = (e & 3) * a.n_expq + slot;
            if (e >= 0 && g < a.G && tile < a.n_exp) ok = df_expand(g, tile);
        }
        if (!ok) break;
    }
    if (threadIdx.x == 0) {
        const DfArgs &a = s_args;
        atomicAdd(&a.stats[KSP_DF_STAT_CYCLES_BG], (unsigned long long) s_sh.cycles[0]);
        atomicAdd(&a.stats[KSP_DF_STAT_CYCLES_NOISE], (unsigned long long) s_sh.cycles[1]);
        atomicAdd(&a.stats[KSP_DF_STAT_CYCLES_THRESHOLD], (unsigned long long) s_sh.cycles[2]);
        atomicAdd(&a.stats[KSP_DF_STAT_CYCLES_EXPAND], (unsigned long long) s_sh.cycles[3]);
        atomicAdd(&a.stats[KSP_DF_STAT_CYCLES_WAIT], (unsigned long long) s_sh.cycles[4]);
        atomicAdd(&a.stats[KSP_DF_STAT_ITEMS], (unsigned long long) s_sh.items);
    }
}

constexpr int DF_SPAN_RUNS = 160;        // runs of 32 channels per staged threshold span: two buffers fit
constexpr size_t df_thr_smem(int T)
{
    return 2 * (size_t) ((((T + 2) * PITCH + 255) / 256) * 256) * 4 + (size_t) (DF_THREADS + 2) * 16 +
           (size_t) (DF_THREADS + 2) * 4 + (size_t) DF_THREADS * 8 + TS_THR_WORDS * 4;
}
constexpr size_t cmax(size_t x, size_t y) { return x > y ? x : y; }
constexpr size_t DF_SMEM = 1024 + cmax(cmax((size_t) TileGeom<BG_TC>::SMEM_BYTES, (size_t) MS_SMEM_WORDS * 4),
                                       cmax(df_thr_smem(DF_SPAN_RUNS), (size_t) EXPAND_SMEM_WORDS * 4));
static_assert(DF_SMEM <= 56 * 1024, "four blocks per SM");

int env_int(const char *name, int def, int lo, int hi)
{
    const char *e = getenv(name);
    if (!e || !*e) return def;
    int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

struct DfConfig { int mode, L1, L2, L3, R, RB; };

// Read once per process: the scratch size and the launch must agree.
const DfConfig &df_config()
{
    static const DfConfig c = [] {
        DfConfig c;
        c.mode = env_int("KSP_DATAFLOW", -1, -1, 1);         // -1 auto, 0 never, 1 wherever legal
        c.L1 = env_int("KSP_DF_L1", 3, 1, 64);
        c.L2 = c.L1 + env_int("KSP_DF_L2", 3, 1, 64);
        c.L3 = c.L2 + 3 + env_int("KSP_DF_L3", 3, 1, 64);
        c.R = env_int("KSP_DF_RING", c.L2 + 6, c.L2 + 2, 1024);
        c.RB = (env_int("KSP_DF_BITS_RING", c.L3 - c.L2 + 16, c.L3 - c.L2 + 8, 4096) + 3) / 4 * 4;
        return c;
    }();
    return c;
}

struct DfLayout {
    int64_t dev_stride, words_stride;
    int S, G, T, n_chunks, chunk_valid, edge;
    size_t ctl_bytes, ring_bytes, bits_bytes;
};

DfLayout df_layout(const ksp_flagger_params *p)
{
    const DfConfig &c = df_config();
    DfLayout l;
    l.dev_stride = p->channels;                               // channels % 32 == 0
    l.words_stride = ksp_divup(p->channels / 32, 4) * 4;
    l.S = (int) ksp_divup(p->baselines, 32);
    l.G = (l.S + 3) / 4;
    const int64_t runs = p->channels / RUN;
    if (runs <= DF_SPAN_RUNS) {
        l.T = (int) (ksp_divup(runs, 32) * 32);
        l.edge = 0;
        l.chunk_valid = l.T * RUN;
        l.n_chunks = 1;
    } else {
        const int reach = (1 << p->n_windows) - p->n_windows - 1;      // influence radius of a sample
        l.edge = (int) (ksp_divup(reach, RUN) * RUN);
        const int max_valid = DF_SPAN_RUNS * RUN - 2 * l.edge;
        int n = (int) ksp_divup(p->channels, max_valid);
        int valid = (int) (ksp_divup(ksp_divup(p->channels, n), RUN) * RUN);
        l.T = (int) (ksp_divup((valid + 2 * l.edge) / RUN, 32) * 32);
        l.chunk_valid = l.T * RUN - 2 * l.edge;
        l.n_chunks = (int) ksp_divup(p->channels, l.chunk_valid);
    }
    l.ctl_bytes = ((size_t) KSP_DF_STATS * 8 + 64 + (size_t) (3 * l.S + l.G) * 4 + 1023) / 1024 * 1024;
    const int R = c.R < l.S ? c.R : l.S, RB = c.RB < 4 * l.G ? c.RB : 4 * l.G;
    l.ring_bytes = ((size_t) R * 32 * (size_t) l.dev_stride * 4 + 1023) / 1024 * 1024;
    l.bits_bytes = (size_t) RB * 32 * (size_t) l.words_stride * 4;
    return l;
}

template <int IN_MODE, int FLAG_MODE>
int df_launch(cudaStream_t s, const DfArgs &a, const CUtensorMap &tmap)
{
    auto kernel = dataflow_kernel<IN_MODE, FLAG_MODE>;
    static bool configured[64];
    int dev = 0;
    KSP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        KSP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DF_SMEM));
        configured[dev] = true;
    }
    int per_sm = 0;
    KSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DF_THREADS, DF_SMEM));
    if (per_sm < 1) return KSP_EINVAL;
    int64_t blocks = (int64_t) per_sm * ksp_sm_count();
    const int64_t items = (int64_t) a.steps * a.K;
    if (blocks > items) blocks = items;
    void *params[2] = {(void *) &a, (void *) &tmap};
    // every block is resident from the start (blocks wait on one another)
    KSP_CUDA(cudaLaunchCooperativeKernel((const void *) kernel, dim3((unsigned) blocks), dim3(DF_THREADS),
                                         params, DF_SMEM, s));
    ksp_count_launch();
    return 0;
}

}  // namespace

bool ksp_dataflow_legal(const ksp_flagger_params *p)
{
    if (p->width != 13 || p->n_windows < 1 || p->n_windows > TS_MAX_WINDOWS) return false;
    if (p->channels % 32 != 0 || p->channels < 32 || p->channels > (int64_t) 1 << 24) return false;
    if (p->baselines < 1 || p->baselines > (int64_t) 1 << 26) return false;
    return tensor_map_encoder() != nullptr;
}

// chunk_baselines: > 0 asks for the chunked four-kernel flagger, < 0 for this one wherever it is
// legal, 0 leaves the choice to the library (KSP_DATAFLOW = 0 / 1 overrides it)
bool ksp_dataflow_applies(const ksp_flagger_params *p)
{
    const DfConfig &c = df_config();
    if (p->chunk_baselines > 0 || !ksp_dataflow_legal(p)) return false;
    if (p->chunk_baselines < 0) return true;
    if (c.mode >= 0) return c.mode == 1;
    // Measured on B200 (profiles/r02*): 1.9 - 2.4 ms per 32768 x 8320 dump against 1.18 ms for the
    // chunked form - four 256-thread item streams per SM leave every item's serial chain (ticket,
    // dependency, load, compute, publish) exposed, where the stand-alone kernels run 4 to 9 smaller
    // blocks per SM.  So the library's own choice is the chunked form; this one is there for callers
    // that would rather spare the device-memory traffic (9 instead of 20 bytes per visibility).
    return false;
}

size_t ksp_dataflow_scratch_bytes(const ksp_flagger_params *p)
{
    const DfLayout l = df_layout(p);
    return l.ctl_bytes + l.ring_bytes + l.bits_bytes;
}

int ksp_dataflow_flagger(cudaStream_t s, const ksp_flagger_params *p, const void *vis,
                         const uint8_t *input_flags, float *noise, uint8_t *flags, void *scratch,
                         size_t scratch_bytes)
{
    const DfConfig &c = df_config();
    const DfLayout l = df_layout(p);
    if (scratch_bytes < l.ctl_bytes + l.ring_bytes + l.bits_bytes) return KSP_ESCRATCH;
    if ((uintptr_t) scratch % 256) return KSP_EALIGN;
    if ((uintptr_t) vis % (p->is_amplitude ? 4 : 8)) return KSP_EALIGN;
    if (p->flag_mode == KSP_FLAGS_FULL && p->input_flags_stride < p->baselines) return KSP_EINVAL;
    if (p->abs_mode != KSP_ABS_NUMPY && p->abs_mode != KSP_ABS_HYPOT) return KSP_EINVAL;

    DfArgs a;
    memset(&a, 0, sizeof(a));
    char *base = (char *) scratch;
    a.stats = (unsigned long long *) base;
    a.ctl = (uint32_t *) (base + (size_t) KSP_DF_STATS * 8);
    a.bg_done = a.ctl + 16;
    a.noise_done = a.bg_done + l.S;
    a.thr_done = a.noise_done + l.S;
    a.exp_done = a.thr_done + l.S;
    float *ring = (float *) (base + l.ctl_bytes);
    a.bits = (uint32_t *) (base + l.ctl_bytes + l.ring_bytes);
    a.bg.vis = vis; a.bg.out = ring; a.bg.flags = input_flags;
    a.bg.channels = p->channels; a.bg.baselines = p->baselines;
    a.bg.vis_stride = p->vis_stride; a.bg.out_stride = l.dev_stride;
    a.bg.flags_stride = p->input_flags_stride; a.bg.seg = 0;
    a.noise = noise; a.flags = flags; a.flags_stride = p->flags_stride;
    a.words_stride = l.words_stride; a.flag_value = p->flag_value;
    a.n_windows = p->n_windows; a.n_sigma = p->n_sigma;
    for (int w = 0; w < TS_MAX_WINDOWS; w++) a.scales[w] = w < p->n_windows ? p->scales[w] : 0.0;
    a.T = l.T; a.n_chunks = l.n_chunks; a.chunk_valid = l.chunk_valid; a.edge = l.edge;
    a.S = l.S; a.G = l.G;
    a.n_bg = (int) ksp_divup(p->channels, BG_TC);
    a.n_exp = (int) ksp_divup(p->channels / 32, 8);
    a.n_expq = (a.n_exp + 3) / 4;
    a.K = a.n_bg + 32 + 32 + a.n_expq;
    a.L1 = c.L1; a.L2 = c.L2; a.L3 = c.L3;
    a.R = c.R < l.S ? c.R : l.S;
    a.RB = c.RB < 4 * l.G ? c.RB : 4 * l.G;
    a.steps = a.L3 + 4 * l.G;
    if ((int64_t) a.steps * a.K >= 0xfff00000ll) return KSP_ETOOLARGE;
    a.total = (uint32_t) a.steps * (uint32_t) a.K;

    // deviations ring viewed as [rows][channels / 32][32] floats; box = one span of T runs,
    // 128-byte swizzle, zero fill outside the band
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    {
        const cuuint64_t dims[3] = {(cuuint64_t) RUN, (cuuint64_t) (p->channels / RUN), (cuuint64_t) a.R * 32};
        const cuuint64_t strides[2] = {RUN * sizeof(float), (cuuint64_t) l.dev_stride * sizeof(float)};
        const cuuint32_t box[3] = {RUN, (cuuint32_t) l.T, 1};
        const cuuint32_t elem[3] = {1, 1, 1};
        CUresult rc = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *) ring, dims,
                                           strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return KSP_EINVAL;
    }
    KSP_CUDA(cudaMemsetAsync(scratch, 0, l.ctl_bytes, s));
    const int in_mode = p->is_amplitude ? IN_AMP : (p->abs_mode == KSP_ABS_NUMPY ? IN_NUMPY : IN_HYPOT);
#define KSP_DF_CASE(IM, FM) \
    if (in_mode == IM && p->flag_mode == FM) return df_launch<IM, FM>(s, a, tmap);
    KSP_DF_CASE(IN_NUMPY, KSP_FLAGS_NONE)
    KSP_DF_CASE(IN_NUMPY, KSP_FLAGS_CHANNEL)
    KSP_DF_CASE(IN_NUMPY, KSP_FLAGS_FULL)
    KSP_DF_CASE(IN_HYPOT, KSP_FLAGS_NONE)
    KSP_DF_CASE(IN_HYPOT, KSP_FLAGS_CHANNEL)
    KSP_DF_CASE(IN_HYPOT, KSP_FLAGS_FULL)
    KSP_DF_CASE(IN_AMP, KSP_FLAGS_NONE)
    KSP_DF_CASE(IN_AMP, KSP_FLAGS_CHANNEL)
    KSP_DF_CASE(IN_AMP, KSP_FLAGS_FULL)
#undef KSP_DF_CASE
    return KSP_EINVAL;
}

int ksp_dataflow_stats(cudaStream_t s, const void *scratch, unsigned long long *out, int n)
{
    unsigned long long buf[KSP_DF_STATS + 1];
    KSP_CUDA(cudaMemcpyAsync(buf, scratch, sizeof(buf), cudaMemcpyDeviceToHost, s));
    KSP_CUDA(cudaStreamSynchronize(s));
    const uint32_t *ctl = (const uint32_t *) &buf[KSP_DF_STATS];
    buf[KSP_DF_STAT_ERROR] = ctl[1];
    for (int i = 0; i < n && i < KSP_DF_STATS; i++) out[i] = buf[i];
    return ctl[1] ? KSP_ETIMEOUT : 0;
}
