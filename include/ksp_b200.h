/*
 * ksp_b200.h -- C ABI of the B200-native katsdpsigproc RFI-flagging hot path.
 *
 * This is the drop-in boundary.  The reference has no native library on this
 * path: every kernel is a Mako template JIT-compiled through PyCUDA/PyOpenCL
 * (reference src/katsdpsigproc/accel.py:165-208, cuda.py:182-187) and launched
 * with CommandQueue.enqueue_kernel (cuda.py:442-459).  Each entry point below
 * replaces one such (template, enqueue_kernel) pair, or one group of PyCUDA
 * runtime calls, and is what the reference-side ctypes binding in
 * INTEGRATION.md binds.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 *   - every launch is asynchronous on `stream`, never synchronises, never
 *     allocates: scratch memory is an explicit argument (the reference's rule
 *     that temporaries are slots, rfi/device.py:1081-1091, fft.py:351-355).
 *   - strides are in ELEMENTS of the array they describe.
 *   - return value: 0 on success; > 0 is a cudaError_t; < 0 is one of KSP_E*.
 *   - element (channel c, baseline b) of a channel-major array is p[c*stride+b];
 *     of a baseline-major ("transposed", suffix _t) array it is p[b*stride+c].
 */
#ifndef KSP_B200_H
#define KSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSP_ABI_VERSION 1

/* argument errors (negative return values) */
#define KSP_EINVAL (-1)      /* bad size / null pointer / unsupported combination */
#define KSP_EALIGN (-2)      /* pointer or stride not aligned as documented */
#define KSP_ETOOLARGE (-3)   /* exceeds a documented limit (channels, width, windows) */
#define KSP_ESCRATCH (-4)    /* scratch buffer too small */
#define KSP_ENOJIT (-5)      /* the run-time compiler (NVRTC) is not available */
#define KSP_EJIT (-6)        /* run-time compilation failed: see ksp_jit_log() */
#define KSP_ETIMEOUT (-7)    /* the dataflow flagger abandoned a launch (see ksp_flagger_stats) */

/* complex64 amplitude rule: which host np.abs the result equals */
#define KSP_ABS_NUMPY 0      /* numpy on AVX-512F hosts: L*sqrt(fma(r,r,1)), r = min/max */
#define KSP_ABS_HYPOT 1      /* correctly rounded hypot (numpy without AVX-512F) */

/* BackgroundFlags (reference rfi/device.py:40-46) */
#define KSP_FLAGS_NONE 0
#define KSP_FLAGS_CHANNEL 1
#define KSP_FLAGS_FULL 2

#define KSP_MAX_WINDOWS 11   /* sum-threshold window sizes 1 .. 2^(n-1): up to 64 (n <= 7) on the
                              * fast kernel, 128 .. 1024 on a general one */
#define KSP_MAX_WIDTH 63     /* median filter width (odd) */

int ksp_abi_version(void);
/* Text for a return code of any function below (static storage). */
const char *ksp_error_string(int code);

/* ------------------------------------------------------------------------
 * Runtime shims: replace the PyCUDA calls made by reference cuda.py
 * (Device :86-158, Context :163-255, CommandQueue :257-479, Event :71-84).
 * ---------------------------------------------------------------------- */
int ksp_device_count(int *count);                                   /* cuda.py:152-155 */
int ksp_device_name(int device, char *buf, int buf_len);            /* cuda.py:98-100 */
int ksp_device_pci_bus_id(int device, char *buf, int buf_len);      /* "0000:1b:00.0"; names the SAME GPU to
                                                                     * NVML, whose indices ignore CUDA_VISIBLE_DEVICES */
int ksp_device_attributes(int device, int *cc_major, int *cc_minor, int *sm_count,
                          int *warp_size, size_t *total_mem, int *l2_bytes); /* :134-141 */
int ksp_device_set(int device);                                     /* Context.__enter__ :243-245 */
int ksp_device_get(int *device);
int ksp_versions(int *runtime_version, int *driver_version);        /* cuda.py:106-110 */

int ksp_malloc(void **ptr, size_t bytes);                           /* Context.allocate_raw :189-191 */
int ksp_free(void *ptr);
int ksp_host_alloc(void **ptr, size_t bytes);                       /* allocate_pinned :202-204 */
int ksp_host_free(void *ptr);

int ksp_stream_create(void **stream);                               /* CommandQueue.__init__ :257-270 */
int ksp_stream_destroy(void *stream);
int ksp_stream_synchronize(void *stream);                           /* finish :477-479 */
int ksp_stream_query(void *stream);                                 /* 0 = idle, cudaErrorNotReady otherwise */
int ksp_stream_wait_event(void *stream, void *event);               /* enqueue_wait_for_events :467-472 */

int ksp_event_create(void **event, int timing);                     /* enqueue_marker :461-465 */
int ksp_event_destroy(void *event);
int ksp_event_record(void *event, void *stream);
int ksp_event_synchronize(void *event);                             /* Event.wait :75-76 */
int ksp_event_elapsed_ms(void *start, void *end, float *ms);        /* Event.time_since :78-81 */

/* kind: 1 = host->device, 2 = device->host, 3 = device->device (cudaMemcpyKind) */
int ksp_memcpy_async(void *dst, const void *src, size_t bytes, int kind, void *stream);
                                                                    /* enqueue_read/write_buffer :272-289 */
int ksp_memcpy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                        size_t width_bytes, size_t height, int kind, void *stream);
                                                                    /* enqueue_*_buffer_rect :299-431 */
int ksp_memset_async(void *dst, int value, size_t bytes, void *stream); /* enqueue_zero_buffer :433-440 */


/* ------------------------------------------------------------------------
 * Kernels
 * ---------------------------------------------------------------------- */

/* Transpose.  dst[c*dst_stride + r] = src[r*src_stride + c], elem_size in
 * {1, 2, 4, 8, 16} bytes.  Replaces transpose.mako:44-73 / transpose.py:146-167. */
int ksp_transpose(void *stream, void *dst, const void *src, int64_t rows, int64_t cols,
                  int64_t dst_stride, int64_t src_stride, int elem_size);

/* Sliding-median background.  dev[c*dev_stride+b] = amp - median of the
 * usable amplitudes in channels [c-width/2, c+width/2] of baseline b (host
 * semantics: clipped windows, even counts averaged in float64, flagged/NaN
 * samples skipped and output 0).  vis is complex64 (float32 pairs), or
 * float32 amplitudes when is_amplitude.  flags: NULL, uint8[channels]
 * (KSP_FLAGS_CHANNEL) or uint8[channels*flags_stride] (KSP_FLAGS_FULL).
 * Replaces rfi/background_median_filter.mako:200-220 / rfi/device.py:311-325. */
int ksp_background_median_filter(void *stream, const void *vis, float *dev, const uint8_t *flags,
                                 int64_t channels, int64_t baselines, int64_t vis_stride,
                                 int64_t dev_stride, int64_t flags_stride, int width,
                                 int is_amplitude, int flag_mode, int abs_mode);

/* Same, but writes baseline-major deviations dev_t[b*dev_t_stride + c]
 * (background + transpose_deviations of rfi/device.py:1152-1157 in one pass). */
int ksp_background_median_filter_t(void *stream, const void *vis, float *dev_t,
                                   const uint8_t *flags, int64_t channels, int64_t baselines,
                                   int64_t vis_stride, int64_t dev_t_stride, int64_t flags_stride,
                                   int width, int is_amplitude, int flag_mode, int abs_mode);

/* noise[b] = float32(1.4826 * median{|dev| : |dev| > 0}) per baseline; NaN if no
 * such sample.  Baseline-major input, rows of any length (streamed once).
 * Replaces rfi/madnz_t.mako:72-87 / rfi/device.py:594-607. */
int ksp_madnz_t(void *stream, const float *dev_t, float *noise, int64_t channels,
                int64_t baselines, int64_t stride);

/* Same statistic on channel-major input (dev[c*stride + b]); no scratch needed
 * (four radix passes over global memory, lane == baseline).  Replaces
 * rfi/madnz.mako:105-123 / rfi/device.py:443-461. */
int ksp_madnz(void *stream, const float *dev, float *noise, int64_t channels, int64_t baselines,
              int64_t stride);

/* flags = dev > float32(n_sigma * noise[b]) ? flag_value : 0.  rows/cols are
 * those of the arrays as stored; noise is indexed by column, or by row when
 * transposed.  Replaces rfi/threshold_simple.mako:27-40, threshold_simple_t.mako:28-42. */
int ksp_threshold_simple(void *stream, const float *dev, const float *noise, uint8_t *flags,
                         int64_t rows, int64_t cols, int64_t dev_stride, int64_t flags_stride,
                         double n_sigma, int flag_value, int transposed);

/* Offringa SumThreshold along channels of baseline-major data, windows
 * 1..2^(n_windows-1).  Per-window threshold = float32((n_sigma*noise[b])*scales[w])
 * with scales[w] = falloff^-w supplied by the caller (host formula,
 * rfi/host.py:215,235).  Only windows fully inside the band are summed.
 * Replaces rfi/threshold_sum.mako:49-132 / rfi/device.py:968-987. */
int ksp_threshold_sum(void *stream, const float *dev_t, const float *noise, uint8_t *flags_t,
                      int64_t channels, int64_t baselines, int64_t dev_stride,
                      int64_t flags_stride, int n_windows, double n_sigma, const double *scales,
                      int flag_value);

/* Percentile5: dest[k*dest_stride + r], k = min, max, 25 %, 75 %, 50 % ("lower")
 * of |src[r, first_col : first_col+n_cols]|; src float32, or complex64 when
 * !is_amplitude (amplitude rule abs_mode, then pure selection).
 * Replaces percentile.mako:115-140 / percentile.py:193-209. */
int ksp_percentile5(void *stream, const void *src, float *dest, int64_t rows, int64_t src_stride,
                    int64_t dest_stride, int64_t first_col, int64_t n_cols, int is_amplitude,
                    int abs_mode);

/* MaskedSum: dest[col] = sum_rows mask[row]*src[row*src_stride+col] (complex64),
 * or of mask[row]*|src| (float32 dest) when use_amplitudes; float64 accumulation,
 * one rounding.  Replaces maskedsum.mako:38-68 / maskedsum.py:141-156. */
int ksp_maskedsum(void *stream, const void *src, const float *mask, void *dest, int64_t rows,
                  int64_t cols, int64_t src_stride, int use_amplitudes, int abs_mode);

/* ------------------------------------------------------------------------
 * Fused flagger: the standard median + MAD + SumThreshold combination of
 * rfi/device.py:1111-1166 (5 launches, 31 B/vis of memory traffic in the reference).
 *
 * Dataflow form (width 13, up to 7 window sizes, channels a multiple of 32; chosen with
 * chunk_baselines < 0 or KSP_DATAFLOW=1): ONE persistent kernel per dump.  Background tiles, noise rows, threshold
 * spans and bit -> byte expansion tiles are work items of one schedule; the deviations pass
 * from item to item through a ring of a few strips of 32 baselines that stays in the L2 cache,
 * so device memory sees the visibilities once and the flags once.
 *
 * Chunked form (everything else, or chunk_baselines > 0): four launches per chunk of
 * baselines, several chunks in flight on internal streams that fork from / join into `stream`.
 * The scratch holds the ring, or every chunk in flight.
 * ---------------------------------------------------------------------- */
typedef struct ksp_flagger_params {
    int64_t channels, baselines;
    int64_t vis_stride;          /* elements per channel row of vis */
    int64_t flags_stride;        /* bytes per channel row of flags */
    int64_t input_flags_stride;  /* KSP_FLAGS_FULL only */
    int width;                   /* median width (odd) */
    int is_amplitude;
    int flag_mode;               /* KSP_FLAGS_* */
    int abs_mode;                /* KSP_ABS_* */
    int n_windows;               /* sum-threshold window sizes 1 .. 2^(n_windows-1); >= 1 */
    int flag_value;
    double n_sigma;
    double scales[KSP_MAX_WINDOWS];
    int64_t chunk_baselines;     /* 0 = library's choice (the chunked form, one chunk per lane: KSP_LANES,
                                  * default 4; KSP_CHUNK; KSP_DATAFLOW=1: the dataflow form where legal);
                                  * > 0 = chunked form with this chunk; < 0 = dataflow form or KSP_EINVAL */
} ksp_flagger_params;

size_t ksp_flagger_scratch_bytes(const ksp_flagger_params *p);
/* baselines per chunk that ksp_flagger will use for these parameters */
int64_t ksp_flagger_chunk_baselines(const ksp_flagger_params *p);
int ksp_flagger(void *stream, const ksp_flagger_params *p, const void *vis,
                const uint8_t *input_flags, float *noise, uint8_t *flags, void *scratch,
                size_t scratch_bytes);

/* Diagnostics of the last ksp_flagger call that used `scratch` (waits for `stream`).  Dataflow
 * form: SM cycles spent per kind of work item summed over all blocks, cycles spent waiting for
 * other items, items run, noise rows that fell back to the radix select, and whether the launch
 * was abandoned (a wait exceeded its limit: returns KSP_ETIMEOUT, results are invalid).
 * Chunked form: zeros.  out[i] for i < n, indices KSP_DF_STAT_*. */
#define KSP_DF_STATS 8
enum { KSP_DF_STAT_CYCLES_BG = 0, KSP_DF_STAT_CYCLES_NOISE, KSP_DF_STAT_CYCLES_THRESHOLD,
       KSP_DF_STAT_CYCLES_EXPAND, KSP_DF_STAT_CYCLES_WAIT, KSP_DF_STAT_ITEMS,
       KSP_DF_STAT_FALLBACKS, KSP_DF_STAT_ERROR };
int ksp_flagger_stats(void *stream, const ksp_flagger_params *p, const void *scratch,
                      unsigned long long *out, int n);
/* 1 if ksp_flagger runs these parameters in the dataflow form, else 0 */
int ksp_flagger_is_dataflow(const ksp_flagger_params *p);

/* ------------------------------------------------------------------------
 * 2-D flagger: SumThresholdFlagger of rfi/twodflag.py:894-1118 (numba, CPU only in the reference)
 * for a (time, freq, baseline) block of visibilities or magnitudes, baseline fastest.  The
 * caller conditions the parameters as the reference's Python does (twodflag.py:951-1026): window
 * lists clipped and de-duplicated, rho ** log2(window) per window, box radii of the Gaussian
 * approximations per background iteration, frequency-chunk boundaries in averaged channels.
 * Flags are identical to the reference's (the float64 running sums keep its order).
 * ---------------------------------------------------------------------- */
#define KSP_TWOD_MAX_WINDOWS 16      /* window sizes per axis */
#define KSP_TWOD_MAX_WINDOW 64       /* largest window size */
#define KSP_TWOD_MAX_CHUNKS 64       /* frequency chunks */
#define KSP_TWOD_MAX_ITERATIONS 63   /* background iterations */
typedef struct ksp_twodflag_params {
    int64_t n_time, n_freq, n_bl;    /* shape of data / in_flags / out_flags, C order */
    int is_complex;                  /* data is complex64 (else float32 magnitudes) */
    int average_freq;                /* channels averaged before flagging */
    int n_windows_time, n_windows_freq;
    int windows_time[KSP_TWOD_MAX_WINDOWS], windows_freq[KSP_TWOD_MAX_WINDOWS];
    double tf_time[KSP_TWOD_MAX_WINDOWS], tf_freq[KSP_TWOD_MAX_WINDOWS];   /* pow(rho, log2(window)) */
    double outlier_nsigma, background_reject;
    int background_iterations;
    int r_time[KSP_TWOD_MAX_ITERATIONS + 1];   /* box radius for extend factor e = 1 .. iterations */
    int r_freq[KSP_TWOD_MAX_ITERATIONS + 1];   /* (index 0 unused): int(0.5 sqrt(12 (e sigma)^2 / 4 + 1)) */
    int time_extend, freq_extend;
    int n_chunks;
    int64_t chunk_ends[KSP_TWOD_MAX_CHUNKS + 1];   /* 0 = chunk_ends[0] <= ... <= chunk_ends[n_chunks] = averaged channels */
    double flag_all_time_frac, flag_all_freq_frac;
} ksp_twodflag_params;

/* Scratch for `batch_baselines` baselines in flight (0: bad parameters). */
size_t ksp_twodflag_scratch_bytes(const ksp_twodflag_params *p, int64_t batch_baselines);
/* data: complex64 or float32; in_flags: uint8, non-zero = ignore; out_flags: uint8 0 / 1. */
int ksp_twodflag(void *stream, const ksp_twodflag_params *p, const void *data, const uint8_t *in_flags,
                 uint8_t *out_flags, void *scratch, size_t scratch_bytes, int64_t batch_baselines);

/* Baselines the device works on at the same time: the batch size that fills it. */
int ksp_twodflag_resident_baselines(void);
/* Developer aid: SM clock cycles the first block spent in each phase of the ksp_twodflag launches
 * since the last reset (synchronous; indices documented in csrc/twodflag.cu). */
#define KSP_TWOD_PHASES 20
int ksp_twodflag_phases(unsigned long long *out, int n, int reset);

/* ------------------------------------------------------------------------
 * General-purpose operations that sit beside the RFI ones in the reference.
 * ---------------------------------------------------------------------- */
/* Fill.  Replaces fill.mako (reference fill.py:130-139): data[i] = *value for
 * i < elements, where an element is elem_size (1, 2, 4, 8 or 16) opaque bytes. */
int ksp_fill(void *stream, void *data, size_t elements, const void *value, size_t elem_size);

/* HReduce.  Replaces hreduce.mako (reference reduce.py:72-89, 199-214): for every row,
 * dest[row] = op-reduction of src[row, first_col .. first_col + n_cols).  As in the
 * reference the element type and the combining step are C source text (`ctype`, e.g.
 * "unsigned int"; `op`, an expression in a and b, e.g. "a + b"; `identity`, e.g. "0";
 * optional `extra_code` pasted in front), compiled when the template is created - here with
 * NVRTC for sm_100a, loaded lazily with dlopen.  KSP_ENOJIT if NVRTC cannot be loaded,
 * KSP_EJIT if the text does not compile (ksp_jit_log() returns the compiler output of the
 * calling thread's last ksp_hreduce_create).  op must be commutative and associative. */
int ksp_hreduce_create(const char *ctype, const char *op, const char *identity,
                       const char *extra_code, size_t elem_size, void **handle);
int ksp_hreduce(void *stream, void *handle, const void *src, void *dest, int64_t rows,
                int64_t src_stride, int64_t first_col, int64_t n_cols);
int ksp_hreduce_destroy(void *handle);
const char *ksp_jit_log(void);

/* ------------------------------------------------------------------------
 * Introspection (no reference equivalent; used by bench.py and the tests).
 * ---------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process so far. */
int ksp_kernel_launch_count(unsigned long long *count);

/* Per-stage timing of ksp_flagger.  While enabled, ksp_flagger records CUDA events
 * on its own stream around every stage launch; ksp_profile_read waits for them,
 * sums the elapsed milliseconds and launch counts per stage (0 background,
 * 1 noise, 2 threshold, 3 flag expansion; n_stages >= 4) and resets the record.
 * Off by default: the timed path records no events. */
#define KSP_N_STAGES 4
int ksp_profile_enable(int on);
int ksp_profile_read(double *stage_ms, int *stage_launches, int n_stages);

/* Rows for which ksp_madnz_t / ksp_percentile5 had to redo the selection with the slow
 * radix select because the sampled bracket missed the wanted rank (expected: well below
 * 1 % of rows).  Waits for `stream`; `reset` != 0 zeroes the counter afterwards. */
int ksp_selection_fallback_count(void *stream, unsigned long long *count, int reset);

#ifdef __cplusplus
}
#endif
#endif /* KSP_B200_H */
