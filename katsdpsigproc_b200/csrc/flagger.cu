// Fused flagger: the standard median + MAD + SumThreshold combination of
// reference rfi/device.py:1111-1166 (5 launches, 31 B/vis of HBM traffic) as
// 4 launches per CHUNK of baselines whose intermediates stay in L2:
//
//   vis[:, chunk] --bg13_t--> dev_t (chunk x C float32, scratch)
//                 --madnz_t--> noise[chunk]
//                 --threshold_sum (packed)--> bits_t (chunk x C/32 words, scratch)
//                 --expand_flags--> flags[:, chunk]
//
// Compulsory HBM traffic is then the 8 B/vis read of vis and the 1 B/vis write
// of flags; the 4 + 4 B/vis of dev_t and the 1/8 + 1/8 B/vis of bits_t are
// written and re-read while still resident in the 126 MB L2 (the scratch is
// reused chunk after chunk, so its lines are overwritten before eviction).
#include "common.cuh"
#include <stdlib.h>

int ksp_threshold_sum_packed(cudaStream_t s, const float *dev_t, const float *noise,
                             uint32_t *bits_t, int64_t channels, int64_t baselines,
                             int64_t dev_stride, int64_t words_stride, int n_windows,
                             double n_sigma, const double *scales);
int ksp_expand_flags(cudaStream_t s, const uint32_t *bits_t, uint8_t *flags, int64_t channels,
                     int64_t baselines, int64_t words_stride, int64_t flags_stride, int flag_value);

namespace {

struct Layout {
    int64_t chunk;        // baselines per chunk (multiple of 32)
    int64_t dev_stride;   // floats per baseline row of dev_t
    int64_t words_stride; // words per baseline row of bits_t
    size_t dev_bytes, bits_bytes;
};

Layout make_layout(const ksp_flagger_params *p)
{
    Layout l;
    l.dev_stride = ksp_divup(p->channels, 32) * 32;
    l.words_stride = ksp_divup(ksp_divup(p->channels, 32), 4) * 4;
    const int64_t row_bytes = (l.dev_stride + l.words_stride) * 4;
    int64_t chunk = p->chunk_baselines;
    if (chunk <= 0) {
        // Measured on B200 (profiles/): the stages are issue-bound, not HBM-bound, so launch
        // tails cost more than the L2 misses that large chunks cause; take chunks of 32
        // baselines per SM (a 0.6 GB scratch at 32768 channels) unless KSP_CHUNK says otherwise.
        const char *e = getenv("KSP_CHUNK");
        chunk = e ? atoll(e) : 0;
        if (chunk <= 0) chunk = 32 * (int64_t) ksp_sm_count();
        (void) row_bytes;
    }
    chunk = (chunk / 32) * 32;
    if (chunk < 32) chunk = 32;
    const int64_t bl_pad = ksp_divup(p->baselines, 32) * 32;
    if (chunk > bl_pad) chunk = bl_pad > 0 ? bl_pad : 32;
    l.chunk = chunk;
    l.dev_bytes = (size_t) chunk * (size_t) l.dev_stride * 4;
    l.bits_bytes = (size_t) chunk * (size_t) l.words_stride * 4;
    return l;
}

}  // namespace

extern "C" size_t ksp_flagger_scratch_bytes(const ksp_flagger_params *p)
{
    if (!p || p->channels <= 0 || p->baselines <= 0) return 0;
    Layout l = make_layout(p);
    return l.dev_bytes + l.bits_bytes;
}

extern "C" int64_t ksp_flagger_chunk_baselines(const ksp_flagger_params *p)
{
    if (!p || p->channels <= 0 || p->baselines <= 0) return 0;
    return make_layout(p).chunk;
}

extern "C" int ksp_flagger(void *stream, const ksp_flagger_params *p, const void *vis,
                           const uint8_t *input_flags, float *noise, uint8_t *flags, void *scratch,
                           size_t scratch_bytes)
{
    if (!p) return KSP_EINVAL;
    if (p->channels < 0 || p->baselines < 0) return KSP_EINVAL;
    if (p->channels == 0 || p->baselines == 0) return 0;
    if (!vis || !noise || !flags || !scratch) return KSP_EINVAL;
    if (p->n_windows < 1) return KSP_EINVAL;
    if (p->flag_mode != KSP_FLAGS_NONE && !input_flags) return KSP_EINVAL;
    if (p->flags_stride < p->baselines || p->vis_stride < p->baselines) return KSP_EINVAL;
    Layout l = make_layout(p);
    if (scratch_bytes < l.dev_bytes + l.bits_bytes) return KSP_ESCRATCH;
    if ((uintptr_t) scratch % 16) return KSP_EALIGN;
    cudaStream_t s = (cudaStream_t) stream;
    float *dev_t = (float *) scratch;
    uint32_t *bits_t = (uint32_t *) ((char *) scratch + l.dev_bytes);
    const size_t vis_elem = p->is_amplitude ? 4 : 8;

    for (int64_t b0 = 0; b0 < p->baselines; b0 += l.chunk) {
        const int64_t nb = (p->baselines - b0 < l.chunk) ? p->baselines - b0 : l.chunk;
        const void *vis_c = (const char *) vis + (size_t) b0 * vis_elem;
        const uint8_t *in_fl = input_flags;
        if (p->flag_mode == KSP_FLAGS_FULL) in_fl = input_flags + b0;
        ksp_profile_begin(KSP_STAGE_BACKGROUND, s);
        int rc = ksp_background_median_filter_t(s, vis_c, dev_t, in_fl, p->channels, nb,
                                                p->vis_stride, l.dev_stride, p->input_flags_stride,
                                                p->width, p->is_amplitude, p->flag_mode,
                                                p->abs_mode);
        ksp_profile_end(KSP_STAGE_BACKGROUND, s);
        if (rc) return rc;
        ksp_profile_begin(KSP_STAGE_NOISE, s);
        rc = ksp_madnz_t(s, dev_t, noise + b0, p->channels, nb, l.dev_stride);
        ksp_profile_end(KSP_STAGE_NOISE, s);
        if (rc) return rc;
        ksp_profile_begin(KSP_STAGE_THRESHOLD, s);
        rc = ksp_threshold_sum_packed(s, dev_t, noise + b0, bits_t, p->channels, nb, l.dev_stride,
                                      l.words_stride, p->n_windows, p->n_sigma, p->scales);
        ksp_profile_end(KSP_STAGE_THRESHOLD, s);
        if (rc) return rc;
        ksp_profile_begin(KSP_STAGE_EXPAND, s);
        rc = ksp_expand_flags(s, bits_t, flags + b0, p->channels, nb, l.words_stride,
                              p->flags_stride, p->flag_value);
        ksp_profile_end(KSP_STAGE_EXPAND, s);
        if (rc) return rc;
    }
    return 0;
}
