// Internal interface of the dataflow flagger (dataflow.cu), used by flagger.cu.
#pragma once
#include "common.cuh"

// true: ksp_flagger runs these parameters as one persistent dataflow kernel
bool ksp_dataflow_applies(const ksp_flagger_params *p);
// false: these parameters cannot run on the dataflow kernel at all
bool ksp_dataflow_legal(const ksp_flagger_params *p);
size_t ksp_dataflow_scratch_bytes(const ksp_flagger_params *p);
int ksp_dataflow_flagger(cudaStream_t s, const ksp_flagger_params *p, const void *vis,
                         const uint8_t *input_flags, float *noise, uint8_t *flags, void *scratch,
                         size_t scratch_bytes);
int ksp_dataflow_stats(cudaStream_t s, const void *scratch, unsigned long long *out, int n);
