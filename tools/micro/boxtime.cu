// The time-axis / frequency-axis box passes of csrc/twodflag.cu in isolation (developer tool):
// blocks x 128 threads, each on its own T x F arrays, cycles per pass of block 0.
#include "../../katsdpsigproc_b200/csrc/twodflag.cu"
#include <cstdio>
#include <cstdlib>
int ksp_sm_count() { return 148; }
int ksp_l2_bytes() { return 126 << 20; }
void ksp_count_launch() {}
bool ksp_profile_active() { return false; }
void ksp_profile_begin(int, cudaStream_t) {}
void ksp_profile_end(int, cudaStream_t) {}

__global__ void __launch_bounds__(TD_THREADS, TD_BLOCKS_PER_SM)
bench_kernel(const float *data, const uint8_t *flags, float *weight, float *out, int T, int F, int r_t, int r_f,
             long long *cycles)
{
    const size_t A = (size_t) T * F, o = blockIdx.x * A;
    long long t0 = clock64();
    box_time_pass(data + o, flags + o, weight + o, out + o, T, F, r_t);
    long long t1 = clock64();
    box_freq_pass<false>(data + o, flags + o, weight + o, out + o, T, F, r_f);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t1;
    }
}

int main(int argc, char **argv)
{
    const int T = 16, F = 4096, r_t = 10, r_f = 9;
    const size_t A = (size_t) T * F;
    for (int blocks : {148, 148 * 4, 148 * 7, 148 * 8}) {
        float *data, *weight, *out; uint8_t *flags; long long *cycles;
        cudaMalloc(&data, blocks * A * 4); cudaMalloc(&weight, blocks * A * 4); cudaMalloc(&out, blocks * A * 4);
        cudaMalloc(&flags, blocks * A); cudaMallocManaged(&cycles, 16);
        float *h = (float *) malloc(blocks * A * 4); uint8_t *hf = (uint8_t *) malloc(blocks * A);
        for (size_t i = 0; i < blocks * A; i++) { h[i] = 4.0f + (rand() % 1000) * 1e-3f; hf[i] = rand() % 50 == 0; }
        cudaMemcpy(data, h, blocks * A * 4, cudaMemcpyHostToDevice); cudaMemcpy(flags, hf, blocks * A, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; rep++) {
            bench_kernel<<<blocks, TD_THREADS>>>(data, flags, weight, out, T, F, r_t, r_f, cycles);
            cudaDeviceSynchronize();
        }
        printf("%5d blocks: time pass %9lld cycles (%6.1f per line-round iteration), freq pass %9lld (%6.1f per iteration)  %s\n",
               blocks, cycles[0], cycles[0] / (128.0 * (T + 4 * r_t + 3)), cycles[1], cycles[1] / (double) (F + 4 * r_f + 3),
               cudaGetErrorString(cudaGetLastError()));
        unsigned long long ph[20]; cudaMemcpyFromSymbol(ph, td_phase, sizeof(ph));
        printf("      freq pass, both reps: request %llu, request + wait %llu, compute %llu, flush %llu\n", ph[15], ph[14], ph[19], ph[3]);
        unsigned long long z[20] = {0}; cudaMemcpyToSymbol(td_phase, z, sizeof(z));
        cudaFree(data); cudaFree(weight); cudaFree(out); cudaFree(flags); cudaFree(cycles); free(h); free(hf);
    }
    return 0;
}
