// Where a row's time goes in madnz_stream_kernel (developer tool): cycles of thread 0 of every
// block per phase, summed over the rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMS_PHASE_CLOCK \
//        -o tools/micro/madnz_phases tools/micro/madnz_phases.cu
#include "../../katsdpsigproc_b200/csrc/madnz.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
int ksp_sm_count() { return 148; }
int ksp_l2_bytes() { return 126 << 20; }
void ksp_count_launch() {}
bool ksp_profile_active() { return false; }
void ksp_profile_begin(int, cudaStream_t) {}
void ksp_profile_end(int, cudaStream_t) {}

int main(int argc, char **argv)
{
    const int C = 32768, B = argc > 1 ? atoi(argv[1]) : 8320;
    float *dev_t, *noise;
    cudaMalloc(&dev_t, (size_t) C * B * 4);
    cudaMalloc(&noise, B * 4);
    std::vector<float> h((size_t) C * 64);
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.0f, 1.0f);
    for (auto &x : h) x = nd(rng);
    for (int b = 0; b < B; b += 64)
        cudaMemcpy(dev_t + (size_t) b * C, h.data(), (size_t) C * std::min(64, B - b) * 4, cudaMemcpyHostToDevice);
    const char *names[6] = {"zero + sample + histogram", "bracket", "pass over the row", "list histogram",
                            "locate + second walk", "sort + result"};
    for (int rep = 0; rep < 3; rep++) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(ms_phase, z, sizeof(z));
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        int rc = ksp_madnz_t(nullptr, dev_t, noise, C, B, C);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        unsigned long long ph[8];
        cudaMemcpyFromSymbol(ph, ms_phase, sizeof(ph));
        double total = 0;
        for (int k = 0; k < 6; k++) total += (double) ph[k];
        printf("rep %d rc %d %.4f ms (%s); cycles per row %.0f\n", rep, rc, ms, cudaGetErrorString(cudaGetLastError()), total / B);
        for (int k = 0; k < 6; k++) printf("   %-28s %5.1f %%  %7.0f cycles per row\n", names[k], 100.0 * ph[k] / total, (double) ph[k] / B);
    }
    return 0;
}
