// A downstream kernel built with include/ksp_transpose_base.cuh: amplitude of a complex64 array,
// masked by per-row weights, written TRANSPOSED as float32 - load body, store body and nothing
// else (the use the reference documents for transpose_base.mako, doc/user/macros.rst).
// Prints a checksum of the result; tests/test_fused_transpose.py compares with numpy.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ksp_transpose_base.cuh"

using Tile = ksp::TransposeTile<float, 16, 2, 2>;

__global__ void __launch_bounds__(16 * 16)
masked_amplitude_t(float *out, const float2 *in, const float *weight, int rows, int cols,
                   int out_stride, int in_stride)
{
    __shared__ Tile::Values values;
    Tile::Coords at;
    Tile::init_simple(at);
    Tile::load(at, [&](int r, int c, int lr, int lc) {
        if (r < rows && c < cols) {
            const float2 v = in[(size_t) r * in_stride + c];
            values.arr[lr][lc] = weight[r] * sqrtf(v.x * v.x + v.y * v.y);
        }
    });
    __syncthreads();
    Tile::store(at, [&](int r, int c, int lr, int lc) {
        if (r < cols && c < rows) out[(size_t) r * out_stride + c] = values.arr[lr][lc];
    });
}

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 53, cols = argc > 2 ? atoi(argv[2]) : 81;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    const int in_stride = cols + 3, out_stride = rows + 5;
    std::vector<float2> in((size_t) rows * in_stride);
    std::vector<float> weight(rows), out((size_t) cols * out_stride, -1.0f);
    for (int r = 0; r < rows; r++) {
        weight[r] = (r % 7 == 3) ? 0.0f : 1.0f + 0.25f * (float) (r % 3);
        for (int c = 0; c < cols; c++)
            in[(size_t) r * in_stride + c] = make_float2((float) ((r * 31 + c * 17) % 23) - 11.0f,
                                                         (float) ((r * 13 + c * 7) % 19) - 9.0f);
    }
    float2 *d_in;
    float *d_w, *d_out;
    cudaMalloc(&d_in, in.size() * sizeof(float2));
    cudaMalloc(&d_w, weight.size() * sizeof(float));
    cudaMalloc(&d_out, out.size() * sizeof(float));
    cudaMemcpy(d_in, in.data(), in.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, weight.data(), weight.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_out, out.data(), out.size() * sizeof(float), cudaMemcpyHostToDevice);
    dim3 grid((cols + Tile::COLS - 1) / Tile::COLS, (rows + Tile::ROWS - 1) / Tile::ROWS);
    masked_amplitude_t<<<grid, dim3(16, 16)>>>(d_out, d_in, d_w, rows, cols, out_stride, in_stride);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    // out[c][r] for every (c, r), row-major, then the padding must be untouched
    for (int c = 0; c < cols; c++) {
        for (int r = 0; r < rows; r++) printf("%.9g ", out[(size_t) c * out_stride + r]);
        for (int r = rows; r < out_stride; r++)
            if (out[(size_t) c * out_stride + r] != -1.0f) {
                fprintf(stderr, "padding overwritten\n");
                return 2;
            }
    }
    printf("\n");
    return 0;
}
