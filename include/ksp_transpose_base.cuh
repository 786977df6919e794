// ksp_transpose_base.cuh -- tools for kernels that transpose AND do something else.
//
// The CUDA counterpart of reference transpose_base.mako:34-137 (transpose_data_class,
// transpose_coords_class, transpose_load, transpose_store): a square block of BLOCK x BLOCK
// threads moves a tile of (BLOCK * VTY) rows x (BLOCK * VTX) columns through shared memory; the
// caller supplies the two bodies - what to load into the tile and what to do with the
// transposed element - so a kernel can fuse its own arithmetic, masking or type conversion
// with the transposition instead of paying for a separate pass.  Header only; include it in
// any .cu file (nvcc -I<repo>/include).  ksp_transpose's 2- and 16-byte paths are built from
// it (katsdpsigproc_b200/csrc/transpose.cu), which is also the usage example:
//
//     using Tile = ksp::TransposeTile<float, 16, 2, 2>;
//     __global__ void __launch_bounds__(16 * 16) my_kernel(float *out, const float2 *in, int rows, int cols,
//                                                          int out_stride, int in_stride)
//     {
//         __shared__ Tile::Values values;
//         Tile::Coords at;
//         Tile::init_simple(at);
//         Tile::load(at, [&](int r, int c, int lr, int lc) {          // input element (r, c)
//             if (r < rows && c < cols) values.arr[lr][lc] = hypotf(in[r * in_stride + c].x, in[r * in_stride + c].y);
//         });
//         __syncthreads();
//         Tile::store(at, [&](int r, int c, int lr, int lc) {         // output element (r, c) = input (c, r)
//             if (r < cols && c < rows) out[r * out_stride + c] = values.arr[lr][lc];
//         });
//     }
//     // launch: grid (ceil(cols / Tile::COLS), ceil(rows / Tile::ROWS)), block (16, 16)
//
// Both bodies are called with lx fastest: consecutive threads touch consecutive columns of the
// input in load() and consecutive columns of the OUTPUT in store(), so both sides coalesce.
#ifndef KSP_TRANSPOSE_BASE_CUH
#define KSP_TRANSPOSE_BASE_CUH

namespace ksp {

template <typename T, int BLOCK, int VTX = 1, int VTY = 1>
struct TransposeTile {
    static constexpr int ROWS = BLOCK * VTY;     // input rows per tile
    static constexpr int COLS = BLOCK * VTX;     // input columns per tile
    // the row pitch is padded so that column-wise reads hit distinct 4-byte banks
    static constexpr int PAD = sizeof(T) > 4 ? 1 : 4 / (int) sizeof(T);

    /// The tile itself: declare one in shared memory.
    struct Values {
        T arr[ROWS][COLS + PAD];
    };

    /// Where this thread and its block are (registers).
    struct Coords {
        int lx, ly;         ///< thread within the block, lx fastest
        int in_row0;        ///< first input row of the block's tile
        int in_col0;        ///< first input column of the block's tile
    };

    /// Tiles are assigned to blocks diagonally (reference transpose_base.mako:74-76), which
    /// keeps concurrently running blocks off the same memory partitions.
    static __device__ __forceinline__ void init(Coords &at, int local_x, int local_y, int block_x,
                                                int block_y, int blocks_y)
    {
        at.lx = local_x;
        at.ly = local_y;
        at.in_row0 = (block_x + block_y) % blocks_y * ROWS;
        at.in_col0 = block_x * COLS;
    }

    /// From the launch geometry: block (BLOCK, BLOCK), grid (column tiles, row tiles).
    static __device__ __forceinline__ void init_simple(Coords &at)
    {
        init(at, (int) threadIdx.x, (int) threadIdx.y, (int) blockIdx.x, (int) blockIdx.y, (int) gridDim.y);
    }

    /// body(r, c, lr, lc): input element (r, c) belongs in values.arr[lr][lc].
    template <typename Body>
    static __device__ __forceinline__ void load(const Coords &at, Body body)
    {
#pragma unroll
        for (int y = 0; y < VTY; y++)
#pragma unroll
            for (int x = 0; x < VTX; x++)
                body(at.in_row0 + y * BLOCK + at.ly, at.in_col0 + x * BLOCK + at.lx, at.ly + y * BLOCK,
                     at.lx + x * BLOCK);
    }

    /// body(r, c, lr, lc): output element (r, c) - input element (c, r) - is values.arr[lr][lc].
    template <typename Body>
    static __device__ __forceinline__ void store(const Coords &at, Body body)
    {
#pragma unroll
        for (int y = 0; y < VTX; y++)
#pragma unroll
            for (int x = 0; x < VTY; x++)
                body(at.in_col0 + y * BLOCK + at.ly, at.in_row0 + x * BLOCK + at.lx, at.lx + x * BLOCK,
                     at.ly + y * BLOCK);
    }
};

}  // namespace ksp

#endif  // KSP_TRANSPOSE_BASE_CUH
