// 2-D transpose through a padded shared-memory tile.
//
// Replaces reference transpose.mako:44-73 / transpose_base.mako:34-137 (tile in
// local memory, diagonal block order to dodge partition camping on 2014 GPUs).
// On B200 the address hash spreads tiles over the L2 slices by itself, so the
// tile order is the plain one; what matters is that both the loads and the
// stores of a warp cover whole 32-byte sectors.
//
//   * 4-byte elements of 16-byte aligned arrays: 64x64 tile moved with 128-bit loads and
//     stores (transpose_words128_kernel).
//   * otherwise 4/8-byte elements: 32x32-element tile, 32x8 threads, 4 rows per thread;
//     a warp reads 128/256 contiguous bytes and writes the same.
//   * 1/2-byte elements (flags): 128x32... handled by the BYTES variant below: a
//     64x64-element tile moved as 32-bit words (16 lanes per 64-byte row), and
//     re-packed from shared memory so that the stores are 32-bit words too.
#include "common.cuh"
#include "../../include/ksp_transpose_base.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
transpose_tile_kernel(T *__restrict__ dst, const T *__restrict__ src, int64_t rows, int64_t cols,
                      int64_t dst_stride, int64_t src_stride)
{
    __shared__ T tile[32][33];
    const int64_t c0 = (int64_t) blockIdx.x * 32;
    const int64_t r0 = (int64_t) blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * k][tx] = src[r * src_stride + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t c = c0 + ty + 8 * k, r = r0 + tx;  // dst row = src column
        if (c < cols && r < rows) dst[c * dst_stride + r] = tile[tx][ty + 8 * k];
    }
}

// Byte transpose: 64 (src rows) x 64 (src cols) tile per block of 256 threads.
// Loads and stores are 32-bit words when the geometry allows, else bytes.
__global__ void __launch_bounds__(256)
transpose_bytes_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t rows,
                       int64_t cols, int64_t dst_stride, int64_t src_stride, int vec_ok)
{
    __shared__ uint32_t tile[64][17];  // 64 rows x 64 bytes, pitch 68 bytes
    const int64_t c0 = (int64_t) blockIdx.x * 64;
    const int64_t r0 = (int64_t) blockIdx.y * 64;
    const int t = threadIdx.x;
    const int lane16 = t & 15, rowq = t >> 4;  // 16 words per row, 16 rows per pass
    const bool full = vec_ok && (r0 + 64 <= rows) && (c0 + 64 <= cols);
    if (full) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int r = rowq + 16 * k;
            tile[r][lane16] =
                *reinterpret_cast<const uint32_t *>(src + (r0 + r) * src_stride + c0 + 4 * lane16);
        }
    } else {
        uint8_t *tb = reinterpret_cast<uint8_t *>(&tile[0][0]);
        for (int i = t; i < 64 * 64; i += 256) {
            int r = i >> 6, c = i & 63;
            uint8_t v = 0;
            if (r0 + r < rows && c0 + c < cols) v = src[(r0 + r) * src_stride + c0 + c];
            tb[r * 68 + c] = v;
        }
    }
    __syncthreads();
    const uint8_t *tb = reinterpret_cast<const uint8_t *>(&tile[0][0]);
    if (full) {
        // output row = src column c (0..63); 16 words of 4 src rows each
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int c = rowq + 16 * k;
            int r = 4 * lane16;
            uint32_t w = (uint32_t) tb[(r + 0) * 68 + c] | ((uint32_t) tb[(r + 1) * 68 + c] << 8) |
                         ((uint32_t) tb[(r + 2) * 68 + c] << 16) |
                         ((uint32_t) tb[(r + 3) * 68 + c] << 24);
            *reinterpret_cast<uint32_t *>(dst + (c0 + c) * dst_stride + r0 + r) = w;
        }
    } else {
        for (int i = t; i < 64 * 64; i += 256) {
            int c = i >> 6, r = i & 63;
            if (r0 + r < rows && c0 + c < cols) dst[(c0 + c) * dst_stride + r0 + r] = tb[r * 68 + c];
        }
    }
}

// Byte transpose, whole 128 x 128-byte tiles of 16-byte aligned arrays: 128-bit loads (a warp
// reads 4 rows of 128 contiguous bytes), the tile kept as 32-bit words with an XOR swizzle
// (word w of row r at w ^ 4 (r / 16): no padding, no bank conflicts either way), every thread
// transposes a block of 16 rows x 4 columns in registers (4 x 4 byte transposes with PRMT) and
// writes it as 4 rows of 16 bytes - 8 consecutive lanes fill 128 contiguous bytes of a row.
__global__ void __launch_bounds__(256)
transpose_bytes128_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                          int64_t dst_stride, int64_t src_stride)
{
    __shared__ __align__(16) uint32_t tile[128 * 32];
    const int64_t c0 = (int64_t) blockIdx.x * 128;
    const int64_t r0 = (int64_t) blockIdx.y * 128;
    const int t = threadIdx.x;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = t + 256 * k, r = q >> 3, ch = q & 7;
        v[k] = __ldcs(reinterpret_cast<const uint4 *>(src + (r0 + r) * src_stride + c0 + 16 * ch));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = t + 256 * k, r = q >> 3, ch = q & 7;
        *reinterpret_cast<uint4 *>(&tile[r * 32 + ((4 * ch) ^ (4 * ((r >> 4) & 7)))]) = v[k];
    }
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    const int rg = lane & 7, cw = 4 * warp + (lane >> 3);     // 16 rows 16 rg .., columns 4 cw .. 4 cw + 3
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = tile[(16 * rg + i) * 32 + (cw ^ (4 * rg))];
    uint32_t o[4][4];                                         // o[k][g]: column byte k of rows 4 g .. 4 g + 3
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t a01 = __byte_perm(w[4 * g], w[4 * g + 1], 0x5140);
        const uint32_t b01 = __byte_perm(w[4 * g], w[4 * g + 1], 0x7362);
        const uint32_t a23 = __byte_perm(w[4 * g + 2], w[4 * g + 3], 0x5140);
        const uint32_t b23 = __byte_perm(w[4 * g + 2], w[4 * g + 3], 0x7362);
        o[0][g] = __byte_perm(a01, a23, 0x5410);
        o[1][g] = __byte_perm(a01, a23, 0x7632);
        o[2][g] = __byte_perm(b01, b23, 0x5410);
        o[3][g] = __byte_perm(b01, b23, 0x7632);
    }
    uint8_t *out = dst + (c0 + 4 * cw) * dst_stride + r0 + 16 * rg;
#pragma unroll
    for (int k = 0; k < 4; k++)
        __stcs(reinterpret_cast<uint4 *>(out + k * dst_stride), make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]));
}

// 4-byte elements of 16-byte aligned arrays: 64 x 64 tile, 128-bit loads (a warp reads two rows
// of 256 contiguous bytes), the tile stored as 16-byte chunks with an XOR swizzle (chunk ch of
// row r at ch ^ (r / 4 & 7): conflict-free both ways, no padding), every thread takes a 4 x 4
// block as four 128-bit shared loads - the transposition of the block is only a renaming of
// registers - and writes four 128-bit rows, 16 lanes filling 256 contiguous bytes.  Ragged
// tiles at the right and bottom edges go element by element through the same tile.
__global__ void __launch_bounds__(256)
transpose_words128_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int64_t rows,
                          int64_t cols, int64_t dst_stride, int64_t src_stride)
{
    __shared__ __align__(16) uint32_t tile[64 * 64];
    const int64_t c0 = (int64_t) blockIdx.x * 64;
    const int64_t r0 = (int64_t) blockIdx.y * 64;
    const int t = threadIdx.x;
    if (r0 + 64 <= rows && c0 + 64 <= cols) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int q = t + 256 * k, r = q >> 4, ch = q & 15;
            v[k] = __ldcs(reinterpret_cast<const uint4 *>(src + (r0 + r) * src_stride + c0 + 4 * ch));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int q = t + 256 * k, r = q >> 4, ch = q & 15;
            *reinterpret_cast<uint4 *>(&tile[r * 64 + 4 * (ch ^ ((r >> 2) & 7))]) = v[k];
        }
        __syncthreads();
        const int rg = t & 15, cg = t >> 4;               // source rows 4 rg .., columns 4 cg ..
        uint4 w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            w[i] = *reinterpret_cast<const uint4 *>(&tile[(4 * rg + i) * 64 + 4 * (cg ^ (rg & 7))]);
        uint32_t *out = dst + (c0 + 4 * cg) * dst_stride + r0 + 4 * rg;
        __stcs(reinterpret_cast<uint4 *>(out), make_uint4(w[0].x, w[1].x, w[2].x, w[3].x));
        __stcs(reinterpret_cast<uint4 *>(out + dst_stride), make_uint4(w[0].y, w[1].y, w[2].y, w[3].y));
        __stcs(reinterpret_cast<uint4 *>(out + 2 * dst_stride), make_uint4(w[0].z, w[1].z, w[2].z, w[3].z));
        __stcs(reinterpret_cast<uint4 *>(out + 3 * dst_stride), make_uint4(w[0].w, w[1].w, w[2].w, w[3].w));
    } else {
        for (int i = t; i < 64 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            if (r0 + r < rows && c0 + c < cols) tile[r * 64 + ((c + r) & 63)] = src[(r0 + r) * src_stride + c0 + c];
        }
        __syncthreads();
        for (int i = t; i < 64 * 64; i += 256) {
            const int c = i >> 6, r = i & 63;
            if (r0 + r < rows && c0 + c < cols) dst[(c0 + c) * dst_stride + r0 + r] = tile[r * 64 + ((c + r) & 63)];
        }
    }
}

// Plain transposition written with the fusable tools of include/ksp_transpose_base.cuh (the
// reference's transpose.mako:44-73 is the same metakernel with these two bodies); used for the
// element sizes off the flagger's path (2 and 16 bytes).
template <typename T, int BLOCK, int VTX, int VTY>
__global__ void __launch_bounds__(BLOCK * BLOCK)
transpose_base_kernel(T *__restrict__ out, const T *__restrict__ in, int in_rows, int in_cols,
                      int64_t out_stride, int64_t in_stride)
{
    using Tile = ksp::TransposeTile<T, BLOCK, VTX, VTY>;
    __shared__ typename Tile::Values values;
    typename Tile::Coords at;
    Tile::init_simple(at);
    Tile::load(at, [&](int r, int c, int lr, int lc) {
        if (r < in_rows && c < in_cols) values.arr[lr][lc] = in[r * in_stride + c];
    });
    __syncthreads();
    Tile::store(at, [&](int r, int c, int lr, int lc) {
        if (r < in_cols && c < in_rows) out[r * out_stride + c] = values.arr[lr][lc];
    });
}

template <typename T>
int launch_base(cudaStream_t s, void *dst, const void *src, int64_t rows, int64_t cols,
                int64_t dst_stride, int64_t src_stride)
{
    constexpr int BLOCK = 16, VTX = 2, VTY = 2;
    using Tile = ksp::TransposeTile<T, BLOCK, VTX, VTY>;
    if (rows > 0x7fffffff || cols > 0x7fffffff) return KSP_ETOOLARGE;
    dim3 grid((unsigned) ksp_divup(cols, Tile::COLS), (unsigned) ksp_divup(rows, Tile::ROWS));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    transpose_base_kernel<T, BLOCK, VTX, VTY><<<grid, dim3(BLOCK, BLOCK), 0, s>>>(
        (T *) dst, (const T *) src, (int) rows, (int) cols, dst_stride, src_stride);
    KSP_CHECK_LAUNCH();
    return 0;
}

template <typename T>
int launch_tile(cudaStream_t s, void *dst, const void *src, int64_t rows, int64_t cols,
                int64_t dst_stride, int64_t src_stride)
{
    dim3 grid((unsigned) ksp_divup(cols, 32), (unsigned) ksp_divup(rows, 32));
    if (grid.y > 65535) return KSP_ETOOLARGE;
    transpose_tile_kernel<T><<<grid, dim3(32, 8), 0, s>>>((T *) dst, (const T *) src, rows, cols,
                                                          dst_stride, src_stride);
    KSP_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int ksp_transpose(void *stream, void *dst, const void *src, int64_t rows, int64_t cols,
                             int64_t dst_stride, int64_t src_stride, int elem_size)
{
    if (rows < 0 || cols < 0 || dst_stride < rows || src_stride < cols) return KSP_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!dst || !src) return KSP_EINVAL;
    cudaStream_t s = (cudaStream_t) stream;
    if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8 && elem_size != 16)
        return KSP_EINVAL;
    if (((uintptr_t) dst | (uintptr_t) src) % (uintptr_t) elem_size) return KSP_EALIGN;
    switch (elem_size) {
    case 1: {
        const bool a16 = ((uintptr_t) dst % 16 == 0) && ((uintptr_t) src % 16 == 0) &&
                         (dst_stride % 16 == 0) && (src_stride % 16 == 0);
        // whole 128 x 128 tiles with the 128-bit kernel, the ragged right / bottom strips (and
        // arrays that are not 16-byte aligned) with the 64 x 64 one
        const int64_t r_main = a16 ? rows / 128 * 128 : 0, c_main = a16 ? cols / 128 * 128 : 0;
        if (r_main > 0 && c_main > 0) {
            dim3 grid((unsigned) (c_main / 128), (unsigned) (r_main / 128));
            if (grid.y > 65535) return KSP_ETOOLARGE;
            transpose_bytes128_kernel<<<grid, 256, 0, s>>>((uint8_t *) dst, (const uint8_t *) src,
                                                           dst_stride, src_stride);
            KSP_CHECK_LAUNCH();
        }
        int vec_ok = ((uintptr_t) dst % 4 == 0) && ((uintptr_t) src % 4 == 0) &&
                     (dst_stride % 4 == 0) && (src_stride % 4 == 0);
        // region [r_lo, rows) x [c_lo, cols) of the source
        auto rest = [&](int64_t r_lo, int64_t c_lo, int64_t r_hi, int64_t c_hi) -> int {
            if (r_hi <= r_lo || c_hi <= c_lo) return 0;
            dim3 grid((unsigned) ksp_divup(c_hi - c_lo, 64), (unsigned) ksp_divup(r_hi - r_lo, 64));
            if (grid.y > 65535) return KSP_ETOOLARGE;
            transpose_bytes_kernel<<<grid, 256, 0, s>>>(
                (uint8_t *) dst + c_lo * dst_stride + r_lo, (const uint8_t *) src + r_lo * src_stride + c_lo,
                r_hi - r_lo, c_hi - c_lo, dst_stride, src_stride, vec_ok && r_lo % 4 == 0 && c_lo % 4 == 0);
            KSP_CHECK_LAUNCH();
            return 0;
        };
        if (r_main > 0 && c_main > 0) {
            int rc = rest(0, c_main, r_main, cols);            // right strip
            if (rc) return rc;
            return rest(r_main, 0, rows, cols);                // bottom strip (full width)
        }
        return rest(0, 0, rows, cols);
    }
    case 2: return launch_base<uint16_t>(s, dst, src, rows, cols, dst_stride, src_stride);
    case 4: {
        const bool a16 = ((uintptr_t) dst % 16 == 0) && ((uintptr_t) src % 16 == 0) &&
                         (dst_stride % 4 == 0) && (src_stride % 4 == 0);
        if (!a16 || rows < 64 || cols < 64)
            return launch_tile<uint32_t>(s, dst, src, rows, cols, dst_stride, src_stride);
        dim3 grid((unsigned) ksp_divup(cols, 64), (unsigned) ksp_divup(rows, 64));
        if (grid.y > 65535) return KSP_ETOOLARGE;
        transpose_words128_kernel<<<grid, 256, 0, s>>>((uint32_t *) dst, (const uint32_t *) src, rows,
                                                       cols, dst_stride, src_stride);
        KSP_CHECK_LAUNCH();
        return 0;
    }
    case 8: return launch_tile<uint2>(s, dst, src, rows, cols, dst_stride, src_stride);
    case 16: return launch_base<uint4>(s, dst, src, rows, cols, dst_stride, src_stride);
    default: return KSP_EINVAL;
    }
}
