// Fill: set every element of a buffer (padding included) to one value.
// Replaces reference fill.mako (launched from fill.py:130-139).  The element is an opaque bit
// pattern of 1, 2, 4, 8 or 16 bytes, so one pre-built kernel serves every dtype; threads
// write 16-byte vectors, a short scalar head and tail handle unaligned ends.
#include "common.cuh"
#include <string.h>

namespace {

__global__ void __launch_bounds__(256)
fill_kernel(uint8_t *data, size_t bytes, uint4 pattern16)
{
    // pattern16 = the element repeated to 16 bytes; the element size divides 16 and the buffer
    // starts on an element boundary, so byte i of the buffer is pattern byte i mod 16.
    const uintptr_t base = reinterpret_cast<uintptr_t>(data);
    const size_t head = ((16 - (base & 15)) & 15) < bytes ? ((16 - (base & 15)) & 15) : bytes;
    const size_t body = (bytes - head) / 16;
    const size_t tail_start = head + body * 16;
    const uint8_t *pat = reinterpret_cast<const uint8_t *>(&pattern16);
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    // rotate the pattern so that the aligned body starts at the right phase
    uint4 rot;
    uint8_t *r = reinterpret_cast<uint8_t *>(&rot);
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = pat[(head + i) & 15];
    uint4 *body_ptr = reinterpret_cast<uint4 *>(data + head);
    for (size_t i = tid; i < body; i += nthreads) body_ptr[i] = rot;
    for (size_t i = tid; i < head; i += nthreads) data[i] = pat[i % 16];
    for (size_t i = tail_start + tid; i < bytes; i += nthreads) data[i] = pat[i % 16];
}

}  // namespace

extern "C" int ksp_fill(void *stream, void *data, size_t elements, const void *value, size_t elem_size)
{
    if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8 && elem_size != 16)
        return KSP_EINVAL;
    if (elements == 0) return 0;
    if (!data || !value) return KSP_EINVAL;
    if ((uintptr_t) data % elem_size) return KSP_EALIGN;
    uint4 pattern;
    uint8_t *p = reinterpret_cast<uint8_t *>(&pattern);
    for (size_t i = 0; i < 16; i += elem_size) memcpy(p + i, value, elem_size);
    const size_t bytes = elements * elem_size;
    size_t blocks = ksp_divup((int64_t) (bytes / 16 + 1), 256 * 4);
    const size_t cap = (size_t) ksp_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    fill_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>((uint8_t *) data, bytes, pattern);
    KSP_CHECK_LAUNCH();
    return 0;
}
