// `.mofreak` text on the device: the reference writes 22 ostream tokens per row
// (MoFREAKUtilities.cpp:691-719); at device extraction rates that formatting is the whole end-to-end cost, so the
// rows are turned into the same bytes here and leave the GPU as one contiguous text.
//   k_text_lengths   one thread per row: length of its line (formatting into a counting sink)
//   exclusive scan   line offsets, total length
//   k_text_write     one warp per 32 consecutive rows: lanes format into shared memory at their offset inside the
//                    warp's span, then the warp copies the span (contiguous in the output) with coalesced stores
#include "kernels.h"
#include "text_device.cuh"

namespace mfk {

namespace {

constexpr int TW_WARPS = 4;

__global__ void k_text_lengths(const TextRow* __restrict__ rows, uint32_t n, uint32_t* __restrict__ len, uint32_t* __restrict__ flags) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) { len[n] = 0; return; }                  // the scan turns this slot into the total
    CountSink s;
    const TextRow row = rows[r];
    if (!put_row(s, row)) atomicOr(flags, 1u);
    len[r] = (uint32_t)s.n;
}

__global__ void __launch_bounds__(TW_WARPS * 32)
k_text_write(const TextRow* __restrict__ rows, uint32_t n, const uint32_t* __restrict__ off, char* __restrict__ text) {
    __shared__ char span[TW_WARPS][32 * TEXT_ROW_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * TW_WARPS;
    for (uint32_t r0 = (blockIdx.x * TW_WARPS + wib) * 32u; r0 < n; r0 += warps * 32u) {
        const uint32_t r = r0 + lane;
        const uint32_t first = off[r0];
        const uint32_t r_end = min(r0 + 32u, n);
        const uint32_t bytes = off[r_end] - first;
        if (r < n) {
            BufSink s(span[wib] + (off[r] - first));
            const TextRow row = rows[r];
            put_row(s, row);
        }
        __syncwarp();
        char* dst = text + first;
        for (uint32_t i = lane; i < bytes; i += 32) dst[i] = span[wib][i];
        __syncwarp();
    }
}

}  // namespace

void launch_text_lengths(cudaStream_t st, const void* rows, uint32_t n, uint32_t* len, uint32_t* flags) {
    k_text_lengths<<<(n + 1 + 255) / 256, 256, 0, st>>>((const TextRow*)rows, n, len, flags);
}

void launch_text_write(cudaStream_t st, const void* rows, uint32_t n, const uint32_t* off, char* text) {
    if (n == 0) return;
    const uint32_t blocks = min((n + 32u * TW_WARPS - 1) / (32u * TW_WARPS), 148u * 8u);
    k_text_write<<<blocks, TW_WARPS * 32, 0, st>>>((const TextRow*)rows, n, off, text);
}

}  // namespace mfk
