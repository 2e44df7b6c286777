// Bag-of-words assignment: BagOfWordsRepresentation::bruteForceMatch / hammingDistance, reference
// src/MoFREAK/BagOfWordsRepresentation.cpp:22-72 — arg-min over the codebook of the 128-bit Hamming
// distance, lowest index on ties (strict '<', :30) — and the count part of buildHistogram (:107-116).
// The codebook is staged in shared memory (K x 16 bytes: 9.6 KB for the 600 words of KTH, 158 KB for the
// 10100 of UCF101) by ONE bulk copy of the TMA engine per block and chunk (cp.async.bulk, completion on an
// mbarrier: SASS UBLKCP) — the codebook is a dense, 16-byte aligned array, exactly the 1-D case the engine
// takes without a tensor map.  Every thread owns one descriptor in registers and streams the codewords as
// broadcast 128-bit shared loads: 2 XOR + 2 POPC.64 per codeword.  Integer-pipe bound by construction.
#include "kernels.h"
#include "tma_device.cuh"

namespace mfk {

constexpr int BOW_THREADS = 256;
constexpr int BOW_SMEM_WORDS_MAX = 13312;      // 208 KB of codewords per pass

__global__ void __launch_bounds__(BOW_THREADS)
k_bow_assign(const uint8_t* __restrict__ desc, size_t stride, size_t n_max, const uint32_t* __restrict__ d_n,
             const uint4* __restrict__ codebook, int K, int32_t* __restrict__ idx, uint32_t* __restrict__ counts,
             const uint32_t* __restrict__ seg_off, int n_seg) {
    extern __shared__ __align__(128) uint4 words[];
    __shared__ __align__(8) uint64_t bar;
    const size_t n = d_n ? min(n_max, (size_t)*d_n) : n_max;      // live length on the device (rows accumulated so far)
    const int chunk = min(K, BOW_SMEM_WORDS_MAX);
    const bool resident = K <= BOW_SMEM_WORDS_MAX;       // the whole codebook fits: staged once per block
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        if (resident) { mbar_expect_tx(&bar, (uint32_t)K * 16u); bulk_copy_g2s(words, codebook, (uint32_t)K * 16u, &bar); }
    }
    __syncthreads();
    if (resident) { mbar_wait(&bar, phase); phase ^= 1; }
    for (size_t base = (size_t)blockIdx.x * BOW_THREADS; base < n; base += (size_t)gridDim.x * BOW_THREADS) {
        const size_t i = base + threadIdx.x;
        unsigned long long lo = 0, hi = 0;
        if (i < n) {
            const uint8_t* d = desc + i * stride;            // rows are 16-byte aligned (stride 16 or 32)
            const uint4 v = *reinterpret_cast<const uint4*>(d);
            lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
            hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
        }
        int best = 0x7fffffff, best_k = -1;
        for (int k0 = 0; k0 < K; k0 += chunk) {
            const int kn = min(chunk, K - k0);
            if (!resident) {
                __syncthreads();                             // every thread is done with the previous chunk
                if (threadIdx.x == 0) {
                    fence_proxy_async();                                             // generic reads before the engine's writes
                    mbar_expect_tx(&bar, (uint32_t)kn * 16u);
                    bulk_copy_g2s(words, codebook + k0, (uint32_t)kn * 16u, &bar);
                }
                mbar_wait(&bar, phase); phase ^= 1;
            }
#pragma unroll 4
            for (int k = 0; k < kn; k++) {
                const uint4 c = words[k];
                const unsigned long long clo = (unsigned long long)c.x | ((unsigned long long)c.y << 32);
                const unsigned long long chi = (unsigned long long)c.z | ((unsigned long long)c.w << 32);
                const int dist = __popcll(lo ^ clo) + __popcll(hi ^ chi);
                if (dist < best) { best = dist; best_k = k0 + k; }
            }
        }
        if (i < n) {
            if (idx) idx[i] = best_k;
            if (counts && best_k >= 0) {
                size_t at = (size_t)best_k;
                if (seg_off) {                               // the video this row belongs to: last v with seg_off[v] <= i
                    int lo_v = 0, hi_v = n_seg;              // seg_off[n_seg] = total rows > i
                    while (hi_v - lo_v > 1) { const int mid = (lo_v + hi_v) >> 1; if (seg_off[mid] <= (uint32_t)i) lo_v = mid; else hi_v = mid; }
                    at += (size_t)lo_v * K;
                }
                atomicAdd(&counts[at], 1u);
            }
        }
    }
}

cudaError_t launch_bow_assign(cudaStream_t st, const uint8_t* desc, size_t stride_bytes, size_t n, const uint32_t* d_n,
                              const uint8_t* codebook, int K, int32_t* idx, uint32_t* counts, const uint32_t* seg_off, int n_seg) {
    if (n == 0 || K <= 0) return cudaSuccess;
    const int chunk = K < BOW_SMEM_WORDS_MAX ? K : BOW_SMEM_WORDS_MAX;
    const size_t smem = (size_t)chunk * 16;
    // a per-DEVICE attribute (a process may hold contexts on several GPUs): set it for the current device on every launch
    cudaError_t e = cudaFuncSetAttribute(k_bow_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, BOW_SMEM_WORDS_MAX * 16);
    if (e != cudaSuccess) return e;
    size_t blocks = (n + BOW_THREADS - 1) / BOW_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_bow_assign<<<(unsigned)blocks, BOW_THREADS, smem, st>>>(desc, stride_bytes, n, d_n, reinterpret_cast<const uint4*>(codebook), K, idx, counts,
                                                                  seg_off, n_seg);
    return cudaGetLastError();
}

// rows per video from the per-pair row counts: video v owns the queued pairs [pair_start[v], pair_start[v+1]); slot n_videos = 0
// (the exclusive scan that follows turns the array into row offsets, slot n_videos = total)
__global__ void k_video_rows(const uint32_t* __restrict__ pair_rows, const uint32_t* __restrict__ pair_start, int n_videos,
                             uint32_t* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > n_videos) return;
    uint32_t sum = 0;
    if (v < n_videos) for (uint32_t p = pair_start[v]; p < pair_start[v + 1]; p++) sum += pair_rows[p];
    out[v] = sum;
}

void launch_video_rows(cudaStream_t st, const uint32_t* pair_rows, const uint32_t* pair_start, int n_videos, uint32_t* out) {
    k_video_rows<<<(n_videos + 1 + 127) / 128, 128, 0, st>>>(pair_rows, pair_start, n_videos, out);
}

}  // namespace mfk
