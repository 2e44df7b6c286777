// Bag-of-words assignment: BagOfWordsRepresentation::bruteForceMatch / hammingDistance, reference
// src/MoFREAK/BagOfWordsRepresentation.cpp:22-72 — arg-min over the codebook of the 128-bit Hamming
// distance, lowest index on ties (strict '<', :30) — and the count part of buildHistogram (:107-116).
// The codebook is staged in shared memory once per block (K x 16 bytes: 9.6 KB for the 600 words of KTH,
// 158 KB for the 10100 of UCF101); every thread owns one descriptor in registers and streams the codewords
// as broadcast 128-bit shared loads: 2 XOR + 2 POPC.64 per codeword.  Integer-pipe bound by construction.
#include "kernels.h"

namespace mfk {

constexpr int BOW_THREADS = 256;
constexpr int BOW_SMEM_WORDS_MAX = 13312;      // 208 KB of codewords per pass

__global__ void __launch_bounds__(BOW_THREADS)
k_bow_assign(const uint8_t* __restrict__ desc, size_t stride, size_t n, const uint4* __restrict__ codebook, int K,
             int32_t* __restrict__ idx, uint32_t* __restrict__ counts) {
    extern __shared__ uint4 words[];
    const int chunk = min(K, BOW_SMEM_WORDS_MAX);
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
            __syncthreads();
            for (int k = threadIdx.x; k < kn; k += BOW_THREADS) words[k] = codebook[k0 + k];
            __syncthreads();
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
            if (counts && best_k >= 0) atomicAdd(&counts[best_k], 1u);
        }
    }
}

void launch_bow_assign(cudaStream_t st, const uint8_t* desc, size_t stride_bytes, size_t n, const uint8_t* codebook, int K,
                       int32_t* idx, uint32_t* counts) {
    if (n == 0 || K <= 0) return;
    const int chunk = K < BOW_SMEM_WORDS_MAX ? K : BOW_SMEM_WORDS_MAX;
    const size_t smem = (size_t)chunk * 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_bow_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, BOW_SMEM_WORDS_MAX * 16);
        attr_set = true;
    }
    size_t blocks = (n + BOW_THREADS - 1) / BOW_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_bow_assign<<<(unsigned)blocks, BOW_THREADS, smem, st>>>(desc, stride_bytes, n, reinterpret_cast<const uint4*>(codebook), K, idx, counts);
}

}  // namespace mfk
