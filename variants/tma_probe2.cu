#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int V>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int fence, int cx, int cy) {
    __shared__ __align__(128) float raw[64 * 32];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(raw);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        if (fence == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (fence != 2) __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(64 * 32 * 4) : "memory");
        if (V == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(cx), "r"(cy), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(cx), "r"(cy), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(b) : "memory");
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = raw[i];
}
int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    const int wb = argc > 2 ? atoi(argv[2]) : 4096; const int fence = argc > 3 ? atoi(argv[3]) : 0; const int u8 = argc > 4 ? atoi(argv[4]) : 0;
    const int cx = argc > 5 ? atoi(argv[5]) : 0, cy = argc > 6 ? atoi(argv[6]) : 0;
    const int w = 1024, h = 256;
    float* img; cudaMalloc(&img, (size_t)w * h * 4);
    float* himg = (float*)malloc((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) himg[i] = (float)i;
    cudaMemcpy(img, himg, (size_t)w * h * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)(u8 ? wb : wb / 4), (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)wb};
    const cuuint32_t box[2] = {(cuuint32_t)(u8 ? 256 : 64), 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = ((EncodeTiledFn)p)(&tm, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d\n", (int)rc);
    float* out; cudaMalloc(&out, 64 * 32 * 4);
    if (v == 0) probe<0><<<1, 128>>>(tm, out, fence, cx, cy); else probe<1><<<1, 128>>>(tm, out, fence, cx, cy);
    cudaError_t e = cudaDeviceSynchronize();
    printf("xy %d %d variant %d wb %d fence %d u8 %d: %s\n", cx, cy, v, wb, fence, u8, cudaGetErrorString(e));
    if (e == cudaSuccess) { float ho[4]; cudaMemcpy(ho, out + 64, 16, cudaMemcpyDeviceToHost); printf("row1: %g %g (want 1024 1025)\n", ho[0], ho[1]); }
    return 0;
}
