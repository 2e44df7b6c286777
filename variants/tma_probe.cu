// Probe: which way of handing a 2-D tensor map to cp.async.bulk.tensor works on this box (param direct, param array with
// dynamic index, global memory).  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "../mofreak_b200/csrc/tma_device.cuh"
using namespace mfk;
struct Maps { CUtensorMap map[6]; int use[6]; };
__device__ __forceinline__ void tload(void* dst, const void* tm, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(smem_addr(bar)) : "memory");
}
template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap one, const __grid_constant__ Maps arr, const CUtensorMap* gm, int l, int x, int y, uint8_t* out, int BX, int BY) {
    __shared__ __align__(128) uint8_t raw[64 * 256];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, BX * BY);
        const void* tm = MODE == 0 ? (const void*)&one : MODE == 1 ? (const void*)&arr.map[l] : (const void*)(gm + l);
        tload(raw, tm, x, y, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = raw[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int BX = argc > 1 ? atoi(argv[1]) : 80, BY = argc > 2 ? atoi(argv[2]) : 38;
    const int w = 720, h = 576 * 4;
    uint8_t* img; cudaMalloc(&img, (size_t)w * h + 256);
    uint8_t* himg = (uint8_t*)malloc((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) himg[i] = (uint8_t)(i * 7 + (i >> 9));
    cudaMemcpy(img, himg, (size_t)w * h, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %p q %d\n", p, (int)q);
    Maps M; memset(&M, 0, sizeof(M));
    for (int l = 0; l < 6; l++) {
        const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
        const cuuint64_t strides[1] = {(cuuint64_t)w};
        const cuuint32_t box[2] = {(cuuint32_t)BX, (cuuint32_t)BY};
        const cuuint32_t estr[2] = {1, 1};
        CUresult rc = ((EncodeTiledFn)p)(&M.map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (l == 0) printf("encode rc %d\n", (int)rc);
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(M.map)); cudaMemcpy(gm, M.map, sizeof(M.map), cudaMemcpyHostToDevice);
    uint8_t* out; cudaMalloc(&out, 64 * 256);
    static uint8_t hout[64 * 256];
    const int xs[3] = {60, -4, 700}, ys[3] = {100, -3, 2290};
    const int only = argc > 3 ? atoi(argv[3]) : -1;
    for (int mode = 0; mode < 3; mode++) if (only < 0 || only == mode)
        for (int t = 0; t < 3; t++) {
            cudaMemset(out, 0xEE, 64 * 256);
            if (mode == 0) probe<0><<<1, 128>>>(M.map[0], M, gm, 3, xs[t], ys[t], out, BX, BY);
            if (mode == 1) probe<1><<<1, 128>>>(M.map[0], M, gm, 3, xs[t], ys[t], out, BX, BY);
            if (mode == 2) probe<2><<<1, 128>>>(M.map[0], M, gm, 3, xs[t], ys[t], out, BX, BY);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("box %dx%d mode %d case %d: %s\n", BX, BY, mode, t, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hout, out, sizeof(hout), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < BY; r++) for (int c = 0; c < BX; c++) {
                const int gx = xs[t] + c, gy = ys[t] + r;
                const uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? himg[(size_t)gy * w + gx] : 0;
                bad += hout[r * BX + c] != want;
            }
            printf("mode %d case %d: %d mismatches\n", mode, t, bad);
        }
    return 0;
}
