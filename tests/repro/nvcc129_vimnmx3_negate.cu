// Reproducer (not part of the product): nvcc 12.9 / ptxas for sm_100a miscompiles the scalar formulation v_a below
// (max(best, max(mn, -mx)) is fused into VIMNMX3 and the negation is dropped); v_b and v_c are correct.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o repro nvcc129_vimnmx3_negate.cu ; prints a!=b count > 0 on B200.
#include <cstdio>
#include <cstdint>
__device__ __noinline__ int v_a(const uint8_t* img, int w, int x, int y) {   // original
    const uint8_t* p = img + y * w + x; const int c = *p;
    int d[8] = {p[-1]-c, p[-w-1]-c, p[-w]-c, p[-w+1]-c, p[1]-c, p[w+1]-c, p[w]-c, p[w-1]-c};
    int best=-256;
    for (int s=0;s<8;s++){ int mn=255,mx=-255; for(int k=0;k<5;k++){ int v=d[(s+k)&7]; mn=min(mn,v); mx=max(mx,v);} best=max(best,max(mn,-mx)); }
    return min(254,max(0,best-1));
}
__device__ __noinline__ int v_b(const uint8_t* img, int w, int x, int y) {   // explicit bright and dark differences
    const uint8_t* p = img + y * w + x; const int c = *p;
    int b[8] = {p[-1]-c, p[-w-1]-c, p[-w]-c, p[-w+1]-c, p[1]-c, p[w+1]-c, p[w]-c, p[w-1]-c};
    int e[8] = {c-p[-1], c-p[-w-1], c-p[-w], c-p[-w+1], c-p[1], c-p[w+1], c-p[w], c-p[w-1]};
    int best=-256;
    for (int s=0;s<8;s++){ int mb=255,md=255; for(int k=0;k<5;k++){ mb=min(mb,b[(s+k)&7]); md=min(md,e[(s+k)&7]);} best=max(best,max(mb,md)); }
    return min(254,max(0,best-1));
}
__device__ __noinline__ int v_c(const uint8_t* img, int w, int x, int y) {   // packed u16x2 like the 9/16 kernel
    const uint8_t* p = img + y * w + x; const unsigned c = *p;
    const int o[8] = {-1, -w-1, -w, -w+1, 1, w+1, w, w-1};
    unsigned r[8];
    for (int i=0;i<8;i++){ unsigned v=p[o[i]]; r[i]=v|((255u-v)<<16); }
    unsigned best=0;
    for (int s=0;s<8;s++){ unsigned m=__vimin3_u16x2(r[s],r[(s+1)&7],r[(s+2)&7]); m=__vimin3_u16x2(m,r[(s+3)&7],r[(s+4)&7]); best=__vmaxu2(best,m); }
    int br=(int)(best&0xffff)-(int)c, dk=(int)(best>>16)-(255-(int)c);
    return min(254,max(0,max(br,dk)-1));
}
__global__ void k(const uint8_t* img, int w, int h, int* bad){
    int x = 2 + blockIdx.x*blockDim.x+threadIdx.x, y = 2 + blockIdx.y;
    if (x>=w-2||y>=h-2) return;
    int a=v_a(img,w,x,y), b=v_b(img,w,x,y), c=v_c(img,w,x,y);
    if (a!=b) atomicAdd(&bad[0],1);
    if (b!=c) atomicAdd(&bad[1],1);
    if (x==117&&y==4) printf("a=%d b=%d c=%d\n",a,b,c);
}
int main(){
    const int w=160,h=120; static uint8_t im[w*h]; srand(3);
    for(int i=0;i<w*h;i++) im[i]=rand()%256;
    int rows[5][5]={{0,3,100,105,108},{1,3,101,108,109},{1,4,105,103,105},{0,1,100,101,103},{4,1,3,1,0}};
    for(int yy=0;yy<5;yy++)for(int xx=0;xx<5;xx++) im[(3+yy)*w+116+xx]=rows[yy][xx];
    uint8_t* d; cudaMalloc(&d,w*h); cudaMemcpy(d,im,w*h,cudaMemcpyHostToDevice);
    int* bad; cudaMalloc(&bad,8); cudaMemset(bad,0,8);
    k<<<dim3(2,h),128>>>(d,w,h,bad); cudaDeviceSynchronize(); int hb[2]; cudaMemcpy(hb,bad,8,cudaMemcpyDeviceToHost);
    printf("a!=b: %d  b!=c: %d  %s\n",hb[0],hb[1],cudaGetErrorString(cudaGetLastError())); return 0; }
