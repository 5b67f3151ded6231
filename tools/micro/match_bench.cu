// micro-benchmark (development aid): cost of __match_any_sync vs a shared-memory atomicOr emulation, one warp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const int *vals, int iters, long long *out, int distinct)
{
    __shared__ unsigned mask[512];
    const int lane = threadIdx.x;
    for (int i = lane; i < 512; i += 32) mask[i] = 0;
    __syncwarp();
    int v = vals[lane] % distinct;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const unsigned p = __match_any_sync(0xffffffffu, v);
        acc += __popc(p);
        v = (v + (p & 1)) % distinct;
    }
    long long t1 = clock64();
    v = vals[lane] % distinct;
    for (int i = 0; i < iters; i++) {
        atomicOr(&mask[v], 1u << lane);
        __syncwarp();
        const unsigned p = mask[v];
        __syncwarp();
        mask[v] = 0;
        __syncwarp();
        acc += __popc(p);
        v = (v + (p & 1)) % distinct;
    }
    long long t2 = clock64();
    if (lane == 0) { out[0] = (t1 - t0) / iters; out[1] = (t2 - t1) / iters; out[2] = acc; }
}
int main()
{
    int h[32]; for (int i = 0; i < 32; i++) h[i] = (i * 2654435761u) >> 7 & 0xFFFF;
    int *d; long long *o; cudaMalloc(&d, 128); cudaMalloc(&o, 64); cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
    for (int distinct : {1, 2, 4, 8, 16, 32, 365}) {
        k<<<1, 32>>>(d, 2000, o, distinct); long long r[3]; cudaMemcpy(r, o, 24, cudaMemcpyDeviceToHost);
        printf("distinct<=%3d: match_any %lld cycles/iter, smem atomicOr emulation %lld cycles/iter\n", distinct, r[0], r[1]);
    }
    return 0;
}
