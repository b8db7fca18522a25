// Microbenchmark: issue rate of the legacy mma.sync tensor path on sm_100a (cycles per instruction per SM sub-partition).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int KIND, int ILP>
__global__ void k(long long* out, float* sink, int iters) {
    float acc[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 5u, 7u}; uint32_t b0 = threadIdx.x, b1 = 11u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { if (KIND == 0) mma_f16(acc[i], a, b0, b1); else mma_tf32(acc[i], a, b0, b1); }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += acc[i][e];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
}
template <int KIND, int ILP> void run(const char* name, int warps) {
    long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const int iters = 2000;
    k<KIND, ILP><<<148, warps * 32>>>(d, sink, iters); cudaDeviceSynchronize();
    k<KIND, ILP><<<148, warps * 32>>>(d, sink, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)iters * ILP);
    printf("%-22s warps/CTA %2d ILP %d: %7.2f clk per mma per warp  -> %6.2f clk per mma per sub-partition\n", name, warps, ILP, per, per / ((warps + 3) / 4));
    cudaFree(d); cudaFree(sink);
}
int main() {
    run<0, 1>("m16n8k16 f16 (dep chain)", 1); run<0, 4>("m16n8k16 f16", 1); run<0, 8>("m16n8k16 f16", 1);
    run<0, 8>("m16n8k16 f16", 4); run<0, 8>("m16n8k16 f16", 8); run<0, 8>("m16n8k16 f16", 16);
    run<1, 1>("m16n8k8 tf32 (dep chain)", 1); run<1, 8>("m16n8k8 tf32", 1); run<1, 8>("m16n8k8 tf32", 4); run<1, 8>("m16n8k8 tf32", 8); run<1, 8>("m16n8k8 tf32", 16);
    return 0;
}
