// How long do cudaMalloc / cudaMallocAsync take while a long kernel occupies every SM?  (experiment behind the slab allocator)
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
__global__ void spin(long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
}
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    void *p = nullptr;
    cudaMalloc(&p, 1 << 20); cudaFree(p);
    for (int rep = 0; rep < 2; rep++) {
        for (size_t mb : {1, 64, 1024}) {
            spin<<<148 * 4, 256, 0, a>>>(300ll * 1000 * 1000);
            double t0 = now(); cudaMalloc(&p, mb << 20); double t1 = now();
            printf("busy GPU: cudaMalloc %4zu MB %.3f ms\n", mb, t1 - t0);
            t0 = now(); cudaFree(p); t1 = now();
            printf("busy GPU: cudaFree   %4zu MB %.3f ms\n", mb, t1 - t0);
            cudaStreamSynchronize(a);
            spin<<<148 * 4, 256, 0, a>>>(300ll * 1000 * 1000);
            t0 = now(); cudaMallocAsync(&p, mb << 20, b); cudaStreamSynchronize(b); t1 = now();
            printf("busy GPU: cudaMallocAsync %4zu MB (pool %s) %.3f ms\n", mb, rep ? "warm" : "cold", t1 - t0);
            t0 = now(); cudaFreeAsync(p, b); cudaStreamSynchronize(b); t1 = now();
            printf("busy GPU: cudaFreeAsync  %4zu MB %.3f ms\n", mb, t1 - t0);
            t0 = now(); cudaMallocAsync(&p, (mb << 20) / 2, b); cudaStreamSynchronize(b); t1 = now();
            printf("busy GPU: cudaMallocAsync %4zu MB again (half size) %.3f ms\n", mb / 2, t1 - t0);
            cudaFreeAsync(p, b);
            cudaStreamSynchronize(a);
        }
    }
    double t0 = now(); cudaMalloc(&p, 1024 << 20); double t1 = now();
    printf("idle GPU: cudaMalloc 1024 MB %.3f ms\n", t1 - t0);
    return 0;
}
