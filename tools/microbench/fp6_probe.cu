// Upper bound for the field layer: a register-resident loop of fp6_mul / fp6_sqr (the real headers),
// no point-formula wrappers, no local memory.   nvcc -arch=sm_100a -O3 -I../../schnorr-sig_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "fp6.cuh"
using namespace sb;

template <int MODE>
__global__ void __launch_bounds__(128) probe(const fp6* a, fp6* r, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fp6 x = a[i], y = a[i + 1];
    for (int t = 0; t < iters; t++) {
        if (MODE == 0) { x = fp6_mul(x, y); y = fp6_mul(y, x); }
        if (MODE == 1) { x = fp6_sqr(x); y = fp6_sqr(y); }
        if (MODE == 2) { x = fp6_sqr(fp6_add(x, y)); y = fp6_sub(fp6_sqr(y), x); }   // dbl-like mix
    }
    r[i] = fp6_add(x, y);
}
template <int MODE>
void run(const char* name, double w_per_iter, int sms, int blocks_per_sm) {
    int blocks = sms * blocks_per_sm, iters = 2000;
    fp6 *a, *r;
    cudaMalloc(&a, sizeof(fp6) * (blocks * 128 + 1));
    cudaMalloc(&r, sizeof(fp6) * blocks * 128);
    cudaMemset(a, 0x5a, sizeof(fp6) * (blocks * 128 + 1));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 128>>>(a, r, 100);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 128>>>(a, r, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 128 * iters * 2;
    printf("%-28s blocks/SM %d: %8.3f ms  %.3e fp6-ops/s  %.3e W/s (%.0f%% of 8.55e12)\n", name, blocks_per_sm, ms,
           ops / (ms * 1e-3), ops * w_per_iter / (ms * 1e-3), 100.0 * ops * w_per_iter / (ms * 1e-3) / 8.55e12);
    cudaFree(a); cudaFree(r);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int b = 2; b <= 8; b += 2) {
        run<0>("fp6_mul chain", 154, p.multiProcessorCount, b);
        run<1>("fp6_sqr chain", 93, p.multiProcessorCount, b);
        run<2>("sqr + add/sub mix", 93, p.multiProcessorCount, b);
    }
    return 0;
}
