// Do two warps of one scheduler run faster when they are OUT of phase (one in the multiply-heavy head of fp6_mul,
// the other in its ALU-only reduction tail)?  Same register-resident fp6_mul chain as fp6_probe.cu; odd warps are
// delayed by `delay` clock cycles before the loop.   nvcc -arch=sm_100a -O3 -I../../schnorr-sig_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "fp6.cuh"
using namespace sb;

template <int MODE>
__global__ void __launch_bounds__(128) probe(const fp6* a, fp6* r, int iters, int delay, int which) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fp6 x = a[i], y = a[i + 1];
    int warp = threadIdx.x >> 5;
    bool late = which == 0 ? (warp & 1) : (which == 1 ? (blockIdx.x & 1) : ((warp + blockIdx.x) & 1));
    if (late && delay > 0) {
        long long t0 = clock64();
        while (clock64() - t0 < delay) { }
    }
    for (int t = 0; t < iters; t++) {
        if (MODE == 0) { x = fp6_mul(x, y); y = fp6_mul(y, x); }
        if (MODE == 1) { x = fp6_sqr(x); y = fp6_sqr(y); }
    }
    r[i] = fp6_add(x, y);
}
template <int MODE>
void run(const char* name, int sms, int blocks_per_sm, int delay, int which) {
    int blocks = sms * blocks_per_sm, iters = 2000;
    fp6 *a, *r;
    cudaMalloc(&a, sizeof(fp6) * (blocks * 128 + 1));
    cudaMalloc(&r, sizeof(fp6) * blocks * 128);
    cudaMemset(a, 0x5a, sizeof(fp6) * (blocks * 128 + 1));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 128>>>(a, r, 100, delay, which);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 128>>>(a, r, iters, delay, which);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 128 * iters * 2;
    double cyc = ms * 1e-3 * 1.965e9 / (iters * 2.0 * blocks_per_sm);   // SMSP cycles per warp-level fp6 op (4 warps per block, 1 per SMSP)
    printf("%-10s blocks/SM %d delay %5d (%s): %8.3f ms  %.3e fp6-ops/s  %.0f cycles per warp-op per SMSP\n", name, blocks_per_sm, delay,
           which == 0 ? "odd warps" : (which == 1 ? "odd blocks" : "checker"), ms, ops / (ms * 1e-3), cyc);
    cudaFree(a); cudaFree(r);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int b = 2; b <= 4; b += 2)
        for (int which = 1; which <= 2; which++)
            for (int d = 0; d <= 1200; d += 200) {
                run<0>("fp6_mul", sms, b, d, which);
                if (which == 2 && d == 0) continue;
            }
    for (int d = 0; d <= 800; d += 200) run<1>("fp6_sqr", sms, 2, d, 1);
    return 0;
}
