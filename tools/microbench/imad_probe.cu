// Integer-pipe probe for the roofline denominator: how fast do IMAD.WIDE / IMAD issue on sm_100a,
// and how many ALU instructions co-issue for free?   nvcc -arch=sm_100a -O3 -o imad_probe imad_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, uint64_t* sink) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint64_t acc[8];
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = k; x[k] = k * 77u + a; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0 || MODE >= 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
            if (MODE == 1) { uint32_t lo = (uint32_t)acc[k]; asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a + k), "r"(b)); acc[k] = lo; }
            if (MODE == 3 || MODE == 4 || MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(b));
            if (MODE == 4 || MODE == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[(k + 1) & 7]) : "r"(a));
            if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(k + 3) & 7]) : "r"(a));
        }
        if (MODE == 6) {  // carry-chain form used by the field code: 4 IMAD.WIDE + 3 addc per 64x64 product, two products
            uint32_t e0 = (uint32_t)acc[0], e1 = (uint32_t)(acc[0] >> 32), e2 = (uint32_t)acc[1], e3 = (uint32_t)(acc[1] >> 32), e4 = x[0];
            uint32_t o1 = (uint32_t)acc[2], o2 = (uint32_t)(acc[2] >> 32), o3 = x[1];
            asm volatile("mad.lo.cc.u32 %0, %8, %10, %0;\n\tmadc.hi.cc.u32 %1, %8, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\t"
                         "madc.hi.cc.u32 %3, %9, %11, %3;\n\taddc.u32 %4, %4, 0;\n\tmad.lo.cc.u32 %5, %8, %11, %5;\n\t"
                         "madc.hi.cc.u32 %6, %8, %11, %6;\n\taddc.u32 %7, %7, 0;\n\tmad.lo.cc.u32 %5, %9, %10, %5;\n\t"
                         "madc.hi.cc.u32 %6, %9, %10, %6;\n\taddc.u32 %7, %7, 0;"
                         : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o1), "+r"(o2), "+r"(o3)
                         : "r"(a), "r"(b), "r"(a + 1), "r"(b + 1));
            acc[0] = ((uint64_t)e1 << 32) | e0; acc[1] = ((uint64_t)e3 << 32) | e2; acc[2] = ((uint64_t)o2 << 32) | o1; x[0] = e4; x[1] = o3;
        }
        b += 3;
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k] + x[k];
    if (s == 0x123456789abcdefULL) sink[0] = s;
}

template <int MODE>
double run(const char* name, int sms, double w_per_iter, uint64_t* sink) {
    int blocks = sms * 8, iters = 1 << 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(iters / 16, sink);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * 256 * iters * w_per_iter / (ms * 1e-3);
    printf("%-44s %8.3f ms  %.3e mul/s  = %.2f per clk per SM @1.965GHz\n", name, ms, rate, rate / sms / 1.965e9);
    return rate;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint64_t* sink; cudaMalloc(&sink, 64);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("mad.wide.u32 x8 chains", p.multiProcessorCount, 8, sink);
    run<1>("mad.lo.u32 x8 chains", p.multiProcessorCount, 8, sink);
    run<3>("mad.wide + 1 ALU per mul", p.multiProcessorCount, 8, sink);
    run<4>("mad.wide + 2 ALU per mul", p.multiProcessorCount, 8, sink);
    run<5>("mad.wide + 3 ALU per mul", p.multiProcessorCount, 8, sink);
    run<6>("8 mad.wide + carry-chain product (4W+3addc)", p.multiProcessorCount, 12, sink);
    run<0>("mad.wide.u32 x8 chains (again, warm)", p.multiProcessorCount, 8, sink);
    return 0;
}
