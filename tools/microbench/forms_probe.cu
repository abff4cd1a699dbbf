#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// Throughput of the integer-multiply instruction FORMS the field code compiles to.
// One multiplicand of every product is the low word of an accumulator, so nothing is loop-invariant.
template <int MODE>
__global__ void __launch_bounds__(128) probe(int iters, uint64_t* sink, uint32_t seed) {
    uint32_t y[8], x[8];
    uint64_t A[8];   // 64-bit accumulators (aligned register pairs)
    uint32_t T[8];   // third words
#pragma unroll
    for (int k = 0; k < 8; k++) {
        y[k] = blockIdx.x * 40503u + 977u * (k + 3) + threadIdx.x;
        A[k] = (uint64_t)(threadIdx.x * 2654435761u + seed * (k + 1)) * 0x9E3779B97F4A7C15ULL;
        T[k] = k;
        x[k] = y[k] * 3u + 1u;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] += y[(k + 1) & 7];   // one IADD per product keeps the multiplicands loop-variant
        if (MODE == 0) {  // IMAD.WIDE RZ, result xored into the accumulator (2 LOP3)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[k]), "r"(y[k]));
                A[k] ^= p;
            }
        }
        if (MODE == 1) {  // IMAD.WIDE with 64-bit addend
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(A[k]) : "r"(x[k]), "r"(y[k]));
        }
        if (MODE == 2) {  // IMAD.WIDE addend + carry-out, carry add into a third word
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %3, %2, lo;\n\tmadc.hi.cc.u32 hi, %3, %2, hi;\n\t"
                             "addc.u32 %1, %1, 0;\n\tmov.b64 %0, {lo, hi};\n\t}"
                             : "+l"(A[k]), "+r"(T[k]) : "r"(y[k]), "r"(x[k]));
        }
        if (MODE == 3) {  // even chain: WIDE(P out) -> WIDE.X(P in/out) -> carry add     [4 chains, 2 wide each]
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("{\n\t.reg .u32 a, b, c, d;\n\tmov.b64 {a, b}, %0;\n\tmov.b64 {c, d}, %1;\n\t"
                             "mad.lo.cc.u32 a, %3, %4, a;\n\tmadc.hi.cc.u32 b, %3, %4, b;\n\tmadc.lo.cc.u32 c, %5, %6, c;\n\t"
                             "madc.hi.cc.u32 d, %5, %6, d;\n\taddc.u32 %2, %2, 0;\n\tmov.b64 %0, {a, b};\n\tmov.b64 %1, {c, d};\n\t}"
                             : "+l"(A[2 * k]), "+l"(A[2 * k + 1]), "+r"(T[k])
                             : "r"(x[2 * k]), "r"(y[2 * k]), "r"(x[2 * k + 1]), "r"(y[2 * k + 1]));
        }
        if (MODE == 5) {  // WIDE RZ pairs summed into 96-bit accumulators with the two-carry IADD3 (3 ALU per 2 wide)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint64_t p, q;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[2 * k]), "r"(y[2 * k]));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(x[2 * k + 1]), "r"(y[2 * k + 1]));
                unsigned __int128 acc = ((unsigned __int128)T[k] << 64) | A[2 * k];
                acc += (unsigned __int128)p + q;
                A[2 * k] = (uint64_t)acc; T[k] = (uint32_t)(acc >> 64);
            }
        }
        if (MODE == 6) {  // MODE 0 + one FFMA per multiply
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[k]), "r"(y[k]));
                A[k] ^= p;
                float f = __uint_as_float(T[k]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
                T[k] = __float_as_uint(f);
            }
        }
        if (MODE == 7) {  // IMAD lo + IMAD.HI instead of one WIDE
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t lo = (uint32_t)A[k], hi = (uint32_t)(A[k] >> 32);
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(x[k]), "r"(y[k]));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x[k]), "r"(y[k]));
                A[k] = ((uint64_t)hi << 32) | lo;
            }
        }
        if (MODE == 8) {  // 32-bit IMAD only
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(T[k]) : "r"(x[k]), "r"(y[k]));
        }
        if (MODE == 9) {  // IMAD.HI only
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(T[k]) : "r"(x[k]), "r"(y[k]));
        }
        if (MODE == 10) {  // MODE 0 with 4 extra LOP3 per multiply (6 ALU per wide)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[k]), "r"(y[k]));
                A[k] ^= p;
                asm volatile("xor.b32 %0, %0, %1;\n\txor.b32 %0, %0, %2;\n\txor.b32 %0, %0, %3;\n\txor.b32 %0, %0, %4;"
                             : "+r"(T[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]), "r"(y[(k + 2) & 7]), "r"(y[(k + 3) & 7]));
            }
        }
        if (MODE == 11) {  // MODE 0 with 2 extra LOP3 per multiply (4 ALU per wide)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[k]), "r"(y[k]));
                A[k] ^= p;
                asm volatile("xor.b32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(T[k]) : "r"(y[k]), "r"(y[(k + 1) & 7]));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s = s * 0x9E3779B97F4A7C15ULL + A[k] + T[k];
    if (s == 0x123456789abcdefULL) sink[0] = s;
}
template <int MODE>
void run(const char* name, int sms, double w_per_iter, uint64_t* sink, int bps) {
    int blocks = sms * bps, iters = 1 << 15;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 128>>>(iters / 16, sink, 1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 128>>>(iters, sink, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * 128 * iters * w_per_iter / (ms * 1e-3);
    printf("%-52s warps/SMSP %d %8.3f ms  %.2f mul/clk/SM -> %.2f cycles per warp-multiply per SMSP\n", name, bps, ms,
           rate / sms / 1.965e9, 128.0 / (rate / sms / 1.965e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint64_t* sink; cudaMalloc(&sink, 64);
    for (int bps = 2; bps <= 8; bps *= 4) {
        run<0>("IMAD.WIDE RZ + 2 LOP3", p.multiProcessorCount, 8, sink, bps);
        run<11>("IMAD.WIDE RZ + 4 LOP3", p.multiProcessorCount, 8, sink, bps);
        run<10>("IMAD.WIDE RZ + 6 LOP3", p.multiProcessorCount, 8, sink, bps);
        run<1>("IMAD.WIDE 64-bit addend", p.multiProcessorCount, 8, sink, bps);
        run<2>("IMAD.WIDE addend+carry-out, + carry add", p.multiProcessorCount, 8, sink, bps);
        run<3>("even chain: WIDE(P) WIDE.X(P) addc", p.multiProcessorCount, 8, sink, bps);
        run<5>("WIDE RZ pairs + two-carry IADD3 into 96 bits", p.multiProcessorCount, 8, sink, bps);
        run<6>("IMAD.WIDE RZ + 2 LOP3 + FFMA", p.multiProcessorCount, 8, sink, bps);
        run<7>("IMAD lo + IMAD.HI pairs (count pairs)", p.multiProcessorCount, 8, sink, bps);
        run<8>("IMAD lo x8 (count instr)", p.multiProcessorCount, 8, sink, bps);
        run<9>("IMAD.HI x8 (count instr)", p.multiProcessorCount, 8, sink, bps);
    }
    return 0;
}
