// tools/pipeprobe.cu -- which integer instructions share an issue pipe on B200 (development aid).
// Each test runs 8 (or 8+8) independent chains per thread, fully unrolled, and reports
// lane-operations per clock per SM for every instruction class in the mix.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipeprobe tools/pipeprobe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

#define LOP(x)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c))
#define ADD(x)   asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define SHF(x)   asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x) : "r"(b))
#define MAD(x)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c))
#define MHI(x)   asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define MWD(x)   asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(x) : "r"(b))
#define FMA(x)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(fb), "f"(fc))
#define SHL(x)   asm volatile("shl.b32 %0, %0, 1;" : "+r"(x))
#define POPC(x)  asm volatile("popc.b32 %0, %0;" : "+r"(x))
#define PRMT(x)  asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(x) : "r"(b))

#define REP8(OP) OP(a0); OP(a1); OP(a2); OP(a3); OP(a4); OP(a5); OP(a6); OP(a7)
#define REP8B(OP) OP(d0); OP(d1); OP(d2); OP(d3); OP(d4); OP(d5); OP(d6); OP(d7)
#define REP8F(OP) OP(f0); OP(f1); OP(f2); OP(f3); OP(f4); OP(f5); OP(f6); OP(f7)
#define REP8W(OP) OP(w0); OP(w1); OP(w2); OP(w3); OP(w4); OP(w5); OP(w6); OP(w7)
#define REP4W(OP) OP(w0); OP(w1); OP(w2); OP(w3)
#define REP4B(OP) OP(d0); OP(d1); OP(d2); OP(d3)
#define REP2B(OP) OP(d0); OP(d1)

#define KERNEL(NAME, BODY)                                                                      \
    __global__ void __launch_bounds__(256) NAME(uint32_t* out, int iters, uint32_t b, uint32_t c) { \
        uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; \
        uint32_t d0 = a0 * 3, d1 = a1 * 3, d2 = a2 * 3, d3 = a3 * 3, d4 = a4 * 3, d5 = a5 * 3, d6 = a6 * 3, d7 = a7 * 3; \
        float f0 = a0, f1 = a1, f2 = a2, f3 = a3, f4 = a4, f5 = a5, f6 = a6, f7 = a7;          \
        float fb = __uint_as_float(b) , fc = __uint_as_float(c);                                \
        uint32_t b2 = b, junk = 0;                                                              \
        uint64_t w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;       \
        _Pragma("unroll 1") for (int i = 0; i < iters; ++i) {                                   \
            _Pragma("unroll") for (int u = 0; u < 4; ++u) { BODY }                              \
        }                                                                                       \
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ d0 ^ d1 ^ d2 ^ d3 ^ d4 ^ d5 ^ d6 ^ d7 ^ junk ^ b2 \
            ^ (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 7) \
            ^ __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);                           \
    }

KERNEL(k_lop, REP8(LOP);)
KERNEL(k_add, REP8(ADD);)
KERNEL(k_shf, REP8(SHF);)
KERNEL(k_shl, REP8(SHL);)
KERNEL(k_mad, REP8(MAD);)
KERNEL(k_mhi, REP8(MHI);)
KERNEL(k_mwd, REP8W(MWD);)
KERNEL(k_lop_mwd, REP8(LOP); REP8W(MWD);)
KERNEL(k_lop_mwd2, REP8(LOP); REP4W(MWD);)
KERNEL(k_fma, REP8F(FMA);)
KERNEL(k_popc, REP8(POPC);)
KERNEL(k_prmt, REP8(PRMT);)
KERNEL(k_lop_mad, REP8(LOP); REP8B(MAD);)
KERNEL(k_lop_mad2, REP8(LOP); REP4B(MAD);)
KERNEL(k_lop_mad4, REP8(LOP); REP2B(MAD);)
KERNEL(k_lop_mhi4, REP8(LOP); REP2B(MHI);)
KERNEL(k_lop_mhi2, REP8(LOP); REP4B(MHI);)
KERNEL(k_lop_fma, REP8(LOP); REP8F(FMA);)
KERNEL(k_lop_shf, REP8(LOP); REP8B(SHF);)
KERNEL(k_lop_add, REP8(LOP); REP8B(ADD);)
KERNEL(k_lop_popc4, REP8(LOP); REP2B(POPC);)

typedef void (*kern_t)(uint32_t*, int, uint32_t, uint32_t);

static void run(const char* name, kern_t k, int ops_a, int ops_b, int sms, double mhz_hint) {
    const int blocks = sms * 8, iters = 8192;
    uint32_t* out; CK(cudaMalloc(&out, (size_t)blocks * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<blocks, 256>>>(out, iters, 0x9e3779b9u, 0x85ebca6bu);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) best = ms < best ? ms : best;
    }
    const double lanes = (double)blocks * 256 * iters * 4.0;
    const double per_clk_sm = 1.0 / (best * 1e-3) / (mhz_hint * 1e6) / sms;
    printf("%-14s %8.3f ms   A: %6.1f lane-ops/clk/SM", name, best, lanes * ops_a * per_clk_sm);
    if (ops_b) printf("   B: %6.1f   total %6.1f", lanes * ops_b * per_clk_sm, lanes * (ops_a + ops_b) * per_clk_sm);
    printf("\n");
    CK(cudaFree(out));
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
    printf("%s, %d SMs; per-clock figures assume %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    const int s = p.multiProcessorCount;
    run("LOP3", k_lop, 8, 0, s, mhz);
    run("IADD", k_add, 8, 0, s, mhz);
    run("SHF.L.W", k_shf, 8, 0, s, mhz);
    run("SHL", k_shl, 8, 0, s, mhz);
    run("IMAD", k_mad, 8, 0, s, mhz);
    run("IMAD.HI", k_mhi, 8, 0, s, mhz);
    run("IMAD.WIDE", k_mwd, 8, 0, s, mhz);
    run("FFMA", k_fma, 8, 0, s, mhz);
    run("POPC", k_popc, 8, 0, s, mhz);
    run("PRMT", k_prmt, 8, 0, s, mhz);
    run("LOP3+IMAD 8:8", k_lop_mad, 8, 8, s, mhz);
    run("LOP3+IMAD 8:4", k_lop_mad2, 8, 4, s, mhz);
    run("LOP3+IMAD 8:2", k_lop_mad4, 8, 2, s, mhz);
    run("LOP3+WIDE 8:8", k_lop_mwd, 8, 8, s, mhz);
    run("LOP3+WIDE 8:4", k_lop_mwd2, 8, 4, s, mhz);
    run("LOP3+MHI 8:2", k_lop_mhi4, 8, 2, s, mhz);
    run("LOP3+MHI 8:4", k_lop_mhi2, 8, 4, s, mhz);
    run("LOP3+FFMA 8:8", k_lop_fma, 8, 8, s, mhz);
    run("LOP3+SHF 8:8", k_lop_shf, 8, 8, s, mhz);
    run("LOP3+IADD 8:8", k_lop_add, 8, 8, s, mhz);
    run("LOP3+POPC 8:2", k_lop_popc4, 8, 2, s, mhz);
    return 0;
}
