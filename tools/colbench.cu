// tools/colbench.cu -- micro-benchmark of Myers column-update variants on B200 (development aid).
//
// Every thread owns one pair (like a lane of nn_tile_kernel), keeps W window words of Pv/Mv in
// registers, reads Peq from shared memory and walks `cols` columns.  The variants differ only in
// which pipe executes the cross-word shifts / the carry add; all must print the same checksum.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gpurun_out/colbench tools/colbench.cu
//   gpurun -- ./gpurun_out/colbench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t d; asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d;
}

template <int W, int V>
struct Band {
    uint32_t Pv[W], Mv[W];
    uint32_t accP, accM;
    int score;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int w = 0; w < W; ++w) { Pv[w] = 0xffffffffu; Mv[w] = 0u; }
        accP = accM = 0u; score = 32 * W;
    }
    __device__ __forceinline__ void column(const uint32_t* __restrict__ eq, uint32_t two, uint32_t one) {
        if constexpr (V == 0) {                      // shipped code: everything on the ALU pipe
            uint32_t ph_below = 0x80000000u, mh_below = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t Eq = eq[4 * w];
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t t = Eq & pv;
                uint32_t s;
                if (w == 0) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                else        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                const uint32_t Xh = (s ^ pv) | Eq;
                const uint32_t Ph = mv | ~(Xh | pv);
                const uint32_t Mh = pv & Xh;
                const uint32_t Xv = Eq | mv;
                const uint32_t Phs = __funnelshift_l(ph_below, Ph, 1);
                const uint32_t Mhs = __funnelshift_l(mh_below, Mh, 1);
                Pv[w] = Mhs | ~(Xv | Phs);
                Mv[w] = Phs & Xv;
                ph_below = Ph; mh_below = Mh;
            }
            accP = __funnelshift_l(ph_below, accP, 1);
            accM = __funnelshift_l(mh_below, accM, 1);
        } else if constexpr (V == 1 || V == 2) {     // shifts on the FMA pipe: x*2 + carry_bit, carry_bit = mulhi(x, 2)
            uint32_t hp = 1u, hm = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t Eq = eq[4 * w];
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t t = Eq & pv;
                uint32_t s;
                if (w == 0) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                else        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                const uint32_t Xh = (s ^ pv) | Eq;
                const uint32_t Ph = mv | ~(Xh | pv);
                const uint32_t Mh = pv & Xh;
                const uint32_t Xv = Eq | mv;
                uint32_t Phs, Mhs;
                if constexpr (V == 1) {
                    Phs = mad_lo(Ph, two, hp); Mhs = mad_lo(Mh, two, hm);
                    hp = mul_hi(Ph, two); hm = mul_hi(Mh, two);
                } else {
                    const uint64_t a = mad_wide(Ph, two, (uint64_t)hp), b = mad_wide(Mh, two, (uint64_t)hm);
                    Phs = (uint32_t)a; hp = (uint32_t)(a >> 32);
                    Mhs = (uint32_t)b; hm = (uint32_t)(b >> 32);
                }
                Pv[w] = Mhs | ~(Xv | Phs);
                Mv[w] = Phs & Xv;
            }
            accP = mad_lo(accP, two, hp);
            accM = mad_lo(accM, two, hm);
        } else if constexpr (V == 3) {               // V1 + the carry add as two mad.wide per word
            uint32_t hp = 1u, hm = 0u, cin = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t Eq = eq[4 * w];
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t t = Eq & pv;
                uint64_t x = mad_wide(t, one, (uint64_t)pv);
                if (w > 0) x = mad_wide(cin, one, x);
                const uint32_t s = (uint32_t)x; cin = (uint32_t)(x >> 32);
                const uint32_t Xh = (s ^ pv) | Eq;
                const uint32_t Ph = mv | ~(Xh | pv);
                const uint32_t Mh = pv & Xh;
                const uint32_t Xv = Eq | mv;
                const uint32_t Phs = mad_lo(Ph, two, hp), Mhs = mad_lo(Mh, two, hm);
                hp = mul_hi(Ph, two); hm = mul_hi(Mh, two);
                Pv[w] = Mhs | ~(Xv | Phs);
                Mv[w] = Phs & Xv;
            }
            accP = mad_lo(accP, two, hp);
            accM = mad_lo(accM, two, hm);
        } else if constexpr (V == 5) {               // shifts on the FMA pipe: IMAD (low half) + IMAD.WIDE (carry bit)
            uint32_t hp = 1u, hm = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t Eq = eq[4 * w];
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t t = Eq & pv;
                uint32_t s;
                if (w == 0) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                else        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                const uint32_t Xh = (s ^ pv) | Eq;
                const uint32_t Ph = mv | ~(Xh | pv);
                const uint32_t Mh = pv & Xh;
                const uint32_t Xv = Eq | mv;
                const uint64_t a = mad_wide(Ph, two, 0ull), b = mad_wide(Mh, two, 0ull);
                const uint32_t Phs = mad_lo((uint32_t)a, one, hp), Mhs = mad_lo((uint32_t)b, one, hm);
                hp = (uint32_t)(a >> 32); hm = (uint32_t)(b >> 32);
                Pv[w] = Mhs | ~(Xv | Phs);
                Mv[w] = Phs & Xv;
            }
            accP = mad_lo(accP, two, hp);
            accM = mad_lo(accM, two, hm);
        } else if constexpr (V == 4) {               // only the Mh shift on the FMA pipe (balance point?)
            uint32_t ph_below = 0x80000000u, hm = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t Eq = eq[4 * w];
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t t = Eq & pv;
                uint32_t s;
                if (w == 0) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                else        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
                const uint32_t Xh = (s ^ pv) | Eq;
                const uint32_t Ph = mv | ~(Xh | pv);
                const uint32_t Mh = pv & Xh;
                const uint32_t Xv = Eq | mv;
                const uint32_t Phs = __funnelshift_l(ph_below, Ph, 1);
                const uint32_t Mhs = mad_lo(Mh, two, hm);
                hm = mul_hi(Mh, two);
                Pv[w] = Mhs | ~(Xv | Phs);
                Mv[w] = Phs & Xv;
                ph_below = Ph;
            }
            accP = __funnelshift_l(ph_below, accP, 1);
            accM = mad_lo(accM, two, hm);
        }
    }
    __device__ __forceinline__ void flush() { score += __popc(accP) - __popc(accM); accP = accM = 0u; }
};

template <int W, int V>
__global__ void __launch_bounds__(256) colbench(const uint32_t* __restrict__ peq_g, int peq_words,
                                               const uint32_t* __restrict__ tgt, int chunks, uint32_t two,
                                               uint32_t one, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t peq[];
    for (int i = threadIdx.x; i < peq_words * 4; i += blockDim.x) peq[i] = peq_g[i];
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    Band<W, V> B;
    B.init();
    int first = 0;
    for (int c = 0; c < chunks; ++c) {
        const uint32_t lo = tgt[(size_t)(2 * c) * stride + tid], hi = tgt[(size_t)(2 * c + 1) * stride + tid];
        const uint32_t* prow = peq + 4 * first;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const uint32_t cur = h ? hi : lo;
#pragma unroll
            for (int i = 0; i < 16; ++i) B.column(prow + ((cur >> (2 * i)) & 3u), two, one);
        }
        B.flush();
        first = (first + 1) % (peq_words - W);
    }
    uint32_t x = (uint32_t)B.score;
#pragma unroll
    for (int w = 0; w < W; ++w) x = x * 31u + B.Pv[w] * 7u + B.Mv[w];
    out[tid] = x;
}

template <int W, int V>
void run(const char* name, const uint32_t* d_peq, int peq_words, const uint32_t* d_tgt, int chunks, uint32_t* d_out,
         int grid) {
    const size_t smem = (size_t)peq_words * 16;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        colbench<W, V><<<grid, 256, smem>>>(d_peq, peq_words, d_tgt, chunks, 2u, 1u, d_out);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) best = ms < best ? ms : best;
    }
    std::vector<uint32_t> h((size_t)grid * 256);
    CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
    uint64_t sum = 0;
    for (uint32_t v : h) sum = sum * 1000003ull + v;
    const double wc = (double)grid * 256 * chunks * 32.0 * W;
    printf("%-28s W=%d  %8.3f ms  %7.1f G lane-word-columns/s  %9.1f GCUPS(band)  checksum %016llx\n", name, W, best,
           wc / best / 1e6, wc * 32 / best / 1e6, (unsigned long long)sum);
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    const int grid = p.multiProcessorCount * 2, chunks = 480, peq_words = 64;
    printf("%s, %d SMs, grid %d x 256\n", p.name, p.multiProcessorCount, grid);
    std::vector<uint32_t> peq((size_t)peq_words * 4), tgt((size_t)2 * chunks * grid * 256);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    for (int w = 0; w < peq_words; ++w) {   // a random query: every bit set in exactly one of the 4 masks
        uint32_t m[4] = {0, 0, 0, 0};
        for (int b = 0; b < 32; ++b) m[rnd() & 3] |= 1u << b;
        for (int c = 0; c < 4; ++c) peq[4 * w + c] = m[c];
    }
    for (auto& v : tgt) v = rnd() ^ (rnd() << 16);
    uint32_t *d_peq, *d_tgt, *d_out;
    CK(cudaMalloc(&d_peq, peq.size() * 4)); CK(cudaMalloc(&d_tgt, tgt.size() * 4)); CK(cudaMalloc(&d_out, (size_t)grid * 256 * 4));
    CK(cudaMemcpy(d_peq, peq.data(), peq.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tgt, tgt.data(), tgt.size() * 4, cudaMemcpyHostToDevice));
#define RUNW(W) \
    run<W, 0>("V0 alu shifts (shipped)", d_peq, peq_words, d_tgt, chunks, d_out, grid); \
    run<W, 1>("V1 imad+mulhi shifts", d_peq, peq_words, d_tgt, chunks, d_out, grid); \
    run<W, 2>("V2 mad.wide shifts", d_peq, peq_words, d_tgt, chunks, d_out, grid); \
    run<W, 3>("V3 V1 + mad.wide carry add", d_peq, peq_words, d_tgt, chunks, d_out, grid); \
    run<W, 4>("V4 only Mh shift on fma", d_peq, peq_words, d_tgt, chunks, d_out, grid); \
    run<W, 5>("V5 imad + imad.wide shifts", d_peq, peq_words, d_tgt, chunks, d_out, grid);
    RUNW(2) RUNW(4) RUNW(6) RUNW(8)
    return 0;
}
