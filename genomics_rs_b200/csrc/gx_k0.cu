// gx_k0.cu -- K0: INT32 / DPX issue-rate micro-benchmark (roofline denominator).
// Dependency-free unrolled chains, 32 warps per SM; every rate is  warp-instructions / SM cycles of the WHOLE launch,
// the launch timed with CUDA events and converted to cycles with the SM clock measured in the same call
// (clock64 against globaltimer).  (The round-1 version read clock64() per CTA, which over-counts: CTAs that start late
// see a shorter interval than the launch really took.)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gxalign.h"

namespace {

constexpr int CHAINS = 12;   // independent chains per thread
constexpr int ITERS = 8192;

template <int OP>
__global__ void __launch_bounds__(256) k0_kernel(int *sink, int seed) {
    int v[CHAINS], w[CHAINS];
    unsigned acc2[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        v[c] = seed + threadIdx.x * 7 + c;
        w[c] = seed * 3 + c * 5 + threadIdx.x;
        acc2[c] = 0u;
    }
    const int g = seed - 3, hg = seed - 9;
    const unsigned one = (unsigned)(seed > 0);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) {          // VIADDMNMX (ALU pipe)
                v[c] = __viaddmax_s32(v[c], -1, w[c]);
            } else if (OP == 1) {   // VIMNMX3 (ALU pipe)
                v[c] = __vimax3_s32(v[c], w[c], g);
            } else if (OP == 2) {   // IMAD (FMA pipe)
                asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(v[c]) : "r"(w[c]));
            } else if (OP == 3) {   // IDP.4A (the match/mismatch add of the one-hot path)
                v[c] = __dp4a(w[c], g, v[c]);
            } else if (OP == 4) {   // VIADDMNMX + IMAD: both pipes
                v[c] = __viaddmax_s32(v[c], -1, w[c]);
                asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(w[c]) : "r"(g));
            } else if (OP == 5) {   // the score-only cell: 3 ALU-pipe + 2 FMA-pipe instructions
                const int In = __viaddmax_s32(v[c], g, w[c]);
                const int Sn = __dp4a(w[c], hg, v[c]);
                const int Dn = __viaddmax_s32(w[c], g, In);
                const int Vn = __vimax3_s32(In, Dn, Sn);
                asm volatile("mad.lo.s32 %0, %1, 1, %2;" : "=r"(v[c]) : "r"(Vn), "r"(hg));
                w[c] = Dn;
            } else if (OP == 6) {   // the traceback cell: 5 ALU-pipe + 4 FMA-pipe instructions
                const int In = __viaddmax_s32(v[c], g, w[c]);
                const int Sn = __dp4a(w[c], hg, v[c]);
                const int Dn = __viaddmax_s32(w[c], g, In);
                const int Vn = __vimax3_s32(In, Dn, Sn);
                asm volatile("{\n\t.reg .pred p1, p2;\n\tsetp.ne.s32 p1, %1, %3;\n\tsetp.ne.and.s32 p2, %2, %3, p1;\n\t"
                             "@p1 mad.lo.u32 %0, %4, 4, %0;\n\t@p2 mad.lo.u32 %0, %4, 4, %0;\n\t}"
                             : "+r"(acc2[c]) : "r"(Sn), "r"(In), "r"(Vn), "r"(one));
                asm volatile("mad.lo.s32 %0, %1, 1, %2;" : "=r"(v[c]) : "r"(Vn), "r"(hg));
                w[c] = Dn;
            }
        }
    }
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= v[c] ^ w[c] ^ (int)acc2[c];
    if (acc == 0x7fffffff) sink[0] = acc;
}

// SM clock against wall clock: one warp spins for ~2 ms
__global__ void k0_calib(long long *out) {
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long c0 = clock64();
    long long c1 = c0;
    while (c1 - c0 < 4000000) c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) {
        out[0] = c1 - c0;
        out[1] = (long long)(g1 - g0);
    }
}

template <int OP>
double run(int sms, int *sink, double ghz, int per_iter, cudaEvent_t e0, cudaEvent_t e1) {
    const int cps = 4;   // 4 CTAs x 8 warps = 32 warps per SM
    const int grid = sms * cps;
    k0_kernel<OP><<<grid, 256>>>(sink, 11);  // warm-up
    cudaEventRecord(e0);
    k0_kernel<OP><<<grid, 256>>>(sink, 13);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = (double)ms * 1e-3 * ghz * 1e9;                           // SM cycles of the whole launch
    const double winstr = (double)ITERS * CHAINS * per_iter * 8 * cps;            // warp-instructions per SM
    return winstr / cycles;
}

}  // namespace

// out[0..] = warp-instructions per clock per SM of: VIADDMNMX, VIMNMX3, IMAD, IDP.4A, VIADDMNMX+IMAD (both pipes), the
// 5-instruction score-only cell, the 9-instruction traceback cell; then the SM clock (GHz) during the probe and the SM count.
extern "C" int gx_k0_measure(double *out, int n) {
    if (!out || n < 9) return -GX_ERR_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -GX_ERR_NO_DEVICE;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int *sink = nullptr;
    long long *cal = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess || cudaMalloc(&cal, 16) != cudaSuccess) return -GX_ERR_NOMEM;
    k0_calib<<<1, 32>>>(cal);
    k0_calib<<<1, 32>>>(cal);
    long long h[2] = {0, 1};
    if (cudaMemcpy(h, cal, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaFree(sink);
        cudaFree(cal);
        return -GX_ERR_CUDA;
    }
    const double ghz = (double)h[0] / (double)h[1];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    out[0] = run<0>(sms, sink, ghz, 1, e0, e1);
    out[1] = run<1>(sms, sink, ghz, 1, e0, e1);
    out[2] = run<2>(sms, sink, ghz, 1, e0, e1);
    out[3] = run<3>(sms, sink, ghz, 1, e0, e1);
    out[4] = run<4>(sms, sink, ghz, 2, e0, e1);
    out[5] = run<5>(sms, sink, ghz, 5, e0, e1);
    out[6] = run<6>(sms, sink, ghz, 9, e0, e1);
    out[7] = ghz;
    out[8] = sms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cal);
    return 9;
}
