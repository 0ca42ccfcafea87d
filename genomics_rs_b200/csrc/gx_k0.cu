// gx_k0.cu -- K0: INT32 / DPX issue-rate micro-benchmark (roofline denominator).
// Dependency-free unrolled chains, enough warps per SM to saturate the pipes; reports
// warp-instructions per clock per SM (x32 = lanes/clk/SM) measured with clock64() per CTA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gxalign.h"

namespace {

constexpr int CHAINS = 8;    // independent chains per thread
constexpr int ITERS = 4096;

template <int OP>
__global__ void __launch_bounds__(256) k0_kernel(int *sink, long long *cycles, int seed) {
    int v[CHAINS], w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        v[c] = seed + threadIdx.x * 7 + c;
        w[c] = seed * 3 + c * 5 + threadIdx.x;
    }
    const int g = seed - 3, hg = seed - 9, ap = seed + 5, bp = seed - 4;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) {  // IADD3 (kept on the ALU pipe by the xor the compiler cannot fold into IMAD)
                asm volatile("add.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(w[c]));
            } else if (OP == 1) {  // VIADDMNMX
                v[c] = __viaddmax_s32(v[c], g, w[c]);
            } else if (OP == 2) {  // VIMNMX3
                v[c] = __vimax3_s32(v[c], w[c], g);
                w[c] ^= it;  // not counted; keeps the chain from collapsing (1 LOP3 per VIMNMX3, see host)
            } else if (OP == 3) {  // ISETP + SEL
                v[c] = (v[c] == w[c]) ? ap : bp;
                w[c] += v[c];  // IADD, counted as third op of the triple
            } else if (OP == 4) {  // IMAD
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(g), "r"(w[c]));
            } else if (OP == 5) {  // one NW cell (7 ops): v = E diag / running state, w = D
                const int In = __viaddmax_s32(v[c], g, w[c]);
                const int Dn = __viaddmax_s32(w[c], g, v[c]);
                const int Sn = v[c] + ((it == w[c]) ? ap : bp);
                const int Vn = __vimax3_s32(In, Dn, Sn);
                v[c] = Vn + hg;
                w[c] = Dn;
            }
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= v[c] ^ w[c];
    if (acc == 0x7fffffff) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
double run(int sms, int *sink, long long *cycles_d, int ctas_per_sm) {
    const int grid = sms * ctas_per_sm;
    k0_kernel<OP><<<grid, 256>>>(sink, cycles_d, 11);  // warm-up
    k0_kernel<OP><<<grid, 256>>>(sink, cycles_d, 13);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
    long long *h = new long long[grid];
    cudaMemcpy(h, cycles_d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int k = 0; k < grid; ++k) avg += (double)h[k];
    avg /= grid;
    delete[] h;
    // warp-instructions issued per SM while those CTAs were co-resident
    const double winstr = (double)ITERS * CHAINS * (256 / 32) * ctas_per_sm;
    return winstr / avg;
}

}  // namespace

extern "C" int gx_k0_measure(double *out, int n) {
    if (!out || n < 8) return -GX_ERR_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -GX_ERR_NO_DEVICE;
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    int *sink = nullptr;
    long long *cyc = nullptr;
    const int cps = 4;  // 4 CTAs x 8 warps = 32 warps per SM
    if (cudaMalloc(&sink, 64) != cudaSuccess || cudaMalloc(&cyc, sizeof(long long) * sms * cps) != cudaSuccess) return -GX_ERR_NOMEM;
    out[0] = run<0>(sms, sink, cyc, cps);
    out[1] = run<1>(sms, sink, cyc, cps);
    out[2] = run<2>(sms, sink, cyc, cps) * 2.0;   // VIMNMX3 + LOP3 per iteration
    out[3] = run<3>(sms, sink, cyc, cps) * 3.0;   // ISETP + SEL + IADD per iteration
    out[4] = run<4>(sms, sink, cyc, cps);
    out[5] = run<5>(sms, sink, cyc, cps) * 32.0;  // cells per clock per SM (one cell per chain step per lane)
    out[6] = khz / 1000.0;
    out[7] = sms;
    cudaFree(sink);
    cudaFree(cyc);
    return 8;
}
