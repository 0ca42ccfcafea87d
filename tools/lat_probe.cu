// lat_probe.cu -- dependent-issue latencies (one warp, one SM) of the instructions on the DP row chain.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat_probe lat_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int UNR = 16;

template <int OP>
__global__ void __launch_bounds__(32) probe(int *sink, long long *cycles, int seed) {
    int v = seed + threadIdx.x, w = seed * 3 + threadIdx.x, u = seed * 5;
    const int g = seed - 3, hg = seed - 9;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < UNR; ++c) {
            if (OP == 0) v = __viaddmax_s32(v, g, w);                                   // VIADDMNMX -> VIADDMNMX
            if (OP == 1) v = __vimax3_s32(v, w, u);                                     // VIMNMX3 -> VIMNMX3
            if (OP == 2) asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(v) : "r"(w)); // IMAD -> IMAD
            if (OP == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(w));       // IADD3 -> IADD3
            if (OP == 4) {                                                              // classic cell: I -> V -> E(IMAD) -> I
                const int I = __viaddmax_s32(u, g, v);
                const int V = __vimax3_s32(I, w, hg);
                asm volatile("mad.lo.s32 %0, %1, 1, %2;" : "=r"(v) : "r"(V), "r"(hg));
                u = I;
            }
            if (OP == 5) {                                                              // classic cell with the add on the ALU pipe
                const int I = __viaddmax_s32(u, g, v);
                const int V = __vimax3_s32(I, w, hg);
                asm volatile("add.s32 %0, %1, %2;" : "=r"(v) : "r"(V), "r"(hg));
                u = I;
            }
            if (OP == 6) v = __shfl_up_sync(0xffffffffu, v, 1);                         // SHFL -> SHFL
            if (OP == 7) { v = __shfl_up_sync(0xffffffffu, v, 1); v = __viaddmax_s32(v, g, w); }   // SHFL -> VIADDMNMX -> SHFL
            if (OP == 8) v = max(v, w) + hg;                                            // VIMNMX -> IADD
            if (OP == 9) { v = __viaddmax_s32(v, g, w); w = __viaddmax_s32(w, hg, v); } // two alternating chains
        }
    }
    const long long t1 = clock64();
    if ((v ^ w ^ u) == 0x7fffffff) sink[0] = v;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
void run(const char *name, int ops, int *sink, long long *cyc) {
    probe<OP><<<1, 32>>>(sink, cyc, 11);
    probe<OP><<<1, 32>>>(sink, cyc, 13);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-52s %.2f clk per dependent op (%.2f per unit)\n", name, (double)h / ((double)ITERS * UNR * ops), (double)h / ((double)ITERS * UNR));
}

int main() {
    int *sink; long long *cyc;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 64);
    run<0>("VIADDMNMX -> VIADDMNMX", 1, sink, cyc);
    run<1>("VIMNMX3 -> VIMNMX3", 1, sink, cyc);
    run<2>("IMAD -> IMAD", 1, sink, cyc);
    run<3>("IADD -> IADD", 1, sink, cyc);
    run<4>("cell: VIADDMNMX -> VIMNMX3 -> IMAD (3 ops)", 3, sink, cyc);
    run<5>("cell: VIADDMNMX -> VIMNMX3 -> IADD (3 ops)", 3, sink, cyc);
    run<6>("SHFL.UP -> SHFL.UP", 1, sink, cyc);
    run<7>("SHFL.UP -> VIADDMNMX (2 ops)", 2, sink, cyc);
    run<8>("VIMNMX -> IADD (2 ops)", 2, sink, cyc);
    run<9>("VIADDMNMX <-> VIADDMNMX (2 ops)", 2, sink, cyc);
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
