// Microbenchmark: issue cost of packed fp32 (FFMA2 / FMUL2 / FADD2) against scalar FFMA on sm_100a, alone and mixed with
// ALU-pipe integer work, at the step kernel's occupancy (one 512-thread CTA per SM = 4 warps per scheduler).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NALU>
__global__ void __launch_bounds__(512, 1) k(float2* x, unsigned* y, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float2 a[8]; float2 b = x[t], c = x[t + 1];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = x[t + 2 + i]; u[i] = y[t + i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }   // 2 FFMA
            else a[i] = __ffma2_rn(a[i], b, c);                                                       // 1 FFMA2
        }
#pragma unroll
        for (int j = 0; j < NALU; j++) u[j & 7] = (u[j & 7] ^ (u[(j + 1) & 7] >> 3)) + 0x9e3779b9u;   // LOP3/SHF/IADD3: alu pipe
    }
    float2 s = a[0]; unsigned v = u[0];
#pragma unroll
    for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; v ^= u[i]; }
    x[t] = s; y[t] = v;
}

template <int MODE, int NALU> float run(float2* x, unsigned* y, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NALU><<<148, 512>>>(x, y, 16);
    cudaEventRecord(e0);
    k<MODE, NALU><<<148, 512>>>(x, y, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float2* x; unsigned* y; cudaMalloc(&x, 148 * 512 * 16 * sizeof(float2)); cudaMalloc(&y, 148 * 512 * 16 * 4);
    cudaMemset(x, 0, 148 * 512 * 16 * sizeof(float2)); cudaMemset(y, 1, 148 * 512 * 16 * 4);
    const int it = 1 << 16;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
#define R(M, N) { float ms = run<M, N>(x, y, it); double cyc = ms * 1e-3 * clk * 1e3 / it; \
    printf("mode %s  alu/iter %2d : %8.3f ms  %6.2f cycles/iter/SM  (16 fp32 FMA-lanes x warp: %.2f warp-FMA-pairs/clk/SMSP)\n", M ? "FFMA2" : "FFMA ", N, ms, cyc, 8.0 * 4 / cyc); }
    R(0, 0) R(1, 0) R(0, 8) R(1, 8) R(0, 16) R(1, 16) R(0, 32) R(1, 32)
    printf("clock %d kHz\n", clk);
    return 0;
}
