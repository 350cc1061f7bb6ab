#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ inline void tm_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, int iters, long long* cyc) {
    __shared__ uint32_t tbase;
    int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t base = tbase;
    // this warp's lanes: 32*(warp%4); its private columns: (warp/4)*64
    uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
    uint32_t v[8];
    for (int j = 0; j < 8; j++) v[j] = tid * 100 + j;
    for (int c = 0; c < 3; c++) { uint32_t w[8]; for (int j = 0; j < 8; j++) w[j] = v[j] + 1000000 * c; tm_st8(taddr + 8 * c, w); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        for (int c = 0; c < 3; c++) {
            uint32_t w[8];
            tm_ld8(taddr + 8 * c, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; j++) { w[j] += 1; acc += w[j]; }
            tm_st8(taddr + 8 * c, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    uint32_t w[8];
    tm_ld8(taddr + 8, w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[blockIdx.x * 1024 + tid] = w[3];   // expect tid*100+3 + 1000000 + iters
    if (tid == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = acc; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}
int main() {
    uint32_t* d; long long* c; int nb = 148 * 2;
    cudaMalloc(&d, nb * 1024 * 4); cudaMalloc(&c, 16);
    int iters = 1000;
    k<<<nb, 1024>>>(d, iters, c);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    static uint32_t h[148 * 2 * 1024]; long long hc[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int b = 0; b < nb; b++) for (int t = 0; t < 1024; t++) if (h[b * 1024 + t] != (uint32_t)(t * 100 + 3 + 1000000 + iters)) { if (bad < 5) printf("bad b %d t %d got %u\n", b, t, h[b * 1024 + t]); bad++; }
    printf("bad %d, cycles per iter (3x ld8+st8 per thread, 1024 thr) %.1f\n", bad, (double)hc[0] / iters);
    return bad != 0;
}
