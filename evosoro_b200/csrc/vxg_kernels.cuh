// vxg_kernels.cuh -- sm_100a kernels of the Voxelyze timestep engine.
//
//   vx_step_kernel<Real, NV, VPT, SMEM, COMPACT>  one CTA per robot; advances the robot `n_steps` timesteps.
//       NV           : voxel stride of the blob (compile-time: every array offset is an immediate);
//                     the CTA has NV / VPT threads.
//       SMEM = true : the robot blob is staged into shared memory once, all steps run on chip, the
//                     dynamic section is written back at the end (HBM traffic = blob in + blob out per
//                     LAUNCH, not per step).
//       SMEM = false: same code against the blob in global memory (robots larger than ~227 KB of
//                     state, e.g. the fp64 mode at 10^3); the working set of the 148 resident robots
//                     lives in L2.
//       COMPACT      : bond classes in the two-vector compact form (BondKc, vx_math.cuh) instead of the wide table; this
//                     production family also leaves out the boundary-condition code (pack_compact sends such robots to the wide one).
//                     Small robots run with fewer threads per CTA; the launch bounds ask for 512 resident threads per SM
//                     whatever the CTA size (<= 128 registers per thread), so that several small CTAs share an SM.
//       VPT          : voxels per thread (thread t owns voxels t, t+T, ...), so a 10^3 robot (NV = 1024) runs
//                     with 512 threads x 2 voxels and <=128 registers per thread.  Robots above 2048 voxels (strides 8192 and
//                     32768, blob in global memory) run 16 / 64 voxels per thread as rolled loops: the per-voxel working set of a
//                     step (own state, force sums, pending pushes) is then a thread-local array, not registers -- a functional
//                     path for sizes the reference has no limit on (VX_Sim.cpp:530-600), not a tuned one.
//   vx_stats_kernel<Real>            one CTA per robot; reductions for the result writer.
//
// Replaces CVX_Sim::Integrate (VX_Sim.cpp:1763-1933); see vx_robot.cuh / vx_math.cuh for the
// per-bond and per-voxel arithmetic and the reference lines they restate.
#pragma once
#include <cuda_runtime.h>
#include "vx_robot.cuh"

namespace vx {

#ifndef VX_CTAS_512
#define VX_CTAS_512 2
#endif
#define VXG_ST_DIVERGED_ 0x1u
#define VXG_ST_FINISHED_ 0x2u
#define VXG_F_FLOOR_ 0x0001u
#define VXG_F_TEMPERATURE_ 0x0002u
#define VXG_F_LW_SCALE_ 0x0004u
#define VXG_F_STICKY_FLOOR_ 0x0008u
#define VXG_F_FLUID_ 0x0010u
#define VXG_F_COLLISIONS_ 0x0020u
#define VXG_ST_COL_OVERFLOW_ 0x4u

template <typename Real> struct BatchDev {
    unsigned char* blobs;       // [n_robots][L.total_bytes]
    BlobLayout L;
    RobotCtl* ctl;              // [n_robots]
    unsigned char* bk;          // [n_robots][max_bclass] BondK<Real> (wide) or BondKc<Real> (compact) records
    double* bk_dist;            // [n_robots][max_bclass]
    VoxK<Real>* vk;             // [n_robots][max_vclass]
    ScaleK* sk;                 // [n_robots][max_vclass]
    double* scale_c;            // [n_robots][max_vclass]  Scale per class as left by the last step
    uint32_t* bpair;            // [n_robots][max_bclass] voxel-class pair of each re-keyed bond class (compact family)
    Real* ext_disp;             // [n_robots][max_vclass][3]
    Real* ext_tdisp;            // [n_robots][max_vclass][3]
    double* nom;                // [n_robots][3][nv]
    Real* bond_rec;             // [n_robots][39][nv] or nullptr
    int max_vclass, max_bclass, n_robots;
    // self-collision: contact-bond list (persists between launches) and the static broad-phase tables
    ColPair* col_pairs;         // [n_robots][max_col]
    int16_t* col_surf;          // [n_robots][nv]
    int8_t* col_lat4;           // [n_robots][nv][4]
    unsigned long long* col_near;   // [n_robots][nv]
    int max_col;
};

// shared-memory carve-up after the blob
template <typename Real> struct SmemTables {
    unsigned char* bk;   // BondK<Real> or BondKc<Real> records
    double* bk_dist; VoxK<Real>* vk; ScaleK* sk;
    double* scale_c;  // [2][max_vclass]
    Real* dsc_c;      // [2][max_vclass]
    Real* ext_disp; Real* ext_tdisp;
    double* red;      // [4][32] reduction scratch
    unsigned long long* mbar;   // split-phase barrier between the axis sub-phases
    unsigned int* maxvel;       // [2] max |Vel|^2 of a step as float bits (double-buffered by step parity)
    unsigned int* arrcnt;       // warps that have arrived at the current axis barrier (VX_BAR_MODE 2)
    unsigned int* colflag;      // [2] thread 0's verdict for this step: rebuild the contact-bond list (by step parity)
    double* colstate;           // [4] MaxDispSinceLastBondUpdate, SS.MaxVoxVel, CollisionHorizon, lattice pitch: thread 0's working set,
                                //     kept out of every other thread's registers
    uint32_t* bpair;            // [max_bclass] voxel-class pair of each bond class (compact family)
    Real* dnom_bc;              // [2][max_bclass] rest-length deviation per bond class, double-buffered like dsc_c
    ScaleSim* sl;               // the robot's scale-law constants (read by the few threads that evaluate the law)
    ColPair* pairs;             // [max_col] contact bonds
};

template <typename Real>
__host__ __device__ inline size_t tables_bytes(int max_vclass, int max_bclass, bool compact = false, int max_col = 0) {
    size_t o = 0;
    o = vx_align16(o + (compact ? sizeof(BondKc<Real>) : sizeof(BondK<Real>)) * max_bclass);
    o = vx_align16(o + sizeof(double) * max_bclass);
    o = vx_align16(o + sizeof(VoxK<Real>) * max_vclass);
    o = vx_align16(o + sizeof(ScaleK) * max_vclass);
    o = vx_align16(o + sizeof(double) * 2 * max_vclass);
    o = vx_align16(o + sizeof(Real) * 2 * max_vclass);
    o = vx_align16(o + sizeof(Real) * 3 * max_vclass);
    o = vx_align16(o + sizeof(Real) * 3 * max_vclass);
    o = vx_align16(o + sizeof(double) * 4 * 32);
    o = vx_align16(o + sizeof(unsigned long long) + 6 * sizeof(unsigned int));
    o = vx_align16(o + 4 * sizeof(double));
    o = vx_align16(o + sizeof(ScaleSim));
    if (compact) { o = vx_align16(o + sizeof(uint32_t) * max_bclass); o = vx_align16(o + sizeof(Real) * 2 * max_bclass); }
    o = vx_align16(o + sizeof(ColPair) * max_col);
    return o;
}

template <typename Real>
__device__ inline SmemTables<Real> carve_tables(unsigned char* p, int max_vclass, int max_bclass, bool compact, int max_col) {
    SmemTables<Real> t; size_t o = 0;
    t.bk = p + o;                   o = vx_align16(o + (compact ? sizeof(BondKc<Real>) : sizeof(BondK<Real>)) * max_bclass);
    t.bk_dist = (double*)(p + o);   o = vx_align16(o + sizeof(double) * max_bclass);
    t.vk = (VoxK<Real>*)(p + o);    o = vx_align16(o + sizeof(VoxK<Real>) * max_vclass);
    t.sk = (ScaleK*)(p + o);        o = vx_align16(o + sizeof(ScaleK) * max_vclass);
    t.scale_c = (double*)(p + o);   o = vx_align16(o + sizeof(double) * 2 * max_vclass);
    t.dsc_c = (Real*)(p + o);       o = vx_align16(o + sizeof(Real) * 2 * max_vclass);
    t.ext_disp = (Real*)(p + o);    o = vx_align16(o + sizeof(Real) * 3 * max_vclass);
    t.ext_tdisp = (Real*)(p + o);   o = vx_align16(o + sizeof(Real) * 3 * max_vclass);
    t.red = (double*)(p + o);       o = vx_align16(o + sizeof(double) * 4 * 32);
    t.mbar = (unsigned long long*)(p + o);
    t.maxvel = (unsigned int*)(p + o + sizeof(unsigned long long));
    t.colflag = t.maxvel + 2;
    t.arrcnt = t.maxvel + 4;
    o = vx_align16(o + sizeof(unsigned long long) + 6 * sizeof(unsigned int));
    t.colstate = (double*)(p + o);  o = vx_align16(o + 4 * sizeof(double));
    t.sl = (ScaleSim*)(p + o);      o = vx_align16(o + sizeof(ScaleSim));
    t.bpair = nullptr; t.dnom_bc = nullptr;
    if (compact) {
        t.bpair = (uint32_t*)(p + o);   o = vx_align16(o + sizeof(uint32_t) * max_bclass);
        t.dnom_bc = (Real*)(p + o);     o = vx_align16(o + sizeof(Real) * 2 * max_bclass);
    }
    t.pairs = (ColPair*)(p + o);
    return t;
}

__device__ inline void copy16(void* dst, const void* src, size_t bytes, int tid, int nthreads) {
    const uint4* s = (const uint4*)src; uint4* d = (uint4*)dst;
    size_t n = bytes >> 4;
    for (size_t i = tid; i < n; i += nthreads) d[i] = s[i];
}
__device__ inline void copy4(void* dst, const void* src, size_t bytes, int tid, int nthreads) {
    const uint32_t* s = (const uint32_t*)src; uint32_t* d = (uint32_t*)dst;
    size_t n = bytes >> 2;
    for (size_t i = tid; i < n; i += nthreads) d[i] = s[i];
}

// Split-phase CTA barrier (mbarrier): a warp ARRIVES once its accumulator pushes of one axis are done and
// only WAITS right before its pushes of the next axis, so the whole bond evaluation in between hides the
// barrier latency.  One arrival per warp (after __syncwarp, which orders the other lanes' stores).
__device__ inline unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ inline void mbar_arrive_warp(unsigned long long* bar, bool lane0) {
    __syncwarp();
    if (lane0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint, in ns, runs out)
// instead of polling -- without the hint a waiting warp re-issues try_wait + branch ~20 times per wait and takes those
// issue slots from the warps that still work (8 % of all executed instructions in the 1024-thread kernel, ncu)
__device__ inline void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VX_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra VX_MBAR_DONE;\n"
        "bra VX_MBAR_WAIT;\n"
        "VX_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}

// VX_BAR_MODE selects how the axis sub-phases are ordered (tuning switch; 0 = split-phase mbarrier with one arrival per warp):
//   1 : plain CTA barrier at the wait point (hardware BAR.SYNC, no split phase)
//   2 : split phase with ONE mbarrier arrival per phase: the warps count themselves in on a shared-memory counter and the last
//       one arrives, so that sleeping waiters are woken once per phase instead of once per arriving warp
#ifndef VX_BAR_MODE
#define VX_BAR_MODE 0
#endif
__device__ inline void axis_arrive(unsigned long long* bar, unsigned int* cnt, int nwarps, bool lane0) {
#if VX_BAR_MODE == 1
    (void)bar; (void)cnt; (void)nwarps; (void)lane0;
#elif VX_BAR_MODE == 2
    __syncwarp();
    if (lane0) {
        __threadfence_block();
        const unsigned int old = atomicAdd(cnt, 1u);
        if (old == (unsigned int)(nwarps - 1)) {
            *cnt = 0u;
            __threadfence_block();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
    }
#else
    (void)cnt; (void)nwarps;
    mbar_arrive_warp(bar, lane0);
#endif
}
__device__ inline void axis_wait(unsigned long long* bar, unsigned int parity) {
#if VX_BAR_MODE == 1
    (void)bar; (void)parity;
    __syncthreads();
#else
    mbar_wait(bar, parity);
#endif
}

// ---- bulk asynchronous copies (TMA engine, 1-D form): a robot's sections travel between HBM and shared memory as a few large
// transfers issued by one thread and tracked by an mbarrier, instead of a loop of 128-bit loads and stores through registers.
__device__ inline void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ inline void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned int bytes, unsigned long long* bar) {   // bytes % 16 == 0, 16 B aligned
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ inline void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ inline void bulk_s2g_finish() {   // the source may be reused / the CTA may exit once the engine has read it
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ inline void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // generic-proxy writes -> async proxy

__device__ inline double warp_max(double x) {
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ inline double warp_sum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block-wide sum of 4 doubles; result valid in every thread. `red` = [4][32] shared doubles.
__device__ inline void block_sum4(double& a, double& b, double& c, double& d, double* red) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
    __syncthreads();
    if (lane == 0) { red[wid] = a; red[32 + wid] = b; red[64 + wid] = c; red[96 + wid] = d; }
    __syncthreads();
    a = b = c = d = 0;
    for (int i = 0; i < nw; i++) { a += red[i]; b += red[32 + i]; c += red[64 + i]; d += red[96 + i]; }
}

// CVX_Sim::CalcL1Bonds (VX_Sim.cpp:2357-2413): rebuild the contact-bond list from the current positions.
// The surface-voxel pair triangle (i < j), flattened in the reference's creation order (i, then j), is cut into T equal runs of
// pairs, one per thread -- not into runs of rows: row i has S - 1 - i pairs, and with about one row per thread (S ~ T for the GA's
// robots) the thread that owned row 0 tested three times the average number of pairs while half the CTA waited.  Two passes
// (count, block-wide exclusive scan in thread order, write) keep the list in creation order.
// Returns the number of contact bonds found (the caller treats > cap as an overflow error).
template <typename Real>
__device__ __noinline__ int col_rebuild(const RobotView<Real>& R, const RobotConsts<Real>& C, const ColTables& ct, Real horizon, Real lattice,
                                        ColPair* pairs, int cap, int* scratch /* [33] ints of shared memory */) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    const Real filt = horizon * (Real)1.5 * lattice;
    const Real filt2 = filt * filt;
    const int S = ct.n_surf;
    const long long P = (long long)S * (S - 1) / 2;
    const int p0 = (int)(P * tid / T), p1 = (int)(P * (tid + 1) / T);
    // row of pair p0: the largest i with i (2S - i - 1) / 2 <= p0
    int i0 = 0, j0 = 1;
    if (p0 < p1) {
        const double b = 2.0 * S - 1.0;
        i0 = (int)((b - sqrt(b * b - 8.0 * (double)p0)) * 0.5);
        if (i0 < 0) i0 = 0;
        while (i0 > 0 && (long long)i0 * (2 * S - i0 - 1) / 2 > p0) i0--;
        while ((long long)(i0 + 1) * (2 * S - i0 - 2) / 2 <= p0) i0++;
        j0 = i0 + 1 + (p0 - (int)((long long)i0 * (2 * S - i0 - 1) / 2));
    }
    int cnt = 0;
    {
        int i = i0, j = j0, vi = p0 < p1 ? ct.surf[i0] : 0;
        for (int p = p0; p < p1; p++) {
            cnt += col_pair_hit<Real>(R, C, ct, vi, ct.surf[j], horizon, filt2) ? 1 : 0;
            if (++j == S) { i++; j = i + 1; if (p + 1 < p1) vi = ct.surf[i]; }
        }
    }
    int inc = cnt;   // inclusive scan across the block, in thread (= creation) order
    for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    __syncthreads();
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < nw; w++) { int x = scratch[w]; if (w < wid) base += x; total += x; }
    int off = base + inc - cnt;
    if (cnt > 0) {
        int i = i0, j = j0, vi = ct.surf[i0];
        for (int p = p0; p < p1; p++) {
            const int vj = ct.surf[j];
            if (col_pair_hit<Real>(R, C, ct, vi, vj, horizon, filt2)) { if (off < cap) { ColPair q; q.v1 = (int16_t)vi; q.v2 = (int16_t)vj; pairs[off] = q; } off++; }
            if (++j == S) { i++; j = i + 1; if (p + 1 < p1) vi = ct.surf[i]; }
        }
    }
    __syncthreads();
    return total;
}

// Resident CTAs per SM the register allocation is sized for.  Stride 1024 and up: one robot fills the SM's shared memory -> 1
// (128 registers at 512 threads, 64 at 1024).  Stride 512: two robots fit -> 2 x 512 threads, 64 registers (the fluid and
// self-collision variants need more registers and room: 1).  Smaller robots: 768 resident threads, i.e. 80 registers -- what one voxel per thread needs without spilling.
__host__ __device__ constexpr int vx_min_ctas(int nv, int threads, bool extras, bool fp64) {   // extras: fluid / self-collision variants
    return fp64 ? 1                                        // the logic-parity mode needs its registers, not occupancy
         : nv > 512 ? 1                                    // one robot per SM (shared memory)
         : nv == 512 ? (extras ? 1 : VX_CTAS_512)
         : (threads == 224 && !extras) ? 4                 // 4 x 224 threads: 72 registers
         : 768 / threads;
}
//       FLUID        : blob layout with the land/water surface-mesh arrays; robots with VXG_F_FLUID get the drag phase.
//       COLLIDE      : self-collision support compiled in (contact-bond list, rebuild, max-velocity reduction); kept out
//                     of the other instantiations so that they do not pay its registers.
template <typename Real, int NV, int VPT, bool SMEM, bool COMPACT, bool FLUID = false, bool COLLIDE = false>
__global__ void __launch_bounds__(NV / VPT, vx_min_ctas(NV, NV / VPT, FLUID || COLLIDE, sizeof(Real) == 8)) vx_step_kernel(BatchDev<Real> B, int n_steps, int record_bonds) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = NV / VPT;
    constexpr bool HOIST = !(COMPACT && (NV == 1024 || NV == 512));   // bond_force: as in the gather form of the same stride, bit for bit
    constexpr BlobLayout LC = make_layout(NV, (int)sizeof(Real));
    const BlobLayout L = FLUID ? B.L : LC;   // with the mesh arrays the capacity is a run-time size
    const int robot = blockIdx.x;
    int tid;   // read once: an asm the compiler cannot re-materialise (it otherwise re-reads %tid.x ~9x per voxel-step)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    RobotCtl* ctlp = B.ctl + robot;
    const unsigned int status0 = ctlp->status;
    if (status0 & (VXG_ST_DIVERGED_ | VXG_ST_FINISHED_)) return;
    const long long step0 = ctlp->step, total = ctlp->total_steps, latch = ctlp->latch_step, eol = ctlp->eol_step;
    long long remaining = total - step0;
    int nsteps = remaining < (long long)n_steps ? (int)remaining : n_steps;
    if (nsteps <= 0) { if (tid == 0) ctlp->status = status0 | VXG_ST_FINISHED_; return; }

    unsigned char* gblob = B.blobs + (size_t)robot * L.total_bytes;
    unsigned char* base = SMEM ? smem : gblob;
    unsigned char* tab = SMEM ? smem + L.total_bytes : smem;
    SmemTables<Real> S = carve_tables<Real>(tab, B.max_vclass, B.max_bclass, COMPACT, B.max_col);
    const int n_vox = ctlp->n_vox, n_vc = ctlp->n_vclass, n_bc = ctlp->n_bclass;
    const unsigned int flags = ctlp->flags;

    // ---- stage in ---------------------------------------------------------------------------
    if (SMEM) {
        copy16(smem, gblob, L.stat_bytes, tid, T);
        // zero the accumulators
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* a = (uint4*)(smem + L.off_accA);
        for (size_t i = tid; i < ((L.total_bytes - L.off_accA) >> 4); i += T) a[i] = z;
    }
    constexpr size_t BREC = COMPACT ? sizeof(BondKc<Real>) : sizeof(BondK<Real>);
    copy4(S.bk, B.bk + (size_t)robot * B.max_bclass * BREC, BREC * n_bc, tid, T);
    copy4(S.vk, B.vk + (size_t)robot * B.max_vclass, sizeof(VoxK<Real>) * n_vc, tid, T);
    copy4(S.sk, B.sk + (size_t)robot * B.max_vclass, sizeof(ScaleK) * n_vc, tid, T);
    for (int i = tid; i < n_bc; i += T) S.bk_dist[i] = B.bk_dist[(size_t)robot * B.max_bclass + i];
    for (int i = tid; i < n_vc; i += T) {
        S.scale_c[i] = B.scale_c[(size_t)robot * B.max_vclass + i];
    }
    for (int i = tid; i < 3 * n_vc; i += T) {
        S.ext_disp[i] = B.ext_disp[(size_t)robot * B.max_vclass * 3 + i];
        S.ext_tdisp[i] = B.ext_tdisp[(size_t)robot * B.max_vclass * 3 + i];
    }
    if (tid == 0) { mbar_init(S.mbar, VX_BAR_MODE == 2 ? 1 : T / 32); *S.arrcnt = 0u; }
    __syncthreads();
    for (int i = tid; i < n_vc; i += T) S.dsc_c[i] = (Real)(S.scale_c[i] - S.sk[i].nom_size);
    if (COMPACT) for (int i = tid; i < n_bc; i += T) S.bpair[i] = B.bpair[(size_t)robot * B.max_bclass + i];
    __syncthreads();
    if (COMPACT) {
        for (int i = tid; i < n_bc; i += T) { const uint32_t pr = S.bpair[i]; S.dnom_bc[i] = (Real)0.5 * (S.dsc_c[pr & 0xFFFFu] + S.dsc_c[pr >> 16]); }
        __syncthreads();
    }
    unsigned int mpar = 0;

    RobotView<Real> R = make_view<Real>(base, L, n_vox);
    RobotConsts<Real> C;
    C.bk = COMPACT ? nullptr : (const BondK<Real>*)S.bk;
    C.bkc = COMPACT ? (const BondKc<Real>*)S.bk : nullptr;
    C.geo = make_geo<Real>(ctlp->bond_len);
    C.bk_dist = S.bk_dist; C.vk = S.vk;
    C.nom = B.nom + (size_t)robot * 3 * NV;
    C.ext_disp = S.ext_disp; C.ext_tdisp = S.ext_tdisp;
    C.half_bond_z = (Real)ctlp->half_bond_z;
    const double dt = ctlp->dt;
    C.dt = (Real)dt;
    const bool fluid = (flags & VXG_F_FLUID_) != 0;
    C.fluid = FLUID && fluid; C.drag_coeff = (Real)ctlp->drag_coeff; C.n_vert = ctlp->n_vert; C.n_facet = ctlp->n_facet;
    C.floor_on = (flags & VXG_F_FLOOR_) && !fluid;
    C.sticky = (flags & VXG_F_STICKY_FLOOR_) != 0;
    C.sym_mass = ctlp->sym_mass != 0;
    const bool temp_on = (flags & VXG_F_TEMPERATURE_) != 0, lw = (flags & VXG_F_LW_SCALE_) != 0;
    if (tid == 0) {
        ScaleSim SL;
        SL.init_cm_time = ctlp->init_cm_time; SL.temp_period = ctlp->temp_period; SL.temp_amplitude = ctlp->temp_amplitude;
        SL.min_fact = ctlp->min_fact; SL.max_fact = ctlp->max_fact; SL.growth_amp = ctlp->growth_amp;
        SL.midlife_freeze = ctlp->midlife_freeze; SL.stop_value = ctlp->stop_value; SL.temp_on = temp_on; SL.lw = lw;
        *S.sl = SL;
    }
    __syncthreads();
    const Real dt_inv_run = (Real)(1.0 / dt);
    // bond outputs are recorded by the wide-table families only (the engine switches a recording batch to them): the production
    // family carries neither the pointer nor the per-bond test
    Real* rec_base = (!COMPACT && record_bonds && B.bond_rec) ? B.bond_rec + (size_t)robot * 39 * NV : nullptr;

    // self-collision state (CVX_Sim::UpdateCollisions): the list lives in shared memory during the launch
    const bool col_on = COLLIDE && (flags & VXG_F_COLLISIONS_) != 0 && B.max_col > 0;
    ColTables CT;
    CT.surf = B.col_surf + (size_t)robot * NV; CT.lat4 = B.col_lat4 + (size_t)robot * NV * 4; CT.near = B.col_near + (size_t)robot * NV;
    CT.n_surf = ctlp->n_surf; CT.hops = ctlp->col_hops;
    int n_col = col_on ? ctlp->n_col : 0;
    bool col_first = ctlp->col_first != 0;
    const bool col_host = ctlp->col_host != 0;
    bool col_overflow = false;
    if (COLLIDE && col_on) {
        for (int i = tid; i < n_col; i += T) S.pairs[i] = B.col_pairs[(size_t)robot * B.max_col + i];
        if (tid == 0) {
            S.maxvel[0] = 0u; S.maxvel[1] = 0u; S.colflag[0] = 0u; S.colflag[1] = 0u;
            S.colstate[0] = ctlp->col_acc; S.colstate[1] = ctlp->max_vel; S.colstate[2] = ctlp->col_horizon; S.colstate[3] = ctlp->lattice_dim;
        }
        __syncthreads();
    }

    double t = ctlp->cur_time;
    int cur = 0;
    int done = 0;
    bool diverged = false;
    for (int k = 0; k < nsteps; ++k) {
        const long long s = step0 + k;
        if (s == latch) {   // IniCM = SS.CurCM at TimeStep entry (VX_Sim.cpp:1064-1075): COM after the previous step
            double mx = 0, my = 0, mz = 0, mm = 0;
            for (int v = tid; v < n_vox; v += T) {
                double m = (double)S.vk[R.stat[v].vcls].mass;
                mx += R.pos[v] * m; my += R.pos[NV + v] * m; mz += R.pos[2 * NV + v] * m; mm += m;
            }
            block_sum4(mx, my, mz, mm, S.red);
            if (tid == 0) { ctlp->ini_com[0] = mx / mm; ctlp->ini_com[1] = my / mm; ctlp->ini_com[2] = mz / mm; ctlp->ini_latched = 1; }
        }
        if (s == eol) {     // SS.EndOfLifetimePosteriorY = getPosteriorY() at TimeStep entry (VX_Sim.cpp:1078-1082)
            double lo = 1e300;
            for (int v = tid; v < n_vox; v += T) if (S.sk[R.stat[v].vcls].mat_index != 5) lo = fmin(lo, R.pos[NV + v]);
            lo = -warp_max(-lo);
            __syncthreads();
            if ((tid & 31) == 0) S.red[tid >> 5] = lo;
            __syncthreads();
            if (tid == 0) { for (int i = 1; i < T / 32; i++) lo = fmin(lo, S.red[i]); ctlp->eol_post_y = lo; }
        }
        C.scale_c = S.scale_c + cur * B.max_vclass;
        C.dsc_c = S.dsc_c + cur * B.max_vclass;
        if (COMPACT) C.dnom_bc = S.dnom_bc + cur * B.max_bclass;
        const Real dt_inv = (s == 0) ? (Real)0 : dt_inv_run;   // dt == 0 during the very first bond pass (VX_Sim.cpp:373)
        Real* rec = (!COMPACT && k == nsteps - 1) ? rec_base : nullptr;

        if (COLLIDE && col_on && tid == 0) {
            // UpdateCollisions (VX_Sim.cpp:1729-1755), the bookkeeping: one thread keeps MaxDispSinceLastBondUpdate in fp64 and posts
            // its verdict; the CTA acts on it after the first axis barrier below (dt is still 0 in the very first step)
            const unsigned int mvb = S.maxvel[(k + 1) & 1];            // written by the previous step's voxel phase
            double col_acc = S.colstate[0], max_vel = S.colstate[1];
            if (k > 0) max_vel = sqrt((double)__uint_as_float(mvb));
            S.maxvel[k & 1] = 0u;                                      // this step's slot; next read after two barriers
            col_acc += fabs(max_vel * (s == 0 ? 0.0 : dt) / S.colstate[3]);
            const bool rebuild = !col_host && (col_acc > (S.colstate[2] - 1.0) / 2 || col_first);
            if (rebuild) { col_first = false; col_acc = 0.0; }
            S.colstate[0] = col_acc; S.colstate[1] = max_vel;
            S.colflag[k & 1] = rebuild ? 1u : 0u;
        }
        V3<Real> F[VPT], M[VPT];
        VoxStat st[VPT];
        OwnState<Real> own[VPT];
        BondPush<Real> P[VPT];
        uint32_t fl[VPT];
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) {
            F[j] = mk3<Real>(0, 0, 0); M[j] = mk3<Real>(0, 0, 0);
            st[j] = R.stat[tid + j * T];
            fl[j] = R.flags[tid + j * T];
            own[j] = load_own<Real>(R, tid + j * T);
        }
        if (FLUID && C.fluid) {   // DragForce of this step from the state and bond strains of the previous one
            for (int u = tid; u < C.n_vert; u += T) mesh_vertex<Real>(R, C, u);
            __syncthreads();
            for (int i = tid; i < C.n_facet; i += T) facet_task<Real>(R, C, i);
            __syncthreads();
#pragma unroll(VPT <= 4 ? VPT : 1)
            for (int j = 0; j < VPT; j++) if (tid + j * T < n_vox) F[j] = voxel_drag<Real>(R, tid + j * T, st[j]);
        }
        // Scale of every class for the NEXT step, from CurTime before the increment (VXS_Voxel.cpp:224-336).
        // It depends on time only, so it is evaluated here, off the critical path, by the highest threads
        // of the CTA (the last warp owns the fewest voxels); it lands in the other half of the double buffer,
        // which nobody reads before the end-of-step barrier.
        for (int i = T - 1 - tid; i < n_vc; i += T) {
            double last = S.scale_c[cur * B.max_vclass + i];
            double ns = scale_law(S.sk[i], last, t, *S.sl);
            S.scale_c[(cur ^ 1) * B.max_vclass + i] = ns;
            S.dsc_c[(cur ^ 1) * B.max_vclass + i] = (Real)(ns - S.sk[i].nom_size);
        }
        // Bond phase.  Evaluations read only state of the previous step; the pushes of one axis must be
        // complete before the pushes of the next axis start (same accumulator, different writer), which
        // the split-phase barrier orders without stalling the evaluations.
        int div = 0;
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) { P[j].w = -1; if (tid + j * T < n_vox) div |= bond_eval<Real, 0, COMPACT, FLUID, HOIST>(R, C, tid + j * T, st[j], own[j], fl[j], dt_inv, F[j], M[j], rec, P[j]) ? 1 : 0; }
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) bond_push<Real>(R, P[j]);
        axis_arrive(S.mbar, S.arrcnt, T / 32, (tid & 31) == 0);
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) { P[j].w = -1; if (tid + j * T < n_vox) div |= bond_eval<Real, 1, COMPACT, FLUID, HOIST>(R, C, tid + j * T, st[j], own[j], fl[j], dt_inv, F[j], M[j], rec, P[j]) ? 1 : 0; }
        axis_wait(S.mbar, mpar); mpar ^= 1u;
        if (COLLIDE && col_on && S.colflag[k & 1]) {   // rebuild the contact-bond list (positions are still those of the previous step)
            C.scale_c = S.scale_c + cur * B.max_vclass;
            // (copies: the out-of-line rebuild takes its views by reference; the step loop's own R and C stay in registers)
            const RobotView<Real> R2 = R; const RobotConsts<Real> C2 = C; const ColTables CT2 = CT;
            n_col = col_rebuild<Real>(R2, C2, CT2, (Real)S.colstate[2], (Real)S.colstate[3], S.pairs, B.max_col, (int*)S.red);
            if (n_col > B.max_col) { col_overflow = true; break; }
        }
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) bond_push<Real>(R, P[j]);
        axis_arrive(S.mbar, S.arrcnt, T / 32, (tid & 31) == 0);
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) { P[j].w = -1; if (tid + j * T < n_vox) div |= bond_eval<Real, 2, COMPACT, FLUID, HOIST>(R, C, tid + j * T, st[j], own[j], fl[j], dt_inv, F[j], M[j], rec, P[j]) ? 1 : 0; }
        axis_wait(S.mbar, mpar); mpar ^= 1u;
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) bond_push<Real>(R, P[j]);
        if (COLLIDE && col_on && n_col > 0) {   // contact bonds: positions are still those of the previous step here
#pragma unroll(VPT <= 4 ? VPT : 1)
            for (int j = 0; j < VPT; j++) if (tid + j * T < n_vox) F[j] = F[j] + col_force_on<Real>(R, C, tid + j * T, st[j], own[j], S.pairs, n_col);
        }
        if (__syncthreads_or(div)) { diverged = true; break; }   // Integrate() returns before the voxel loop (VX_Sim.cpp:1777)

        if (COMPACT) {   // rest-length deviation of every bond class for the NEXT step: the classes' new Scale was written before the barrier
            for (int i = tid; i < n_bc; i += T) {
                const uint32_t pr = S.bpair[i];
                const Real* dn = S.dsc_c + (cur ^ 1) * B.max_vclass;
                S.dnom_bc[(cur ^ 1) * B.max_bclass + i] = (Real)0.5 * (dn[pr & 0xFFFFu] + dn[pr >> 16]);
            }
        }
        Real sp2 = (Real)0;
#pragma unroll(VPT <= 4 ? VPT : 1)
        for (int j = 0; j < VPT; j++) {
            int v = tid + j * T;
            if (v < n_vox) { Real x = voxel_task<Real, !COMPACT>(R, C, v, st[j], own[j], fl[j], F[j], M[j]); sp2 = x > sp2 ? x : sp2; }
        }
        if (COLLIDE && col_on) {   // SS.MaxVoxVel for the next UpdateCollisions (VX_Sim.cpp:1626-1649)
            float m = (float)sp2;
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((tid & 31) == 0) atomicMax(&S.maxvel[k & 1], __float_as_uint(m));
        }
        __syncthreads();
        cur ^= 1;
        t += dt;                                                 // CurTime += dt (VX_Sim.cpp:1929)
        done = k + 1;
    }

    // ---- stage out --------------------------------------------------------------------------
    __syncthreads();
    if (SMEM) copy16(gblob, smem, L.dyn_bytes, tid, T);
    for (int i = tid; i < n_vc; i += T) B.scale_c[(size_t)robot * B.max_vclass + i] = S.scale_c[cur * B.max_vclass + i];
    if (COLLIDE && col_on) {
        const int nc = n_col < B.max_col ? n_col : B.max_col;
        for (int i = tid; i < nc; i += T) B.col_pairs[(size_t)robot * B.max_col + i] = S.pairs[i];
    }
    if (tid == 0) {
        ctlp->cur_time = t;
        ctlp->step = step0 + done;
        if (COLLIDE && col_on) {
            ctlp->n_col = n_col < B.max_col ? n_col : B.max_col; ctlp->col_first = col_first ? 1 : 0; ctlp->col_acc = S.colstate[0];
            ctlp->max_vel = done > 0 ? sqrt((double)__uint_as_float(S.maxvel[(done - 1) & 1])) : S.colstate[1];
        }
        unsigned int st2 = status0;
        if (col_overflow) st2 |= VXG_ST_COL_OVERFLOW_ | VXG_ST_DIVERGED_;
        if (diverged) st2 |= VXG_ST_DIVERGED_;
        if (step0 + done >= total) st2 |= VXG_ST_FINISHED_;
        ctlp->status = st2;
    }
}

struct StatsDev {   // mirrors vxg_robot_stats
    double com[3], ini_com[3], max_vel, min_pos[3], max_pos[3], time;
    long long steps;
    int n_touching_floor, ini_com_latched;
    double anterior_dist, posterior_dist, anterior_y, posterior_y, eol_posterior_y;
    int n_nonfeet_touching_floor, reserved;
    double mesh_volume;
};


// CVX_Sim::GetCM (VX_Sim.cpp:2415-2430), GetNumTouchingFloor (:2432-2441), SS.MaxVoxVel (:1511-1690)
template <typename Real>
__global__ void __launch_bounds__(256) vx_stats_kernel(BatchDev<Real> B, StatsDev* out) {
    __shared__ double red[4 * 32];
    __shared__ double redm[11 * 8];
    __shared__ int redi[2 * 8];
    const int robot = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const RobotCtl* ctl = B.ctl + robot;
    const BlobLayout L = B.L;
    RobotView<Real> R = make_view<Real>(B.blobs + (size_t)robot * L.total_bytes, L, ctl->n_vox);
    const VoxK<Real>* vk = B.vk + (size_t)robot * B.max_vclass;
    const ScaleK* sk = B.sk + (size_t)robot * B.max_vclass;
    const double* scale_c = B.scale_c + (size_t)robot * B.max_vclass;
    const int nv = R.nv;
    const double icx = ctl->ini_com[0], icy = ctl->ini_com[1];
    double mx = 0, my = 0, mz = 0, mm = 0, vmax = 0;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    double dmax = 0, dmin = 1e300, ymax = -1e300, ymin = 1e300;
    int touching = 0, feet = 0;
    for (int v = tid; v < R.n_vox; v += T) {
        int vc = R.stat[v].vcls;
        double m = (double)vk[vc].mass;
        double px = R.pos[v], py = R.pos[nv + v], pz = R.pos[2 * nv + v];
        mx += px * m; my += py * m; mz += pz * m; mm += m;
        const R4<Real> lm = R.momA[v];
        double vx_ = (double)lm.x * (double)vk[vc].mass_inv, vy_ = (double)lm.y * (double)vk[vc].mass_inv,
               vz_ = (double)lm.z * (double)vk[vc].mass_inv;
        vmax = fmax(vmax, sqrt(vx_ * vx_ + vy_ * vy_ + vz_ * vz_));
        lo[0] = fmin(lo[0], px); lo[1] = fmin(lo[1], py); lo[2] = fmin(lo[2], pz);
        hi[0] = fmax(hi[0], px); hi[1] = fmax(hi[1], py); hi[2] = fmax(hi[2], pz);
        const bool touch = 0.5 * scale_c[vc] - pz > 0;
        if (touch) touching++;
        const long long mat = sk[vc].mat_index;
        if (touch && mat == 6) feet++;                                   // GetNumNonFeetTouchingFloor (VX_Sim.cpp:2690-2704)
        double dd = sqrt((px - icx) * (px - icx) + (py - icy) * (py - icy));   // getAnteriorDist / getPosteriorDist
        dmax = fmax(dmax, dd); dmin = fmin(dmin, dd);
        if (mat != 5) { ymax = fmax(ymax, py); ymin = fmin(ymin, py); }  // getAnteriorY / getPosteriorY
    }
    // RobotVolumeEnd (land/water result file): rebuild the deformed surface mesh from the final state in the blob's own
    // vertex scratch (global memory here) and sum the signed tetrahedra of its facets
    double vol = -1.0;
    if (L.nvert_cap > 0) {
        RobotConsts<Real> C;
        C.vk = vk;
        for (int u = tid; u < ctl->n_vert; u += T) mesh_vertex<Real>(R, C, u);
        __syncthreads();
        const int nc = L.nvert_cap;
        double v6 = 0, z1 = 0, z2 = 0, z3 = 0;
        for (int i = tid; i < ctl->n_facet; i += T) {
            const Facet f = R.facet[i];
            const double ax = R.vpos[f.a], ay = R.vpos[nc + f.a], az = R.vpos[2 * nc + f.a];
            const double bx = R.vpos[f.b], by = R.vpos[nc + f.b], bz = R.vpos[2 * nc + f.b];
            const double cx = R.vpos[f.c], cy = R.vpos[nc + f.c], cz = R.vpos[2 * nc + f.c];
            v6 += (1.0 / 6.0) * ((ay * bz - az * by) * cx + (az * bx - ax * bz) * cy + (ax * by - ay * bx) * cz);
        }
        block_sum4(v6, z1, z2, z3, red);
        vol = v6;
        __syncthreads();
    }
    block_sum4(mx, my, mz, mm, red);
    double m11[11] = {vmax, -lo[0], -lo[1], -lo[2], hi[0], hi[1], hi[2], dmax, -dmin, ymax, -ymin};
    int lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
    for (int i = 0; i < 11; i++) m11[i] = warp_max(m11[i]);
    for (int o = 16; o > 0; o >>= 1) { touching += __shfl_xor_sync(0xffffffffu, touching, o); feet += __shfl_xor_sync(0xffffffffu, feet, o); }
    if (lane == 0) { for (int i = 0; i < 11; i++) redm[i * 8 + wid] = m11[i]; redi[wid] = touching; redi[8 + wid] = feet; }
    __syncthreads();
    if (tid == 0) {
        StatsDev s;
        for (int i = 0; i < 11; i++) { double x = redm[i * 8]; for (int w = 1; w < nw; w++) x = fmax(x, redm[i * 8 + w]); m11[i] = x; }
        int tc = 0, fc = 0; for (int w = 0; w < nw; w++) { tc += redi[w]; fc += redi[8 + w]; }
        s.com[0] = mx / mm; s.com[1] = my / mm; s.com[2] = mz / mm;
        for (int i = 0; i < 3; i++) { s.ini_com[i] = ctl->ini_com[i]; s.min_pos[i] = -m11[1 + i]; s.max_pos[i] = m11[4 + i]; }
        s.max_vel = m11[0]; s.time = ctl->cur_time; s.steps = ctl->step;
        s.n_touching_floor = tc; s.ini_com_latched = ctl->ini_latched;
        s.anterior_dist = m11[7]; s.posterior_dist = -m11[8]; s.anterior_y = m11[9]; s.posterior_y = -m11[10];
        s.eol_posterior_y = ctl->eol_post_y; s.n_nonfeet_touching_floor = fc; s.reserved = 0;
        s.mesh_volume = vol;
        out[robot] = s;
    }
}

}  // namespace vx
