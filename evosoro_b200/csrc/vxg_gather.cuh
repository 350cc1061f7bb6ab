// vxg_gather.cuh -- the production step kernel in its "gather" form (fp32, compact bond table, one voxel per thread).
//
// Same arithmetic as vx_step_kernel<float, NV, 1, true, true> (bond_force / voxel_integrate of vx_math.cuh, same summation
// order, bit-identical results), different data movement:
//
//   * Tensor memory as per-thread scratch.  The damping history of a voxel's three owned bonds (_LastPos2, _LastAngle1,
//     _LastAngle2: 27 words, half of the robot blob) is read and written by its owner only, so it does not need to be in
//     shared memory at all: it lives in TMEM (tcgen05.alloc; every warp owns 32 lanes x 32 columns, one column set per
//     thread) and moves with tcgen05.ld / tcgen05.st (LDTM / STTM), which neither occupy shared-memory bandwidth nor cost
//     address arithmetic.  That frees 108 B per voxel of shared memory for ...
//   * ... a staging area for the neighbour's share of every bond (Force1, Moment2 of the owned +X/+Y/+Z bonds, 72 B per voxel):
//     the -axis neighbour GATHERS them in the voxel phase instead of the owner pushing them into the neighbour's accumulator.
//     No read-modify-write, no accumulator zeroing, and above all no ordering between the three axes: a timestep has two
//     CTA-wide barriers (bonds -> voxels, voxels -> bonds, both hardware BAR) instead of four, two of which were mbarrier
//     waits whose sleep/wake traffic alone was 10 % of all issued instructions (profiles/r01_ncu_step_kernel_*.txt: SYNCS +
//     NANOSLEEP).
//   * A voxel's own position, orientation, static word and flags stay in registers from step to step.
//   * Voxels without a +axis neighbour evaluate a "zero class" bond (all stiffnesses 0, pitch 0: every output is exactly 0), so
//     the bond phase has no per-lane branch around the warp-collective TMEM operations; warps in which no lane has a bond on an
//     axis skip it.
//
// Replaces CVX_Sim::Integrate (VX_Sim.cpp:1763-1933) for robots of the production family (pack_compact, vx_pack.h).
#pragma once
#include <cstdio>
#include <cstdlib>
#include "vxg_kernels.cuh"
#include "vxg_launch.cuh"

namespace vx {

// shared-memory image of one robot in the gather kernel (byte offsets; the dynamic part keeps the blob's order so that it
// stages in and out with two straight copies)
struct GLayout {
    size_t off_pos, off_quat, off_momA, off_momB;   // as in the blob: [off_pos, dyn_front) is one contiguous copy
    size_t dyn_front;                               // bytes of pos .. momB
    size_t off_stat;                                // VoxStat [nv]
    size_t off_nbm;                                 // int16 [nv][4]: the -X/-Y/-Z neighbour of each voxel (nv = "none": the zero slot)
    size_t off_stA[3], off_stB[3];                  // per axis: R4 [nv + 1] (Force1.xyz, Moment2.x), R2 [nv + 1] (Moment2.y, Moment2.z)
    size_t total;
};
VX_HD constexpr GLayout make_glayout(int nv_in) {
    GLayout G{};
    const BlobLayout L = make_layout(nv_in, 4);
    const size_t nv = (size_t)nv_in;
    G.off_pos = L.off_pos; G.off_quat = L.off_quat; G.off_momA = L.off_momA; G.off_momB = L.off_momB;
    G.dyn_front = L.off_hist;
    size_t o = G.dyn_front;
    G.off_stat = o; o = vx_align32(o + nv * sizeof(VoxStat));
    G.off_nbm = o;  o = vx_align32(o + nv * 8);
    for (int a = 0; a < 3; a++) {
        G.off_stA[a] = o; o = vx_align32(o + (nv + 1) * 16);
        G.off_stB[a] = o; o = vx_align32(o + (nv + 1) * 8);
    }
    G.total = o;
    return G;
}

// A bond class as the gather kernel reads it: the constants bond_force needs when both voxels of every bond have the same mass
// (the production family), already expanded from the compact form (expand_bond) once per launch -- three 128-bit loads per
// bond, no per-bond expansion arithmetic, no geometry constants held in registers.
struct alignas(16) BondKm {
    float a2, b1, b2, b3;
    float inv_len, k_axial, nom_dist, len_minus_dist;
    float s1, c_a2i1, c_b2fm1, c_b3i1;     // 2 sqrt(a1 m) = 2 sqrt(b1 m), 2 sqrt(a2 I), 2 sqrt(b2 FM), 2 sqrt(b3 I)
};
__device__ __forceinline__ BondK<float> bond_from_km(const BondKm& m) {
    BondK<float> k;
    k.a2 = m.a2; k.b1 = m.b1; k.b2 = m.b2; k.b3 = m.b3;
    k.inv_len = m.inv_len; k.k_axial = m.k_axial; k.nom_dist = m.nom_dist; k.len_minus_dist = m.len_minus_dist;
    k.c_a1m1 = m.s1; k.c_a1m2 = m.s1; k.c_b1m1 = m.s1; k.c_b1m2 = m.s1;
    k.c_a2i1 = m.c_a2i1; k.c_a2i2 = m.c_a2i1; k.c_b2fm1 = m.c_b2fm1; k.c_b2fm2 = m.c_b2fm1; k.c_b3i1 = m.c_b3i1; k.c_b3i2 = m.c_b3i1;
    k.sf1 = 1.0f; k.sf2 = 1.0f; k.homogeneous = 0;   // (the per-voxel strain shares feed the land/water surface mesh only)
    return k;
}

// Resident CTAs per SM the register allocation of the gather kernels is sized for: one 10^3 robot, two 8^3 robots (with or without
// the land/water mesh, which shares memory with the staging slots), four robots of 129 .. 224 voxels, 768 threads' worth of smaller ones.
__host__ __device__ constexpr int vxg_ctas(int nv, bool collide, bool fluid) {
    return nv > 512 ? 1 : nv == 512 ? 2 : nv >= 160 && nv <= 256 ? ((collide || fluid) && nv == 256 ? 3 : 4) : 768 / nv;
}

// ---- tensor memory as scratch: 32 lanes x 32 columns per warp, lane = thread, columns = that thread's words -----------------
#define VXG_TM_COLS_PER_WARP 32
__device__ inline void tm_ld8(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr) : "memory");
}
__device__ inline void tm_ld1(uint32_t taddr, float& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(taddr) : "memory");
}
__device__ inline void tm_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ inline void tm_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(v) : "memory");
}
__device__ inline void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ inline void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// columns a CTA of `threads` threads needs: its warps share the four 32-lane quarters, each warp gets its own 32 columns
__host__ __device__ constexpr int tm_cols_for(int threads) {
    const int per_quarter = ((threads + 31) / 32 + 3) / 4;
    const int need = per_quarter * VXG_TM_COLS_PER_WARP;
    return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

// history words of the bond on axis AX of this thread's voxel: columns [8 AX, 8 AX + 8) and 24 + AX
template <int AX> __device__ inline void hist_load(uint32_t tm, BondHist<float>& h) {
    float a[8], c;
    tm_ld8(tm + 8 * AX, a);
    tm_ld1(tm + 24 + AX, c);
    tm_wait_ld();
    h.p2 = mk3<float>(a[0], a[1], a[2]); h.a1 = mk3<float>(a[3], a[4], a[5]); h.a2 = mk3<float>(a[6], a[7], c);
}
// the same in two halves: issue the loads, do something else, wait
template <int AX> __device__ inline void hist_issue(uint32_t tm, float (&a)[8], float& c) { tm_ld8(tm + 8 * AX, a); tm_ld1(tm + 24 + AX, c); }
__device__ inline void hist_take(const float (&a)[8], float c, BondHist<float>& h) {
    tm_wait_ld();
    h.p2 = mk3<float>(a[0], a[1], a[2]); h.a1 = mk3<float>(a[3], a[4], a[5]); h.a2 = mk3<float>(a[6], a[7], c);
}
template <int AX> __device__ inline void hist_store(uint32_t tm, const BondHist<float>& h) {
    const float a[8] = {h.p2.x, h.p2.y, h.p2.z, h.a1.x, h.a1.y, h.a1.z, h.a2.x, h.a2.y};
    tm_st8(tm + 8 * AX, a);
    tm_st1(tm + 24 + AX, h.a2.z);
}

// One owned bond in the gather form.  `st` carries the zero class for a missing neighbour (patched once per launch), so every
// lane runs the same code; the warp skips the axis when none of its lanes has a bond there.
//   F_REGS : the own share of the force accumulates in registers (else it is re-read from this thread's staging slots in the voxel
//            phase: three registers less across the bond phase).  Measured: registers +0.7 % at 10^3, re-read +0.5 % at 8^3.
template <int AX, bool HOIST, bool F_REGS, bool FLUID>
__device__ __forceinline__ bool bond_eval_g(unsigned char* smem, const GLayout& G, int NV, const RobotConsts<float>& C, const BondKm* bkm, float* strain, int tid, const VoxStat& st,
                                            const OwnState<float>& own, uint32_t& fl, float dt_inv, uint32_t tm, V3<float>& F, V3<float>& M) {
    const int w = st.nbr[AX];
    if (!__any_sync(0xffffffffu, w >= 0)) return false;
    const int wq = w >= 0 ? w : tid;
    const int bc = st.bcls[AX];
    float ha[8], hc;
    hist_issue<AX>(tm, ha, hc);   // the history is on its way from tensor memory while the neighbour's state is fetched

    BondK<float> k = bond_from_km(bkm[bc]);
    if (FLUID) { k.sf1 = C.bkc[bc].sf1; k.sf2 = 2.0f - k.sf1; }   // per-voxel shares of the axial strain: they shape the surface mesh
    const double* pos = (const double*)(smem + G.off_pos);
    double dpa[3] = {pos[wq] - own.px, pos[NV + wq] - own.py, pos[2 * NV + wq] - own.pz};
    dpa[AX] -= C.bk_dist[bc];
    const V3<float> e = mk3<float>((float)dpa[0], (float)dpa[1], (float)dpa[2]);
    const R4<float> qa = own.q, qb = ((const R4<float>*)(smem + G.off_quat))[wq];
    Q4<float> q1, q2;
    q1.w = qa.x; q1.x = qa.y; q1.y = qa.z; q1.z = qa.w;
    q2.w = qb.x; q2.x = qb.y; q2.y = qb.z; q2.z = qb.w;
    bool small = (fl >> AX) & 1u;
    BondHist<float> h;
    hist_take(ha, hc, h);
    BondOut<float> o;
    bond_force<AX, float, true, HOIST>(k, e, q1, q2, C.dnom_bc[bc], C.half_bond_z, dt_inv, true, small, h, o);
    // _LastPos2/_LastAngle* keep their values on a mode-flip step (VXS_BondInternal.cpp:157): rare, so the old values are
    // fetched again (they are still in tensor memory) behind a warp-uniform test instead of being kept in registers
    if (__any_sync(0xffffffffu, o.changed)) { BondHist<float> old; hist_load<AX>(tm, old); if (o.changed) h = old; }
    hist_store<AX>(tm, h);
    fl = (fl & ~(1u << AX)) | ((small ? 1u : 0u) << AX);
    if (FLUID && w >= 0) {   // SetStrainDir(BD_P<axis>, CurStrainV1) on voxel 1, (BD_N<axis>, CurStrainV2) on voxel 2: one writer per slot
        strain[AX * NV + tid] = o.strain1;
        strain[(3 + AX) * NV + w] = o.strain2;
    }
    if (F_REGS) F = F + o.f1;   // (else the own share of the force is read back from the staging slot in the voxel phase)
    M = M + o.m1;
    R4<float> a; a.x = o.f1.x; a.y = o.f1.y; a.z = o.f1.z; a.w = o.m2.x;
    R2<float> b; b.x = o.m2.y; b.y = o.m2.z;
    ((R4<float>*)(smem + G.off_stA[AX]))[tid] = a;
    ((R2<float>*)(smem + G.off_stB[AX]))[tid] = b;
    return o.strain > 100.0f;
}

// Voxel phase of the gather form: the shares of the -X/-Y/-Z bonds (Force2 = -Force1, Moment2) from their owners' staging slots,
// then EulerStep; the new position / orientation go to shared memory for the neighbours and stay in `own` for the next step.
//   F_EXTRA : with !F_REGS, F arrives holding what joined the own share after the bonds (the contact forces) instead of zero
template <bool F_REGS, bool F_EXTRA = false>
__device__ __forceinline__ float voxel_phase_g(unsigned char* smem, const GLayout& G, int NV, const RobotConsts<float>& C, const RobotView<float>& R, int tid,
                                              const VoxStat& st, uint2 nb, OwnState<float>& own, uint32_t& fl, V3<float> F, V3<float> M) {
    const int mX = nb.x & 0xFFFFu, mY = nb.x >> 16, mZ = nb.y & 0xFFFFu;
    const R4<float> ax = ((const R4<float>*)(smem + G.off_stA[0]))[mX], ay = ((const R4<float>*)(smem + G.off_stA[1]))[mY], az = ((const R4<float>*)(smem + G.off_stA[2]))[mZ];
    const R2<float> bx = ((const R2<float>*)(smem + G.off_stB[0]))[mX], by = ((const R2<float>*)(smem + G.off_stB[1]))[mY], bz = ((const R2<float>*)(smem + G.off_stB[2]))[mZ];
    // own shares: Force1 of the owned X, Y, Z bonds, as staged by this very thread (axes the voxel has no bond on staged zeros)
    if (!F_REGS) {
        const R4<float> ox = ((const R4<float>*)(smem + G.off_stA[0]))[tid], oy = ((const R4<float>*)(smem + G.off_stA[1]))[tid], oz = ((const R4<float>*)(smem + G.off_stA[2]))[tid];
        const V3<float> own_f = mk3<float>(((0.0f + ox.x) + oy.x) + oz.x, ((0.0f + ox.y) + oy.y) + oz.y, ((0.0f + ox.z) + oy.z) + oz.z);
        F = F_EXTRA ? own_f + F : own_f;
    }
    // the accumulator of the push form received -Force1 of the X, Y, Z bond in that order, starting from zero
    F = F + mk3<float>(((0.0f - ax.x) - ay.x) - az.x, ((0.0f - ax.y) - ay.y) - az.y, ((0.0f - ax.z) - ay.z) - az.z);
    M = M + mk3<float>(((0.0f + ax.w) + ay.w) + az.w, ((0.0f + bx.x) + by.x) + bz.x, ((0.0f + bx.y) + by.y) + bz.y);
    const int vc = st.vcls;
    const VoxK<float> kv = C.vk[vc];
    VoxState<float> vs;
    vs.px = own.px; vs.py = own.py; vs.pz = own.pz;
    const R4<float> ma = R.momA[tid]; const R2<float> mb = R.momB[tid];
    vs.q.w = own.q.x; vs.q.x = own.q.y; vs.q.y = own.q.z; vs.q.z = own.q.w;
    vs.lm = mk3<float>(ma.x, ma.y, ma.z);
    vs.am = mk3<float>(ma.w, mb.x, mb.y);
    const bool sf = voxel_integrate<float, false>(kv, F, M, 0.5 * C.scale_c[vc], C.floor_on, false, C.dt, nullptr, nullptr, nullptr, vs);
    own.px = vs.px; own.py = vs.py; own.pz = vs.pz;
    own.q.x = vs.q.w; own.q.y = vs.q.x; own.q.z = vs.q.y; own.q.w = vs.q.z;
    R.pos[tid] = vs.px; R.pos[NV + tid] = vs.py; R.pos[2 * NV + tid] = vs.pz;
    R.quat[tid] = own.q;
    R4<float> na; na.x = vs.lm.x; na.y = vs.lm.y; na.z = vs.lm.z; na.w = vs.am.x;
    R2<float> nb2; nb2.x = vs.am.y; nb2.y = vs.am.z;
    R.momA[tid] = na; R.momB[tid] = nb2;
    fl = sf ? (fl | VX_FL_STATIC_FRIC) : (fl & ~VX_FL_STATIC_FRIC);
    const V3<float> vel = vs.lm * kv.mass_inv;
    return dot(vel, vel);   // |Vel|^2 after the step (SS.MaxVoxVel, VX_Sim.cpp:1626-1649)
}

// do the land/water scratch arrays (vertex positions fp64, facet drags) fit into the staging region they can share memory with?
__host__ __device__ inline bool fluid_scratch_aliases(int nv, int nvert_cap, int nfacet_cap) {
    const GLayout G = make_glayout(nv);
    return vx_align16(sizeof(double) * 3 * (size_t)nvert_cap) + vx_align16(sizeof(float) * 3 * (size_t)nfacet_cap) <= G.total - G.off_stA[0];
}

// Per-voxel view of the contact-bond list (CSR): seg[v] = begin | end << 16 into adj[], the partners of voxel v in list order (the
// list is sorted by (v1, v2) with v1 < v2, so that is ascending partner index).  Built once per list rebuild -- a fraction of a
// percent to a few percent of the steps -- so that a step costs a voxel one word and its own contacts instead of a scan of every
// pair of the robot (CVXS_Voxel's ColBondInds, VXS_Voxel.cpp:516-528).  One pair per thread and round: shared-memory atomics count
// and place, then every voxel sorts its own few partners -- O(pairs / threads) per thread where a scan of the whole list by every
// thread cost ten timesteps' worth of instructions per rebuild on a dense robot (a thousand pairs inside the horizon).
static __device__ __noinline__ void col_csr_build(const ColPair* pairs, int n_col, uint32_t* seg, uint16_t* adj, int* scratch /* [33] */) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    seg[tid] = 0u;
    __syncthreads();
    for (int k = tid; k < n_col; k += T) { const ColPair p = pairs[k]; atomicAdd(&seg[p.v1], 1u); atomicAdd(&seg[p.v2], 1u); }
    __syncthreads();
    const int cnt = (int)seg[tid];
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid && w < nw; w++) base += scratch[w];
    const int off = base + inc - cnt;
    seg[tid] = (uint32_t)off | ((uint32_t)off << 16);   // the end doubles as the fill cursor
    __syncthreads();
    for (int k = tid; k < n_col; k += T) {
        const ColPair p = pairs[k];
        adj[atomicAdd(&seg[p.v1], 0x10000u) >> 16] = (uint16_t)p.v2;
        adj[atomicAdd(&seg[p.v2], 0x10000u) >> 16] = (uint16_t)p.v1;
    }
    __syncthreads();
    for (int a = off + 1; a < off + cnt; a++) {   // list order = ascending partner: insertion sort of this voxel's segment
        const uint16_t x = adj[a];
        int b = a - 1;
        while (b >= off && adj[b] > x) { adj[b + 1] = adj[b]; b--; }
        adj[b + 1] = x;
    }
    __syncthreads();
}

// Dataflow form (FLOW): within a window of up to VXG_FLOW_R timesteps the warps are not held together by CTA-wide barriers at all.
// A warp's bond phase needs the positions its +axis neighbours' owners wrote in their last voxel phase, its voxel phase needs the
// shares its -axis neighbours' owners staged in their bond phase -- a handful of warps in either case (at 10^3: the warp itself, the
// next / previous one and the ones 100 voxels away).  Every warp therefore has two mbarriers (bond phase done, voxel phase done; one
// arrival per phase) and waits only on those of the warps its voxels' neighbours live in.  The same waits cover the write-after-read
// hazards (the dependency sets are each other's mirror image), which also keeps any two dependent warps within one phase of each
// other -- the condition for parity waits.  Warps drift apart by whole phases and fill each other's stalls: the per-step CTA
// barriers cost 29 % of the issue slots in the barrier form (ncu: smsp__issue_active 71 %; profiles/).
// What needs the whole CTA -- the per-class Scale tables, the IniCM / end-of-life latches, the divergence verdict -- happens once
// per window: Scale is computed VXG_FLOW_R steps ahead into a ring, the rest-length table of a window is derived between the two
// barriers that open it, latches end a window early, and a diverged bond is recorded with its step and acted on at the window's end.
#define VXG_FLOW_R 8
#define VXG_FLOW_RING (2 * VXG_FLOW_R)
__host__ __device__ inline size_t flow_bytes(int max_vclass, int mb) {
    return vx_align16(sizeof(double) * VXG_FLOW_RING * max_vclass) + vx_align16(sizeof(float) * VXG_FLOW_RING * max_vclass) +
           vx_align16(sizeof(float) * VXG_FLOW_RING * mb) + 64 * sizeof(unsigned long long) + 16 + vx_align16(sizeof(double) * (2 + max_vclass));
}
// wait until, for every warp u in `mask`, bar[u] has completed the phase of the given parity (lane u waits on bar[u])
__device__ inline void flow_wait(unsigned long long* bar, uint32_t mask, uint32_t parity, int lane) {
    if ((mask >> lane) & 1u) mbar_wait(bar + lane, parity);
    __syncwarp();
}

//   CV, CB : compile-time capacities of the voxel / bond class tables in shared memory (0 = the batch's run-time limits).  With them
//            every shared-memory address in the step loop is the CTA's base plus an immediate; with run-time capacities the
//            table pointers were rematerialised over and over at 64 registers (5 % of the executed instructions, ncu).
//   COLLIDE: self-collision (contact-bond list in shared memory, rebuilt on the device, contact forces through the per-voxel CSR)
//   FLUID  : land/water layouts: the deformable surface mesh (vertex -> voxels, facets) and the per-voxel strains live in shared memory
//            behind the class tables (run-time capacities), fluid drag is evaluated at the start of every step
//   NVC    : land/water layouts: compile-time capacity of the surface-mesh arrays in shared memory (vertices; 0 = the batch's).  With CV, CB
//            and NVC fixed every shared-memory address of the mesh phases is an immediate as well
template <int NV_, bool FLOW, int CV, int CB, bool COLLIDE = false, bool FLUID = false, int NVC = 0>
__global__ void __launch_bounds__(NV_ ? NV_ : 256, NV_ ? vxg_ctas(NV_, COLLIDE, FLUID) : 3) vx_step_kernel_g(BatchDev<float> B, int n_steps) {
    static_assert(!(FLOW && (COLLIDE || FLUID)), "the dataflow form has no per-step CTA-wide phases");
    extern __shared__ __align__(128) unsigned char smem[];
    // NV_ == 0: the voxel stride is a RUN-TIME value (a multiple of 32 up to 256, = the CTA's thread count): one kernel image for
    // every small robot, each with exactly the warps it needs (several images resident on an SM evict each other from the
    // instruction cache -- profiles/r02_sizes.txt).  NV_ > 0: compile-time stride, every offset an immediate.
    const int NV = NV_ ? NV_ : B.L.nv;
    const int T = NV;
    const BlobLayout L = NV_ ? make_layout(NV_, 4) : make_layout(NV, 4);
    const GLayout G = make_glayout(NV);
    const int TM_COLS = tm_cols_for(NV);
    constexpr bool HOIST = !(NV_ == 1024 || NV_ == 512);   // bond_force: only where there are more than 64 registers per thread
    // (the drag force is the START of a swimmer's force sum, so there the sum stays in registers; contact forces join the finished own
    // share, so the self-collision variants of the 64-register strides re-read it like the plain ones and add the contacts to it: the
    // sum's three registers across the bond phase were spilled -- a local load and store per axis, profiles/r02b: 42 LDL per voxel-step)
    constexpr bool F_REGS = FLUID || (COLLIDE ? !(NV_ == 1024 || NV_ == 512) : NV_ == 1024);
    constexpr bool F_EXTRA = COLLIDE && !F_REGS;
    // PERM (-DVXG_PERM, off): thread t of the 1024-thread kernel owns voxel (t % 32) * 32 + t / 32 -- the transpose of the warp x lane
    // grid -- and every on-chip array is indexed by THREAD, so that a warp's voxels are spread over the whole robot instead of being
    // 32 consecutive ones.  Motivation: the layers of a 10^3 robot differ in how much they deform, and the calm top-layer warps reach
    // the bonds -> voxels barrier thousands of cycles before the others (clock64 trace).  Measured: balanced, bit-identical, and 8 %
    // SLOWER (1.82e10 vs 1.97e10): the calm warps were cheap precisely because all their lanes were in small-angle mode and some
    // skipped an axis; spread out, every warp pays for both modes and all three axes.  Kept as the record of that experiment.
#ifdef VXG_PERM
    constexpr bool PERM = NV_ == 1024 && !FLOW && !COLLIDE && !FLUID;
#else
    constexpr bool PERM = false;
#endif
    const int robot = blockIdx.x;
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int vme = PERM ? (((tid & 31) << 5) | (tid >> 5)) : tid;   // the voxel this thread owns
    RobotCtl* ctlp = B.ctl + robot;
    const unsigned int status0 = ctlp->status;
    if (status0 & (VXG_ST_DIVERGED_ | VXG_ST_FINISHED_)) return;
    const long long step0 = ctlp->step, total = ctlp->total_steps, latch = ctlp->latch_step, eol = ctlp->eol_step;
    long long remaining = total - step0;
    int nsteps = remaining < (long long)n_steps ? (int)remaining : n_steps;
    if (nsteps <= 0) { if (tid == 0) ctlp->status = status0 | VXG_ST_FINISHED_; return; }

    unsigned char* gblob = B.blobs + (size_t)robot * (FLUID ? B.L.total_bytes : L.total_bytes);
    const int MB = CB > 0 ? CB : B.max_bclass + 1;   // one more bond class than the batch's limit: the zero class
    const int mv = CV > 0 ? CV : B.max_vclass;
    BondKm* bkm = (BondKm*)(smem + G.total);
    unsigned char* tabp = smem + G.total + sizeof(BondKm) * (size_t)MB;
    const int max_col = COLLIDE ? B.max_col : 0;
    SmemTables<float> S = carve_tables<float>(tabp, mv, MB, true, max_col);
    unsigned char* fp = tabp + tables_bytes<float>(mv, MB, true, max_col);
    uint32_t* c_seg = (uint32_t*)fp;                    // self-collision: per-voxel segments of c_adj (the flow ring shares the address: never both)
    uint16_t* c_adj = (uint16_t*)(fp + sizeof(uint32_t) * NV);
    unsigned char* flp = fp + (max_col > 0 ? vx_align16(sizeof(uint32_t) * NV + sizeof(uint16_t) * 2 * (size_t)max_col) : 0);   // land/water arrays
    const int n_vox = ctlp->n_vox, n_vc = ctlp->n_vclass, n_bc = ctlp->n_bclass;
    const unsigned int flags = ctlp->flags;
    const int ZC = n_bc;               // index of the zero class

    // ---- tensor memory -------------------------------------------------------------------------------------------------
    uint32_t* tm_slot = (uint32_t*)S.red;   // the reduction scratch is idle until the first latch
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tm_slot)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- stage in ------------------------------------------------------------------------------------------------------
    // positions .. momenta (one contiguous run of the blob) and the static words: two bulk copies by the TMA engine, in flight while
    // the CTA fills the class tables and allocates tensor memory; waited for right before the first read
    if (!PERM && tid == 0) {
        mbar_init(S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(S.mbar, (unsigned int)(G.dyn_front + (size_t)NV * sizeof(VoxStat)));
        bulk_g2s(smem, gblob, (unsigned int)G.dyn_front, S.mbar);
        bulk_g2s(smem + G.off_stat, gblob + (FLUID ? B.L.off_stat : L.off_stat), (unsigned int)((size_t)NV * sizeof(VoxStat)), S.mbar);   // (the strains sit before it in a land/water blob)
    }
    {
        uint2 none; none.x = none.y = (uint32_t)NV | ((uint32_t)NV << 16);
        ((uint2*)(smem + G.off_nbm))[tid] = none;
        // staging slots start as zeros and stay so where nothing is ever staged: the "no neighbour" slot [NV] and the slots of
        // warps that skip an axis altogether (the voxel phase reads its own slots unconditionally)
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (size_t i = tid; i < ((G.total - G.off_stA[0]) >> 4); i += T) ((uint4*)(smem + G.off_stA[0]))[i] = z;
    }
    copy4(S.bk, B.bk + (size_t)robot * B.max_bclass * sizeof(BondKc<float>), sizeof(BondKc<float>) * n_bc, tid, T);
    copy4(S.vk, B.vk + (size_t)robot * B.max_vclass, sizeof(VoxK<float>) * n_vc, tid, T);
    copy4(S.sk, B.sk + (size_t)robot * B.max_vclass, sizeof(ScaleK) * n_vc, tid, T);
    for (int i = tid; i < n_bc; i += T) { S.bk_dist[i] = B.bk_dist[(size_t)robot * B.max_bclass + i]; S.bpair[i] = B.bpair[(size_t)robot * B.max_bclass + i]; }
    for (int i = tid; i < n_vc; i += T) S.scale_c[i] = B.scale_c[(size_t)robot * B.max_vclass + i];
    if (tid < (int)(sizeof(BondKc<float>) / 4)) ((uint32_t*)((BondKc<float>*)S.bk + ZC))[tid] = 0u;
    if (tid == 0) { S.bk_dist[ZC] = 0.0; S.dnom_bc[ZC] = 0.0f; S.dnom_bc[MB + ZC] = 0.0f; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!PERM) mbar_wait(S.mbar, 0);   // the bulk copies have landed
    const uint32_t tm = *tm_slot + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)(tid >> 7) * VXG_TM_COLS_PER_WARP;   // lane (warp % 4) * 32, this warp's columns
    for (int i = tid; i < n_vc; i += T) S.dsc_c[i] = (float)(S.scale_c[i] - S.sk[i].nom_size);
    for (int i = tid; i <= n_bc; i += T) {   // expanded bond classes (a robot with a class per voxel has more of them than threads); [n_bc] = the zero class
        BondKm m; memset(&m, 0, sizeof m);
        if (i < n_bc) {
            const BondK<float> k = expand_bond<float>(((const BondKc<float>*)S.bk)[i], make_geo<float>(ctlp->bond_len));
            m.a2 = k.a2; m.b1 = k.b1; m.b2 = k.b2; m.b3 = k.b3;
            m.inv_len = k.inv_len; m.k_axial = k.k_axial; m.nom_dist = k.nom_dist; m.len_minus_dist = k.len_minus_dist;
            m.s1 = k.c_a1m1; m.c_a2i1 = k.c_a2i1; m.c_b2fm1 = k.c_b2fm1; m.c_b3i1 = k.c_b3i1;
        }
        bkm[i] = m;
    }
    // own static word (zero class for the missing bonds), own state, and the -axis neighbour table
    const bool active = vme < n_vox;
    VoxStat st = PERM ? ((const VoxStat*)(gblob + L.off_stat))[vme] : ((const VoxStat*)(smem + G.off_stat))[tid];
    if (!active) st.nbr[0] = st.nbr[1] = st.nbr[2] = -1;
    if (PERM) {   // neighbour indices become thread indices; this voxel's state goes to this thread's slots (gathered from the blob once per launch)
#pragma unroll
        for (int a = 0; a < 3; a++) if (st.nbr[a] >= 0) st.nbr[a] = (int16_t)(((st.nbr[a] & 31) << 5) | (st.nbr[a] >> 5));
        const RobotView<float> Rg = make_view<float>(gblob, L, n_vox);
        double* sp = (double*)(smem + G.off_pos);
        sp[tid] = Rg.pos[vme]; sp[NV + tid] = Rg.pos[NV + vme]; sp[2 * NV + tid] = Rg.pos[2 * NV + vme];
        ((R4<float>*)(smem + G.off_quat))[tid] = Rg.quat[vme];
        ((R4<float>*)(smem + G.off_momA))[tid] = Rg.momA[vme];
        ((R2<float>*)(smem + G.off_momB))[tid] = Rg.momB[vme];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (st.nbr[a] < 0) st.bcls[a] = (uint16_t)ZC;
        else ((int16_t*)(smem + G.off_nbm))[4 * st.nbr[a] + a] = (int16_t)tid;
    }
    uint32_t fl = ((const uint32_t*)(gblob + L.off_flags))[vme];
    {   // damping history: blob -> this thread's tensor-memory columns
        const RobotView<float> Rg = make_view<float>(gblob, L, n_vox);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const R4<float> x = Rg.histA(a)[vme], y = Rg.histB(a)[vme]; const float z = Rg.histC(a)[vme];
            BondHist<float> h; h.p2 = mk3<float>(x.x, x.y, x.z); h.a1 = mk3<float>(x.w, y.x, y.y); h.a2 = mk3<float>(y.z, y.w, z);
            if (a == 0) hist_store<0>(tm, h); else if (a == 1) hist_store<1>(tm, h); else hist_store<2>(tm, h);
        }
        tm_wait_st();
    }
    __syncthreads();
    for (int i = tid; i < n_bc; i += T) { const uint32_t pr = S.bpair[i]; S.dnom_bc[i] = 0.5f * (S.dsc_c[pr & 0xFFFFu] + S.dsc_c[pr >> 16]); }

    RobotView<float> R = make_view<float>(smem, L, n_vox);   // pos / quat / momA / momB views are valid (same offsets as the blob)
    R.stat = (VoxStat*)(smem + G.off_stat);
    const BlobLayout LF = B.L;   // (the batch's layout: the blob of a land/water batch carries the mesh arrays and the strains)
    const int nvert_cap = NVC ? NVC : LF.nvert_cap, nfacet_cap = NVC ? ((2 * NVC + 64 + 31) & ~31) : LF.nfacet_cap;   // (make_layout's rule)
    const bool fluid_alias = FLUID && fluid_scratch_aliases(NV, nvert_cap, nfacet_cap);
    if (FLUID) {
        const size_t nvc = (size_t)nvert_cap, nfc = (size_t)nfacet_cap;
        unsigned char* q = flp;
        R.strain = (float*)q;     q += vx_align16(sizeof(float) * 6 * NV);
        R.fbase = (uint16_t*)q;   q += vx_align16(sizeof(uint16_t) * NV);
        R.facet = (Facet*)q;      q += vx_align16(sizeof(Facet) * nfc);
        R.vcon = (Corner8*)q;     q += vx_align16(sizeof(Corner8) * nvc);
        // The deformed vertex positions and the per-facet drag only live from the start of a step to the per-voxel drag sum, the
        // staging slots from the bond phase to the voxel phase: they share memory where they fit (a barrier after the drag sum keeps
        // a fast warp's staging writes off a slow warp's reads).  That is what lets two 8^3 swimmers share an SM.
        if (fluid_alias) { R.vpos = (double*)(smem + G.off_stA[0]); R.fdrag = (float*)(smem + G.off_stA[0] + vx_align16(sizeof(double) * 3 * nvc)); }
        else { R.vpos = (double*)q; q += vx_align16(sizeof(double) * 3 * nvc); R.fdrag = (float*)q; }
        R.nvert_cap = nvert_cap; R.nfacet_cap = nfacet_cap;   // (the shared-memory arrays' strides; the blob keeps the batch's)
        unsigned char* gb = B.blobs + (size_t)robot * LF.total_bytes;
        copy4(R.strain, gb + LF.off_strain, sizeof(float) * 6 * NV, tid, T);
        copy4(R.fbase, gb + LF.off_fbase, (sizeof(uint16_t) * NV) & ~(size_t)3, tid, T);
        copy4(R.facet, gb + LF.off_facet, sizeof(Facet) * (size_t)LF.nfacet_cap, tid, T);
        copy4(R.vcon, gb + LF.off_vcon, sizeof(Corner8) * (size_t)LF.nvert_cap, tid, T);
    }
    RobotConsts<float> C;
    C.bk = nullptr; C.bkc = (const BondKc<float>*)S.bk;
    C.geo = make_geo<float>(1.0);
    C.bk_dist = S.bk_dist; C.vk = S.vk;
    C.nom = nullptr; C.ext_disp = S.ext_disp; C.ext_tdisp = S.ext_tdisp;
    C.half_bond_z = (float)ctlp->half_bond_z;
    const double dt = ctlp->dt;
    C.dt = (float)dt;
    const bool fluid = FLUID && (flags & VXG_F_FLUID_) != 0;
    C.fluid = fluid; C.drag_coeff = (float)ctlp->drag_coeff; C.n_vert = ctlp->n_vert; C.n_facet = ctlp->n_facet;
    C.floor_on = (flags & VXG_F_FLOOR_) != 0 && !fluid;
    C.sticky = false; C.sym_mass = true;
    if (tid == 0) {
        ScaleSim SL;
        SL.init_cm_time = ctlp->init_cm_time; SL.temp_period = ctlp->temp_period; SL.temp_amplitude = ctlp->temp_amplitude;
        SL.min_fact = ctlp->min_fact; SL.max_fact = ctlp->max_fact; SL.growth_amp = ctlp->growth_amp;
        SL.midlife_freeze = ctlp->midlife_freeze; SL.stop_value = ctlp->stop_value;
        SL.temp_on = (flags & VXG_F_TEMPERATURE_) != 0; SL.lw = (flags & VXG_F_LW_SCALE_) != 0;
        *S.sl = SL;
    }
    OwnState<float> own = load_own<float>(R, tid);
    __syncthreads();
    const uint2 nb = ((const uint2*)(smem + G.off_nbm))[tid];   // -X | -Y << 16, -Z
    const float dt_inv_run = (float)(1.0 / dt);
    // ---- dataflow form: table ring, per-warp phase barriers, dependency masks --------------------------------------------
    double* r_scale = (double*)fp;                      fp += vx_align16(sizeof(double) * VXG_FLOW_RING * mv);
    float* r_dsc = (float*)fp;                          fp += vx_align16(sizeof(float) * VXG_FLOW_RING * mv);
    float* r_dnom = (float*)fp;                         fp += vx_align16(sizeof(float) * VXG_FLOW_RING * MB);
    unsigned long long* bondbar = (unsigned long long*)fp; unsigned long long* voxbar = bondbar + 32;   fp += 64 * sizeof(unsigned long long);
    int* f_div = (int*)fp;                              fp += 16;
    double* f_time = (double*)fp;                       // [0] CurTime (thread 0's), [2 + c] time of the step after class c's newest ring entry
    uint32_t plus_mask = 0, minus_mask = 0;
    const int my_class = T - 1 - tid;                   // the voxel class whose Scale this thread computes (if < n_vc)
    if (FLOW) {
        const int lane = tid & 31, warp = tid >> 5;
        uint32_t pm = 0, mm = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) if (st.nbr[a] >= 0) pm |= 1u << (st.nbr[a] >> 5);
        if (active) {
            const int m0 = nb.x & 0xFFFFu, m1 = nb.x >> 16, m2 = nb.y & 0xFFFFu;
            if (m0 < NV) mm |= 1u << (m0 >> 5);
            if (m1 < NV) mm |= 1u << (m1 >> 5);
            if (m2 < NV) mm |= 1u << (m2 >> 5);
        }
        plus_mask = __reduce_or_sync(0xffffffffu, pm) & ~(1u << warp);
        minus_mask = __reduce_or_sync(0xffffffffu, mm) & ~(1u << warp);
        if (lane == 0) { mbar_init(bondbar + warp, 1); mbar_init(voxbar + warp, 1); }
        if (tid == 0) *f_div = 0x7fffffff;
        for (int i = tid; i < VXG_FLOW_RING; i += T) r_dnom[i * MB + ZC] = 0.0f;
        if (my_class < n_vc) {   // ring entries of the first window: Scale of step j from the Scale and the time of step j - 1
            double sc_last = S.scale_c[my_class], sc_t = ctlp->cur_time;
            r_scale[my_class] = sc_last; r_dsc[my_class] = (float)(sc_last - S.sk[my_class].nom_size);
            for (int j = 1; j < VXG_FLOW_R; j++) {
                sc_last = scale_law(S.sk[my_class], sc_last, sc_t, *S.sl); sc_t += dt;
                r_scale[j * mv + my_class] = sc_last; r_dsc[j * mv + my_class] = (float)(sc_last - S.sk[my_class].nom_size);
            }
            f_time[2 + my_class] = sc_t;
        }
        if (tid == 0) f_time[0] = ctlp->cur_time;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }

    // self-collision state (CVX_Sim::UpdateCollisions, VX_Sim.cpp:1729-1755): the list lives in shared memory during the launch
    const bool col_on = COLLIDE && (flags & VXG_F_COLLISIONS_) != 0 && max_col > 0;
    // (what only the rare rebuild or thread 0's bookkeeping needs is not carried in registers through the loop: the 1024-thread variant
    // runs at 64 registers, and its spills were 4 % of the executed instructions and most of its long-scoreboard stalls)
    int n_col = col_on ? ctlp->n_col : 0;
    bool col_overflow = false;
    if (COLLIDE && col_on) {
        for (int i = tid; i < n_col; i += T) S.pairs[i] = B.col_pairs[(size_t)robot * max_col + i];
        if (tid == 0) {
            S.maxvel[0] = 0u; S.maxvel[1] = 0u; S.colflag[0] = 0u;
            S.colflag[1] = (ctlp->col_first != 0 ? 1u : 0u) | (ctlp->col_host != 0 ? 2u : 0u);   // ColEnableChanged; the host manages the list
            S.colstate[0] = ctlp->col_acc; S.colstate[1] = ctlp->max_vel; S.colstate[2] = ctlp->col_horizon; S.colstate[3] = ctlp->lattice_dim;
        }
        __syncthreads();
        if (n_col > 0) col_csr_build(S.pairs, n_col, c_seg, c_adj, (int*)S.red);
    }

    double t = ctlp->cur_time;
    int cur = 0, done = 0;
    bool diverged = false;
    if (FLOW) {
        const int lane = tid & 31, warp = tid >> 5;
        int k = 0;
        while (k < nsteps) {
            const long long s0 = step0 + k;
            __syncthreads();                                 // every warp has finished the previous window
            if (*f_div != 0x7fffffff) break;
            if (tid == 0) { f_time[1] = f_time[0]; f_div[1] = k; }   // where this window starts, for the divergence rewind
            if (s0 == latch) {   // IniCM = SS.CurCM at TimeStep entry (VX_Sim.cpp:1064-1075): COM after the previous step
                double mx = 0, my = 0, mz = 0, mm = 0;
                if (active) { const double m = (double)S.vk[st.vcls].mass; mx = own.px * m; my = own.py * m; mz = own.pz * m; mm = m; }
                block_sum4(mx, my, mz, mm, S.red);
                if (tid == 0) { ctlp->ini_com[0] = mx / mm; ctlp->ini_com[1] = my / mm; ctlp->ini_com[2] = mz / mm; ctlp->ini_latched = 1; }
            }
            if (s0 == eol) {     // SS.EndOfLifetimePosteriorY = getPosteriorY() at TimeStep entry (VX_Sim.cpp:1078-1082)
                double lo = 1e300;
                if (active && S.sk[st.vcls].mat_index != 5) lo = own.py;
                lo = -warp_max(-lo);
                __syncthreads();
                if ((tid & 31) == 0) S.red[tid >> 5] = lo;
                __syncthreads();
                if (tid == 0) { for (int i = 1; i < T / 32; i++) lo = fmin(lo, S.red[i]); ctlp->eol_post_y = lo; }
            }
            int wlen = nsteps - k < VXG_FLOW_R ? nsteps - k : VXG_FLOW_R;          // a latch ends the window before its step
            if (latch > s0 && latch - s0 < (long long)wlen) wlen = (int)(latch - s0);
            if (eol > s0 && eol - s0 < (long long)wlen) wlen = (int)(eol - s0);
            // rest-length deviation of every bond class for the steps of this window (their Scale entries were written >= one window ago)
            for (int i = tid; i < wlen * n_bc; i += T) {
                const int j = i / n_bc, b = i - j * n_bc, e = (k + j) & (VXG_FLOW_RING - 1);
                const uint32_t pr = S.bpair[b];
                r_dnom[e * MB + b] = 0.5f * (r_dsc[e * mv + (pr & 0xFFFFu)] + r_dsc[e * mv + (pr >> 16)]);
            }
            __syncthreads();
            for (int j = 0; j < wlen; ++j, ++k) {
                const int e = k & (VXG_FLOW_RING - 1);
                C.scale_c = r_scale + e * mv; C.dsc_c = r_dsc + e * mv; C.dnom_bc = r_dnom + e * MB;
                const float dt_inv = (step0 + k == 0) ? 0.0f : dt_inv_run;   // dt == 0 during the very first bond pass (VX_Sim.cpp:373)
                // +axis neighbours' owners must have integrated step k - 1 (and read the shares this warp is about to overwrite)
                if (j > 0) flow_wait(voxbar, plus_mask, (uint32_t)(k - 1) & 1u, lane);
                if (my_class < n_vc) {   // Scale of step k + R (VXS_Voxel.cpp:224-336) into the ring, R steps ahead of its readers
                    const int ea = (k + VXG_FLOW_R) & (VXG_FLOW_RING - 1), el = (k + VXG_FLOW_R - 1) & (VXG_FLOW_RING - 1);
                    const double sc_t = f_time[2 + my_class];
                    const double ns = scale_law(S.sk[my_class], r_scale[el * mv + my_class], sc_t, *S.sl);
                    f_time[2 + my_class] = sc_t + dt;
                    r_scale[ea * mv + my_class] = ns; r_dsc[ea * mv + my_class] = (float)(ns - S.sk[my_class].nom_size);
                }
                V3<float> F = mk3<float>(0, 0, 0), M = mk3<float>(0, 0, 0);
                bool dv = bond_eval_g<0, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M);
                dv |= bond_eval_g<1, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M);
                dv |= bond_eval_g<2, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M);
                if (dv) atomicMin(f_div, k);                 // Integrate() fails in this step (VX_Sim.cpp:1775-1777): acted on at the window's end
                tm_wait_st();
                mbar_arrive_warp(bondbar + warp, lane == 0);
                // -axis neighbours' owners must have staged their shares of step k (and read the positions this warp is about to overwrite)
                flow_wait(bondbar, minus_mask, (uint32_t)k & 1u, lane);
                if (active) voxel_phase_g<F_REGS>(smem, G, NV, C, R, tid, st, nb, own, fl, F, M);
                mbar_arrive_warp(voxbar + warp, lane == 0);
                if (tid == 0) f_time[0] += dt;               // CurTime += dt (VX_Sim.cpp:1929)
            }
            done = k;
        }
        const int dstep = *f_div;
        if (dstep != 0x7fffffff) {   // the robot stops where the reference's Integrate() first returned false
            diverged = true; done = dstep;
            if (tid == 0) { double tt = f_time[1]; for (int i = f_div[1]; i < dstep; i++) tt += dt; f_time[0] = tt; }
        }
        if (tid == 0) t = f_time[0];
        cur = done & (VXG_FLOW_RING - 1);
    } else {
        for (int k = 0; k < nsteps; ++k) {
            const long long s = step0 + k;
            if (s == latch) {   // IniCM = SS.CurCM at TimeStep entry (VX_Sim.cpp:1064-1075): COM after the previous step
                double mx = 0, my = 0, mz = 0, mm = 0;
                if (active) { const double m = (double)S.vk[st.vcls].mass; mx = own.px * m; my = own.py * m; mz = own.pz * m; mm = m; }
                block_sum4(mx, my, mz, mm, S.red);
                if (tid == 0) { ctlp->ini_com[0] = mx / mm; ctlp->ini_com[1] = my / mm; ctlp->ini_com[2] = mz / mm; ctlp->ini_latched = 1; }
            }
            if (s == eol) {     // SS.EndOfLifetimePosteriorY = getPosteriorY() at TimeStep entry (VX_Sim.cpp:1078-1082)
                double lo = 1e300;
                if (active && S.sk[st.vcls].mat_index != 5) lo = own.py;
                lo = -warp_max(-lo);
                __syncthreads();
                if ((tid & 31) == 0) S.red[tid >> 5] = lo;
                __syncthreads();
                if (tid == 0) { for (int i = 1; i < T / 32; i++) lo = fmin(lo, S.red[i]); ctlp->eol_post_y = lo; }
            }
            C.scale_c = S.scale_c + cur * mv;
            C.dsc_c = S.dsc_c + cur * mv;
            C.dnom_bc = S.dnom_bc + cur * MB;
            const float dt_inv = (s == 0) ? 0.0f : dt_inv_run;   // dt == 0 during the very first bond pass (VX_Sim.cpp:373)

            // Scale of every class for the NEXT step (VXS_Voxel.cpp:224-336), by the highest threads, into the other buffer half
            for (int i = T - 1 - tid; i < n_vc; i += T) {
                const double last = S.scale_c[cur * mv + i];
                const double ns = scale_law(S.sk[i], last, t, *S.sl);
                S.scale_c[(cur ^ 1) * mv + i] = ns;
                S.dsc_c[(cur ^ 1) * mv + i] = (float)(ns - S.sk[i].nom_size);
            }
            // ---- bond phase: reads positions / orientations of the previous step, writes only what this thread owns --------------
            if (COLLIDE && col_on && tid == 0) {
                // UpdateCollisions, the bookkeeping (VX_Sim.cpp:1729-1755): one thread keeps MaxDispSinceLastBondUpdate in fp64 and posts
                // its verdict, tagged with the step, where the others pick it up once their bonds are done (dt is still 0 in step 0)
                const unsigned int mvb = S.maxvel[(k + 1) & 1];            // written by the previous step's voxel phase
                double col_acc = S.colstate[0], max_vel = S.colstate[1];
                if (k > 0) max_vel = sqrt((double)__uint_as_float(mvb));
                S.maxvel[k & 1] = 0u;                                      // this step's slot; next read after two barriers
                col_acc += fabs(max_vel * (s == 0 ? 0.0 : dt) / S.colstate[3]);
                const unsigned int cf = S.colflag[1];
                const bool rebuild = !(cf & 2u) && (col_acc > (S.colstate[2] - 1.0) / 2 || (cf & 1u));
                if (rebuild) { S.colflag[1] = cf & ~1u; col_acc = 0.0; }
                S.colstate[0] = col_acc; S.colstate[1] = max_vel;
                *(volatile unsigned int*)S.colflag = ((unsigned int)(k + 1) << 1) | (rebuild ? 1u : 0u);
            }
            V3<float> F = mk3<float>(0, 0, 0), M = mk3<float>(0, 0, 0);
            int div = 0;
            if (FLUID && fluid) {   // DragForce of this step from the state and bond strains of the previous one (LW/VX_Sim.cpp:1516-1597)
                for (int u = tid; u < C.n_vert; u += T) mesh_vertex<float>(R, C, u);
                __syncthreads();
                for (int i = tid; i < C.n_facet; i += T) facet_task<float>(R, C, i);
                __syncthreads();
                if (active) F = voxel_drag<float>(R, tid, st);
                if (fluid_alias) {   // the facet drags share memory with the staging slots the bond phase writes next
                    __syncthreads();
                    if (tid == 0) {  // ... and with the "no neighbour" slots, which have to read as zeros again
                        R4<float> za; za.x = za.y = za.z = za.w = 0.0f; R2<float> zb; zb.x = zb.y = 0.0f;
#pragma unroll
                        for (int a = 0; a < 3; a++) { ((R4<float>*)(smem + G.off_stA[a]))[NV] = za; ((R2<float>*)(smem + G.off_stB[a]))[NV] = zb; }
                    }
                }
            }
#ifdef VXG_TRACE
            const long long tr0 = clock64();
#endif
            div |= bond_eval_g<0, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M) ? 1 : 0;
            div |= bond_eval_g<1, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M) ? 1 : 0;
            div |= bond_eval_g<2, HOIST, F_REGS, FLUID>(smem, G, NV, C, bkm, R.strain, tid, st, own, fl, dt_inv, tm, F, M) ? 1 : 0;
            tm_wait_st();
            if (COLLIDE && col_on) {   // positions are still those of the previous step here
                unsigned int verdict;
                do { verdict = *(volatile unsigned int*)S.colflag; } while ((verdict >> 1) != (unsigned int)(k + 1));
                if (verdict & 1u) {    // CalcL1Bonds (VX_Sim.cpp:2357-2413), by the whole CTA
                    // (copies: the out-of-line rebuild takes its views by reference, and what it is handed must live in memory --
                    // the step loop's own R and C stay in registers)
                    // (views for the out-of-line rebuild, which takes them by reference so that they live in memory: built here from
                    // scratch with the few members it reads -- a copy of the step loop's own R and C made ptxas keep THOSE in local
                    // memory and reload their members all through the loop: 42 local loads per voxel-timestep at 10^3)
                    RobotView<float> R2; memset(&R2, 0, sizeof R2);
                    R2.n_vox = n_vox; R2.nv = NV; R2.pos = (double*)(smem + G.off_pos); R2.stat = (VoxStat*)(smem + G.off_stat);
                    RobotConsts<float> C2; memset(&C2, 0, sizeof C2);
                    C2.scale_c = S.scale_c + cur * mv;
                    ColTables CT2;
                    CT2.surf = B.col_surf + (size_t)blockIdx.x * NV; CT2.lat4 = B.col_lat4 + (size_t)blockIdx.x * NV * 4; CT2.near = B.col_near + (size_t)blockIdx.x * NV;
                    CT2.n_surf = ctlp->n_surf; CT2.hops = ctlp->col_hops;
                    n_col = col_rebuild<float>(R2, C2, CT2, (float)S.colstate[2], (float)S.colstate[3], S.pairs, max_col, (int*)S.red);
                    if (n_col > max_col) { col_overflow = true; break; }
                    if (n_col > 0) col_csr_build(S.pairs, n_col, c_seg, c_adj, (int*)S.red);
                }
                if (n_col > 0 && active) {   // CVXS_BondCollision::CalcContactForce for this voxel's contact bonds, in list order
                    const uint32_t sg = c_seg[tid];
                    V3<float> fc = mk3<float>(0, 0, 0);
#pragma unroll 1   // (four inlined copies of the contact force were 7 KB of the step loop: the SM's instruction cache holds 32 KB)
                    for (uint32_t i = sg & 0xFFFFu; i < (sg >> 16); i++) col_contact<float>(R, C, st, own, c_adj[i], fc);
                    F = F + fc;
                }
            }
#ifdef VXG_TRACE
            const long long tr1 = clock64();
#endif
            if (__syncthreads_or(div)) { diverged = true; break; }   // Integrate() returns before the voxel loop (VX_Sim.cpp:1777)
#ifdef VXG_TRACE
            const long long tr2 = clock64();
#endif

            // rest-length deviation of every bond class for the NEXT step
            for (int i = tid; i < n_bc; i += T) {
                const uint32_t pr = S.bpair[i];
                const float* dn = S.dsc_c + (cur ^ 1) * mv;
                S.dnom_bc[(cur ^ 1) * MB + i] = 0.5f * (dn[pr & 0xFFFFu] + dn[pr >> 16]);
            }
            // ---- voxel phase: gather the shares of the -X/-Y/-Z bonds (Force2 = -Force1, Moment2), integrate -------------------
            float sp2 = 0.0f;
            if (active) sp2 = voxel_phase_g<F_REGS, F_EXTRA>(smem, G, NV, C, R, tid, st, nb, own, fl, F, M);
            if (COLLIDE && col_on) {   // SS.MaxVoxVel for the next UpdateCollisions
                for (int o = 16; o > 0; o >>= 1) sp2 = fmaxf(sp2, __shfl_xor_sync(0xffffffffu, sp2, o));
                if ((tid & 31) == 0) atomicMax(&S.maxvel[k & 1], __float_as_uint(sp2));
            }
#ifdef VXG_TRACE
            const long long tr3 = clock64();
#endif
            __syncthreads();
#ifdef VXG_TRACE
            if (robot == 0 && (tid & 31) == 0 && (k == 50 || k == 51) && step0 > 2500)
                printf("TRACE k %d warp %2d bond_start %lld bond_end %lld bar1_exit %lld vox_end %lld bar2_exit %lld\n", k, tid >> 5, tr0 % 10000000, tr1 % 10000000, tr2 % 10000000, tr3 % 10000000, clock64() % 10000000);
#endif
            cur ^= 1;
            t += dt;                                                 // CurTime += dt (VX_Sim.cpp:1929)
            done = k + 1;
        }
    }

    // ---- stage out -------------------------------------------------------------------------------------------------------
    __syncthreads();
    if (PERM) {
        if (active) {   // this thread's voxel back to its place in the blob
            const RobotView<float> Rg = make_view<float>(gblob, L, n_vox);
            Rg.pos[vme] = own.px; Rg.pos[NV + vme] = own.py; Rg.pos[2 * NV + vme] = own.pz;
            Rg.quat[vme] = own.q;
            Rg.momA[vme] = R.momA[tid]; Rg.momB[vme] = R.momB[tid];
        }
    } else if (tid == 0) { fence_async_smem(); bulk_s2g(gblob, smem, (unsigned int)G.dyn_front); bulk_s2g_finish(); }   // positions .. momenta back to the blob
    if (active) ((uint32_t*)(gblob + L.off_flags))[vme] = fl;
    {
        const RobotView<float> Rg = make_view<float>(gblob, L, n_vox);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            BondHist<float> h;
            if (a == 0) hist_load<0>(tm, h); else if (a == 1) hist_load<1>(tm, h); else hist_load<2>(tm, h);
            if (st.nbr[a] >= 0) {   // slots of bonds that do not exist were never touched
                R4<float> x, y;
                x.x = h.p2.x; x.y = h.p2.y; x.z = h.p2.z; x.w = h.a1.x;
                y.x = h.a1.y; y.y = h.a1.z; y.z = h.a2.x; y.w = h.a2.y;
                Rg.histA(a)[vme] = x; Rg.histB(a)[vme] = y; Rg.histC(a)[vme] = h.a2.z;
            }
        }
    }
    for (int i = tid; i < n_vc; i += T) B.scale_c[(size_t)robot * B.max_vclass + i] = FLOW ? r_scale[cur * mv + i] : S.scale_c[cur * mv + i];
    if (FLUID) copy4(B.blobs + (size_t)robot * LF.total_bytes + LF.off_strain, R.strain, sizeof(float) * 6 * NV, tid, T);
    if (COLLIDE && col_on) {
        const int nc = n_col < max_col ? n_col : max_col;
        for (int i = tid; i < nc; i += T) B.col_pairs[(size_t)robot * max_col + i] = S.pairs[i];
    }
    if (tid == 0) {
        ctlp->cur_time = t;
        ctlp->step = step0 + done;
        unsigned int st2 = status0;
        if (COLLIDE && col_on) {
            ctlp->n_col = n_col < max_col ? n_col : max_col; ctlp->col_first = (S.colflag[1] & 1u) ? 1 : 0; ctlp->col_acc = S.colstate[0];
            ctlp->max_vel = done > 0 ? sqrt((double)__uint_as_float(S.maxvel[(done - 1) & 1])) : S.colstate[1];
        }
        if (col_overflow) st2 |= VXG_ST_COL_OVERFLOW_ | VXG_ST_DIVERGED_;
        if (diverged) st2 |= VXG_ST_DIVERGED_;
        if (step0 + done >= total) st2 |= VXG_ST_FINISHED_;
        ctlp->status = st2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // warp 0's own address IS the allocation's base (lane quarter 0, column block 0); the slot it was published through has been
    // reused as reduction scratch since
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(TM_COLS) : "memory");
}

// shared memory of one CTA: mv / mb = capacities of the class tables as the kernel instantiation has them (mb includes the zero class)
inline size_t gather_smem_bytes(int nv, int mv, int mb, bool flow, int max_col = 0, int nvert_cap = 0, int nfacet_cap = 0) {
    size_t fl = 0;
    if (nvert_cap > 0)
        fl = vx_align16(sizeof(float) * 6 * nv) + vx_align16(sizeof(uint16_t) * nv) + vx_align16(sizeof(Facet) * (size_t)nfacet_cap) + vx_align16(sizeof(Corner8) * (size_t)nvert_cap) +
             (fluid_scratch_aliases(nv, nvert_cap, nfacet_cap) ? 0 : vx_align16(sizeof(double) * 3 * (size_t)nvert_cap) + vx_align16(sizeof(float) * 3 * (size_t)nfacet_cap));
    return make_glayout(nv).total + sizeof(BondKm) * (size_t)mb + tables_bytes<float>(mv, mb, true, max_col) + (flow ? flow_bytes(mv, mb) : 0) +
           (max_col > 0 ? vx_align16(sizeof(uint32_t) * nv + sizeof(uint16_t) * 2 * (size_t)max_col) : 0) + fl;
}

// ---- launch plumbing (host) ----------------------------------------------------------------------------------------------------
// class-table capacities the step loop is specialised for (palette robots: <= 8 voxel classes, <= 63 bond classes after
// re-keying by voxel-class pair); larger tables run the instantiation with run-time capacities
#define VXG_CV 8
#define VXG_CB 64
template <int NV, bool FLOW, int CV, int CB, bool COLLIDE = false, bool FLUID = false, int NVC = 0> inline cudaError_t launch_gather_form(const BatchDev<float>& B, const LaunchCfg& c) {
    auto kern = vx_step_kernel_g<NV, FLOW, CV, CB, COLLIDE, FLUID, NVC>;
    const size_t smem = gather_smem_bytes(NV ? NV : c.nv, CV > 0 ? CV : B.max_vclass, CB > 0 ? CB : B.max_bclass + 1, FLOW, COLLIDE ? B.max_col : 0,
                                          FLUID ? (NVC ? NVC : B.L.nvert_cap) : 0, FLUID ? (NVC ? ((2 * NVC + 64 + 31) & ~31) : B.L.nfacet_cap) : 0);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<c.n_robots, NV ? NV : c.nv, smem, c.stream>>>(B, c.n_steps);
    return cudaGetLastError();
}
// does a CTA of this form fit next to the others the launch bounds ask for (227 KB per CTA, 228 KB per SM with 1 KB reserved per CTA)?
// (the launch bounds size the registers for vxg_ctas() CTAs per SM; when shared memory allows fewer, fewer are resident -- no error)
inline bool gather_fits(int nv, int mv, int mb, bool flow, int max_col = 0, int nvert_cap = 0, int nfacet_cap = 0) {
    return gather_smem_bytes(nv, mv, mb, flow, max_col, nvert_cap, nfacet_cap) <= 227u * 1024u;
}
// VXG_RT_STRIDE=1: small robots run the run-time-stride image (NV = 0): one image, exactly the warps a robot needs.  Measured
// SLOWER than the compile-time strides (dense 6^3: 1.80e10 vs 2.13e10; CPPN-like blobs with self-collision at their exact
// strides 8.3e9 vs 1.00e10 at one shared compile-time stride): the run-time offsets cost registers and address arithmetic in
// every access, and a CTA's idle warps were never the expensive part.  Kept selectable as the record of that experiment.
inline bool vxg_use_rt(int nv) {
    const char* e = getenv("VXG_RT_STRIDE");
    return e && e[0] == '1' && nv <= 256 && nv % 32 == 0;
}
template <int NV> inline cudaError_t launch_gather(const BatchDev<float>& B, const LaunchCfg& c, bool flow) {
    const bool fixed = B.max_vclass <= VXG_CV && B.max_bclass + 1 <= VXG_CB && gather_fits(NV ? NV : c.nv, VXG_CV, VXG_CB, false);
    if (flow) {   // experiment form, run-time capacities only
        if (gather_fits(NV ? NV : c.nv, B.max_vclass, B.max_bclass + 1, true)) return launch_gather_form<NV, true, 0, 0>(B, c);
    }
    return fixed ? launch_gather_form<NV, false, VXG_CV, VXG_CB>(B, c) : launch_gather_form<NV, false, 0, 0>(B, c);
}
// the same with self-collision: 0 = launched, 1 = this stride / table size has no gather instantiation that fits (caller falls back)
template <int NV> inline int launch_gather_collide(const BatchDev<float>& B, const LaunchCfg& c, cudaError_t* err) {
    if (B.max_vclass <= VXG_CV && B.max_bclass + 1 <= VXG_CB && gather_fits(NV ? NV : c.nv, VXG_CV, VXG_CB, false, B.max_col)) { *err = launch_gather_form<NV, false, VXG_CV, VXG_CB, true>(B, c); return 0; }
    if (gather_fits(NV ? NV : c.nv, B.max_vclass, B.max_bclass + 1, false, B.max_col)) { *err = launch_gather_form<NV, false, 0, 0, true>(B, c); return 0; }
    return 1;
}
// land/water layouts (with or without self-collision); run-time table capacities (measured: with the compile-time capacities of the
// plain kernels -- 64 bond classes' worth of tables -- an 8^3 swimmer no longer fits twice on an SM: 1.12e10 instead of 1.40e10;
// with 32, which do fit, 1.40e10: no gain over the run-time capacities here)
#define VXG_CB_FLUID 32
#define VXG_NVC_512 416    // mesh vertices of an 8^3 cube's surface (386), rounded up to make_layout's granularity
template <int NV> inline int launch_gather_fluid(const BatchDev<float>& B, const LaunchCfg& c, bool collide, cudaError_t* err) {
    if constexpr (NV == 512) {   // 8^3 swimmers (BASELINE config 5): everything compile-time, sized so that two robots still share an SM
        const size_t two_per_sm = (228u * 1024u) / 2 - 1024u;
        if (!collide && B.max_vclass <= VXG_CV && B.max_bclass + 1 <= VXG_CB_FLUID && B.L.nvert_cap <= VXG_NVC_512 &&
            gather_smem_bytes(NV, VXG_CV, VXG_CB_FLUID, false, 0, VXG_NVC_512, (2 * VXG_NVC_512 + 64 + 31) & ~31) <= two_per_sm) {
            *err = launch_gather_form<NV, false, VXG_CV, VXG_CB_FLUID, false, true, VXG_NVC_512>(B, c);
            return 0;
        }
    }
    if (!gather_fits(NV ? NV : c.nv, B.max_vclass, B.max_bclass + 1, false, collide ? B.max_col : 0, B.L.nvert_cap, B.L.nfacet_cap)) return 1;
    *err = collide ? launch_gather_form<NV, false, 0, 0, true, true>(B, c) : launch_gather_form<NV, false, 0, 0, false, true>(B, c);
    return 0;
}

}  // namespace vx
