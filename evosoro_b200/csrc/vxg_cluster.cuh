// vxg_cluster.cuh -- the production step kernel for robots of 513 .. 1024 voxels (stride 1024: the 10^3 lattices of the headline
// benchmark), split over a thread-block CLUSTER of two CTAs.
//
// Same arithmetic, same order, bit-identical results as vx_step_kernel_g<1024, ...> (vxg_gather.cuh).  What changes is who waits for whom:
//
//   * A 1024-voxel robot at one voxel per thread needs every register of an SM, so the single-CTA form has ONE barrier domain per
//     SM: whenever the warps of the robot wait for each other (twice per timestep), the SM's issue slots idle -- 13 % of a step at
//     the bonds -> voxels barrier alone (clock64 trace, profiles/r02_sizes.txt), and the 8^3 robots, two per SM, show what a second
//     domain recovers (0.90 -> 1.01 of the roofline).
//   * Here CTA 0 of the cluster owns voxels [0, split) and CTA 1 the rest (split: RobotCtl::c2_split, chosen on the host so that the two
//     halves have the same work: a dense 10^3 robot is cut 488 | 512 because its top layer has no +Z bonds) -- state, staging slots, tensor-memory history and all -- so an SM
//     holds HALVES OF TWO DIFFERENT ROBOTS: two independent barrier domains that fill each other's waits.  The two halves of one
//     robot sit on two SMs and meet at the hardware cluster barrier (barrier.cluster.arrive.release / wait.acquire) where the
//     single-CTA form has its two CTA barriers.
//   * The only data that crosses: a bond whose +axis neighbour lives in the partner CTA reads that voxel's position / orientation
//     through distributed shared memory (ld.shared::cluster), and the -axis neighbour gathers the bond's share from the owner's
//     staging slot the same way -- at 10^3 the 100 bonds through the plane between z = 4 and z = 5 plus a dozen in-plane ones.
//     Which warps have such lanes is static: a warp-uniform bit per axis selects the remote-capable loads, every other warp runs the
//     plain shared-memory path.
//
// MEASURED (B200, 10^3 x 1024 robots x 1000 timesteps): 1.85e10 voxel-timesteps/s with the release / acquire cluster barrier, 1.91e10
// with a CTA barrier + relaxed arrival (VXG_C2_SYNC=2) -- against 1.97e10 for the single-CTA kernel.  The premise was wrong: two
// 16-warp domains per SM issue no more than one 32-warp domain (70.9 % vs 72.9 % issue-active), the split costs 4 % more
// instructions, and a CTA now waits for a partner whose pace depends on what it shares its own SM with (28 % of warp samples at the
// cluster barrier; profiles/r02_ncu_cluster10.txt).  NOT the default; VXG_KERNEL_FORM=cluster selects it, the bit-identity tests cover it.
//
// Replaces CVX_Sim::Integrate (VX_Sim.cpp:1763-1933) for robots of the production family (pack_compact, vx_pack.h).
#pragma once
#include "vxg_gather.cuh"

namespace vx {

#define VXG_C2_HV 512          // voxels = threads per CTA
#define VXG_C2_NV 1024         // voxel stride of the blob

// shared memory of one CTA (byte offsets): its half of positions / orientations / momenta, the -axis neighbour table, staging slots
struct CLayout {
    size_t off_pos, off_quat, off_momA, off_momB, off_nbm, off_stA[3], off_stB[3], total;
};
VX_HD constexpr CLayout make_clayout() {
    CLayout G{};
    const size_t hv = VXG_C2_HV;
    size_t o = 0;
    G.off_pos = o;  o = vx_align32(o + 3 * hv * sizeof(double));
    G.off_quat = o; o = vx_align32(o + hv * 16);
    G.off_momA = o; o = vx_align32(o + hv * 16);
    G.off_momB = o; o = vx_align32(o + hv * 8);
    G.off_nbm = o;  o = vx_align32(o + hv * 8);
    for (int a = 0; a < 3; a++) {
        G.off_stA[a] = o; o = vx_align32(o + (hv + 1) * 16);
        G.off_stB[a] = o; o = vx_align32(o + (hv + 1) * 8);
    }
    G.total = o;
    return G;
}

// ---- cluster primitives ----------------------------------------------------------------------------------------------------------
__device__ inline uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ inline uint32_t cl_mapa(uint32_t addr, uint32_t rank) {   // shared::cta address -> the same location in CTA `rank`'s window
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ inline void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ inline void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ inline void cl_sync() { cl_arrive(); cl_wait(); }
// The per-step form.  VXG_C2_SYNC 1: the same release / acquire pair (ptxas: MEMBAR.ALL.GPU + UCGABAR_ARV / _WAIT + CCTL.IVALL).  2: a CTA
// barrier -- after which every shared-memory store of this CTA has been performed -- and a RELAXED arrival: what crosses is shared
// memory only, which no cache sits in front of, so the GPU-scope fence of the release form orders nothing this kernel needs.
#ifndef VXG_C2_SYNC
#define VXG_C2_SYNC 1
#endif
__device__ inline void cl_step_sync() {
#if VXG_C2_SYNC == 2
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    cl_wait();
#else
    cl_sync();
#endif
}
__device__ inline double cl_ld_f64(uint32_t a) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ inline R4<float> cl_ld_r4(uint32_t a) {
    R4<float> v; asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory"); return v;
}
__device__ inline R2<float> cl_ld_r2(uint32_t a) {
    R2<float> v; asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory"); return v;
}
__device__ inline void cl_st_u16(uint32_t a, uint16_t v) { asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ inline void cl_st_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ inline void cl_st_f64(uint32_t a, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// Per-thread / per-warp static facts ride in the unused bits of the voxel's flags word (bits 0..3 are the state the blob keeps):
//   8 + a : this voxel's +axis neighbour lives in the partner CTA          12 + a : its -axis neighbour's owner does
//   16 + a: some lane of this warp has a bond on axis a                    20 + a : some lane's +axis neighbour is in the partner CTA
//   24    : some lane's -axis neighbour's owner is in the partner CTA
#define VXG_C2_FL_STATE 0xFu

// One owned bond (bond_eval_g of vxg_gather.cuh with the neighbour fetched from either CTA of the cluster).
template <int AX>
__device__ __forceinline__ bool bond_eval_c2(unsigned char* smem, const CLayout& G, const RobotConsts<float>& C, const BondKm* bkm, int tid, uint32_t rank,
                                             const VoxStat& st, const OwnState<float>& own, uint32_t& fl, float dt_inv, uint32_t tm, V3<float>& F, V3<float>& M) {
    if (!(fl & (0x10000u << AX))) return false;          // warp-uniform: no lane has a bond on this axis
    const int w = st.nbr[AX];                            // index inside the CTA that holds the neighbour
    const int wq = w >= 0 ? w : tid;
    const int bc = st.bcls[AX];
    float ha[8], hc;
    hist_issue<AX>(tm, ha, hc);   // the history is on its way from tensor memory while the neighbour's state is fetched

    BondK<float> k = bond_from_km(bkm[bc]);
    double nx, ny, nz;
    R4<float> qb;
    if (fl & (0x100000u << AX)) {                        // warp-uniform: a lane's neighbour lives in the partner CTA
        const uint32_t tr = ((fl >> (8 + AX)) & 1u) ^ rank;
        const uint32_t ap = cl_mapa(smem_u32(smem + G.off_pos) + 8u * (uint32_t)wq, tr), aq = cl_mapa(smem_u32(smem + G.off_quat) + 16u * (uint32_t)wq, tr);
        nx = cl_ld_f64(ap); ny = cl_ld_f64(ap + 8u * VXG_C2_HV); nz = cl_ld_f64(ap + 16u * VXG_C2_HV);
        qb = cl_ld_r4(aq);
    } else {
        const double* pos = (const double*)(smem + G.off_pos);
        nx = pos[wq]; ny = pos[VXG_C2_HV + wq]; nz = pos[2 * VXG_C2_HV + wq];
        qb = ((const R4<float>*)(smem + G.off_quat))[wq];
    }
    double dpa[3] = {nx - own.px, ny - own.py, nz - own.pz};
    dpa[AX] -= C.bk_dist[bc];
    const V3<float> e = mk3<float>((float)dpa[0], (float)dpa[1], (float)dpa[2]);
    const R4<float> qa = own.q;
    Q4<float> q1, q2;
    q1.w = qa.x; q1.x = qa.y; q1.y = qa.z; q1.z = qa.w;
    q2.w = qb.x; q2.x = qb.y; q2.y = qb.z; q2.z = qb.w;
    bool small = (fl >> AX) & 1u;
    BondHist<float> h;
    hist_take(ha, hc, h);
    BondOut<float> o;
    bond_force<AX, float, true, false>(k, e, q1, q2, C.dnom_bc[bc], C.half_bond_z, dt_inv, true, small, h, o);
    // _LastPos2/_LastAngle* keep their values on a mode-flip step (VXS_BondInternal.cpp:157)
    if (__any_sync(0xffffffffu, o.changed)) { BondHist<float> old; hist_load<AX>(tm, old); if (o.changed) h = old; }
    hist_store<AX>(tm, h);
    fl = (fl & ~(1u << AX)) | ((small ? 1u : 0u) << AX);
    F = F + o.f1;
    M = M + o.m1;
    R4<float> a; a.x = o.f1.x; a.y = o.f1.y; a.z = o.f1.z; a.w = o.m2.x;
    R2<float> b; b.x = o.m2.y; b.y = o.m2.z;
    ((R4<float>*)(smem + G.off_stA[AX]))[tid] = a;
    ((R2<float>*)(smem + G.off_stB[AX]))[tid] = b;
    return o.strain > 100.0f;
}

// Voxel phase (voxel_phase_g with the shares gathered from either CTA).
__device__ __forceinline__ void voxel_phase_c2(unsigned char* smem, const CLayout& G, const RobotConsts<float>& C, int tid, uint32_t rank, const VoxStat& st, uint2 nb,
                                               OwnState<float>& own, uint32_t& fl, V3<float> F, V3<float> M) {
    const int mX = nb.x & 0xFFFFu, mY = nb.x >> 16, mZ = nb.y & 0xFFFFu;
    R4<float> ax, ay, az; R2<float> bx, by, bz;
    if (fl & 0x1000000u) {                               // warp-uniform: a lane's -axis neighbour is owned by the partner CTA
        const uint32_t rx = ((fl >> 12) & 1u) ^ rank, ry = ((fl >> 13) & 1u) ^ rank, rz = ((fl >> 14) & 1u) ^ rank;
        ax = cl_ld_r4(cl_mapa(smem_u32(smem + G.off_stA[0]) + 16u * (uint32_t)mX, rx)); bx = cl_ld_r2(cl_mapa(smem_u32(smem + G.off_stB[0]) + 8u * (uint32_t)mX, rx));
        ay = cl_ld_r4(cl_mapa(smem_u32(smem + G.off_stA[1]) + 16u * (uint32_t)mY, ry)); by = cl_ld_r2(cl_mapa(smem_u32(smem + G.off_stB[1]) + 8u * (uint32_t)mY, ry));
        az = cl_ld_r4(cl_mapa(smem_u32(smem + G.off_stA[2]) + 16u * (uint32_t)mZ, rz)); bz = cl_ld_r2(cl_mapa(smem_u32(smem + G.off_stB[2]) + 8u * (uint32_t)mZ, rz));
    } else {
        ax = ((const R4<float>*)(smem + G.off_stA[0]))[mX]; ay = ((const R4<float>*)(smem + G.off_stA[1]))[mY]; az = ((const R4<float>*)(smem + G.off_stA[2]))[mZ];
        bx = ((const R2<float>*)(smem + G.off_stB[0]))[mX]; by = ((const R2<float>*)(smem + G.off_stB[1]))[mY]; bz = ((const R2<float>*)(smem + G.off_stB[2]))[mZ];
    }
    // the accumulator of the push form received -Force1 of the X, Y, Z bond in that order, starting from zero
    F = F + mk3<float>(((0.0f - ax.x) - ay.x) - az.x, ((0.0f - ax.y) - ay.y) - az.y, ((0.0f - ax.z) - ay.z) - az.z);
    M = M + mk3<float>(((0.0f + ax.w) + ay.w) + az.w, ((0.0f + bx.x) + by.x) + bz.x, ((0.0f + bx.y) + by.y) + bz.y);
    const int vc = st.vcls;
    const VoxK<float> kv = C.vk[vc];
    R4<float>* momA = (R4<float>*)(smem + G.off_momA); R2<float>* momB = (R2<float>*)(smem + G.off_momB);
    VoxState<float> vs;
    vs.px = own.px; vs.py = own.py; vs.pz = own.pz;
    const R4<float> ma = momA[tid]; const R2<float> mb = momB[tid];
    vs.q.w = own.q.x; vs.q.x = own.q.y; vs.q.y = own.q.z; vs.q.z = own.q.w;
    vs.lm = mk3<float>(ma.x, ma.y, ma.z);
    vs.am = mk3<float>(ma.w, mb.x, mb.y);
    const bool sf = voxel_integrate<float, false>(kv, F, M, 0.5 * C.scale_c[vc], C.floor_on, false, C.dt, nullptr, nullptr, nullptr, vs);
    own.px = vs.px; own.py = vs.py; own.pz = vs.pz;
    own.q.x = vs.q.w; own.q.y = vs.q.x; own.q.z = vs.q.y; own.q.w = vs.q.z;
    double* pos = (double*)(smem + G.off_pos);
    pos[tid] = vs.px; pos[VXG_C2_HV + tid] = vs.py; pos[2 * VXG_C2_HV + tid] = vs.pz;
    ((R4<float>*)(smem + G.off_quat))[tid] = own.q;
    R4<float> na; na.x = vs.lm.x; na.y = vs.lm.y; na.z = vs.lm.z; na.w = vs.am.x;
    R2<float> nb2; nb2.x = vs.am.y; nb2.y = vs.am.z;
    momA[tid] = na; momB[tid] = nb2;
    fl = sf ? (fl | VX_FL_STATIC_FRIC) : (fl & ~VX_FL_STATIC_FRIC);
}

//   CV, CB : compile-time capacities of the voxel / bond class tables (0 = the batch's run-time limits), as in vx_step_kernel_g
template <int CV, int CB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(VXG_C2_HV, 2) vx_step_kernel_c2(BatchDev<float> B, int n_steps) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int HV = VXG_C2_HV, NV = VXG_C2_NV, T = VXG_C2_HV;
    constexpr BlobLayout L = make_layout(NV, 4);
    constexpr CLayout G = make_clayout();
    constexpr int TM_COLS = tm_cols_for(HV);
    const int robot = blockIdx.x >> 1;
    const uint32_t rank = cl_rank();
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    RobotCtl* ctlp = B.ctl + robot;
    const int split = ctlp->c2_split;                   // CTA 0: voxels [0, split), CTA 1: [split, n_vox)
    const int vbase = rank ? split : 0;
    const int vme = vbase + tid;                        // the voxel this thread owns
    const unsigned int status0 = ctlp->status;
    if (status0 & (VXG_ST_DIVERGED_ | VXG_ST_FINISHED_)) return;   // (both CTAs read the same word: nothing writes it before the last cluster barrier)
    const long long step0 = ctlp->step, total = ctlp->total_steps, latch = ctlp->latch_step, eol = ctlp->eol_step;
    long long remaining = total - step0;
    int nsteps = remaining < (long long)n_steps ? (int)remaining : n_steps;
    if (nsteps <= 0) { if (rank == 0 && tid == 0) ctlp->status = status0 | VXG_ST_FINISHED_; return; }

    unsigned char* gblob = B.blobs + (size_t)robot * L.total_bytes;
    const int MB = CB > 0 ? CB : B.max_bclass + 1;   // one more bond class than the batch's limit: the zero class
    const int mv = CV > 0 ? CV : B.max_vclass;
    BondKm* bkm = (BondKm*)(smem + G.total);
    unsigned char* tabp = smem + G.total + sizeof(BondKm) * (size_t)MB;
    SmemTables<float> S = carve_tables<float>(tabp, mv, MB, true, 0);
    const int n_vox = ctlp->n_vox, n_vc = ctlp->n_vclass, n_bc = ctlp->n_bclass;
    const unsigned int flags = ctlp->flags;
    const int ZC = n_bc;               // index of the zero class
    unsigned int* divflag = S.arrcnt;  // a bond of either half diverged (posted into both CTAs)

    // ---- tensor memory: this CTA's 16 warps x 32 columns ------------------------------------------------------------------
    uint32_t* tm_slot = (uint32_t*)S.red;   // the reduction scratch is idle until the first latch
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tm_slot)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- stage in: this half's slices of the position planes, orientations and momenta, six bulk copies by the TMA engine --------
    if (tid == 0) {
        mbar_init(S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(S.mbar, (unsigned int)(HV * (3 * 8 + 16 + 16 + 8)));
        const size_t h8 = (size_t)vbase * 8, h16 = (size_t)vbase * 16;   // (CTA 0 loads 512 entries too: the ones past its share are never read)
        for (int p = 0; p < 3; p++) bulk_g2s(smem + G.off_pos + (size_t)p * HV * 8, gblob + L.off_pos + (size_t)p * NV * 8 + h8, HV * 8, S.mbar);
        bulk_g2s(smem + G.off_quat, gblob + L.off_quat + h16, HV * 16, S.mbar);
        bulk_g2s(smem + G.off_momA, gblob + L.off_momA + h16, HV * 16, S.mbar);
        bulk_g2s(smem + G.off_momB, gblob + L.off_momB + h8, HV * 8, S.mbar);
        *divflag = 0u;
    }
    {
        uint2 none; none.x = none.y = (uint32_t)HV | ((uint32_t)HV << 16);
        ((uint2*)(smem + G.off_nbm))[tid] = none;
        // staging slots start as zeros and stay so where nothing is ever staged: the "no neighbour" slot [HV] and the slots of warps
        // that skip an axis altogether
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
    cl_sync();   // both CTAs run and have initialised their neighbour tables (remote stores follow)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(S.mbar, 0);   // the bulk copies have landed
    const uint32_t tm = *tm_slot + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)(tid >> 7) * VXG_TM_COLS_PER_WARP;   // lane (warp % 4) * 32, this warp's columns
    for (int i = tid; i < n_vc; i += T) S.dsc_c[i] = (float)(S.scale_c[i] - S.sk[i].nom_size);
    for (int i = tid; i <= n_bc; i += T) {   // expanded bond classes; [n_bc] = the zero class
        BondKm m; memset(&m, 0, sizeof m);
        if (i < n_bc) {
            const BondK<float> k = expand_bond<float>(((const BondKc<float>*)S.bk)[i], make_geo<float>(ctlp->bond_len));
            m.a2 = k.a2; m.b1 = k.b1; m.b2 = k.b2; m.b3 = k.b3;
            m.inv_len = k.inv_len; m.k_axial = k.k_axial; m.nom_dist = k.nom_dist; m.len_minus_dist = k.len_minus_dist;
            m.s1 = k.c_a1m1; m.c_a2i1 = k.c_a2i1; m.c_b2fm1 = k.c_b2fm1; m.c_b3i1 = k.c_b3i1;
        }
        bkm[i] = m;
    }
    // own static word (zero class for the missing bonds; neighbour indices become indices inside the CTA that holds the neighbour), and
    // the -axis neighbour tables: the owner of a bond enters itself into its +axis neighbour's row, in whichever CTA that row lives
    const bool active = rank ? vme < n_vox : tid < split;
    VoxStat st = ((const VoxStat*)(gblob + L.off_stat))[vme];
    if (!active) st.nbr[0] = st.nbr[1] = st.nbr[2] = -1;
    uint32_t fl = ((const uint32_t*)(gblob + L.off_flags))[vme] & VXG_C2_FL_STATE;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int w = st.nbr[a];
        bool remote = false;
        if (w < 0) st.bcls[a] = (uint16_t)ZC;
        else {
            const uint32_t wr = w >= split ? 1u : 0u;   // the CTA that holds the neighbour
            const int wl = w - (wr ? split : 0);        // ... and the neighbour's index there
            remote = wr != rank;
            st.nbr[a] = (int16_t)wl;
            cl_st_u16(cl_mapa(smem_u32(smem + G.off_nbm) + 8u * (uint32_t)wl + 2u * a, wr), (uint16_t)(tid | (remote ? 0x8000 : 0)));
            if (remote) fl |= 0x100u << a;
        }
        if (__any_sync(0xffffffffu, w >= 0)) fl |= 0x10000u << a;
        if (__any_sync(0xffffffffu, remote)) fl |= 0x100000u << a;
    }
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
    cl_sync();   // neighbour tables complete in both CTAs
    for (int i = tid; i < n_bc; i += T) { const uint32_t pr = S.bpair[i]; S.dnom_bc[i] = 0.5f * (S.dsc_c[pr & 0xFFFFu] + S.dsc_c[pr >> 16]); }

    RobotConsts<float> C;
    C.bk = nullptr; C.bkc = (const BondKc<float>*)S.bk;
    C.geo = make_geo<float>(1.0);
    C.bk_dist = S.bk_dist; C.vk = S.vk;
    C.nom = nullptr; C.ext_disp = S.ext_disp; C.ext_tdisp = S.ext_tdisp;
    C.half_bond_z = (float)ctlp->half_bond_z;
    const double dt = ctlp->dt;
    C.dt = (float)dt;
    C.fluid = false; C.drag_coeff = 0.0f; C.n_vert = 0; C.n_facet = 0;
    C.floor_on = (flags & VXG_F_FLOOR_) != 0;
    C.sticky = false; C.sym_mass = true;
    if (tid == 0) {
        ScaleSim SL;
        SL.init_cm_time = ctlp->init_cm_time; SL.temp_period = ctlp->temp_period; SL.temp_amplitude = ctlp->temp_amplitude;
        SL.min_fact = ctlp->min_fact; SL.max_fact = ctlp->max_fact; SL.growth_amp = ctlp->growth_amp;
        SL.midlife_freeze = ctlp->midlife_freeze; SL.stop_value = ctlp->stop_value;
        SL.temp_on = (flags & VXG_F_TEMPERATURE_) != 0; SL.lw = (flags & VXG_F_LW_SCALE_) != 0;
        *S.sl = SL;
    }
    OwnState<float> own;
    {
        const double* pos = (const double*)(smem + G.off_pos);
        own.px = pos[tid]; own.py = pos[HV + tid]; own.pz = pos[2 * HV + tid];
        own.q = ((const R4<float>*)(smem + G.off_quat))[tid];
    }
    uint2 nb = ((const uint2*)(smem + G.off_nbm))[tid];   // -X | -Y << 16, -Z; bit 15 of an entry: the owner is in the partner CTA
    {
        const uint32_t rm = ((nb.x >> 15) & 1u) | ((nb.x >> 30) & 2u) | ((nb.y >> 13) & 4u);
        fl |= rm << 12;
        if (__any_sync(0xffffffffu, rm != 0u)) fl |= 0x1000000u;
        nb.x &= 0x7FFF7FFFu; nb.y &= 0x7FFF7FFFu;
    }
    __syncthreads();
    const float dt_inv_run = (float)(1.0 / dt);
    const uint32_t red0 = cl_mapa(smem_u32(S.red), 0);   // the reduction scratch of CTA 0, where the latches of both halves meet

    double t = ctlp->cur_time;
    int cur = 0, done = 0;
    bool diverged = false;
    for (int k = 0; k < nsteps; ++k) {
        const long long s = step0 + k;
        if (s == latch) {   // IniCM = SS.CurCM at TimeStep entry (VX_Sim.cpp:1064-1075): COM after the previous step
            double mx = 0, my = 0, mz = 0, mm = 0;
            if (active) { const double m = (double)S.vk[st.vcls].mass; mx = own.px * m; my = own.py * m; mz = own.pz * m; mm = m; }
            mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz); mm = warp_sum(mm);
            cl_sync();
            if ((tid & 31) == 0) {   // per-warp sums of both halves into CTA 0, added up in warp order (as the 1024-thread kernel does)
                const uint32_t slot = red0 + 8u * (rank * 16u + (uint32_t)(tid >> 5));
                cl_st_f64(slot, mx); cl_st_f64(slot + 8u * 32u, my); cl_st_f64(slot + 8u * 64u, mz); cl_st_f64(slot + 8u * 96u, mm);
            }
            cl_sync();
            if (rank == 0 && tid == 0) {
                mx = my = mz = mm = 0;
                for (int i = 0; i < 32; i++) { mx += S.red[i]; my += S.red[32 + i]; mz += S.red[64 + i]; mm += S.red[96 + i]; }
                ctlp->ini_com[0] = mx / mm; ctlp->ini_com[1] = my / mm; ctlp->ini_com[2] = mz / mm; ctlp->ini_latched = 1;
            }
        }
        if (s == eol) {     // SS.EndOfLifetimePosteriorY = getPosteriorY() at TimeStep entry (VX_Sim.cpp:1078-1082)
            double lo = 1e300;
            if (active && S.sk[st.vcls].mat_index != 5) lo = own.py;
            lo = -warp_max(-lo);
            cl_sync();
            if ((tid & 31) == 0) cl_st_f64(red0 + 8u * (rank * 16u + (uint32_t)(tid >> 5)), lo);
            cl_sync();
            if (rank == 0 && tid == 0) { lo = S.red[0]; for (int i = 1; i < 32; i++) lo = fmin(lo, S.red[i]); ctlp->eol_post_y = lo; }
        }
        C.scale_c = S.scale_c + cur * mv;
        C.dsc_c = S.dsc_c + cur * mv;
        C.dnom_bc = S.dnom_bc + cur * MB;
        const float dt_inv = (s == 0) ? 0.0f : dt_inv_run;   // dt == 0 during the very first bond pass (VX_Sim.cpp:373)

        // Scale of every class for the NEXT step (VXS_Voxel.cpp:224-336), by the highest threads, into the other buffer half (each CTA its own copy)
        for (int i = T - 1 - tid; i < n_vc; i += T) {
            const double last = S.scale_c[cur * mv + i];
            const double ns = scale_law(S.sk[i], last, t, *S.sl);
            S.scale_c[(cur ^ 1) * mv + i] = ns;
            S.dsc_c[(cur ^ 1) * mv + i] = (float)(ns - S.sk[i].nom_size);
        }
        // ---- bond phase: reads positions / orientations of the previous step, writes only what this thread owns ------------------
        V3<float> F = mk3<float>(0, 0, 0), M = mk3<float>(0, 0, 0);
        bool div = bond_eval_c2<0>(smem, G, C, bkm, tid, rank, st, own, fl, dt_inv, tm, F, M);
        div |= bond_eval_c2<1>(smem, G, C, bkm, tid, rank, st, own, fl, dt_inv, tm, F, M);
        div |= bond_eval_c2<2>(smem, G, C, bkm, tid, rank, st, own, fl, dt_inv, tm, F, M);
        tm_wait_st();
        if (div) { const uint32_t a = smem_u32(divflag); cl_st_u32(cl_mapa(a, 0), 1u); cl_st_u32(cl_mapa(a, 1), 1u); }
        cl_step_sync();                                      // bonds -> voxels, both halves
        if (*(volatile unsigned int*)divflag) { diverged = true; break; }   // Integrate() returns before the voxel loop (VX_Sim.cpp:1777)

        // rest-length deviation of every bond class for the NEXT step
        for (int i = tid; i < n_bc; i += T) {
            const uint32_t pr = S.bpair[i];
            const float* dn = S.dsc_c + (cur ^ 1) * mv;
            S.dnom_bc[(cur ^ 1) * MB + i] = 0.5f * (dn[pr & 0xFFFFu] + dn[pr >> 16]);
        }
        // ---- voxel phase: gather the shares of the -X/-Y/-Z bonds (Force2 = -Force1, Moment2), integrate -------------------------
        if (active) voxel_phase_c2(smem, G, C, tid, rank, st, nb, own, fl, F, M);
        cl_step_sync();                                      // voxels -> bonds
        cur ^= 1;
        t += dt;                                             // CurTime += dt (VX_Sim.cpp:1929)
        done = k + 1;
    }

    // ---- stage out ---------------------------------------------------------------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        fence_async_smem();
        const size_t h8 = (size_t)vbase * 8, h16 = (size_t)vbase * 16;
        const unsigned int cnt = rank ? (unsigned int)HV : (unsigned int)split;   // CTA 0 returns its own share only (the rest of its slots never held live data)
        if (cnt > 0) {
            for (int p = 0; p < 3; p++) bulk_s2g(gblob + L.off_pos + (size_t)p * NV * 8 + h8, smem + G.off_pos + (size_t)p * HV * 8, cnt * 8);
            bulk_s2g(gblob + L.off_quat + h16, smem + G.off_quat, cnt * 16);
            bulk_s2g(gblob + L.off_momA + h16, smem + G.off_momA, cnt * 16);
            bulk_s2g(gblob + L.off_momB + h8, smem + G.off_momB, cnt * 8);
            bulk_s2g_finish();
        }
    }
    if (active) ((uint32_t*)(gblob + L.off_flags))[vme] = fl & VXG_C2_FL_STATE;
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
    if (rank == 0) for (int i = tid; i < n_vc; i += T) B.scale_c[(size_t)robot * B.max_vclass + i] = S.scale_c[cur * mv + i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cl_sync();   // (no CTA leaves while its partner could still read its shared memory; every read of the control block is behind us)
    if (rank == 0 && tid == 0) {
        ctlp->cur_time = t;
        ctlp->step = step0 + done;
        unsigned int st2 = status0;
        if (diverged) st2 |= VXG_ST_DIVERGED_;
        if (step0 + done >= total) st2 |= VXG_ST_FINISHED_;
        ctlp->status = st2;
    }
    // warp 0's own address IS the allocation's base (lane quarter 0, column block 0)
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(TM_COLS) : "memory");
}

inline size_t cluster_smem_bytes(int mv, int mb) { return make_clayout().total + sizeof(BondKm) * (size_t)mb + tables_bytes<float>(mv, mb, true, 0); }
template <int CV, int CB> inline cudaError_t launch_cluster_form(const BatchDev<float>& B, const LaunchCfg& c) {
    auto kern = vx_step_kernel_c2<CV, CB>;
    const size_t smem = cluster_smem_bytes(CV > 0 ? CV : B.max_vclass, CB > 0 ? CB : B.max_bclass + 1);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<2 * c.n_robots, VXG_C2_HV, smem, c.stream>>>(B, c.n_steps);
    return cudaGetLastError();
}
// 0 = launched; 1 = the class tables do not fit next to a second CTA on the SM (the caller runs the single-CTA form)
inline int launch_cluster(const BatchDev<float>& B, const LaunchCfg& c, cudaError_t* err) {
    if (B.max_vclass <= VXG_CV && B.max_bclass + 1 <= VXG_CB) { *err = launch_cluster_form<VXG_CV, VXG_CB>(B, c); return 0; }
    if (cluster_smem_bytes(B.max_vclass, B.max_bclass + 1) <= 113u * 1024u) { *err = launch_cluster_form<0, 0>(B, c); return 0; }
    return 1;
}

}  // namespace vx
