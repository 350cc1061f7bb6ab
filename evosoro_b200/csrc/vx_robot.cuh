// vx_robot.cuh -- per-robot state layout ("blob") and the per-voxel tasks of one timestep.
//
// One robot slot = one contiguous, fixed-stride blob in HBM so that a CTA can bulk-copy it into shared
// memory, advance it K steps there and bulk-copy it back (DESIGN.md "Data layout").  The same task
// functions run against a view of global memory when the robot does not fit on chip.
//
// Inside the blob every array has stride `nv` (a power of two >= 64 chosen at vxg_create, a compile-time
// constant in the kernels) and the words a thread needs together are grouped into 128-bit / 64-bit
// vectors, so one LDS.128 / LDG.128 replaces four scalar loads and consecutive threads touch consecutive
// vectors (bank-conflict free, fully coalesced).
//
// Task decomposition (replaces the serial bond loop + voxel loop of CVX_Sim::Integrate,
// VX_Sim.cpp:1773-1776 and :1913): voxel v OWNS its +X, +Y, +Z bonds.  In the bond phase the owner
// evaluates the bond, keeps its own share (Force1, Moment1) in registers and adds the neighbour's
// share (Force2, Moment2) to the neighbour's accumulator in shared memory; the three axes run as
// three sub-phases separated by barriers, so every accumulator word has exactly one writer per
// sub-phase -- no atomics, deterministic summation order.  In the voxel phase the owner integrates.
#pragma once
#include <stdlib.h>
#include "vx_math.cuh"

namespace vx {

// flags word per voxel
#define VX_FL_SMALL_X 0x1u     // SmallAngle of the owned +X bond (VXS_BondInternal.h: SmallAngle)
#define VX_FL_SMALL_Y 0x2u
#define VX_FL_SMALL_Z 0x4u
#define VX_FL_SMALL_ALL 0x7u
#define VX_FL_STATIC_FRIC 0x8u // StaticFricFlag of the last step (VXS_Voxel.h:129)

template <typename T> struct alignas(4 * sizeof(T)) R4 { T x, y, z, w; };
template <typename T> struct alignas(2 * sizeof(T)) R2 { T x, y; };

// static description of one voxel: one 128-bit word
struct alignas(16) VoxStat {
    int16_t nbr[3];    // index of the +X/+Y/+Z neighbour, -1 = no bond
    uint16_t vcls;     // voxel class
    uint16_t bcls[3];  // class of the owned +axis bond
    uint16_t faces;    // bit d set: no voxel at the lattice neighbour in direction d (+X,-X,+Y,-Y,+Z,-Z): an exposed face
};
// surface-mesh tables of the land/water fork (LW/VX_MeshUtil.cpp:110-276)
struct alignas(16) Corner8 { int16_t i[8]; };   // per mesh vertex: the voxels that share it, compacted: entry = (code << 12) | voxel with
                                                // code = 4*X+ + 2*Y+ + Z+ of the corner the vertex is for that voxel, ascending code, -1 ends
struct alignas(8) Facet { int16_t a, b, c, v; };   // one triangle of an exposed face: its three mesh vertices and the voxel it belongs to

struct BlobLayout {            // byte offsets inside one robot blob; all multiples of 32
    int nv;                    // voxel stride
    int real_size;
    // dynamic: copied in and out
    size_t off_pos;            // double [3][nv]
    size_t off_quat;           // R4 [nv]                     Angle (w,x,y,z) stored as (x=w, y=x, z=y, w=z)
    size_t off_momA;           // R4 [nv]  LinMom.xyz, AngMom.x
    size_t off_momB;           // R2 [nv]  AngMom.y, AngMom.z
    size_t off_hist;           // per axis: R4 A[nv] (p2.xyz, a1.x), R4 B[nv] (a1.y, a1.z, a2.x, a2.y), Real C[nv] (a2.z)
    size_t hist_axis_bytes;
    size_t off_flags;          // uint32 [nv]
    size_t off_strain;         // fluid only: Real [6][nv]  StrainPosDirsCur.xyz, StrainNegDirsCur.xyz of the last bond pass
    size_t dyn_bytes;
    // static: copied in
    size_t off_stat;           // VoxStat [nv]
    size_t off_fbase;          // fluid only: uint16 [nv]           first facet of each voxel (its facets are consecutive: face ascending, 2 each)
    size_t off_facet;          // fluid only: Facet [nfacet_cap]    the triangles of all exposed faces, ordered by voxel
    size_t off_vcon;           // fluid only: Corner8 [nvert_cap]   mesh vertex -> its <= 8 voxels
    size_t stat_bytes;         // dynamic + static
    // scratch: zero-initialised on chip, never written back
    size_t off_accA;           // R4 [nv]  Force.xyz, Moment.x received from the -axis neighbours
    size_t off_accB;           // R2 [nv]  Moment.y, Moment.z
    size_t off_vpos;           // fluid only: double [3][nvert_cap]  deformed mesh vertex positions of this step
    size_t off_fdrag;          // fluid only: Real [3][nfacet_cap]   drag of every facet of this step
    size_t total_bytes;
    int nvert_cap;             // mesh vertex capacity (0 = no fluid-drag support in this batch)
    int nfacet_cap;            // facet capacity, derived: a closed quad surface has V - 2 + 2g faces, i.e. 2V - 4 + 4g triangles
};

VX_HD constexpr size_t vx_align16(size_t x) { return (x + 15) & ~(size_t)15; }
VX_HD constexpr size_t vx_align32(size_t x) { return (x + 31) & ~(size_t)31; }

// voxel stride used for a robot of up to max_vox voxels: the kernels are instantiated for these strides
// (every multiple of a warp up to 256 but the last one -- CPPN morphologies in evosoro's default 6^3 workspace have 65 .. 216 voxels,
// and a CTA's idle warps cost the SM resident robots --, then 352 / 736 so that 7^3 and 9^3 robots do not pay for 512 / 1024 slots;
// above 2048 voxels two strides, 8192 and 32768, on the blob in global memory)
inline int stride_for(int max_vox) {
    static const int strides[] = {64, 96, 128, 160, 192, 224, 256, 352, 512, 736, 1024, 2048};
    if (const char* e = getenv("VXG_STRIDES")) {   // tuning: restrict the strides in use, e.g. "128,224,512,1024" (must be instantiated ones)
        for (const char* p = e; *p;) { const int v = atoi(p); if (v >= max_vox) { for (int s : strides) if (s == v) return v; } while (*p && *p != ',') p++; if (*p) p++; }
    }
    for (int s : strides) if (max_vox <= s) return s;
    return max_vox <= 8192 ? 8192 : 32768;   // global-memory path, 16 / 64 voxels per thread (voxel indices are 16-bit: 32767 voxels at most)
}

VX_HD constexpr BlobLayout make_layout(int nv_in, int real_size, int nvert_cap = 0) {
    BlobLayout L{};
    L.nv = nv_in;
    L.real_size = real_size;
    L.nvert_cap = nvert_cap;
    L.nfacet_cap = nvert_cap ? ((2 * nvert_cap + 64 + 31) & ~31) : 0;
    const size_t nvc = (size_t)nvert_cap, nfc = (size_t)L.nfacet_cap;
    size_t nv = (size_t)nv_in, rs = (size_t)real_size, o = 0;
    L.off_pos = o;   o = vx_align32(o + 3 * nv * sizeof(double));
    L.off_quat = o;  o = vx_align32(o + 4 * nv * rs);
    L.off_momA = o;  o = vx_align32(o + 4 * nv * rs);
    L.off_momB = o;  o = vx_align32(o + 2 * nv * rs);
    L.hist_axis_bytes = vx_align32(9 * nv * rs);
    L.off_hist = o;  o = o + 3 * L.hist_axis_bytes;
    L.off_flags = o; o = vx_align32(o + nv * sizeof(uint32_t));
    L.off_strain = o; if (nvc) o = vx_align32(o + 6 * nv * rs);
    L.dyn_bytes = o;
    L.off_stat = o;  o = vx_align32(o + nv * sizeof(VoxStat));
    L.off_fbase = o; if (nvc) o = vx_align32(o + nv * sizeof(uint16_t));
    L.off_facet = o; if (nvc) o = vx_align32(o + nfc * sizeof(Facet));
    L.off_vcon = o;  if (nvc) o = vx_align32(o + nvc * sizeof(Corner8));
    L.stat_bytes = o;
    L.off_accA = o;  o = vx_align32(o + 4 * nv * rs);
    L.off_accB = o;  o = vx_align32(o + 2 * nv * rs);
    L.off_vpos = o;  if (nvc) o = vx_align32(o + 3 * nvc * sizeof(double));
    L.off_fdrag = o; if (nvc) o = vx_align32(o + 3 * nfc * rs);
    L.total_bytes = o;
    return L;
}

template <typename Real> struct RobotView {
    int n_vox, nv;
    double* pos;          // [3][nv]
    R4<Real>* quat;       // [nv]
    R4<Real>* momA;       // [nv]
    R2<Real>* momB;       // [nv]
    unsigned char* hist;  // 3 axis blocks
    size_t hist_axis_bytes;
    uint32_t* flags;      // [nv]
    VoxStat* stat;        // [nv]
    R4<Real>* accA;       // [nv]
    R2<Real>* accB;       // [nv]
    Real* strain;         // [6][nv]          (fluid layouts only)
    uint16_t* fbase;      // [nv]
    Facet* facet;         // [nfacet_cap]
    Corner8* vcon;        // [nvert_cap]
    double* vpos;         // [3][nvert_cap]
    Real* fdrag;          // [3][nfacet_cap]
    int nvert_cap, nfacet_cap;
    VX_HD R4<Real>* histA(int ax) const { return (R4<Real>*)(hist + (size_t)ax * hist_axis_bytes); }
    VX_HD R4<Real>* histB(int ax) const { return (R4<Real>*)(hist + (size_t)ax * hist_axis_bytes) + nv; }
    VX_HD Real* histC(int ax) const { return (Real*)((R4<Real>*)(hist + (size_t)ax * hist_axis_bytes) + 2 * (size_t)nv); }
};

template <typename Real>
VX_HD RobotView<Real> make_view(unsigned char* base, const BlobLayout& L, int n_vox) {
    RobotView<Real> v;
    v.n_vox = n_vox; v.nv = L.nv;
    v.pos = (double*)(base + L.off_pos);
    v.quat = (R4<Real>*)(base + L.off_quat);
    v.momA = (R4<Real>*)(base + L.off_momA);
    v.momB = (R2<Real>*)(base + L.off_momB);
    v.hist = base + L.off_hist;
    v.hist_axis_bytes = L.hist_axis_bytes;
    v.flags = (uint32_t*)(base + L.off_flags);
    v.stat = (VoxStat*)(base + L.off_stat);
    v.accA = (R4<Real>*)(base + L.off_accA);
    v.accB = (R2<Real>*)(base + L.off_accB);
    v.strain = (Real*)(base + L.off_strain);
    v.fbase = (uint16_t*)(base + L.off_fbase);
    v.facet = (Facet*)(base + L.off_facet);
    v.vcon = (Corner8*)(base + L.off_vcon);
    v.vpos = (double*)(base + L.off_vpos);
    v.fdrag = (Real*)(base + L.off_fdrag);
    v.nvert_cap = L.nvert_cap; v.nfacet_cap = L.nfacet_cap;
    return v;
}

// Per-robot constants the tasks read (class tables live in shared memory in the kernel).
template <typename Real> struct RobotConsts {
    const BondK<Real>* bk;      // [n_bclass] wide bond classes (nullptr in compact mode)
    const BondKc<Real>* bkc;    // [n_bclass] compact bond classes + the shared geometry
    BondGeo<Real> geo;
    const double* bk_dist;      // [n_bclass] pitch along the bond axis, fp64
    const VoxK<Real>* vk;       // [n_vclass]
    const double* scale_c;      // [n_vclass] Scale of each class as left by the previous step
    const Real* dsc_c;          // [n_vclass] Scale - NominalSize, narrowed
    const Real* dnom_bc;        // [n_bclass] 0.5 (dsc_c[class of voxel 1] + dsc_c[class of voxel 2]) of the re-keyed bond classes (compact family)
    const double* nom;          // [3][nv] NominalPosition (global; read for fixed DOFs only)
    const Real* ext_disp;       // [n_vclass][3]
    const Real* ext_tdisp;      // [n_vclass][3]
    Real half_bond_z;           // 0.5 BondDampingZ
    Real dt;
    bool floor_on, sticky;
    bool sym_mass;              // every bond class has equal end-voxel masses: Force2 = -Force1
    bool fluid;                 // land/water fluid environment: surface drag, no gravity / floor (LW/VX_Sim.cpp:1516-1597)
    Real drag_coeff;            // AggregateDragCoefficient
    int n_vert, n_facet;        // mesh vertices / facets of this robot
};

// The owner's own position and orientation, loaded once per step and reused by its three bonds.
template <typename Real> struct OwnState { double px, py, pz; R4<Real> q; };
template <typename Real> VX_HD OwnState<Real> load_own(const RobotView<Real>& R, int v) {
    OwnState<Real> o;
    o.px = R.pos[v]; o.py = R.pos[R.nv + v]; o.pz = R.pos[2 * R.nv + v]; o.q = R.quat[v];
    return o;
}
// The neighbour's share of one bond, waiting to be added to the neighbour's accumulator.
template <typename Real> struct BondPush { int w; V3<Real> f2, m2; };   // w < 0: nothing to push

// Bond (v, AX): evaluate; the owner's share goes into (F, M) +=, the neighbour's share into `p`.
// `st` is the owner's static word, `fl` its flags word, `own` its position/orientation (all kept in
// registers across the three axes by the caller).  rec != nullptr: also store the bond outputs
// ([39][nv] planes, vxg_download_state).  Touches no accumulator, so it needs no ordering against the
// other voxels' pushes.  Returns true when the bond diverged (strain > 100, VX_Sim.cpp:1775).
template <typename Real, int AX, bool COMPACT = false, bool FLUID = false, bool HOIST = true>
VX_HD bool bond_eval(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, const VoxStat& st, const OwnState<Real>& own,
                     uint32_t& fl, Real dt_inv, V3<Real>& F, V3<Real>& M, Real* rec, BondPush<Real>& p) {
    const int nv = R.nv;
    const int w = st.nbr[AX];
    p.w = w;
    if (w < 0) return false;
    const int bc = st.bcls[AX];
    BondK<Real> k;
    if (COMPACT) { const BondKc<Real> kc = C.bkc[bc]; k = expand_bond<Real>(kc, C.geo); }
    else k = C.bk[bc];
    // bond vector in fp64, pitch removed before narrowing (VXS_BondInternal.cpp:58)
    double dpa[3] = {R.pos[w] - own.px, R.pos[nv + w] - own.py, R.pos[2 * nv + w] - own.pz};
    dpa[AX] -= C.bk_dist[bc];
    V3<Real> e = mk3<Real>((Real)dpa[0], (Real)dpa[1], (Real)dpa[2]);
    const R4<Real> qa = own.q, qb = R.quat[w];
    Q4<Real> q1, q2;
    q1.w = qa.x; q1.x = qa.y; q1.y = qa.z; q1.z = qa.w;
    q2.w = qb.x; q2.x = qb.y; q2.y = qb.z; q2.z = qb.w;
    Real dnom;   // NomDistance - L (VXS_BondInternal.cpp:71)
    if (COMPACT) dnom = C.dnom_bc[bc];
    else dnom = (Real)0.5 * (C.dsc_c[st.vcls] + C.dsc_c[R.stat[w].vcls]);
    bool small = (fl >> AX) & 1u;
    R4<Real>* hA = R.histA(AX) + v; R4<Real>* hB = R.histB(AX) + v; Real* hC = R.histC(AX) + v;
    BondHist<Real> h;
    {
        const R4<Real> a = *hA, b = *hB; const Real c = *hC;
        h.p2 = mk3<Real>(a.x, a.y, a.z);
        h.a1 = mk3<Real>(a.w, b.x, b.y);
        h.a2 = mk3<Real>(b.z, b.w, c);
    }
    BondOut<Real> o;
    bond_force<AX, Real, COMPACT, HOIST>(k, e, q1, q2, dnom, C.half_bond_z, dt_inv, C.sym_mass, small, h, o);
    if (!o.changed) {   // _LastPos2/_LastAngle* keep their values on a mode-flip step (VXS_BondInternal.cpp:157)
        R4<Real> a, b;
        a.x = h.p2.x; a.y = h.p2.y; a.z = h.p2.z; a.w = h.a1.x;
        b.x = h.a1.y; b.y = h.a1.z; b.z = h.a2.x; b.w = h.a2.y;
        *hA = a; *hB = b; *hC = h.a2.z;
    }
    fl = (fl & ~(1u << AX)) | ((small ? 1u : 0u) << AX);
    F = F + o.f1; M = M + o.m1;
    p.f2 = o.f2; p.m2 = o.m2;
    if (FLUID) {   // SetStrainDir(BD_P<axis>, CurStrainV1) on voxel 1, (BD_N<axis>, CurStrainV2) on voxel 2: one writer per slot
        R.strain[AX * nv + v] = o.strain1;
        R.strain[(3 + AX) * nv + w] = o.strain2;
    }
    if (rec) {
        Real* r = rec + (size_t)(AX * 13) * nv + v;
        r[0] = o.f1.x; r[nv] = o.f1.y; r[2 * nv] = o.f1.z; r[3 * nv] = o.f2.x; r[4 * nv] = o.f2.y; r[5 * nv] = o.f2.z;
        r[6 * nv] = o.m1.x; r[7 * nv] = o.m1.y; r[8 * nv] = o.m1.z; r[9 * nv] = o.m2.x; r[10 * nv] = o.m2.y; r[11 * nv] = o.m2.z;
        r[12 * nv] = o.strain;
    }
    return o.strain > (Real)100;
}

// Add the neighbour's share to its accumulator.  Within one axis sub-phase every accumulator has exactly
// one writer; sub-phases of different axes must be ordered (barrier) by the caller.
template <typename Real>
VX_HD void bond_push(const RobotView<Real>& R, const BondPush<Real>& p) {
    if (p.w < 0) return;
    R4<Real> a = R.accA[p.w]; R2<Real> b = R.accB[p.w];
    a.x += p.f2.x; a.y += p.f2.y; a.z += p.f2.z; a.w += p.m2.x;
    b.x += p.m2.y; b.y += p.m2.z;
    R.accA[p.w] = a; R.accB[p.w] = b;
}

template <typename Real, int AX, bool COMPACT = false, bool FLUID = false>
VX_HD bool bond_task(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, const VoxStat& st, uint32_t& fl,
                     Real dt_inv, V3<Real>& F, V3<Real>& M, Real* rec) {
    BondPush<Real> p;
    bool d = bond_eval<Real, AX, COMPACT, FLUID>(R, C, v, st, load_own<Real>(R, v), fl, dt_inv, F, M, rec, p);
    bond_push<Real>(R, p);
    return d;
}

// ---- land/water fluid drag (LW/VX_Sim.cpp:1516-1597) ------------------------------------------------
// Deformed position of mesh vertex u: mean over the voxels that share it of Pos + Angle * corner offset,
// corner extents (1 + strain of the bond in that direction) * NominalSize / 2 from the PREVIOUS bond pass
// (CVX_MeshUtil::GetCurVLoc LW/VX_MeshUtil.cpp:388-428, LW/VXS_Voxel.cpp:474-475).
template <typename Real>
VX_HD void mesh_vertex(const RobotView<Real>& R, const RobotConsts<Real>& C, int u) {
    const int nv = R.nv;
    const Corner8 con = R.vcon[u];
    double sx = 0, sy = 0, sz = 0; int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {   // compacted entries: a warp runs as many rounds as its busiest vertex needs (4 on a flat face)
        const int e = con.i[k];
        if (e < 0) break;
        const int vv = e & 0xFFF, c = (e >> 12) & 7;
        const Real h = C.vk[R.stat[vv].vcls].half_nom;
        const Real ox = (c & 4) ? ((Real)1 + R.strain[vv]) * h : -((Real)1 + R.strain[3 * nv + vv]) * h;
        const Real oy = (c & 2) ? ((Real)1 + R.strain[nv + vv]) * h : -((Real)1 + R.strain[4 * nv + vv]) * h;
        const Real oz = (c & 1) ? ((Real)1 + R.strain[2 * nv + vv]) * h : -((Real)1 + R.strain[5 * nv + vv]) * h;
        const R4<Real> q = R.quat[vv];
        const V3<Real> r = rotate<Real>(q.x, mk3<Real>(q.y, q.z, q.w), mk3<Real>(ox, oy, oz));
        sx += R.pos[vv] + (double)r.x; sy += R.pos[nv + vv] + (double)r.y; sz += R.pos[2 * nv + vv] + (double)r.z;
        cnt++;
    }
    const double inv = 1.0 / (double)cnt;
    R.vpos[u] = sx * inv; R.vpos[R.nvert_cap + u] = sy * inv; R.vpos[2 * R.nvert_cap + u] = sz * inv;
}

// drag of one triangular facet (A, B, C) on its voxel moving with `vel`: -C_d * area * (v.n)^2 n when the facet
// faces the flow (acos(v^.n^) < pi/2, LW/VX_Sim.cpp:1559-1577), n = (B-A)x(C-A) normalised (CMesh::CalcFaceNormals)
template <typename Real>
VX_HD V3<Real> facet_drag(const RobotView<Real>& R, int a, int b, int c, V3<Real> vel, Real cd) {
    const int nc = R.nvert_cap;
    const double ax = R.vpos[a], ay = R.vpos[nc + a], az = R.vpos[2 * nc + a];
    const V3<Real> ab = mk3<Real>((Real)(R.vpos[b] - ax), (Real)(R.vpos[nc + b] - ay), (Real)(R.vpos[2 * nc + b] - az));
    const V3<Real> ac = mk3<Real>((Real)(R.vpos[c] - ax), (Real)(R.vpos[nc + c] - ay), (Real)(R.vpos[2 * nc + c] - az));
    const V3<Real> cr = cross(ab, ac);
    const Real l2 = dot(cr, cr);
    if (!(l2 > (Real)0)) return mk3<Real>(0, 0, 0);
    const Real il = vx_rsqrt(l2);
    const V3<Real> n = cr * il;
    const Real vn = dot(vel, n);
    if (!(vn > (Real)0)) return mk3<Real>(0, 0, 0);
    const Real area = (Real)0.5 * l2 * il;
    return n * (-cd * area * vn * vn);
}

// Facet i of the robot (one thread per facet, so the lanes of a warp are all busy whatever the robot's shape): its drag on
// the voxel it belongs to, from that voxel's velocity of the previous step.
template <typename Real>
VX_HD void facet_task(const RobotView<Real>& R, const RobotConsts<Real>& C, int i) {
    const Facet f = R.facet[i];
    const R4<Real> lm = R.momA[f.v];
    const Real mi = C.vk[R.stat[f.v].vcls].mass_inv;
    const V3<Real> d = facet_drag<Real>(R, f.a, f.b, f.c, mk3<Real>(lm.x * mi, lm.y * mi, lm.z * mi), C.drag_coeff);
    R.fdrag[i] = d.x; R.fdrag[R.nfacet_cap + i] = d.y; R.fdrag[2 * R.nfacet_cap + i] = d.z;
}

// DragForce of voxel v: the sum over its facets in the reference's order (faces +X,-X,+Y,-Y,+Z,-Z, two triangles each,
// vertex order as in LinkSimVoxels, LW/VX_MeshUtil.cpp:161-190; FacetToSIndex = the voxel the face belongs to)
template <typename Real>
VX_HD V3<Real> voxel_drag(const RobotView<Real>& R, int v, const VoxStat& st) {
    V3<Real> d = mk3<Real>(0, 0, 0);
    const unsigned faces = st.faces & 63u;
    if (!faces) return d;
#if defined(__CUDA_ARCH__)
    const int n = 2 * __popc(faces);
#else
    const int n = 2 * __builtin_popcount(faces);
#endif
    const int f0 = R.fbase[v], nc = R.nfacet_cap;
    for (int k = 0; k < n; k++) d = d + mk3<Real>(R.fdrag[f0 + k], R.fdrag[nc + f0 + k], R.fdrag[2 * nc + f0 + k]);
    return d;
}

// ---- self-collision (VX_Sim.cpp:1729-1755, :2357-2413; VXS_BondCollision.cpp:41-59) ------------------
struct ColPair { int16_t v1, v2; };
struct ColTables {              // per-robot static tables in global memory, read by the list rebuild only
    const int16_t* surf;        // [n_surf] surface voxels, ascending
    const int8_t* lat4;         // [nv][4] lattice coordinates
    const unsigned long long* near;   // [nv] mask over lattice offsets: within col_hops bond hops (CalcNearby)
    int n_surf, hops;
};
VX_HD int near_code(int dx, int dy, int dz) {   // rank of (dx,dy,dz) among the offsets with |dx|+|dy|+|dz| <= 3, z slowest, then y, then x
    // closed form of the enumeration: a z-layer holds the (x, y) with |x|+|y| <= r = 3-|z|, i.e. 2r^2+2r+1 offsets, a y-row of it 2(r-|y|)+1
    const int az = dz < 0 ? -dz : dz, ay = dy < 0 ? -dy : dy;
    int code = 0;
    for (int z = -3; z < dz; z++) { const int r = 3 - (z < 0 ? -z : z); code += 2 * r * r + 2 * r + 1; }
    const int r = 3 - az;
    for (int y = -r; y < dy; y++) code += 2 * (r - (y < 0 ? -y : y)) + 1;
    return code + dx + (r - ay);
}
// CalcL1Bonds' pair test (VX_Sim.cpp:2376-2386): inside the filter radius, not a lattice-near neighbour, and
// closer than CollisionHorizon * Scale_1 (the reference uses voxel 1's scale twice)
template <typename Real>
VX_HD bool col_pair_hit(const RobotView<Real>& R, const RobotConsts<Real>& C, const ColTables& ct, int vi, int vj, Real horizon, Real filt2) {
    const int nv = R.nv;
    const Real dx = (Real)(R.pos[vi] - R.pos[vj]), dy = (Real)(R.pos[nv + vi] - R.pos[nv + vj]), dz = (Real)(R.pos[2 * nv + vi] - R.pos[2 * nv + vj]);
    const Real d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 < filt2)) return false;
    const int ox = ct.lat4[4 * vj] - ct.lat4[4 * vi], oy = ct.lat4[4 * vj + 1] - ct.lat4[4 * vi + 1], oz = ct.lat4[4 * vj + 2] - ct.lat4[4 * vi + 2];
    const int l1 = (ox < 0 ? -ox : ox) + (oy < 0 ? -oy : oy) + (oz < 0 ? -oz : oz);
    if (l1 <= 3 && ((ct.near[vi] >> near_code(ox, oy, oz)) & 1ULL)) return false;   // IsNearbyVox
    const Real act = horizon * (Real)C.scale_c[R.stat[vi].vcls];
    return d2 < act * act;
}
// CVXS_BondCollision::CalcContactForce, the share of voxel v of its contact bond with voxel o: both ends of a pair are pushed
// apart along the line of centres with a1 * overlap, a1 = composite E * mean size (CVX_Bond::LinkVoxels / UpdateConstants)
template <typename Real>
VX_HD void col_contact(const RobotView<Real>& R, const RobotConsts<Real>& C, const VoxStat& st, const OwnState<Real>& own, int o, V3<Real>& f) {
    const int nv = R.nv;
    const Real dx = (Real)(R.pos[o] - own.px), dy = (Real)(R.pos[nv + o] - own.py), dz = (Real)(R.pos[2 * nv + o] - own.pz);
    const int vco = R.stat[o].vcls;
    const Real nomd = (Real)((C.scale_c[st.vcls] + C.scale_c[vco]) * 0.75);
    const Real len = vx_sqrt(dx * dx + dy * dy + dz * dz);
    const Real rel = nomd - len;
    if (rel > (Real)0) {
        const VoxK<Real>& k1 = C.vk[st.vcls]; const VoxK<Real>& k2 = C.vk[vco];
        const Real a1 = ((k1.e_mod * k2.e_mod / (k1.e_mod + k2.e_mod)) * (Real)2) * (k1.half_nom + k2.half_nom);
        const Real g = a1 * rel / len;
        f.x -= dx * g; f.y -= dy * g; f.z -= dz * g;
    }
}
// ... summed over the contact-bond list in its order (every voxel scans the list: the push-form kernels and the CPU emulation)
template <typename Real>
VX_HD V3<Real> col_force_on(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, const VoxStat& st, const OwnState<Real>& own,
                            const ColPair* pairs, int n_col) {
    V3<Real> f = mk3<Real>(0, 0, 0);
    for (int k = 0; k < n_col; k++) {
        const ColPair p = pairs[k];
        if (p.v1 != v && p.v2 != v) continue;
        col_contact<Real>(R, C, st, own, p.v1 == v ? p.v2 : p.v1, f);
    }
    return f;
}

// Voxel v: total force/moment = own share + received share; integrate (EulerStep).  Returns |Vel|^2 after the step.
template <typename Real, bool BC = true>
VX_HD Real voxel_task(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, const VoxStat& st, const OwnState<Real>& own,
                      uint32_t fl, V3<Real> F, V3<Real> M) {
    const int nv = R.nv;
    {
        const R4<Real> a = R.accA[v]; const R2<Real> b = R.accB[v];
        F = F + mk3<Real>(a.x, a.y, a.z);
        M = M + mk3<Real>(a.w, b.x, b.y);
        R4<Real> za; za.x = za.y = za.z = za.w = (Real)0; R2<Real> zb; zb.x = zb.y = (Real)0;
        R.accA[v] = za; R.accB[v] = zb;
    }
    const int vc = st.vcls;
    const VoxK<Real> k = C.vk[vc];
    VoxState<Real> s;
    s.px = own.px; s.py = own.py; s.pz = own.pz;
    {
        const R4<Real> q = own.q, a = R.momA[v]; const R2<Real> b = R.momB[v];
        s.q.w = q.x; s.q.x = q.y; s.q.y = q.z; s.q.z = q.w;
        s.lm = mk3<Real>(a.x, a.y, a.z);
        s.am = mk3<Real>(a.w, b.x, b.y);
    }
    double nom[3] = {0, 0, 0};
    if (BC && k.has_bc) { nom[0] = C.nom[v]; nom[1] = C.nom[nv + v]; nom[2] = C.nom[2 * nv + v]; }
    bool sf = voxel_integrate<Real, BC>(k, F, M, 0.5 * C.scale_c[vc], C.floor_on, BC && C.sticky, C.dt, nom,
                                    C.ext_disp + 3 * vc, C.ext_tdisp + 3 * vc, s);
    R.pos[v] = s.px; R.pos[nv + v] = s.py; R.pos[2 * nv + v] = s.pz;
    {
        R4<Real> q, a; R2<Real> b;
        q.x = s.q.w; q.y = s.q.x; q.z = s.q.y; q.w = s.q.z;
        a.x = s.lm.x; a.y = s.lm.y; a.z = s.lm.z; a.w = s.am.x;
        b.x = s.am.y; b.y = s.am.z;
        R.quat[v] = q; R.momA[v] = a; R.momB[v] = b;
    }
    R.flags[v] = sf ? (fl | VX_FL_STATIC_FRIC) : (fl & ~VX_FL_STATIC_FRIC);
    const V3<Real> vel = s.lm * k.mass_inv;
    return dot(vel, vel);
}

// Per-robot control block (global memory)
struct RobotCtl {
    double cur_time;        // CurTime (VX_Sim.cpp:1929)
    double dt;              // DtFrac * OptimalDt
    double init_cm_time, temp_period, temp_amplitude, min_fact, max_fact;
    double ini_com[3];      // IniCM (VX_Sim.cpp:1066)
    long long step;         // CurStepCount
    long long total_steps;  // stop after this many (LLONG_MAX = no limit)
    long long latch_step;   // CurStepCount at which IniCM is latched (-1 never)
    unsigned int status;    // VXG_ST_*
    unsigned int flags;     // VXG_F_*
    int n_vox, n_vclass, n_bclass, ini_latched;
    int compact_ok;         // this robot's bond classes have the compact form (pack_compact)
    int sym_mass;           // every bond class has identical voxel-1 / voxel-2 damping constants
    double half_bond_z;     // 0.5 BondDampingZ
    double bond_len;        // the common bond length L of the compact form
    double drag_coeff;      // AggregateDragCoefficient (fluid)
    int n_vert, n_facet;    // mesh vertices / facets (fluid)
    // self-collision (VXG_F_COLLISIONS): CVX_Sim::UpdateCollisions state, VX_Sim.cpp:1729-1755
    double col_horizon, lattice_dim;
    double col_acc;         // MaxDispSinceLastBondUpdate
    double max_vel;         // SS.MaxVoxVel of the last step
    int n_col;              // contact bonds in the list
    int col_first;          // ColEnableChanged: rebuild at the next step
    int col_host;           // the host manages the list (vxg_set_collision_pairs): no device rebuild
    int n_surf, col_hops;
    int c2_split;           // two-CTA cluster form (vxg_cluster.cuh): CTA 0 owns voxels [0, c2_split), CTA 1 the rest (even; <= 512 on either side)
    long long eol_step;     // CurStepCount at which EndOfLifetimePosteriorY is latched (-1 never)
    double eol_post_y;      // min Pos.y over material != 5 at that step (metres); 0 = not latched
    double growth_amp, midlife_freeze, stop_value;   // development terms of the scale law (ScaleSim)
};

}  // namespace vx
