// vx_robot.cuh -- per-robot state layout ("blob") and the per-voxel tasks of one timestep.
//
// One robot slot = one contiguous, SoA, fixed-stride blob in HBM so that a CTA can bulk-copy it into
// shared memory, advance it K steps there and bulk-copy it back (DESIGN.md "Data layout").  The same
// task functions run against a view of global memory when the robot does not fit on chip.
//
// Task decomposition (replaces the serial bond loop + voxel loop of CVX_Sim::Integrate,
// VX_Sim.cpp:1773-1776 and :1913): voxel v OWNS its +X, +Y, +Z bonds.  In the bond phase the owner
// evaluates the bond, keeps its own share (Force1, Moment1) in registers and adds the neighbour's
// share (Force2, Moment2) to the neighbour's accumulator in shared memory; the three axes run as
// three sub-phases separated by barriers, so every accumulator word has exactly one writer per
// sub-phase -- no atomics, deterministic summation order.  In the voxel phase the owner integrates.
#pragma once
#include "vx_math.cuh"

namespace vx {

// flags word per voxel
#define VX_FL_SMALL_X 0x1u     // SmallAngle of the owned +X bond (VXS_BondInternal.h: SmallAngle)
#define VX_FL_SMALL_Y 0x2u
#define VX_FL_SMALL_Z 0x4u
#define VX_FL_SMALL_ALL 0x7u
#define VX_FL_STATIC_FRIC 0x8u // StaticFricFlag of the last step (VXS_Voxel.h:129)

// number of Real planes in the dynamic section after the fp64 positions
#define VX_N_HIST 27           // 3 owned bonds x (_LastPos2, _LastAngle1, _LastAngle2)
#define VX_N_ACC 6

struct BlobLayout {            // byte offsets inside one robot blob; all multiples of 16
    int nv;                    // voxel stride (max_vox rounded up to 8)
    int real_size;
    size_t off_pos, off_quat, off_lm, off_am, off_hist, off_flags;   // dynamic: copied in and out
    size_t dyn_bytes;
    size_t off_nbr, off_bcls, off_vcls;                               // static: copied in
    size_t stat_bytes;         // dynamic + static
    size_t off_accb;           // scratch: zero-initialised on chip, never written back
    size_t total_bytes;
};

VX_HD size_t vx_align16(size_t x) { return (x + 15) & ~(size_t)15; }

inline BlobLayout make_layout(int max_vox, int real_size) {
    BlobLayout L;
    L.nv = (max_vox + 7) & ~7;
    L.real_size = real_size;
    size_t nv = (size_t)L.nv, rs = (size_t)real_size, o = 0;
    L.off_pos = o;   o = vx_align16(o + 3 * nv * sizeof(double));
    L.off_quat = o;  o = vx_align16(o + 4 * nv * rs);
    L.off_lm = o;    o = vx_align16(o + 3 * nv * rs);
    L.off_am = o;    o = vx_align16(o + 3 * nv * rs);
    L.off_hist = o;  o = vx_align16(o + VX_N_HIST * nv * rs);
    L.off_flags = o; o = vx_align16(o + nv * sizeof(uint32_t));
    L.dyn_bytes = o;
    L.off_nbr = o;   o = vx_align16(o + 3 * nv * sizeof(int16_t));
    L.off_bcls = o;  o = vx_align16(o + 3 * nv * sizeof(uint16_t));
    L.off_vcls = o;  o = vx_align16(o + nv * sizeof(uint16_t));
    L.stat_bytes = o;
    L.off_accb = o;  o = vx_align16(o + VX_N_ACC * nv * rs);
    L.total_bytes = o;
    return L;
}

template <typename Real> struct RobotView {
    int n_vox, nv;
    double* pos;          // [3][nv]
    Real* quat;           // [4][nv] w,x,y,z planes
    Real* lm;             // [3][nv]
    Real* am;             // [3][nv]
    Real* hist;           // [27][nv] plane = axis*9 + {p2.xyz, a1.xyz, a2.xyz}
    uint32_t* flags;      // [nv]
    const int16_t* nbr;   // [3][nv] index of the +axis neighbour, -1 = no bond
    const uint16_t* bcls; // [3][nv] class of the owned +axis bond
    const uint16_t* vcls; // [nv]
    Real* accb;           // [6][nv] Force2/Moment2 received from the -axis neighbours
};

template <typename Real>
VX_HD RobotView<Real> make_view(unsigned char* base, const BlobLayout& L, int n_vox) {
    RobotView<Real> v;
    v.n_vox = n_vox; v.nv = L.nv;
    v.pos = (double*)(base + L.off_pos);
    v.quat = (Real*)(base + L.off_quat);
    v.lm = (Real*)(base + L.off_lm);
    v.am = (Real*)(base + L.off_am);
    v.hist = (Real*)(base + L.off_hist);
    v.flags = (uint32_t*)(base + L.off_flags);
    v.nbr = (const int16_t*)(base + L.off_nbr);
    v.bcls = (const uint16_t*)(base + L.off_bcls);
    v.vcls = (const uint16_t*)(base + L.off_vcls);
    v.accb = (Real*)(base + L.off_accb);
    return v;
}

// Per-robot constants the tasks read (class tables live in shared memory in the kernel).
template <typename Real> struct RobotConsts {
    const BondK<Real>* bk;      // [n_bclass]
    const double* bk_dist;      // [n_bclass] pitch along the bond axis, fp64
    const VoxK<Real>* vk;       // [n_vclass]
    const double* scale_c;      // [n_vclass] Scale of each class as left by the previous step
    const Real* dsc_c;          // [n_vclass] Scale - NominalSize, narrowed
    const double* nom;          // [3][nv] NominalPosition (global; read for fixed DOFs only)
    const Real* ext_disp;       // [n_vclass][3]
    const Real* ext_tdisp;      // [n_vclass][3]
    Real half_bond_z;           // 0.5 BondDampingZ
    Real dt;
    bool floor_on, sticky;
};

// optional per-bond outputs of the last step (vxg_download_state): [13][3][nv] planes
template <typename Real> struct BondRecord { Real* out; };   // out == nullptr: do not record

// Bond (v, axis): evaluate, push the neighbour's share, return the owner's share in (F, M) +=.
// Returns true when the bond diverged (strain > 100, VX_Sim.cpp:1775).
template <typename Real>
VX_HD bool bond_task(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, int ax, Real dt_inv,
                     V3<Real>& F, V3<Real>& M, Real* rec) {
    const int nv = R.nv;
    int w = R.nbr[ax * nv + v];
    if (w < 0) return false;
    int bc = R.bcls[ax * nv + v];
    const BondK<Real>& k = C.bk[bc];
    // bond vector in fp64, pitch removed before narrowing (VXS_BondInternal.cpp:58)
    V3<double> dp = mk3<double>(R.pos[w] - R.pos[v], R.pos[nv + w] - R.pos[nv + v], R.pos[2 * nv + w] - R.pos[2 * nv + v]);
    dp = to_xdir(ax, dp);
    V3<Real> e = mk3<Real>((Real)(dp.x - C.bk_dist[bc]), (Real)dp.y, (Real)dp.z);
    Q4<Real> q1, q2;
    {
        V3<Real> u1 = to_xdir(ax, mk3<Real>(R.quat[nv + v], R.quat[2 * nv + v], R.quat[3 * nv + v]));
        V3<Real> u2 = to_xdir(ax, mk3<Real>(R.quat[nv + w], R.quat[2 * nv + w], R.quat[3 * nv + w]));
        q1.w = R.quat[v]; q1.x = u1.x; q1.y = u1.y; q1.z = u1.z;
        q2.w = R.quat[w]; q2.x = u2.x; q2.y = u2.y; q2.z = u2.z;
    }
    Real dnom = (Real)0.5 * (C.dsc_c[R.vcls[v]] + C.dsc_c[R.vcls[w]]);
    uint32_t fl = R.flags[v];
    bool small = (fl >> ax) & 1u;
    BondHist<Real> h;
    Real* hp = R.hist + (size_t)(ax * 9) * nv + v;
    h.p2 = mk3<Real>(hp[0], hp[nv], hp[2 * nv]);
    h.a1 = mk3<Real>(hp[3 * nv], hp[4 * nv], hp[5 * nv]);
    h.a2 = mk3<Real>(hp[6 * nv], hp[7 * nv], hp[8 * nv]);
    BondOut<Real> o;
    bond_force<Real>(k, ax, e, q1, q2, dnom, C.half_bond_z, dt_inv, small, h, o);
    hp[0] = h.p2.x; hp[nv] = h.p2.y; hp[2 * nv] = h.p2.z;
    hp[3 * nv] = h.a1.x; hp[4 * nv] = h.a1.y; hp[5 * nv] = h.a1.z;
    hp[6 * nv] = h.a2.x; hp[7 * nv] = h.a2.y; hp[8 * nv] = h.a2.z;
    R.flags[v] = (fl & ~(1u << ax)) | ((small ? 1u : 0u) << ax);
    F = F + o.f1; M = M + o.m1;
    R.accb[w] += o.f2.x; R.accb[nv + w] += o.f2.y; R.accb[2 * nv + w] += o.f2.z;
    R.accb[3 * nv + w] += o.m2.x; R.accb[4 * nv + w] += o.m2.y; R.accb[5 * nv + w] += o.m2.z;
    if (rec) {
        Real* r = rec + (size_t)(ax * 13) * nv + v;
        r[0] = o.f1.x; r[nv] = o.f1.y; r[2 * nv] = o.f1.z; r[3 * nv] = o.f2.x; r[4 * nv] = o.f2.y; r[5 * nv] = o.f2.z;
        r[6 * nv] = o.m1.x; r[7 * nv] = o.m1.y; r[8 * nv] = o.m1.z; r[9 * nv] = o.m2.x; r[10 * nv] = o.m2.y; r[11 * nv] = o.m2.z;
        r[12 * nv] = o.strain;
    }
    return o.strain > (Real)100;
}

// Voxel v: total force/moment = own share + received share; integrate (EulerStep).
template <typename Real>
VX_HD void voxel_task(const RobotView<Real>& R, const RobotConsts<Real>& C, int v, V3<Real> F, V3<Real> M) {
    const int nv = R.nv;
    F = F + mk3<Real>(R.accb[v], R.accb[nv + v], R.accb[2 * nv + v]);
    M = M + mk3<Real>(R.accb[3 * nv + v], R.accb[4 * nv + v], R.accb[5 * nv + v]);
    R.accb[v] = 0; R.accb[nv + v] = 0; R.accb[2 * nv + v] = 0; R.accb[3 * nv + v] = 0; R.accb[4 * nv + v] = 0; R.accb[5 * nv + v] = 0;
    int vc = R.vcls[v];
    const VoxK<Real>& k = C.vk[vc];
    VoxState<Real> s;
    s.px = R.pos[v]; s.py = R.pos[nv + v]; s.pz = R.pos[2 * nv + v];
    s.q.w = R.quat[v]; s.q.x = R.quat[nv + v]; s.q.y = R.quat[2 * nv + v]; s.q.z = R.quat[3 * nv + v];
    s.lm = mk3<Real>(R.lm[v], R.lm[nv + v], R.lm[2 * nv + v]);
    s.am = mk3<Real>(R.am[v], R.am[nv + v], R.am[2 * nv + v]);
    double nom[3] = {0, 0, 0};
    if (k.has_bc) { nom[0] = C.nom[v]; nom[1] = C.nom[nv + v]; nom[2] = C.nom[2 * nv + v]; }
    bool sf = voxel_integrate<Real>(k, F, M, 0.5 * C.scale_c[vc], C.floor_on, C.sticky, C.dt, nom,
                                    C.ext_disp + 3 * vc, C.ext_tdisp + 3 * vc, s);
    R.pos[v] = s.px; R.pos[nv + v] = s.py; R.pos[2 * nv + v] = s.pz;
    R.quat[v] = s.q.w; R.quat[nv + v] = s.q.x; R.quat[2 * nv + v] = s.q.y; R.quat[3 * nv + v] = s.q.z;
    R.lm[v] = s.lm.x; R.lm[nv + v] = s.lm.y; R.lm[2 * nv + v] = s.lm.z;
    R.am[v] = s.am.x; R.am[nv + v] = s.am.y; R.am[2 * nv + v] = s.am.z;
    uint32_t fl = R.flags[v];
    R.flags[v] = sf ? (fl | VX_FL_STATIC_FRIC) : (fl & ~VX_FL_STATIC_FRIC);
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
    float half_bond_z_pad;  // unused
    int pad;
    double half_bond_z;     // 0.5 BondDampingZ
};

}  // namespace vx
