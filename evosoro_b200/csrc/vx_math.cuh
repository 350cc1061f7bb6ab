// vx_math.cuh -- per-bond and per-voxel arithmetic of the Voxelyze step, templated on the working
// precision `Real` (float = production, double = logic-parity mode).  Positions and time are always
// fp64.  Compiles as CUDA device code (nvcc) and as plain host C++ (tests/_build CPU emulation harness
// uses it to exercise exactly the arithmetic the kernels run -- the harness is test infrastructure,
// the product library contains the device instantiation only).
//
// What is computed follows the reference (file:line cited per function); HOW it is computed does
// not: every quantity that is a small difference of O(lattice) numbers (bond extension, shear offsets,
// relative rotation) is formed from small numbers directly ("delta formulation") so that fp32 keeps
// ~1e-7 RELATIVE accuracy on strains instead of 1e-7 of the lattice pitch:
//   * bond vector      = (pos2 - pos1) in fp64, minus the nominal pitch, then narrowed to Real;
//   * inverse rotation = v + 2(u x (u x v) - w (u x v)) so only the O(angle) correction is rounded;
//   * |dp| - pitch     = (|dp|^2 - pitch^2) / (|dp| + pitch) expanded in the small offsets;
//   * rotation vectors = v * 2 sqrt(2/(1+w))  [== 2 sqrt((2-2w)/(1-w^2)), Vec3D.h:279, without the
//                        cancellation], acos branch via atan2(|v|, w).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define VX_HD __host__ __device__ __forceinline__
#else
#define VX_HD inline
#endif

namespace vx {

// thresholds of the reference (PREC_MED branch)
#define VX_SMALL_ANGLE_RAD 1.732e-2      /* Vec3D.h:49 */
#define VX_DISCARD_ANGLE_RAD 1e-7        /* Vec3D.h:48 */
#define VX_W_THRESH_ACOS2SQRT 0.9988     /* Vec3D.h:50 */
#define VX_HYST 1.1                      /* Vec3D.h:22 */
#define VX_SA_BOND_BEND_RAD 0.05         /* VXS_BondInternal.h:23 */
#define VX_SA_BOND_EXT_PERC 1.30         /* VXS_BondInternal.h:26 */
#define VX_PI 3.14159265358979           /* Vec3D.h:21 */
// sin^2 of the half small-angle thresholds: w > cos(a/2)  <=>  |v|^2 < sin^2(a/2) for unit q, w > 0
#define VX_SIN2_SMALL 7.4993725238740121e-05      /* sin^2(0.5 * 1.732e-2)            (IsSmallAngle,  Vec3D.h:290) */
#define VX_SIN2_SMALLISH 9.0741931167803269e-05   /* sin^2(0.5 * 1.1 * 1.732e-2)      (IsSmallishAngle Vec3D.h:291) */
#define VX_SLTHRESH_ACOS2SQRT 0.0023985599999999385 /* 1 - 0.9988^2                      (Vec3D.h:60) */

template <typename T> struct V3 { T x, y, z; };
template <typename T> struct Q4 { T w, x, y, z; };

template <typename T> VX_HD V3<T> mk3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> VX_HD V3<T> operator+(V3<T> a, V3<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> VX_HD V3<T> operator-(V3<T> a, V3<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> VX_HD V3<T> operator-(V3<T> a) { return mk3<T>(-a.x, -a.y, -a.z); }
template <typename T> VX_HD V3<T> operator*(V3<T> a, T f) { return mk3<T>(a.x * f, a.y * f, a.z * f); }
template <typename T> VX_HD V3<T> cross(V3<T> a, V3<T> b) { return mk3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <typename T> VX_HD T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

VX_HD float vx_sqrt(float x) { return sqrtf(x); }
VX_HD double vx_sqrt(double x) { return sqrt(x); }
VX_HD float vx_abs(float x) { return fabsf(x); }
VX_HD double vx_abs(double x) { return fabs(x); }
VX_HD float vx_atan2(float y, float x) { return atan2f(y, x); }
VX_HD double vx_atan2(double y, double x) { return atan2(y, x); }
VX_HD void vx_sincos(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}
VX_HD void vx_sincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}
// Single-instruction SFU forms on the device for the fp32 instantiation (MUFU.RSQ / MUFU.RCP / MUFU.SQRT,
// <= 2 ulp): the IEEE-rounded sqrtf / division expand to ~10 instructions plus a slow-path branch each,
// and their last-ulp difference is far below the fp32 parity tolerance.  fp64 stays IEEE.
VX_HD float vx_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
VX_HD double vx_rsqrt(double x) { return 1.0 / sqrt(x); }
VX_HD float vx_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}
VX_HD double vx_rcp(double x) { return 1.0 / x; }
VX_HD float vx_fsqrt(float x) {   // fast sqrt (vx_sqrt is the IEEE one)
#if defined(__CUDA_ARCH__)
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
VX_HD double vx_fsqrt(double x) { return sqrt(x); }
#if defined(__CUDACC__)
#define VX_NOINLINE __host__ __device__ __noinline__
#else
#define VX_NOINLINE
#endif
// num/den < c and num/den > c without the division, with the IEEE outcome for den <= 0 (negative ratio,
// +-inf, NaN) that the reference's `(abs(z)+abs(y))/x` and `x/NomDistance` tests would have.
// Written with non-short-circuit predicate logic so that they compile to compare + predicate ops, no branches.
// den == 0 exactly (a bond folded to zero length along its own axis) is not distinguished from den < 0.
template <typename T> VX_HD bool ratio_lt(T num, T den, T c) {   // num >= 0
    const bool pos = den > (T)0;
    return (pos & (num < c * den)) | !pos;      // den < 0: the quotient is negative, hence below any positive bound
}
template <typename T> VX_HD bool ratio_gt(T num, T den, T c) {   // num >= 0
    return (den > (T)0) & (num > c * den);      // den < 0: negative quotient, never above a positive bound
}

// bond axis permutation: the bond is evaluated as if it pointed along +X (VX_Bond.h:45-48)
template <typename T> VX_HD V3<T> to_xdir(int ax, V3<T> v) {
    if (ax == 1) return mk3<T>(v.y, -v.x, v.z);
    if (ax == 2) return mk3<T>(v.z, v.y, -v.x);
    return v;
}
template <typename T> VX_HD V3<T> to_origdir(int ax, V3<T> v) {
    if (ax == 1) return mk3<T>(-v.y, v.x, v.z);
    if (ax == 2) return mk3<T>(-v.z, v.y, v.x);
    return v;
}

// the same permutations with the axis known at compile time (free register renaming in the kernels)
template <int AX, typename T> VX_HD V3<T> to_xdir_c(V3<T> v) {
    if (AX == 1) return mk3<T>(v.y, -v.x, v.z);
    if (AX == 2) return mk3<T>(v.z, v.y, -v.x);
    return v;
}
template <int AX, typename T> VX_HD V3<T> to_origdir_c(V3<T> v) {
    if (AX == 1) return mk3<T>(-v.y, v.x, v.z);
    if (AX == 2) return mk3<T>(-v.z, v.y, v.x);
    return v;
}

// rotation matrix of a unit quaternion, for applying one rotation to several vectors:
// v + 2 (u x (u x v) + w (u x v))  ==  M v  with  M = (1 - 2|u|^2) I + 2 u u^T + 2 w [u]x
template <typename T> struct M33 { T xx, xy, xz, yx, yy, yz, zx, zy, zz; };
template <typename T> VX_HD M33<T> rot_matrix(Q4<T> q) {
    T tx = q.x + q.x, ty = q.y + q.y, tz = q.z + q.z;
    T wx = q.w * tx, wy = q.w * ty, wz = q.w * tz;
    T xx = q.x * tx, xy = q.x * ty, xz = q.x * tz, yy = q.y * ty, yz = q.y * tz, zz = q.z * tz;
    M33<T> m;
    m.xx = (T)1 - (yy + zz); m.xy = xy - wz;          m.xz = xz + wy;
    m.yx = xy + wz;          m.yy = (T)1 - (xx + zz); m.yz = yz - wx;
    m.zx = xz - wy;          m.zy = yz + wx;          m.zz = (T)1 - (xx + yy);
    return m;
}
template <typename T> VX_HD V3<T> mat_apply(const M33<T>& m, V3<T> v) {
    return mk3<T>(m.xx * v.x + m.xy * v.y + m.xz * v.z, m.yx * v.x + m.yy * v.y + m.yz * v.z, m.zx * v.x + m.zy * v.y + m.zz * v.z);
}

// rotate v by unit quaternion (w,u): v + 2 (u x (u x v) + w (u x v))      (CQuat::RotateVec3D Vec3D.h:293-299)
template <typename T> VX_HD V3<T> rot_delta(T w, V3<T> u, V3<T> v) {
    V3<T> c = cross(u, v);
    V3<T> d = cross(u, c) + c * w;
    return d * (T)2;
}
template <typename T> VX_HD V3<T> rotate(T w, V3<T> u, V3<T> v) { return v + rot_delta(w, u, v); }

// quaternion product (CQuat::operator* Vec3D.h:193)
template <typename T> VX_HD Q4<T> qmul(Q4<T> a, Q4<T> f) {
    Q4<T> r;
    r.w = a.w * f.w - a.x * f.x - a.y * f.y - a.z * f.z;
    r.x = a.w * f.x + a.x * f.w + a.y * f.z - a.z * f.y;
    r.y = a.w * f.y - a.x * f.z + a.y * f.w + a.z * f.x;
    r.z = a.w * f.z + a.x * f.y - a.y * f.x + a.z * f.w;
    return r;
}

// The same products when one factor has no x component (FromAngleToPosX's result, Vec3D.h:208-237): the terms dropped are
// exact zeros, so the results are those of qmul bit for bit.
template <typename T> VX_HD Q4<T> qmul_x0_left(Q4<T> a, Q4<T> f) {    // a.x == 0
    Q4<T> r;
    r.w = a.w * f.w - a.y * f.y - a.z * f.z;
    r.x = a.w * f.x + a.y * f.z - a.z * f.y;
    r.y = a.w * f.y + a.y * f.w + a.z * f.x;
    r.z = a.w * f.z - a.y * f.x + a.z * f.w;
    return r;
}
template <typename T> VX_HD Q4<T> qmul_x0_right(Q4<T> a, Q4<T> f) {   // f.x == 0
    Q4<T> r;
    r.w = a.w * f.w - a.y * f.y - a.z * f.z;
    r.x = a.x * f.w + a.y * f.z - a.z * f.y;
    r.y = a.w * f.y - a.x * f.z + a.y * f.w;
    r.z = a.w * f.z + a.x * f.y + a.z * f.w;
    return r;
}

// CQuat::ToRotationVector (Vec3D.h:270-285) in cancellation-free form: the vector part scaled by
// 2 sqrt(2/(1+w)) (sqrt branch) or 2 atan2(|v|, w)/|v| (acos branch); zero when the vector part is zero or
// |w| >= 1.  The acos branch (relative rotation above ~5.6 degrees) is rare and kept out of line.
// theta / |v| for a rotation above ~5.6 degrees (|v| = sin(theta/2) >= 0.049, w = cos(theta/2)): the half angle is
// asin(|v|) below 45 degrees and acos(w) above, each evaluated where it is well conditioned (both are short
// polynomial kernels without slow paths, unlike atan2)
VX_HD float vx_half_angle(float l, float w) { return (w > 0.70710678f) ? asinf(l) : acosf(w < -1.0f ? -1.0f : w); }
VX_HD double vx_half_angle(double l, double w) { return atan2(l, w); }
template <typename T> VX_NOINLINE T rotvec_scale_large(T vv, T w) {
    T l = vx_sqrt(vv);
    return (T)2 * vx_half_angle(l, w) / l;
}
template <typename T> VX_HD V3<T> rotvec(Q4<T> q) {
    T vv = q.x * q.x + q.y * q.y + q.z * q.z;
    // 2 sqrt(2/(1+w)); w is clamped into (-1, 1] so the factor stays finite: for |w| >= 1 the reference returns
    // the zero vector, and there the vector part itself is (numerically) zero
    T d = (T)1 + (q.w > (T)1 ? (T)1 : q.w);
    d = d > (T)1e-12 ? d : (T)1e-12;
    T f = (T)2.8284271247461903 * vx_rsqrt(d);
    if (vv >= (T)VX_SLTHRESH_ACOS2SQRT) f = rotvec_scale_large<T>(vv, q.w);
    return mk3<T>(q.x * f, q.y * f, q.z * f);
}

// CQuat::FromAngleToPosX (Vec3D.h:208-237): quaternion turning `from` onto +X (x component is 0).
// Large-angle branch without trigonometry: with n = from/|from|, cos(theta) = n.x, so
// cos(theta/2) = sqrt((1 + n.x)/2) and sin(theta/2)/(|from| sin(theta)) = 1/(2 |from| cos(theta/2));
// theta > pi - 1e-7 (Vec3D.h:224)  <=>  cos^2(theta/2) < sin^2(0.5e-7) = 2.5e-15.
template <typename T> VX_HD Q4<T> from_angle_to_posx(V3<T> from) {
    // (the reference's early-out for a zero vector, Vec3D.h:210, cannot occur for a bond vector; the clamp below
    // keeps that state finite instead of testing for it on every bond)
    Q4<T> q; q.w = (T)1; q.x = (T)0; q.y = (T)0; q.z = (T)0;
    const T sa = (T)VX_SMALL_ANGLE_RAD;
    T ax = vx_abs(from.x);
    if (vx_abs(from.y) < sa * ax && vx_abs(from.z) < sa * ax) {
        T inv = vx_rcp(from.x);
        q.y = (T)0.5 * from.z * inv; q.z = (T)-0.5 * from.y * inv;
        q.w = (T)1 + (T)0.5 * (-q.y * q.y - q.z * q.z);
        return q;
    }
    T rr = from.y * from.y + from.z * from.z;
    T n2 = from.x * from.x + rr;
    T li = vx_rsqrt(n2 > (T)1e-30 ? n2 : (T)1e-30);
    T c2 = (T)0.5 * ((T)1 + from.x * li);
    if (c2 < (T)2.5e-15) { q.w = 0; q.x = 0; q.y = 1; q.z = 0; return q; }
    T ic = vx_rsqrt(c2);                 // 1 / cos(theta/2)
    T k = (T)0.5 * li * ic;
    q.w = c2 * ic; q.y = from.z * k; q.z = -from.y * k;
    return q;
}

// ---- constants narrowed to the working precision -------------------------------------------------
// Both tables are read with 128-bit loads (one class per thread), hence the alignment and padding.
template <typename T> struct alignas(16) BondK {   // from vxg_bond_class
    T a2, b1, b2, b3;
    T inv_len;                         // 1/L.x
    T k_axial;                         // area * E_eff / L.x : axial force per unit of _Pos2.x (see pack_classes)
    T nom_dist;                        // pitch along the axis (as Real)
    T len_minus_dist;                  // L.x - pitch (0 on regular lattices)
    T c_a1m1, c_a1m2, c_a2i1, c_a2i2;
    T c_b1m1, c_b1m2, c_b2fm1, c_b2fm2;
    T c_b3i1, c_b3i2;
    T sf1, sf2;                        // per-voxel share of the axial strain (UpdateBondStrain: CurStrainV1/V2 = sf * strain)
    int homogeneous;
};   // alignas(16) rounds the size up to whole 128-bit words

// Compact form of a bond class for lattices of equal cubic voxels, where every beam constant of
// CVX_Bond::UpdateConstants (VX_Bond.cpp:125-170, A = L^2, Iy = Iz = L^4/12, J = L^4/6, inertia = m L^2/6,
// first moment = m L/2) is a power of L times one of four class numbers:
//   b1 = a1 = E L   b2 = E L^2/2   b3 = E L^3/6   a2 = b3 / (2(1+nu))
//   2 sqrt(a1 m) = 2 sqrt(b1 m) = s   2 sqrt(b2 FM) = s L/2   2 sqrt(b3 I) = s L^2/6   2 sqrt(a2 I) = s L^2/6 / sqrt(2(1+nu))
// Two 128-bit loads per bond instead of five; the host verifies the relations per class at upload
// (pack_classes) and the batch falls back to the wide table if any class deviates.
template <typename T> struct alignas(16) BondKc {
    T ec, k_axial, s1, sf1;                     // composite E; axial force per unit _Pos2.x; 2 sqrt(E L m) (both voxels of a bond have
                                                // the same mass in this form); voxel 1's share of the axial strain (voxel 2's = 2 - sf1)
    T tnu_half, sq_tnu, nom_dist, len_minus_dist;   // 1/(2(1+nu)) and its square root; pitch; L.x - pitch
};
template <typename T> struct BondGeo { T L, L2h, L3s, Lh, L2s, inv_len; };   // L, L^2/2, L^3/6, L/2, L^2/6, 1/L
template <typename T> VX_HD BondGeo<T> make_geo(double L) {
    BondGeo<T> g;
    g.L = (T)L; g.L2h = (T)(L * L / 2); g.L3s = (T)(L * L * L / 6); g.Lh = (T)(L / 2); g.L2s = (T)(L * L / 6); g.inv_len = (T)(1.0 / L);
    return g;
}
template <typename T> VX_HD BondK<T> expand_bond(const BondKc<T>& c, const BondGeo<T>& g) {
    BondK<T> k;
    k.b1 = c.ec * g.L; k.b2 = c.ec * g.L2h; k.b3 = c.ec * g.L3s; k.a2 = k.b3 * c.tnu_half;
    k.inv_len = g.inv_len; k.k_axial = c.k_axial; k.nom_dist = c.nom_dist; k.len_minus_dist = c.len_minus_dist;
    k.c_a1m1 = c.s1; k.c_a1m2 = c.s1; k.c_b1m1 = c.s1; k.c_b1m2 = c.s1;
    k.c_b2fm1 = c.s1 * g.Lh; k.c_b2fm2 = k.c_b2fm1;
    k.c_b3i1 = c.s1 * g.L2s; k.c_b3i2 = k.c_b3i1;
    k.c_a2i1 = k.c_b3i1 * c.sq_tnu; k.c_a2i2 = k.c_b3i2 * c.sq_tnu;
    k.sf1 = c.sf1; k.sf2 = (T)2 - c.sf1;   // the two shares add up to 2 (UpdateBondStrain's balancing step), both 1 for equal moduli
    k.homogeneous = 0;
    return k;
}

template <typename T> struct alignas(16) VoxK {    // from vxg_vox_class (+ sim constants folded in)
    T mass, mass_inv, inertia_inv;
    T slow_lin;        // SlowDampZ * _2xSqMxExS
    T floor_damp;      // ColDampZ  * _2xSqMxExS
    T ang_fact;        // 1 - 10 SlowDampZ I^-1 _2xSqIxExSxSxS dt
    T k_floor, mu_s, mu_d;
    T weight_z;        // m g  (0 when gravity is off / fluid)
    T ext_f[3], ext_t[3];
    int dof;
    int has_bc;        // any fixed DOF or external load
    T half_nom;        // NominalSize / 2 (corner extents of the surface mesh, LW/VXS_Voxel.cpp:474-475)
    T e_mod;           // Vox_E: composite stiffness of contact bonds (CVX_Bond::LinkVoxels VX_Bond.cpp:84-102)
};

template <typename T> struct BondOut {
    V3<T> f1, f2, m1, m2;   // world frame: force/moment on voxel 1 (negative side) and voxel 2
    T strain;
    T strain1, strain2;     // CurStrainV1 / CurStrainV2: what SetStrainDir hands to the two voxels (VXS_BondInternal.cpp:300-304)
    bool changed;           // the small/large-angle mode flipped this step: the caller keeps the old history (:157)
};

template <typename T> struct BondHist {  // _LastPos2, _LastAngle1, _LastAngle2 (VXS_Bond.h:57)
    V3<T> p2, a1, a2;
};

// One internal bond.  CVXS_BondInternal::CalcLinForce (VXS_BondInternal.cpp:56-187) with
// UpdateBondStrain (:189-307, linear materials) and AddDampForces (:310-346).
//   AX       : bond axis (0/1/2); e, q1, q2 are given in the world frame and permuted here (VX_Bond.h:45-48)
//   e        : (pos2 - pos1) minus the pitch along AX, formed in fp64 by the caller, narrowed to Real
//   dnom     : 0.5 (Scale1 + Scale2) - L.x   (deviation of the rest length from nominal)
//   small    : in/out SmallAngle flag;  hist: in = damping history, out = this step's values (the caller
//              stores them unless out.changed)
//   dt_inv   : 1/dt of the PREVIOUS step (0 on the very first step: no damping, history still stored)
// Only the small/large-angle split is a real branch; everything cheap is evaluated unconditionally so
// the common path is one long basic block the scheduler can interleave across the voxels of a thread.
//   sym_mass : every bond of the robot joins voxels of equal mass/inertia (uniform density, as in all evosoro
//              palettes): Force2 == -Force1 exactly, so its evaluation and rotation are skipped (robot-uniform flag)
//   SYM      : the same known at compile time (production family: pack_compact admits such robots only), so the voxel-2
//              constants are never expanded and Moment2 shares the voxel-1 damping terms
//   HOIST    : take the rotation vector of Angle2 out of the mode branch (one evaluation for a warp whose lanes are in both modes
//              instead of one per mode).  Pays where the kernel has registers to spare (+1.4 % at 72 registers, 6^3); at the
//              64-register cap of the 1024-thread kernels the longer live ranges cost more than the instructions saved (-3 %).
template <int AX, typename T, bool SYM = false, bool HOIST = true>
VX_HD void bond_force(const BondK<T>& k, V3<T> e_w, Q4<T> q1w, Q4<T> q2w, T dnom, T half_bond_z, T dt_inv, bool sym_mass_rt,
                      bool& small, BondHist<T>& hist, BondOut<T>& out) {
    const bool sym_mass = SYM || sym_mass_rt;
    const T c_a2i2 = SYM ? k.c_a2i1 : k.c_a2i2, c_b2fm2 = SYM ? k.c_b2fm1 : k.c_b2fm2, c_b3i2 = SYM ? k.c_b3i1 : k.c_b3i2;
    const T D = k.nom_dist;
    V3<T> e = to_xdir_c<AX>(e_w);
    V3<T> u1 = to_xdir_c<AX>(mk3<T>(q1w.x, q1w.y, q1w.z));
    V3<T> u2 = to_xdir_c<AX>(mk3<T>(q2w.x, q2w.y, q2w.z));
    Q4<T> q1; q1.w = q1w.w; q1.x = u1.x; q1.y = u1.y; q1.z = u1.z;
    Q4<T> q2; q2.w = q2w.w; q2.x = u2.x; q2.y = u2.y; q2.z = u2.z;
    V3<T> dp = mk3<T>(D + e.x, e.y, e.z);
    // Ang1AlignedRelPos - (pitch,0,0)  (:65): inverse rotation = rotation by (w, -u)
    V3<T> rr = e + rot_delta<T>(q1.w, -u1, dp);
    T rx = D + rr.x;
    // NewAng2 = conj(q1) * q2  (:66)
    Q4<T> c1; c1.w = q1.w; c1.x = -q1.x; c1.y = -q1.y; c1.z = -q1.z;
    Q4<T> q12 = qmul(c1, q2);
    T vv12 = q12.x * q12.x + q12.y * q12.y + q12.z * q12.z;
    T nomrel = dnom + k.len_minus_dist;      // NomDistance - pitch  (:71)
    T nom = D + nomrel;

    // small/large angle hysteresis (:74-78); SmallTurn = (|z|+|y|)/x and ExtendPerc = x/NomDistance compared
    // without forming the quotients; non-short-circuit on purpose (no branches)
    T turn_num = vx_abs(rr.z) + vx_abs(rr.y);
    bool is_small_w = (q12.w > (T)0) & (vv12 < (T)VX_SIN2_SMALL);
    bool is_smallish_w = (q12.w > (T)0) & (vv12 < (T)VX_SIN2_SMALLISH);
    // NomDistance = 0.5 (Scale1 + Scale2) is positive, so the extension tests need no sign handling
    bool to_small = !small & is_small_w & ratio_lt<T>(turn_num, rx, (T)VX_SA_BOND_BEND_RAD) & (rx < (T)VX_SA_BOND_EXT_PERC * nom);
    bool to_large = small & (!is_smallish_w | ratio_gt<T>(turn_num, rx, (T)(VX_HYST * VX_SA_BOND_BEND_RAD)) | (rx > (T)(VX_HYST * VX_SA_BOND_EXT_PERC) * nom));
    const bool changed = to_small | to_large;
    small = (small | to_small) & !to_large;

    V3<T> p2, an1, an2;
    Q4<T> back;                               // conj(TotalRot): local -> world rotation
    if (HOIST) {
    // The two modes differ in WHICH quaternion Angle2 is taken from and in Pos2 / Angle1 / the frame; what is then done with
    // them is the same.  Only the mode-specific part sits inside the branch, so a warp whose lanes are in both modes (most warps of
    // an actuated robot) issues the shared part -- the rotation vector of Angle2 -- once, not once per mode.
    p2 = mk3<T>(rr.x - nomrel, rr.y, rr.z); an1 = mk3<T>(0, 0, 0);   // small angle, :87-95
    Q4<T> a2q = q12;
    back = q1;
    if (!small) {                             // :101-118
        Q4<T> pa = from_angle_to_posx(mk3<T>(rx, rr.y, rr.z));
        // |dp| - pitch without cancellation
        T d2 = (T)2 * D * e.x + (e.x * e.x + e.y * e.y + e.z * e.z);
        T len = vx_fsqrt(D * D + d2);
        p2 = mk3<T>(d2 * vx_rcp(len + D) - nomrel, 0, 0);
        an1 = rotvec(pa);
        a2q = qmul_x0_left(pa, q12);          // (pa * conj(q1)) * q2
        Q4<T> cpa; cpa.w = pa.w; cpa.x = (T)0; cpa.y = -pa.y; cpa.z = -pa.z;
        back = qmul_x0_right(q1, cpa);        // conj(pa * conj(q1))
    }
    an2 = rotvec(a2q);
    } else if (small) {                       // :87-95
        an1 = mk3<T>(0, 0, 0);
        an2 = rotvec(q12);
        p2 = mk3<T>(rr.x - nomrel, rr.y, rr.z);
        back = q1;
    } else {                                  // :101-118
        Q4<T> pa = from_angle_to_posx(mk3<T>(rx, rr.y, rr.z));
        T d2 = (T)2 * D * e.x + (e.x * e.x + e.y * e.y + e.z * e.z);
        T len = vx_fsqrt(D * D + d2);
        p2 = mk3<T>(d2 * vx_rcp(len + D) - nomrel, 0, 0);
        an1 = rotvec(pa);
        an2 = rotvec(qmul_x0_left(pa, q12));
        Q4<T> cpa; cpa.w = pa.w; cpa.x = (T)0; cpa.y = -pa.y; cpa.z = -pa.z;
        back = qmul_x0_right(q1, cpa);
    }

    // axial stress (:120, :189-307; MDL_LINEAR: VX_Object.cpp:1472-1482).  For linear materials the
    // heterogeneous fixed-point iteration (:226-258) converges in one pass to sigma = 2 E1 E2/(E1+E2) eps
    // (eps1 <- 2 s2/(s1+s2) eps gives s1 = s2 exactly), so the force is a constant times _Pos2.x.
    out.strain = p2.x * k.inv_len;
    out.strain1 = out.strain * k.sf1; out.strain2 = out.strain * k.sf2;
    out.changed = changed;
    T fx = k.k_axial * p2.x;

    // beam equations (:127-128, :153-154) + damping (:157, :310-346).  The damping terms vanish by
    // themselves when the rates are zero, which is what dt == 0 (first step) and a mode flip ask for.
    const T dti = changed ? (T)0 : dt_inv;
    V3<T> rv = (p2 - hist.p2) * dti;
    V3<T> w1 = (an1 - hist.a1) * dti;
    V3<T> w2 = (an2 - hist.a2) * dti;
    const T bz = half_bond_z, hz = (T)0.5 * half_bond_z;          // 0.5 BondDampingZ, and half of it for the moments
    T ey = k.b1 * p2.y - k.b2 * (an1.z + an2.z), ez = k.b1 * p2.z + k.b2 * (an1.y + an2.y);
    T swz = w1.z + w2.z, swy = w1.y + w2.y;
    V3<T> F1 = mk3<T>(fx + bz * (k.c_a1m1 * rv.x), ey + bz * (k.c_b1m1 * rv.y - k.c_b2fm1 * swz), ez + bz * (k.c_b1m1 * rv.z + k.c_b2fm1 * swy));
    // Force2 = -Force1 for a homogeneous bond (:128, :166); with the voxel-2 damping constants the general
    // expression reduces to exactly that when both voxels have the same mass, so no branch is needed
    V3<T> F2 = mk3<T>(0, 0, 0);
    if (!sym_mass) F2 = mk3<T>(-fx - bz * (k.c_a1m2 * rv.x), -ey - bz * (k.c_b1m2 * rv.y - k.c_b2fm2 * swz), -ez - bz * (k.c_b1m2 * rv.z + k.c_b2fm2 * swy));
    T tw = w2.x - w1.x;
    T my = k.b2 * p2.z, mz = -k.b2 * p2.y;
    V3<T> M1 = mk3<T>(k.a2 * (an1.x - an2.x) - hz * (k.c_a2i1 * tw),
                      my + k.b3 * ((T)2 * an1.y + an2.y) + hz * (k.c_b2fm1 * rv.z + k.c_b3i1 * ((T)2 * w1.y + w2.y)),
                      mz + k.b3 * ((T)2 * an1.z + an2.z) + hz * (-k.c_b2fm1 * rv.y + k.c_b3i1 * ((T)2 * w1.z + w2.z)));
    V3<T> M2 = mk3<T>(k.a2 * (an2.x - an1.x) + hz * (c_a2i2 * tw),
                      my + k.b3 * (an1.y + (T)2 * an2.y) + hz * (c_b2fm2 * rv.z + c_b3i2 * (w1.y + (T)2 * w2.y)),
                      mz + k.b3 * (an1.z + (T)2 * an2.z) + hz * (-c_b2fm2 * rv.y + c_b3i2 * (w1.z + (T)2 * w2.z)));
    hist.p2 = p2; hist.a1 = an1; hist.a2 = an2;

    // back to the world frame (:161-170): one rotation matrix for the four vectors
    const M33<T> Rm = rot_matrix(back);
    out.f1 = to_origdir_c<AX>(mat_apply(Rm, F1));
    out.f2 = sym_mass ? -out.f1 : to_origdir_c<AX>(mat_apply(Rm, F2));
    out.m1 = to_origdir_c<AX>(mat_apply(Rm, M1));
    out.m2 = to_origdir_c<AX>(mat_apply(Rm, M2));
}

// CVXS_BondCollision::CalcContactForce (VXS_BondCollision.cpp:41-59): force on voxel 2; voxel 1 gets the negative
template <typename T>
VX_HD V3<T> contact_force(V3<T> dp, T scale1, T scale2, T a1) {
    T nomd = (scale1 + scale2) * (T)0.75;
    T len = vx_sqrt(dot(dp, dp));
    T rel = nomd - len;
    if (rel > (T)0) return dp * (a1 * rel / len);
    return mk3<T>(0, 0, 0);
}

// Tail of CQuat::NormalizeFast (Vec3D.h:246): `if (w >= 1) q = identity`.  In fp64 that only fires
// when the vector part is below ~1e-8 (w rounds to 1); fp32 rounds w to 1 already at |v| < 2.4e-4,
// where the vector part still carries the whole rotation -- so the fp32 instantiation applies the
// reference's effective threshold (|v|^2 <= 2^-53) to the vector part and merely clamps w otherwise.
VX_HD void clamp_unit_w(Q4<double>& a) {
    if (a.w >= 1.0) { a.w = 1; a.x = 0; a.y = 0; a.z = 0; }
}
VX_HD void clamp_unit_w(Q4<float>& a) {
    if (a.w >= 1.0f) {
        a.w = 1.0f;
        if (a.x * a.x + a.y * a.y + a.z * a.z <= 1.1102230e-16f) { a.x = 0; a.y = 0; a.z = 0; }
    }
}

template <typename T> struct VoxState {
    double px, py, pz;      // Pos (fp64 always)
    Q4<T> q;                // Angle
    V3<T> lm, am;           // LinMom, AngMom
};

// One voxel.  CVXS_Voxel::EulerStep (VXS_Voxel.cpp:169-221) with CalcTotalForce (:482-651),
// CalcFloorEffect (:708-758) and CalcTotalMoment (:653-675).
//   F, M       : sums over the voxel's bonds (F: + Force{1|2}; M: + Moment{1|2}, negated here as :664-665)
//   half_scale : 0.5 * Scale of the PREVIOUS step in fp64 (floor penetration :698)
//   nom        : NominalPosition (only read for fixed DOFs)
//   returns StaticFricFlag
template <typename T, bool BC = true>
VX_HD bool voxel_integrate(const VoxK<T>& k, V3<T> F, V3<T> M, double half_scale, bool floor_on, bool sticky,
                           T dt, const double* nom, const T* ext_disp, const T* ext_tdisp, VoxState<T>& s) {
    bool sfric = false;
    if (BC && k.has_bc && (k.dof & 0x3F) == 0x3F) {          // :174-179 fully fixed
        s.lm = mk3<T>(0, 0, 0); s.am = mk3<T>(0, 0, 0);
        s.px = nom[0] + (double)ext_disp[0]; s.py = nom[1] + (double)ext_disp[1]; s.pz = nom[2] + (double)ext_disp[2];
        T l = vx_sqrt(ext_tdisp[0] * ext_tdisp[0] + ext_tdisp[1] * ext_tdisp[1] + ext_tdisp[2] * ext_tdisp[2]);
        T sn, cs; vx_sincos((T)0.5 * l, &sn, &cs);
        T il = l > (T)0 ? (T)1 / l : (T)0;
        s.q.w = cs; s.q.x = ext_tdisp[0] * il * sn; s.q.y = ext_tdisp[1] * il * sn; s.q.z = ext_tdisp[2] * il * sn;
        return false;
    }
    V3<T> vel = s.lm * k.mass_inv;                      // Vel of the previous step (:415)
    F = F - vel * k.slow_lin;                           // :487
    F.z += k.weight_z;                                  // :536
    if (BC && k.has_bc) { F.x += k.ext_f[0]; F.y += k.ext_f[1]; F.z += k.ext_f[2]; }   // :540
    if (floor_on) {                                     // :635-639, :708-758
        T pen = (T)(half_scale - s.pz);
        if (pen > (T)0) {
            T normal = k.k_floor * pen;
            T fz_floor = normal - k.floor_damp * vel.z;
            T sv2 = vel.x * vel.x + vel.y * vel.y;
            T sf2 = F.x * F.x + F.y * F.y;
            T dfr = k.mu_d * normal;
            if (sticky) { vel.x = 0; vel.y = 0; s.lm.x = 0; s.lm.y = 0; sfric = true; }
            if (vel.x == (T)0 && vel.y == (T)0) {
                T lim = k.mu_s * normal;
                if (sf2 < lim * lim) sfric = true;
            } else {
                T sv = vx_sqrt(sv2);
                if (dfr * dt < k.mass * sv) {           // -mu_d N (cos, sin)(atan2(vy, vx)) == -mu_d N v/|v|
                    T g = dfr / sv;
                    F.x -= g * vel.x; F.y -= g * vel.y;
                } else { sfric = true; s.lm.x = 0; s.lm.y = 0; }
            }
            F.z += fz_floor;
        }
        if (sfric) { F.x = 0; F.y = 0; }
    }
    s.lm = s.lm + F * dt;                               // :184
    T dtm = dt * k.mass_inv;
    s.px += (double)(s.lm.x * dtm); s.py += (double)(s.lm.y * dtm); s.pz += (double)(s.lm.z * dtm);   // :185,:193
    M = -M;                                             // :664-665
    if (BC && k.has_bc) {
        if (k.dof & 0x01) { s.px = nom[0] + (double)ext_disp[0]; s.lm.x = 0; }   // :195-197
        if (k.dof & 0x02) { s.py = nom[1] + (double)ext_disp[1]; s.lm.y = 0; }
        if (k.dof & 0x04) { s.pz = nom[2] + (double)ext_disp[2]; s.lm.z = 0; }
        M.x += k.ext_t[0]; M.y += k.ext_t[1]; M.z += k.ext_t[2];                 // :669
        if (k.dof & 0x08) M.x = 0;
        if (k.dof & 0x10) M.y = 0;
        if (k.dof & 0x20) M.z = 0;
    }
    s.am = (s.am + M * dt) * k.ang_fact;                // :202-207
    V3<T> hw = s.am * (k.inertia_inv * (T)0.5);         // :211-212  Spin = 0.5 (0,w) * Angle
    Q4<T> a = s.q;
    T sw = -hw.x * a.x - hw.y * a.y - hw.z * a.z;
    T sx = hw.x * a.w + hw.y * a.z - hw.z * a.y;
    T sy = -hw.x * a.z + hw.y * a.w + hw.z * a.x;
    T sz = hw.x * a.y - hw.y * a.x + hw.z * a.w;
    a.w += sw * dt; a.x += sx * dt; a.y += sy * dt; a.z += sz * dt;              // :214
    T l2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;                        // NormalizeFast Vec3D.h:243-246
    { T li = vx_rsqrt(l2 > (T)1e-30 ? l2 : (T)1e-30); a.w *= li; a.x *= li; a.y *= li; a.z *= li; }   // `if (l)`: never zero for a unit quaternion
    clamp_unit_w(a);
    s.q = a;
    if (BC && k.has_bc && (k.dof & 0x38) == 0x38) {           // :218-221
        T l = vx_sqrt(ext_tdisp[0] * ext_tdisp[0] + ext_tdisp[1] * ext_tdisp[1] + ext_tdisp[2] * ext_tdisp[2]);
        T sn, cs; vx_sincos((T)0.5 * l, &sn, &cs);
        T il = l > (T)0 ? (T)1 / l : (T)0;
        s.q.w = cs; s.q.x = ext_tdisp[0] * il * sn; s.q.y = ext_tdisp[1] * il * sn; s.q.z = ext_tdisp[2] * il * sn;
        s.am = mk3<T>(0, 0, 0);
    }
    return sfric;
}

// Scale law, evaluated once per voxel CLASS per step in fp64 (every voxel of a class has the same
// Scale history).  _voxcad: VXS_Voxel.cpp:224-336 -- prenatal ramp, mid-life freeze, postnatal linear development of
// size / phase / actuation amplitude, size-adjusted actuation; the velocity-adjusted ("ballistic") development
// windows (:340-381) depend on simulation state and are rejected at import.  Land/water fork: LW/VXS_Voxel.cpp:211-235.
// `t` is CurTime BEFORE the increment.  The development members are `float` upstream, so differences / quotients
// of two of them are float operations there (usual arithmetic conversions) and are restated as such.
// IEEE-rounded float quotient whatever the translation unit's --prec-div setting (the reference divides two float members)
VX_HD float vx_fdiv_ieee(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
struct ScaleK {
    double nom_size, init_size, cte, temp_amp_damp, phase_offset;
    double final_size, growth_time, start_growth, final_phase, final_tad;
    int mat_index;         // palette index of the class (fitness probes key on 5 / 6, VX_Sim.cpp:2626-2705)
    int dev_flags;         // VXG_DEV_*
};
struct ScaleSim {          // what the law reads from CVX_Sim / CVX_Environment
    double init_cm_time, temp_period, temp_amplitude, min_fact, max_fact, growth_amp, midlife_freeze, stop_value;
    bool temp_on, lw;
};
VX_HD double scale_law(const ScaleK& c, double last_scale, double t, const ScaleSim& S) {
    const double two_pi_f = (double)(2 * 3.1415926f);          // the reference's float literal (VXS_Voxel.cpp:302)
    if (S.lw) {
        double tf = 1.0;
        if (S.temp_on && t >= S.init_cm_time) tf = 1 + (S.temp_amplitude * sin(two_pi_f * (t / S.temp_period + c.phase_offset))) * c.cte;
        if (tf < S.min_fact) tf = S.min_fact;
        return tf * c.nom_size;
    }
    double cc = (t >= 0.5 * S.init_cm_time) ? 1.0 : 2 * t / S.init_cm_time;      // :237-238
    double prenatal = cc * ((c.init_size / c.nom_size) - 1);
    double frozen_adj = 0, freeze_init = 1;
    if (S.midlife_freeze > 0) {                                                  // :241-259
        double middle = 0.5 * (S.stop_value - S.init_cm_time);
        double fs = middle - 0.5 * S.midlife_freeze, fe = middle + 0.5 * S.midlife_freeze;
        if (t > fs && t < fe) { frozen_adj = t - fs; if (t < fs + S.init_cm_time) freeze_init = 0; }
        if (t > fe) frozen_adj = S.midlife_freeze;
    }
    double dev = 0, dphase = 0, dtad = 0;
    if (c.dev_flags != 0 && t >= c.start_growth && c.growth_time > 0) {          // :262-291 (without a flag k has no effect)
        double lim = (double)((float)c.start_growth + (float)c.growth_time) + S.midlife_freeze;
        double eff = (t <= lim) ? t : lim;
        eff = eff - frozen_adj;
        double k = (eff - c.start_growth) / c.growth_time;
        if (c.dev_flags & 0x1) dev = k * ((double)vx_fdiv_ieee((float)c.final_size, (float)c.init_size) - 1.0);
        if (c.dev_flags & 0x2) dphase = k * (double)((float)c.final_phase - (float)c.phase_offset);
        if (c.dev_flags & 0x4) dtad = k * (double)((float)c.final_tad - (float)c.temp_amp_damp);
    }
    double ctrl = 0;
    if (S.temp_on && t >= S.init_cm_time)                                        // :294-305
        ctrl = (c.temp_amp_damp + dtad) * (S.temp_amplitude * sin(two_pi_f * (t / S.temp_period + (c.phase_offset + dphase)))) * c.cte * freeze_init;
    if (c.dev_flags & 0x8) {                                                     // :308-318
        double cur_size = (1 + prenatal) * (1 + dev) * c.nom_size;
        double orig = (cur_size / c.nom_size - 1) / S.growth_amp;
        double pos = (orig + 1) * 0.5;
        double capped = (pos > 0.5) ? 0.5 : pos;
        ctrl = ctrl * capped * 2;
    }
    double cur = ctrl * c.nom_size + (1 + prenatal) * (1 + dev) * c.nom_size;    // :321
    if (cur < last_scale && cur < S.min_fact * c.nom_size) cur = last_scale;     // :324-330
    if (cur > last_scale && cur > S.max_fact * c.nom_size) cur = last_scale;
    return cur;
}

}  // namespace vx
