// vx_pack.h -- host-side flattening of a vxg_robot_desc (include/vxgpu.h) into the device layout:
// class tables narrowed to the working precision and the initial state blob
// (CVX_Sim::ResetSimulation, VX_Sim.cpp:839-1007; CVXS_Voxel::ResetVoxel, VXS_Voxel.cpp:95-135;
// CVXS_BondInternal::ResetBond, VXS_BondInternal.cpp:40-54).  Pure host C++ (no CUDA calls).
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>
#include "../../include/vxgpu.h"
#include "vx_robot.cuh"

namespace vx {

// Bond classes re-keyed by (bond class, voxel class of either end): every bond of a class then joins the same pair of voxel
// classes, so NomDistance - L = 0.5 (dScale_1 + dScale_2) is one number per bond class and step (BondPairs / dnom_bc in the
// kernels) instead of two dependent table look-ups per bond.  With one voxel class per material (every evosoro palette robot)
// the count does not grow; robots whose per-voxel fields multiply the voxel classes beyond `max_bclass` keep the original
// classes (ok = false) and run the wide family.
struct RekeyedBonds {
    bool ok = false;
    std::vector<vxg_bond_class> bclasses;
    std::vector<int32_t> bond_class;
    std::vector<uint32_t> pair;      // per new class: voxel class of voxel 1 | voxel class of voxel 2 << 16
    vxg_robot_desc desc;             // the descriptor with the re-keyed arrays swapped in (valid while this object lives)
};
inline void rekey_bond_classes(const vxg_robot_desc* d, int max_bclass, RekeyedBonds& out) {
    out.ok = false; out.bclasses.clear(); out.bond_class.assign(d->n_bond, 0); out.pair.clear();
    out.desc = *d;
    // (bond class, voxel class 1, voxel class 2) -> new class: a direct table while the key space is small (palette robots:
    // 16 x 4 x 4), a map otherwise
    const size_t nvc = (size_t)d->n_vclass, space = (size_t)(d->n_bclass > 0 ? d->n_bclass : 1) * nvc * nvc;
    const bool direct = space <= (1u << 16);
    std::vector<int> table(direct ? space : 0, -1);
    std::map<size_t, int> keys;
    for (int j = 0; j < d->n_bond; j++) {
        const int vc1 = d->vox_class[d->bond_v1[j]], vc2 = d->vox_class[d->bond_v2[j]];
        const size_t key = ((size_t)d->bond_class[j] * nvc + (size_t)vc1) * nvc + (size_t)vc2;
        int id = -1;
        if (direct) id = table[key];
        else { auto it = keys.find(key); if (it != keys.end()) id = it->second; }
        if (id < 0) {
            if ((int)out.bclasses.size() >= max_bclass || vc1 > 0xFFFF || vc2 > 0xFFFF) return;
            id = (int)out.bclasses.size();
            if (direct) table[key] = id; else keys.emplace(key, id);
            out.bclasses.push_back(d->bclasses[d->bond_class[j]]);
            out.pair.push_back((uint32_t)vc1 | ((uint32_t)vc2 << 16));
        }
        out.bond_class[j] = id;
    }
    out.ok = true;
    out.desc.n_bclass = (int)out.bclasses.size();
    out.desc.bclasses = out.bclasses.data();
    out.desc.bond_class = out.bond_class.data();
}

// Compact bond classes (BondKc): returns false when some class is not of the equal-cubic-voxel beam form
// (relative deviation above 1e-9), the classes do not share one bond length, or a voxel class carries boundary conditions;
// *len_out = the bond length.
template <typename Real>
inline bool pack_compact(const vxg_robot_desc* d, BondKc<Real>* bkc, double* len_out) {
    auto near = [](double a, double b) { double m = fabs(a) > fabs(b) ? fabs(a) : fabs(b); return fabs(a - b) <= 1e-9 * m; };
    double L = d->n_bclass > 0 ? d->bclasses[0].len : 1.0;
    *len_out = L;
    bool ok = true;
    // the production family is compiled without the boundary-condition code (fixed DOFs, external loads) and the sticky floor:
    // robots that use them run the wide-table family
    if (d->sim.flags & VXG_F_STICKY_FLOOR) ok = false;
    for (int i = 0; i < d->n_vclass; i++) {
        const vxg_vox_class& c = d->vclasses[i];
        if (c.dof_fixed != 0) ok = false;
        for (int a = 0; a < 3; a++) if (c.ext_force[a] != 0 || c.ext_torque[a] != 0) ok = false;
    }
    for (int i = 0; i < d->n_bclass; i++) {
        const vxg_bond_class& c = d->bclasses[i];
        if (!(c.len > 0) || !near(c.len, L) || !(c.b1 > 0) || !(c.b3 > 0) || !(c.a2 > 0)) { ok = false; continue; }
        // the production family takes Force2 = -Force1 and shared damping constants for granted: both voxels of every bond
        // must have the same mass and inertia (uniform density, as in every evosoro palette)
        if (c.c_a1m1 != c.c_a1m2 || c.c_a2i1 != c.c_a2i2 || c.c_b1m1 != c.c_b1m2 || c.c_b2fm1 != c.c_b2fm2 || c.c_b3i1 != c.c_b3i2) ok = false;
        const double ec = c.b1 / L, th = c.a2 / c.b3, sq = sqrt(th), s1 = c.c_a1m1, s2 = c.c_a1m2;
        ok = ok && near(c.b2, ec * L * L / 2) && near(c.b3, ec * L * L * L / 6);
        ok = ok && near(c.c_b1m1, s1) && near(c.c_b1m2, s2) && near(c.c_b2fm1, s1 * L / 2) && near(c.c_b2fm2, s2 * L / 2);
        ok = ok && near(c.c_b3i1, s1 * L * L / 6) && near(c.c_b3i2, s2 * L * L / 6);
        ok = ok && near(c.c_a2i1, s1 * L * L / 6 * sq) && near(c.c_a2i2, s2 * L * L / 6 * sq);
        BondKc<Real>& k = bkc[i];
        k.ec = (Real)ec; k.s1 = (Real)s1; k.tnu_half = (Real)th; k.sq_tnu = (Real)sq;
        {   // strain share of voxel 1 (-> SetStrainDir, VXS_BondInternal.cpp:300-304): 1 unless the stress-balancing iteration ran
            const double diff = c.e1 >= c.e2 ? c.e1 - c.e2 : c.e2 - c.e1, sum = c.e1 + c.e2;
            k.sf1 = (Real)((!c.homogeneous && diff > sum * .0005) ? 2.0 * c.e2 / sum : 1.0);
        }
        k.nom_dist = (Real)c.nom_dist; k.len_minus_dist = (Real)(c.len - c.nom_dist);
        double e_eff = c.e1;
        if (!c.homogeneous) {
            double diff = c.e1 >= c.e2 ? c.e1 - c.e2 : c.e2 - c.e1, sum = c.e1 + c.e2;
            e_eff = diff > sum * .0005 ? 2.0 * c.e1 * c.e2 / sum : 0.5 * sum;
        }
        k.k_axial = (Real)(c.area * e_eff / c.len);
    }
    return ok;
}

template <typename Real>
inline void pack_classes(const vxg_robot_desc* d, BondK<Real>* bk, double* bk_dist, VoxK<Real>* vk, ScaleK* sk,
                         double* scale_c, Real* ext_disp, Real* ext_tdisp) {
    const vxg_sim_params& P = d->sim;
    for (int i = 0; i < d->n_bclass; i++) {
        const vxg_bond_class& c = d->bclasses[i];
        BondK<Real>& k = bk[i];
        k.a2 = (Real)c.a2; k.b1 = (Real)c.b1; k.b2 = (Real)c.b2; k.b3 = (Real)c.b3;
        k.inv_len = (Real)(1.0 / c.len);
        {   // VXS_BondInternal.cpp:206 (homogeneous: E_mat eps) / :226-258 (heterogeneous: one pass of the stress-balancing iteration)
            double e_eff = c.e1;
            if (!c.homogeneous) {
                double diff = c.e1 >= c.e2 ? c.e1 - c.e2 : c.e2 - c.e1, sum = c.e1 + c.e2;
                e_eff = diff > sum * .0005 ? 2.0 * c.e1 * c.e2 / sum : 0.5 * sum;   // loop entered (one pass converges) / not entered
            }
            k.k_axial = (Real)(c.area * e_eff / c.len);
            // strains the two voxels see (-> SetStrainDir, VXS_BondInternal.cpp:300-304): equal unless the iteration ran
            double diff = c.e1 >= c.e2 ? c.e1 - c.e2 : c.e2 - c.e1, sum = c.e1 + c.e2;
            bool iter = !c.homogeneous && diff > sum * .0005;
            k.sf1 = (Real)(iter ? 2.0 * c.e2 / sum : 1.0); k.sf2 = (Real)(iter ? 2.0 * c.e1 / sum : 1.0);
        }
        k.nom_dist = (Real)c.nom_dist; k.len_minus_dist = (Real)(c.len - c.nom_dist);
        k.c_a1m1 = (Real)c.c_a1m1; k.c_a1m2 = (Real)c.c_a1m2; k.c_a2i1 = (Real)c.c_a2i1; k.c_a2i2 = (Real)c.c_a2i2;
        k.c_b1m1 = (Real)c.c_b1m1; k.c_b1m2 = (Real)c.c_b1m2; k.c_b2fm1 = (Real)c.c_b2fm1; k.c_b2fm2 = (Real)c.c_b2fm2;
        k.c_b3i1 = (Real)c.c_b3i1; k.c_b3i2 = (Real)c.c_b3i2;
        k.homogeneous = c.homogeneous;
        bk_dist[i] = c.nom_dist;
    }
    const bool fluid = (P.flags & VXG_F_FLUID) != 0;
    for (int i = 0; i < d->n_vclass; i++) {
        const vxg_vox_class& c = d->vclasses[i];
        VoxK<Real>& k = vk[i];
        k.mass = (Real)c.mass; k.mass_inv = (Real)c.mass_inv; k.inertia_inv = (Real)c.inertia_inv;
        k.slow_lin = (Real)(P.slow_damp_z * c.damp_lin);                                   // VXS_Voxel.cpp:487
        k.floor_damp = (Real)(P.col_damp_z * c.damp_lin);                                  // VXS_Voxel.cpp:725
        k.ang_fact = (Real)(1 - 10 * P.slow_damp_z * c.inertia_inv * c.damp_ang * P.dt);   // VXS_Voxel.cpp:206
        k.k_floor = (Real)c.k_floor; k.mu_s = (Real)c.mu_static; k.mu_d = (Real)c.mu_dynamic;
        k.weight_z = fluid ? (Real)0 : (Real)(c.mass * P.grav_acc);                        // VXS_Voxel.cpp:536, LW :357
        bool bc = c.dof_fixed != 0;
        for (int a = 0; a < 3; a++) {
            k.ext_f[a] = (Real)c.ext_force[a]; k.ext_t[a] = (Real)c.ext_torque[a];
            ext_disp[3 * i + a] = (Real)c.ext_disp[a]; ext_tdisp[3 * i + a] = (Real)c.ext_tdisp[a];
            if (c.ext_force[a] != 0 || c.ext_torque[a] != 0) bc = true;
        }
        k.dof = c.dof_fixed; k.has_bc = bc ? 1 : 0;
        k.half_nom = (Real)(0.5 * c.nom_size); k.e_mod = (Real)c.e_mod;
        ScaleK& s = sk[i];
        s.mat_index = c.mat_index;
        s.nom_size = c.nom_size; s.init_size = c.init_size; s.cte = c.cte; s.temp_amp_damp = c.temp_amp_damp; s.phase_offset = c.phase_offset;
        s.final_size = c.final_size; s.growth_time = c.growth_time; s.start_growth = c.start_growth_time;
        s.final_phase = c.final_phase_offset; s.final_tad = c.final_temp_amp_damp; s.dev_flags = c.dev_flags;
        scale_c[i] = c.nom_size;   // ResetVoxel: Scale = NominalSize (VXS_Voxel.cpp:110)
    }
}

// The reset image of one robot.  `partial`: only the words that differ from robot to robot are written -- the position planes
// and the static section; the rest (identity orientation, zero momenta / history / accumulators, SmallAngle flags) is what
// vx_reset_kernel writes on the device, or pack_blob_rest on the host when the whole blob has to travel.
template <typename Real>
inline void pack_blob(const vxg_robot_desc* d, const BlobLayout& L, unsigned char* blob, double* nom, bool partial = false) {
    if (partial) { memset(blob + L.off_pos, 0, L.off_quat - L.off_pos); memset(blob + L.off_stat, 0, L.stat_bytes - L.off_stat); }
    else memset(blob, 0, L.total_bytes);
    RobotView<Real> R = make_view<Real>(blob, L, d->n_vox);
    const int nv = R.nv;
    for (int i = 0; i < nv; i++) { R.stat[i].nbr[0] = R.stat[i].nbr[1] = R.stat[i].nbr[2] = -1; }
    for (int i = 0; i < d->n_vox; i++) {
        for (int a = 0; a < 3; a++) { R.pos[a * nv + i] = d->nom_pos[3 * i + a]; nom[a * nv + i] = d->nom_pos[3 * i + a]; }
        if (!partial) {
            R.quat[i].x = (Real)1;          // Angle = identity (VXS_Voxel.cpp:98); stored (w,x,y,z)
            R.flags[i] = VX_FL_SMALL_ALL;   // SmallAngle = true (VXS_BondInternal.cpp:43)
        }
        R.stat[i].vcls = (uint16_t)d->vox_class[i];
    }
    for (int j = 0; j < d->n_bond; j++) {   // voxel bond_v1 (negative side) owns the bond
        int a = d->bond_axis[j], v1 = d->bond_v1[j];
        R.stat[v1].nbr[a] = (int16_t)d->bond_v2[j];
        R.stat[v1].bcls[a] = (uint16_t)d->bond_class[j];
    }
}

// Completes a `partial` reset image on the host (same words vx_reset_kernel writes on the device).
template <typename Real>
inline void pack_blob_rest(const BlobLayout& L, unsigned char* blob, int n_vox) {
    memset(blob + L.off_quat, 0, L.dyn_bytes - L.off_quat);
    memset(blob + L.stat_bytes, 0, L.total_bytes - L.stat_bytes);
    RobotView<Real> R = make_view<Real>(blob, L, n_vox);
    for (int i = 0; i < n_vox; i++) { R.quat[i].x = (Real)1; R.flags[i] = VX_FL_SMALL_ALL; }
}

// Integer lattice coordinates of every voxel, recovered from the nominal positions (cell centres at
// (i + 0.5) * pitch, VX_Object.cpp:543); pitch per axis = the bond pitch along it, else the voxel size.
inline void lattice_coords(const vxg_robot_desc* d, std::vector<int>& ijk) {
    double pitch[3] = {0, 0, 0};
    for (int j = 0; j < d->n_bond; j++) { int a = d->bond_axis[j]; double p = d->bclasses[d->bond_class[j]].nom_dist; if (p > 0 && (pitch[a] == 0 || p < pitch[a])) pitch[a] = p; }
    for (int a = 0; a < 3; a++) if (!(pitch[a] > 0)) pitch[a] = d->vclasses[d->vox_class[0]].nom_size;
    ijk.resize((size_t)3 * d->n_vox);
    for (int i = 0; i < d->n_vox; i++) for (int a = 0; a < 3; a++) ijk[3 * i + a] = (int)llround(d->nom_pos[3 * i + a] / pitch[a] - 0.5);
}

// Dense occupancy grid over the robot's bounding box (cells) and its lattice points (cell corners), z slowest / x fastest
// like the reference's own arrays.
struct LatticeGrid {
    int lo[3], n[3];                 // lowest cell index per axis, cells per axis
    std::vector<int> cell;           // [nz][ny][nx] voxel index or -1
    std::vector<int> ijk;
    explicit LatticeGrid(const vxg_robot_desc* d) {
        lattice_coords(d, ijk);
        int hi[3];
        for (int a = 0; a < 3; a++) { lo[a] = 1 << 30; hi[a] = -(1 << 30); }
        for (int i = 0; i < d->n_vox; i++) for (int a = 0; a < 3; a++) { int c = ijk[3 * i + a]; if (c < lo[a]) lo[a] = c; if (c > hi[a]) hi[a] = c; }
        for (int a = 0; a < 3; a++) n[a] = hi[a] - lo[a] + 1;
        cell.assign((size_t)n[0] * n[1] * n[2], -1);
        for (int i = 0; i < d->n_vox; i++) cell[cidx(ijk[3 * i] - lo[0], ijk[3 * i + 1] - lo[1], ijk[3 * i + 2] - lo[2])] = i;
    }
    size_t cidx(int x, int y, int z) const { return ((size_t)z * n[1] + y) * n[0] + x; }
    int at(int x, int y, int z) const { return (x < 0 || y < 0 || z < 0 || x >= n[0] || y >= n[1] || z >= n[2]) ? -1 : cell[cidx(x, y, z)]; }
    size_t n_points() const { return (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1); }
    size_t pidx(int x, int y, int z) const { return ((size_t)z * (n[1] + 1) + y) * (n[0] + 1) + x; }
};

// Number of surface-mesh vertices of a robot: lattice points touched by 1..7 voxels (LW/VX_MeshUtil.cpp:249-262).
inline int mesh_vertex_count(const vxg_robot_desc* d) {
    LatticeGrid G(d);
    std::vector<unsigned char> cnt(G.n_points(), 0);
    for (int i = 0; i < d->n_vox; i++) {
        const int x = G.ijk[3 * i] - G.lo[0], y = G.ijk[3 * i + 1] - G.lo[1], z = G.ijk[3 * i + 2] - G.lo[2];
        for (int c = 0; c < 8; c++) cnt[G.pidx(x + ((c >> 2) & 1), y + ((c >> 1) & 1), z + (c & 1))]++;
    }
    int n = 0;
    for (unsigned char c : cnt) if (c != 0 && c != 8) n++;
    return n;
}

// Surface-mesh tables of CVX_MeshUtil::LinkSimVoxels (LW/VX_MeshUtil.cpp:110-276): exposed-face mask per voxel,
// voxel -> corner vertices, vertex -> voxels.  Vertices are numbered in lattice order (x fastest), interior ones
// (8 voxels) skipped, exactly like the reference's Map[].  Returns the vertex count, or -1 if it exceeds the capacity.
template <typename Real>
inline int pack_mesh(const vxg_robot_desc* d, const BlobLayout& L, unsigned char* blob, int* n_facet_out) {
    RobotView<Real> R = make_view<Real>(blob, L, d->n_vox);
    LatticeGrid G(d);
    static const int dir[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    Corner8 none; for (int k = 0; k < 8; k++) none.i[k] = -1;
    std::vector<Corner8> verts(G.n_points(), none);   // lattice point -> the voxel that has it as corner c
    for (int i = 0; i < d->n_vox; i++) {
        const int x = G.ijk[3 * i] - G.lo[0], y = G.ijk[3 * i + 1] - G.lo[1], z = G.ijk[3 * i + 2] - G.lo[2];
        unsigned faces = 0;
        for (int f = 0; f < 6; f++) if (G.at(x + dir[f][0], y + dir[f][1], z + dir[f][2]) < 0) faces |= 1u << f;
        R.stat[i].faces = (uint16_t)faces;
        for (int c = 0; c < 8; c++) verts[G.pidx(x + ((c >> 2) & 1), y + ((c >> 1) & 1), z + (c & 1))].i[c] = (int16_t)i;
    }
    int n_vert = 0;
    std::vector<int> id(G.n_points(), -1);
    for (size_t p = 0; p < verts.size(); p++) {   // ascending p = (z, y, x) lattice order
        int cnt = 0; for (int k = 0; k < 8; k++) cnt += verts[p].i[k] >= 0;
        if (cnt == 0 || cnt == 8) continue;
        if (n_vert >= L.nvert_cap) return -1;
        Corner8 e = none; int k = 0;   // compacted (code << 12) | voxel, ascending corner code
        for (int c = 0; c < 8; c++) if (verts[p].i[c] >= 0) e.i[k++] = (int16_t)((c << 12) | verts[p].i[c]);
        R.vcon[n_vert] = e;
        id[p] = n_vert++;
    }
    // facets: for every voxel (ascending) its exposed faces +X,-X,+Y,-Y,+Z,-Z, two triangles each; corner codes of the
    // triangles as CVX_MeshUtil::LinkSimVoxels lists them (LW/VX_MeshUtil.cpp:161-190)
    static const int T1[6][3] = {{4, 6, 7}, {0, 3, 2}, {2, 3, 7}, {0, 5, 1}, {1, 5, 7}, {0, 6, 4}};
    static const int T2[6][3] = {{4, 7, 5}, {0, 1, 3}, {2, 7, 6}, {0, 4, 5}, {1, 7, 3}, {0, 2, 6}};
    int n_facet = 0;
    for (int i = 0; i < d->n_vox; i++) {
        const int x = G.ijk[3 * i] - G.lo[0], y = G.ijk[3 * i + 1] - G.lo[1], z = G.ijk[3 * i + 2] - G.lo[2];
        int cv[8];
        for (int c = 0; c < 8; c++) cv[c] = id[G.pidx(x + ((c >> 2) & 1), y + ((c >> 1) & 1), z + (c & 1))];
        R.fbase[i] = (uint16_t)n_facet;
        for (int f = 0; f < 6; f++) {
            if (!((R.stat[i].faces >> f) & 1u)) continue;
            if (n_facet + 2 > L.nfacet_cap || n_facet + 2 > 65535) return -2;
            Facet a; a.a = (int16_t)cv[T1[f][0]]; a.b = (int16_t)cv[T1[f][1]]; a.c = (int16_t)cv[T1[f][2]]; a.v = (int16_t)i;
            Facet b; b.a = (int16_t)cv[T2[f][0]]; b.b = (int16_t)cv[T2[f][1]]; b.c = (int16_t)cv[T2[f][2]]; b.v = (int16_t)i;
            R.facet[n_facet++] = a; R.facet[n_facet++] = b;
        }
    }
    if (n_facet_out) *n_facet_out = n_facet;
    return n_vert;
}

// Static tables of the self-collision broad phase (CVX_Sim::Import VX_Sim.cpp:675-686, CVX_Voxel::CalcNearby
// VX_Voxel.cpp:171-209): the surface voxels (any of the 6 faces without a bond), every voxel's lattice
// coordinates, and per voxel a 64-bit mask over the lattice offsets with |dx|+|dy|+|dz| <= 3 telling which of
// those voxels are within `hops` bond hops (NearbyVoxInds).  Returns false if hops > 3.
inline int near_code_host(int dx, int dy, int dz) {   // rank of (dx,dy,dz) among the offsets with |dx|+|dy|+|dz| <= 3, z slowest (as near_code on the device)
    struct Table {
        signed char code[7][7][7];
        Table() {
            int c = 0;
            for (int z = -3; z <= 3; z++) for (int y = -3; y <= 3; y++) for (int x = -3; x <= 3; x++)
                code[z + 3][y + 3][x + 3] = (abs(x) + abs(y) + abs(z) > 3) ? (signed char)-1 : (signed char)c++;
        }
    };
    static const Table t;
    if (dx < -3 || dx > 3 || dy < -3 || dy > 3 || dz < -3 || dz > 3) return -1;
    return t.code[dz + 3][dy + 3][dx + 3];
}
inline bool pack_collision_tables(const vxg_robot_desc* d, int hops, std::vector<int16_t>& surf, std::vector<int8_t>& lat4, std::vector<unsigned long long>& near) {
    if (hops > 3) return false;
    std::vector<int> ijk; lattice_coords(d, ijk);
    int lo[3] = {1 << 30, 1 << 30, 1 << 30};
    for (int i = 0; i < d->n_vox; i++) for (int a = 0; a < 3; a++) if (ijk[3 * i + a] < lo[a]) lo[a] = ijk[3 * i + a];
    std::vector<int> adj((size_t)6 * d->n_vox, -1), nb(d->n_vox, 0);   // a voxel has at most one bond per direction
    for (int j = 0; j < d->n_bond; j++) { const int a = d->bond_v1[j], b = d->bond_v2[j]; adj[6 * a + nb[a]++] = b; adj[6 * b + nb[b]++] = a; }
    surf.clear(); lat4.assign((size_t)4 * d->n_vox, 0); near.assign(d->n_vox, 0ULL);
    std::vector<int> mark(d->n_vox, -1), seen;
    seen.reserve(128);
    for (int i = 0; i < d->n_vox; i++) {
        if (nb[i] < 6) surf.push_back((int16_t)i);   // IsSurfaceVoxel (VX_Voxel.h:80)
        for (int a = 0; a < 3; a++) { int c = ijk[3 * i + a] - lo[a]; if (c > 120) return false; lat4[4 * i + a] = (int8_t)c; }
        seen.clear(); seen.push_back(i); mark[i] = i;   // breadth-first over the permanent bonds, `hops` levels
        size_t begin = 0;
        for (int h = 0; h < hops; h++) {
            const size_t end = seen.size();
            for (size_t k = begin; k < end; k++) {
                const int u = seen[k];
                for (int e = 0; e < nb[u]; e++) { const int w = adj[6 * u + e]; if (mark[w] != i) { mark[w] = i; seen.push_back(w); } }
            }
            begin = end;
        }
        for (int w : seen) {
            int code = near_code_host(ijk[3 * w] - ijk[3 * i], ijk[3 * w + 1] - ijk[3 * i + 1], ijk[3 * w + 2] - ijk[3 * i + 2]);
            if (code >= 0) near[i] |= 1ULL << code;   // a voxel h hops away is at most h lattice steps away, so it always has a code
        }
    }
    return true;
}

inline void init_ctl(const vxg_robot_desc* d, RobotCtl& c) {
    const vxg_sim_params& P = d->sim;
    memset(&c, 0, sizeof c);
    c.cur_time = 0; c.dt = P.dt; c.init_cm_time = P.init_cm_time; c.temp_period = P.temp_period; c.temp_amplitude = P.temp_amplitude;
    c.min_fact = P.min_scale_fact; c.max_fact = P.max_scale_fact;
    c.growth_amp = P.growth_amplitude; c.midlife_freeze = P.midlife_freeze_time; c.stop_value = P.stop_condition_value;
    c.step = 0; c.total_steps = 0x7fffffffffffffffLL; c.latch_step = -1; c.eol_step = -1; c.eol_post_y = 0; c.status = 0; c.flags = P.flags;
    c.n_vox = d->n_vox; c.n_vclass = d->n_vclass; c.n_bclass = d->n_bclass; c.ini_latched = 0;
    c.half_bond_z = 0.5 * P.bond_damp_z;
    c.drag_coeff = P.drag_coeff; c.n_vert = 0;
    c.col_horizon = P.col_horizon; c.lattice_dim = P.lattice_dim; c.col_acc = 0; c.max_vel = 0;
    c.n_col = 0; c.col_first = 1; c.col_host = 0; c.n_surf = 0; c.col_hops = (int)(P.col_horizon * 1.5);
    {   // where the two-CTA cluster form cuts the robot: the work of a step -- ~6 units per owned bond (v1 owns it), ~5 per voxel --
        // as even as the 512 slots per CTA allow (a dense 10^3 robot: 488 | 512, the top layer has no +Z bonds); even, because the
        // halves travel as bulk copies of 8-byte words on 16-byte boundaries
        std::vector<int> cost(d->n_vox, 5);
        for (int j = 0; j < d->n_bond; j++) cost[d->bond_v1[j]] += 6;
        long long tot = 0; for (int v = 0; v < d->n_vox; v++) tot += cost[v];
        int lo = d->n_vox > 512 ? ((d->n_vox - 512 + 1) & ~1) : 0, hi = (d->n_vox < 512 ? d->n_vox : 512) & ~1;
        long long pre = 0, best = -1; int split = lo;
        for (int v = 0; v <= hi; v++) {
            if (v >= lo && (v & 1) == 0) { const long long diff = pre > tot - pre ? 2 * pre - tot : tot - 2 * pre; if (best < 0 || diff < best) { best = diff; split = v; } }
            if (v < d->n_vox) pre += cost[v];
        }
        c.c2_split = split;
    }
    c.sym_mass = 1;
    for (int i = 0; i < d->n_bclass; i++) {
        const vxg_bond_class& b = d->bclasses[i];
        if (b.c_a1m1 != b.c_a1m2 || b.c_a2i1 != b.c_a2i2 || b.c_b1m1 != b.c_b1m2 || b.c_b2fm1 != b.c_b2fm2 || b.c_b3i1 != b.c_b3i2) c.sym_mass = 0;
    }
}

}  // namespace vx
