// vx_emul.cpp -- CPU emulation harness for the device arithmetic.  TEST INFRASTRUCTURE ONLY.
//
// Compiles evosoro_b200/csrc/vx_math.cuh + vx_robot.cuh + vx_pack.h as plain host C++ and runs the
// kernel's per-voxel tasks serially in the kernel's phase order (3 bond sub-phases, voxel phase,
// class-scale update), so the fp32/fp64 "delta formulation" can be checked against the oracle on a
// machine without a GPU.  It is never linked into libvxgpu.so and nothing in the product calls it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../evosoro_b200/csrc/vx_pack.h"

using namespace vx;

struct EmulBase {
    virtual ~EmulBase() {}
    virtual int64_t step(int64_t n) = 0;
    virtual void get_state(vxg_robot_state* o) = 0;
    virtual double time() = 0;
    virtual int diverged() = 0;
};

template <typename Real, bool COMPACT = false, bool FLUID = false, bool COLLIDE = false> struct Emul : EmulBase {
    BlobLayout L;
    std::vector<unsigned char> blob;
    std::vector<double> nom, bk_dist, scale_c[2];
    std::vector<BondK<Real>> bk;
    std::vector<BondKc<Real>> bkc;
    double bond_len = 1.0;
    bool compact_ok = false;
    std::vector<VoxK<Real>> vk;
    std::vector<ScaleK> sk;
    std::vector<Real> dsc_c[2], dnom_bc[2], ext_disp, ext_tdisp, rec;
    RekeyedBonds rk;
    std::vector<int32_t> bv1, bax;
    RobotCtl ctl;
    int cur = 0;
    int div = 0;
    // self-collision (mirrors the COLLIDE kernel: UpdateCollisions state + CalcL1Bonds tables)
    std::vector<int16_t> surf; std::vector<int8_t> lat4; std::vector<unsigned long long> near;
    std::vector<ColPair> pairs;
    double col_acc = 0, max_vel = 0; bool col_first = true;
    int n_rebuilds = 0;

    Emul(const vxg_robot_desc* d) {
        if (COMPACT) {   // the compact family works on bond classes re-keyed by voxel-class pair (vx_pack.h), like the engine
            rekey_bond_classes(d, 1 << 20, rk);
            d = &rk.desc;
        }
        L = make_layout(stride_for(d->n_vox), (int)sizeof(Real), FLUID ? ((mesh_vertex_count(d) + 31) & ~31) : 0);
        blob.resize(L.total_bytes);
        nom.resize(3 * (size_t)L.nv);
        int nvc = d->n_vclass, nbc = d->n_bclass > 0 ? d->n_bclass : 1;
        bk.resize(nbc); bk_dist.resize(nbc); vk.resize(nvc); sk.resize(nvc);
        scale_c[0].resize(nvc); scale_c[1].resize(nvc); dsc_c[0].resize(nvc); dsc_c[1].resize(nvc);
        ext_disp.resize(3 * nvc); ext_tdisp.resize(3 * nvc);
        pack_classes<Real>(d, bk.data(), bk_dist.data(), vk.data(), sk.data(), scale_c[0].data(), ext_disp.data(), ext_tdisp.data());
        pack_blob<Real>(d, L, blob.data(), nom.data());
        bkc.resize(nbc);
        compact_ok = pack_compact<Real>(d, bkc.data(), &bond_len);
        init_ctl(d, ctl);
        if (FLUID) ctl.n_vert = pack_mesh<Real>(d, L, blob.data(), &ctl.n_facet);
        if (COLLIDE && (ctl.flags & VXG_F_COLLISIONS)) pack_collision_tables(d, ctl.col_hops, surf, lat4, near);
        for (int i = 0; i < nvc; i++) dsc_c[0][i] = (Real)(scale_c[0][i] - sk[i].nom_size);
        dnom_bc[0].assign(nbc, (Real)0); dnom_bc[1].assign(nbc, (Real)0);
        if (COMPACT) for (size_t i = 0; i < rk.pair.size(); i++) dnom_bc[0][i] = (Real)0.5 * (dsc_c[0][rk.pair[i] & 0xFFFFu] + dsc_c[0][rk.pair[i] >> 16]);
        rec.assign((size_t)39 * L.nv, (Real)0);
        bv1.assign(d->bond_v1, d->bond_v1 + d->n_bond);
        bax.assign(d->bond_axis, d->bond_axis + d->n_bond);
    }

    int64_t step(int64_t n) override {
        RobotView<Real> R = make_view<Real>(blob.data(), L, ctl.n_vox);
        RobotConsts<Real> C;
        C.bk = bk.data(); C.bkc = bkc.data(); C.geo = make_geo<Real>(bond_len); C.bk_dist = bk_dist.data(); C.vk = vk.data(); C.nom = nom.data();
        C.ext_disp = ext_disp.data(); C.ext_tdisp = ext_tdisp.data();
        C.half_bond_z = (Real)ctl.half_bond_z; C.dt = (Real)ctl.dt;
        const bool fluid = (ctl.flags & VXG_F_FLUID) != 0;
        C.floor_on = (ctl.flags & VXG_F_FLOOR) && !fluid;
        C.sticky = (ctl.flags & VXG_F_STICKY_FLOOR) != 0;
        C.sym_mass = ctl.sym_mass != 0;
        C.fluid = FLUID && fluid; C.drag_coeff = (Real)ctl.drag_coeff; C.n_vert = ctl.n_vert; C.n_facet = ctl.n_facet;
        const bool temp_on = (ctl.flags & VXG_F_TEMPERATURE) != 0, lw = (ctl.flags & VXG_F_LW_SCALE) != 0;
        const Real dt_inv_run = (Real)(1.0 / ctl.dt);
        std::vector<V3<Real>> F(ctl.n_vox), M(ctl.n_vox);
        int64_t k;
        for (k = 0; k < n; k++) {
            if (div) break;
            C.scale_c = scale_c[cur].data(); C.dsc_c = dsc_c[cur].data(); C.dnom_bc = dnom_bc[cur].data();
            const Real dt_inv = ctl.step == 0 ? (Real)0 : dt_inv_run;
            const bool col_on = COLLIDE && (ctl.flags & VXG_F_COLLISIONS);
            if (col_on) {   // UpdateCollisions (VX_Sim.cpp:1729-1755), same order of operations as the kernel
                col_acc += fabs(max_vel * (ctl.step == 0 ? 0.0 : ctl.dt) / ctl.lattice_dim);
                if (col_acc > (ctl.col_horizon - 1.0) / 2 || col_first) {
                    col_first = false; col_acc = 0; n_rebuilds++;
                    ColTables ct; ct.surf = surf.data(); ct.lat4 = lat4.data(); ct.near = near.data(); ct.n_surf = (int)surf.size(); ct.hops = ctl.col_hops;
                    const Real hz = (Real)ctl.col_horizon, filt = hz * (Real)1.5 * (Real)ctl.lattice_dim;
                    pairs.clear();
                    for (int i = 0; i < ct.n_surf; i++) for (int j = i + 1; j < ct.n_surf; j++)
                        if (col_pair_hit<Real>(R, C, ct, surf[i], surf[j], hz, filt * filt)) { ColPair p; p.v1 = surf[i]; p.v2 = surf[j]; pairs.push_back(p); }
                }
            }
            for (int v = 0; v < ctl.n_vox; v++) { F[v] = mk3<Real>(0, 0, 0); M[v] = mk3<Real>(0, 0, 0); }
            if (C.fluid) {
                for (int u = 0; u < C.n_vert; u++) mesh_vertex<Real>(R, C, u);
                for (int i = 0; i < C.n_facet; i++) facet_task<Real>(R, C, i);
                for (int v = 0; v < ctl.n_vox; v++) F[v] = voxel_drag<Real>(R, v, R.stat[v]);
            }
            bool d = false;
            std::vector<uint32_t> fl(ctl.n_vox);
            for (int v = 0; v < ctl.n_vox; v++) fl[v] = R.flags[v];
            for (int v = 0; v < ctl.n_vox; v++) d |= bond_task<Real, 0, COMPACT, FLUID>(R, C, v, R.stat[v], fl[v], dt_inv, F[v], M[v], rec.data());
            for (int v = 0; v < ctl.n_vox; v++) d |= bond_task<Real, 1, COMPACT, FLUID>(R, C, v, R.stat[v], fl[v], dt_inv, F[v], M[v], rec.data());
            for (int v = 0; v < ctl.n_vox; v++) d |= bond_task<Real, 2, COMPACT, FLUID>(R, C, v, R.stat[v], fl[v], dt_inv, F[v], M[v], rec.data());
            if (col_on && !pairs.empty())
                for (int v = 0; v < ctl.n_vox; v++) F[v] = F[v] + col_force_on<Real>(R, C, v, R.stat[v], load_own<Real>(R, v), pairs.data(), (int)pairs.size());
            if (d) { div = 1; break; }
            Real sp2 = 0;
            for (int v = 0; v < ctl.n_vox; v++) { Real x = voxel_task<Real>(R, C, v, R.stat[v], load_own<Real>(R, v), fl[v], F[v], M[v]); if (x > sp2) sp2 = x; }
            max_vel = sizeof(Real) == 8 ? sqrt((double)sp2) : sqrt((double)(float)sp2);
            for (int i = 0; i < ctl.n_vclass; i++) {
                ScaleSim SL;
                SL.init_cm_time = ctl.init_cm_time; SL.temp_period = ctl.temp_period; SL.temp_amplitude = ctl.temp_amplitude;
                SL.min_fact = ctl.min_fact; SL.max_fact = ctl.max_fact; SL.growth_amp = ctl.growth_amp;
                SL.midlife_freeze = ctl.midlife_freeze; SL.stop_value = ctl.stop_value; SL.temp_on = temp_on; SL.lw = lw;
                double ns = scale_law(sk[i], scale_c[cur][i], ctl.cur_time, SL);
                scale_c[cur ^ 1][i] = ns;
                dsc_c[cur ^ 1][i] = (Real)(ns - sk[i].nom_size);
            }
            if (COMPACT) for (size_t i = 0; i < rk.pair.size(); i++) dnom_bc[cur ^ 1][i] = (Real)0.5 * (dsc_c[cur ^ 1][rk.pair[i] & 0xFFFFu] + dsc_c[cur ^ 1][rk.pair[i] >> 16]);
            cur ^= 1;
            ctl.cur_time += ctl.dt;
            ctl.step++;
        }
        return k;
    }

    void get_state(vxg_robot_state* o) override {
        RobotView<Real> R = make_view<Real>(blob.data(), L, ctl.n_vox);
        const int nv = R.nv;
        for (int i = 0; i < ctl.n_vox; i++) {
            const R4<Real> q = R.quat[i], ma = R.momA[i]; const R2<Real> mb = R.momB[i];
            if (o->pos) for (int a = 0; a < 3; a++) o->pos[3 * i + a] = R.pos[a * nv + i];
            if (o->quat) { o->quat[4 * i] = (double)q.x; o->quat[4 * i + 1] = (double)q.y; o->quat[4 * i + 2] = (double)q.z; o->quat[4 * i + 3] = (double)q.w; }
            if (o->lin_mom) { o->lin_mom[3 * i] = (double)ma.x; o->lin_mom[3 * i + 1] = (double)ma.y; o->lin_mom[3 * i + 2] = (double)ma.z; }
            if (o->ang_mom) { o->ang_mom[3 * i] = (double)ma.w; o->ang_mom[3 * i + 1] = (double)mb.x; o->ang_mom[3 * i + 2] = (double)mb.y; }
            if (o->scale) o->scale[i] = scale_c[cur][R.stat[i].vcls];
            if (o->static_fric) o->static_fric[i] = (R.flags[i] & VX_FL_STATIC_FRIC) ? 1 : 0;
        }
        for (size_t j = 0; j < bv1.size(); j++) {
            int a = bax[j], v = bv1[j];
            const Real* r = rec.data() + (size_t)(a * 13) * nv + v;
            for (int x = 0; x < 3; x++) {
                if (o->bond_force1) o->bond_force1[3 * j + x] = (double)r[(size_t)x * nv];
                if (o->bond_force2) o->bond_force2[3 * j + x] = (double)r[(size_t)(3 + x) * nv];
                if (o->bond_moment1) o->bond_moment1[3 * j + x] = (double)r[(size_t)(6 + x) * nv];
                if (o->bond_moment2) o->bond_moment2[3 * j + x] = (double)r[(size_t)(9 + x) * nv];
            }
            if (o->bond_strain) o->bond_strain[j] = (double)r[(size_t)12 * nv];
            if (o->bond_small_angle) o->bond_small_angle[j] = (R.flags[v] >> a) & 1u;
        }
    }
    double time() override { return ctl.cur_time; }
    int diverged() override { return div; }
};

extern "C" {
void* vxe_create(const vxg_robot_desc* d, int prec) {
    if (prec == VXG_PREC_F64) return new Emul<double>(d);
    if (prec == 7) return new Emul<double, false, true, true>(d);    // fp64, land/water drag + self-collision
    if (prec == 8) return new Emul<float, false, true, true>(d);     // fp32, same
    if (prec == 5) return new Emul<double, false, false, true>(d);   // fp64 with self-collision
    if (prec == 6) return new Emul<float, false, false, true>(d);    // fp32 with self-collision
    if (prec == 3) return new Emul<float, false, true>(d);    // fp32, land/water layout with the fluid-drag phase
    if (prec == 4) return new Emul<double, false, true>(d);   // fp64, same (logic parity of the drag phase on the CPU)
    if (prec == 9 || prec == 10 || prec == 11) {   // fp32 compact form with the land/water drag phase (10: and the self-collision list; 11: list only)
        EmulBase* e; bool ok;
        if (prec == 9) { auto* x = new Emul<float, true, true>(d); ok = x->compact_ok; e = x; }
        else if (prec == 10) { auto* x = new Emul<float, true, true, true>(d); ok = x->compact_ok; e = x; }
        else { auto* x = new Emul<float, true, false, true>(d); ok = x->compact_ok; e = x; }
        if (!ok) { delete e; return nullptr; }
        return e;
    }
    if (prec == 2) {   // fp32 with the compact bond table (what the production kernel uses when every class qualifies)
        Emul<float, true>* e = new Emul<float, true>(d);
        if (!e->compact_ok) { delete e; return nullptr; }
        return e;
    }
    return new Emul<float>(d);
}
void vxe_destroy(void* h) { delete (EmulBase*)h; }
int64_t vxe_step(void* h, int64_t n) { return ((EmulBase*)h)->step(n); }
void vxe_get_state(void* h, vxg_robot_state* o) { ((EmulBase*)h)->get_state(o); }
double vxe_time(void* h) { return ((EmulBase*)h)->time(); }
int vxe_diverged(void* h) { return ((EmulBase*)h)->diverged(); }
}
