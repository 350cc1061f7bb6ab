// vx_integrate_gpu.cpp -- seam B2 of SURVEY.md section 8b: the reference-side binding of libvxgpu.so.
//
// This translation unit DEFINES `bool CVX_Sim::Integrate()` -- the one member function of the reference whose
// body the engine replaces (VX_Sim.cpp:1763-1933) -- against the reference's own, unmodified headers.  Linked
// with the reference's objects (whose own Integrate symbol is weakened by the build recipe, oracle/Makefile),
// everything else stays the reference's code: LoadVXAFile / Import, TimeStep's latches, UpdateCollisions,
// UpdateStats, StopConditionMet, WriteResultFile and voxelyzeMain/main.cpp itself.  So `voxelyze -f x.vxa`,
// VoxCad's QVX_Interface and any other CVX_Sim client drive the CUDA engine without knowing it.
//
// Per call: one timestep on the device (vxg_step(…, 1)), then the device state is written back into
// VoxArray / BondArrayInternal so every getter the GUI, CVX_MeshUtil and UpdateStats read
// (VXS_Voxel.h:46-59, VXS_Bond.h:31-43) still answers.  That per-step round trip is the price of keeping the
// class API; the batched CLI (csrc/voxelyze_main.cpp) is the fast path and never writes back.
//
// No reference source is copied here; the file only calls the reference's API.  It needs write access to
// state the reference keeps private/protected without setters (VXS_Voxel.h:122-138, VXS_Bond.h:54-64), so the
// reference headers -- and only they -- are included with the access keywords neutralised (GCC does not
// reorder members by access, the layout is unchanged).  There is NO CPU fallback: a model that needs a
// feature the kernels lack, or a missing GPU, aborts the process with the engine's message.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "VX_Sim.h"
#include "VX_Environment.h"
#include "VX_Object.h"
#undef private
#undef protected

#include "vxgpu.h"

namespace {

struct Attached {
    vxg_batch* batch = nullptr;
    int n_vox = 0, n_bond = 0;
    bool collisions = false;
    std::vector<int32_t> col_v1, col_v2;      // the contact-bond list last handed to the device
    std::vector<double> pos, quat, lin_mom, ang_mom, scale, f1, f2, m1, m2, strain, vstrain;
    std::vector<int32_t> sfric, small;
};

std::map<const CVX_Sim*, Attached>& registry() { static std::map<const CVX_Sim*, Attached> r; return r; }

[[noreturn]] void die(const char* what, const char* detail) {
    fprintf(stderr, "voxelyze (GPU engine): %s: %s\n", what, detail ? detail : "");
    exit(3);   // neither of the CLI's own exit codes (0 / 1, voxelyzeMain/main.cpp:28,57,132)
}

template <typename T> int intern(std::vector<T>& tab, const T& c) {
    for (size_t i = 0; i < tab.size(); i++) if (!memcmp(&tab[i], &c, sizeof(T))) return (int)i;
    tab.push_back(c);
    return (int)tab.size() - 1;
}

// What the kernels do not implement must stop the run (the product importer does the same on the .vxa text,
// csrc/vxh_import.cpp:check_supported).
void check_supported(CVX_Sim& S) {
    if (S.IsFeatureEnabled(VXSFEAT_VOLUME_EFFECTS)) die("unsupported feature", "VolumeEffectsEnabled");
    if (S.IsFeatureEnabled(VXSFEAT_EQUILIBRIUM_MODE)) die("unsupported feature", "EquilibriumModeEnabled");
    if (S.IsFeatureEnabled(VXSFEAT_MAX_VELOCITY)) die("unsupported feature", "MaxVelLimitEnabled");
    if (S.IsFeatureEnabled(VXSFEAT_BLENDING)) die("unsupported feature", "BlendingEnabled");
    if (S.IsFeatureEnabled(VXSFEAT_PLASTICITY)) die("unsupported feature", "non-linear (plastic) materials");
#ifndef REF_LW
    CVX_Environment* E = S.pEnv;
    if (E->getSourcesPresent()) die("unsupported feature", "environmental <Sources>");
    if (E->getFloorRadius() > 0) die("unsupported feature", "FloorRadius (limited floor)");
    if (E->isFallingProhibited()) die("unsupported feature", "FallingProhibited");
    if (E->GetNumTimeStepsInWindow() > 0) die("unsupported feature", "velocity-adjusted development windows");
#endif
    for (int i = 0; i < S.NumVox(); i++) {
        int model = S.VoxArray[i].GetpMaterial()->GetMatModel();
        if (model != MDL_LINEAR && model != MDL_LINEAR_FAIL) die("unsupported feature", "non-linear material model");
    }
}

// CVX_Sim -> flat descriptor: the constants are the reference's own cached members (CVX_Voxel::SetMaterial
// VX_Voxel.cpp:109-125, CVX_Bond::LinkVoxels / UpdateConstants VX_Bond.cpp:65-173), read as they stand after
// Import + ResetSimulation -- so whatever the reference derived them from (palette or evolved stiffness) carries over.
void attach(CVX_Sim& S, Attached& A) {
    if (vxg_device_count() <= 0) die("no CUDA device", "the engine has no CPU path");
    check_supported(S);
    const int nv = S.NumVox(), nb = S.NumBond();
    std::vector<double> nom(3 * (size_t)nv);
    std::vector<int32_t> vcls(nv), b1(nb), b2(nb), bax(nb), bcls(nb);
    std::vector<vxg_vox_class> vc; std::vector<vxg_bond_class> bc;
    for (int i = 0; i < nv; i++) {
        CVXS_Voxel& v = S.VoxArray[i];
        const Vec3D<> p = v.GetNominalPosition();
        nom[3 * i] = p.x; nom[3 * i + 1] = p.y; nom[3 * i + 2] = p.z;
        vxg_vox_class c; memset(&c, 0, sizeof c);
        const double s = v.GetNominalSize(), E = v.GetEMod();
        c.mass = v.GetMass(); c.mass_inv = v._massInv; c.inertia = v.GetInertia(); c.inertia_inv = v._inertiaInv;
        c.damp_lin = v._2xSqMxExS; c.damp_ang = v._2xSqIxExSxSxS;
        c.k_floor = v.GetLinearStiffness();
        c.mu_static = v.GetpMaterial()->GetuStatic(); c.mu_dynamic = v.GetpMaterial()->GetuDynamic();
        c.nom_size = s; c.cte = v.GetCTE(); c.phase_offset = v.phaseOffset; c.e_mod = E;
#ifndef REF_LW
        c.temp_amp_damp = v.TempAmpDamp; c.init_size = v.initialVoxelSize;
        c.final_size = v.finalVoxelSize; c.growth_time = v.growthTime; c.start_growth_time = v.startGrowthTime;
        c.final_phase_offset = v.finalPhaseOffset; c.final_temp_amp_damp = v.finalTempAmpDamp;
        {
            CVX_Object* O = S.pEnv->pObj;
            c.dev_flags = (O->GetUsingFinalVoxelSize() ? VXG_DEV_FINAL_SIZE : 0) | (O->GetUsingFinalPhaseOffset() ? VXG_DEV_FINAL_PHASE : 0) |
                          (O->GetUsingFinalTempAmpDamp() ? VXG_DEV_FINAL_TAD : 0) |
                          ((O->GetUsingInitialVoxelSize() || O->GetUsingFinalVoxelSize()) ? VXG_DEV_SIZE_ADJUST : 0);
        }
#else
        c.temp_amp_damp = 1.0; c.init_size = s;
#endif
        c.dof_fixed = (unsigned char)v.GetDofFixed(); c.mat_index = v.GetMaterialIndex();
        const Vec3D<> ef = v.GetExternalForce(), et = v.GetExternalTorque(), ed = v.ExternalDisp, ea = v.ExternalTDisp;
        c.ext_force[0] = ef.x; c.ext_force[1] = ef.y; c.ext_force[2] = ef.z;
        c.ext_torque[0] = et.x; c.ext_torque[1] = et.y; c.ext_torque[2] = et.z;
        c.ext_disp[0] = ed.x; c.ext_disp[1] = ed.y; c.ext_disp[2] = ed.z;
        c.ext_tdisp[0] = ea.x; c.ext_tdisp[1] = ea.y; c.ext_tdisp[2] = ea.z;
        vcls[i] = intern(vc, c);
    }
    for (int j = 0; j < nb; j++) {
        CVXS_BondInternal& b = S.BondArrayInternal[j];
        b1[j] = b.GetVox1SInd(); b2[j] = b.GetVox2SInd(); bax[j] = (int)b.GetBondAxis() - 1;   // AXIS_X = 1 (VX_Enums.h:66-70)
        CVXS_Voxel& v1 = S.VoxArray[b1[j]]; CVXS_Voxel& v2 = S.VoxArray[b2[j]];
        vxg_bond_class k; memset(&k, 0, sizeof k);
        k.a1 = b.a1; k.a2 = b.a2; k.b1 = b.b1y; k.b2 = b.b2y; k.b3 = b.b3y;      // cubic voxels: the y and z constants coincide
        k.len = b.L.x; k.area = v1.GetNominalSize() * v1.GetNominalSize();       // CSArea reset to s^2 every step (VXS_Voxel.cpp:625-631)
#ifndef REF_LW   // the stress model's modulus: the palette material's (VX_Object.cpp:1477); the land/water fork passes the voxel's own
        k.e1 = v1.GetpMaterial()->GetElasticMod(); k.e2 = v2.GetpMaterial()->GetElasticMod();
#else            // ... through a `float voxelEmod` parameter (LW/VX_Object.cpp:1472-1474)
        k.e1 = (double)(float)v1.GetEMod(); k.e2 = (double)(float)v2.GetEMod();
#endif
        k.homogeneous = b.HomogenousBond ? 1 : 0;
        const Vec3D<> p1 = v1.GetNominalPosition(), p2 = v2.GetNominalPosition();
        k.nom_dist = bax[j] == 0 ? p2.x - p1.x : (bax[j] == 1 ? p2.y - p1.y : p2.z - p1.z);
        k.c_a1m1 = b._2xSqA1xM1; k.c_a1m2 = b._2xSqA1xM2; k.c_a2i1 = b._2xSqA2xI1; k.c_a2i2 = b._2xSqA2xI2;
        k.c_b1m1 = b._2xSqB1YxM1; k.c_b1m2 = b._2xSqB1YxM2; k.c_b2fm1 = b._2xSqB2YxFM1; k.c_b2fm2 = b._2xSqB2YxFM2;
        k.c_b3i1 = b._2xSqB3YxI1; k.c_b3i2 = b._2xSqB3YxI2;
        bcls[j] = intern(bc, k);
    }
    vxg_robot_desc d; memset(&d, 0, sizeof d);
    d.n_vox = nv; d.n_bond = nb; d.n_vclass = (int)vc.size(); d.n_bclass = (int)bc.size();
    d.nom_pos = nom.data(); d.vox_class = vcls.data(); d.bond_v1 = b1.data(); d.bond_v2 = b2.data();
    d.bond_axis = bax.data(); d.bond_class = bcls.data(); d.vclasses = vc.data(); d.bclasses = bc.data();
    vxg_sim_params& sp = d.sim;
    sp.dt = S.DtFrac * S.OptimalDt;                                 // constant without volume effects (VX_Sim.cpp:1792-1793)
    sp.bond_damp_z = S.GetBondDampZ(); sp.col_damp_z = S.GetCollisionDampZ(); sp.slow_damp_z = S.GetSlowDampZ();
    sp.grav_acc = S.IsFeatureEnabled(VXSFEAT_GRAVITY) ? S.pEnv->GetGravityAccel() : 0.0;
    sp.init_cm_time = S.GetInitCmTime();
    sp.temp_period = (double)(float)S.pEnv->GetTempPeriod(); sp.temp_amplitude = (double)(float)S.pEnv->GetTempAmplitude();   // per-voxel floats upstream
    sp.lattice_dim = S.LocalVXC.GetLatticeDim(); sp.col_horizon = S.CollisionHorizon;
    uint32_t fl = 0;
    if (S.IsFeatureEnabled(VXSFEAT_FLOOR)) fl |= VXG_F_FLOOR;
    if (S.IsFeatureEnabled(VXSFEAT_TEMPERATURE)) fl |= VXG_F_TEMPERATURE;
    A.collisions = S.IsFeatureEnabled(VXSFEAT_COLLISIONS);
    if (A.collisions) fl |= VXG_F_COLLISIONS;
#ifndef REF_LW
    sp.min_scale_fact = S.getMinTempFact(); sp.max_scale_fact = 1.0 + S.pEnv->getGrowthAmplitude();
    sp.growth_amplitude = S.pEnv->getGrowthAmplitude(); sp.midlife_freeze_time = S.GetMidLifeFreezeTime(); sp.stop_condition_value = S.GetStopConditionValue();
    if (S.pEnv->getUsingStickyFloor()) fl |= VXG_F_STICKY_FLOOR;
#else
    sp.min_scale_fact = MIN_TEMP_FACTOR; sp.max_scale_fact = 1.0;
    fl |= VXG_F_LW_SCALE;
    if (S.pEnv->GetFluidEnvironment()) fl |= VXG_F_FLUID;
    sp.drag_coeff = S.pEnv->GetAggregateDragCoefficient();
#endif
    sp.flags = fl;

    vxg_limits lim; memset(&lim, 0, sizeof lim);
    lim.max_vox = nv; lim.max_vclass = d.n_vclass; lim.max_bclass = d.n_bclass > 0 ? d.n_bclass : 1;
    lim.precision = getenv("VXG_ADAPTER_F64") ? VXG_PREC_F64 : VXG_PREC_F32;
    lim.max_col_pairs = A.collisions ? 4 * nv + 64 : 0;             // the host's list (UpdateCollisions stays the reference's)
#ifdef REF_LW   // the fork's deformable surface mesh (fluid drag, RobotVolumeEnd) reads per-voxel strains after every step
    lim.max_mesh_verts = vxg_mesh_vertex_count(&d);
#else
    lim.max_mesh_verts = 0;
#endif
    const char* dev = getenv("VXG_DEVICE");
    if (A.batch) { vxg_destroy(A.batch); A.batch = nullptr; }
    int rc = vxg_create(&A.batch, dev ? atoi(dev) : 0, 1, &lim);
    if (!rc) rc = vxg_upload(A.batch, 0, &d);
    if (!rc) rc = vxg_enable_bond_record(A.batch, 1);
    if (!rc && A.collisions) rc = vxg_set_collision_pairs(A.batch, 0, 0, nullptr, nullptr, nullptr);   // host-managed from the start
    if (rc) die(rc > 0 ? "unsupported feature" : "engine error", vxg_last_error());
    A.n_vox = nv; A.n_bond = nb; A.col_v1.clear(); A.col_v2.clear();
    A.pos.resize(3 * (size_t)nv); A.quat.resize(4 * (size_t)nv); A.lin_mom.resize(3 * (size_t)nv); A.ang_mom.resize(3 * (size_t)nv);
    A.scale.resize(nv); A.sfric.resize(nv);
    A.f1.resize(3 * (size_t)nb); A.f2.resize(3 * (size_t)nb); A.m1.resize(3 * (size_t)nb); A.m2.resize(3 * (size_t)nb);
    A.strain.resize(nb); A.small.resize(nb); A.vstrain.resize(6 * (size_t)nv);
}

// device state -> VoxArray / BondArrayInternal (what EulerStep / UpdateBond leave behind: VXS_Voxel.cpp:184-215, :414-417;
// VXS_BondInternal.cpp:120-170)
void write_back(CVX_Sim& S, Attached& A) {
    vxg_robot_state st; memset(&st, 0, sizeof st);
    st.pos = A.pos.data(); st.quat = A.quat.data(); st.lin_mom = A.lin_mom.data(); st.ang_mom = A.ang_mom.data();
    st.scale = A.scale.data(); st.static_fric = A.sfric.data();
    if (A.n_bond) { st.bond_force1 = A.f1.data(); st.bond_force2 = A.f2.data(); st.bond_moment1 = A.m1.data(); st.bond_moment2 = A.m2.data();
                    st.bond_strain = A.strain.data(); st.bond_small_angle = A.small.data(); }
#ifdef REF_LW
    st.vox_strain = A.vstrain.data();
#endif
    if (vxg_download_state(A.batch, 0, &st)) die("engine error", vxg_last_error());
    const bool kin = (S.StatToCalc & CALCSTAT_KINE) != 0;
    for (int i = 0; i < A.n_vox; i++) {
        CVXS_Voxel& v = S.VoxArray[i];
        v.Pos = Vec3D<double>(A.pos[3 * i], A.pos[3 * i + 1], A.pos[3 * i + 2]);
        v.Angle = CQuat<double>(A.quat[4 * i], A.quat[4 * i + 1], A.quat[4 * i + 2], A.quat[4 * i + 3]);
        v.LinMom = Vec3D<double>(A.lin_mom[3 * i], A.lin_mom[3 * i + 1], A.lin_mom[3 * i + 2]);
        v.AngMom = Vec3D<double>(A.ang_mom[3 * i], A.ang_mom[3 * i + 1], A.ang_mom[3 * i + 2]);
        v.Scale = A.scale[i];
#ifndef REF_LW
        v.lastScale = A.scale[i];
#endif
#ifdef REF_LW   // SetStrainDir of every bond + the corner extents CalcTotalForce derives from them (LW/VXS_Voxel.cpp:258-268, :474-475)
        v.StrainPosDirsCur = Vec3D<>(A.vstrain[6 * i], A.vstrain[6 * i + 1], A.vstrain[6 * i + 2]);
        v.StrainNegDirsCur = Vec3D<>(A.vstrain[6 * i + 3], A.vstrain[6 * i + 4], A.vstrain[6 * i + 5]);
        v.CornerPosCur = (Vec3D<>(1, 1, 1) + v.StrainPosDirsCur) * v.NominalSize / 2;
        v.CornerNegCur = -(Vec3D<>(1, 1, 1) + v.StrainNegDirsCur) * v.NominalSize / 2;
#endif
        v.StaticFricFlag = A.sfric[i] != 0;
        v.Vel = v.LinMom * v._massInv; v.AngVel = v.AngMom * v._inertiaInv;
        if (kin) v.KineticEnergy = 0.5 * v.Mass * v.Vel.Length2() + 0.5 * v.Inertia * v.AngVel.Length2();
    }
    const bool stre = (S.StatToCalc & CALCSTAT_STRAINE) != 0;
    for (int j = 0; j < A.n_bond; j++) {
        CVXS_BondInternal& b = S.BondArrayInternal[j];
        b.Force1 = Vec3D<>(A.f1[3 * j], A.f1[3 * j + 1], A.f1[3 * j + 2]); b.Force2 = Vec3D<>(A.f2[3 * j], A.f2[3 * j + 1], A.f2[3 * j + 2]);
        b.Moment1 = Vec3D<>(A.m1[3 * j], A.m1[3 * j + 1], A.m1[3 * j + 2]); b.Moment2 = Vec3D<>(A.m2[3 * j], A.m2[3 * j + 1], A.m2[3 * j + 2]);
        b.CurStrainTot = A.strain[j]; b.SmallAngle = A.small[j] != 0;
        if (stre) b.StrainEnergy = b.CalcStrainEnergy();
    }
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------
// The replaced member.  Contract of the original (VX_Sim.cpp:1763-1933): advance every bond and voxel one
// step; false when a bond strain exceeded 100 (TimeStep then reports "Simulation Diverged", :1147-1150);
// dt = DtFrac * OptimalDt; CurTime += dt; CurStepCount++.
// ----------------------------------------------------------------------------------------------------------
bool CVX_Sim::Integrate()
{
    Attached& A = registry()[this];
    if (!A.batch || CurStepCount == 0 || A.n_vox != NumVox() || A.n_bond != NumBond()) attach(*this, A);   // first step after Import / ResetSimulation

    if (A.collisions) {   // UpdateCollisions ran in TimeStep (the reference's, on the host): hand its list over when it changed
        const int n = NumColBond();
        bool same = (int)A.col_v1.size() == n;
        for (int i = 0; same && i < n; i++) same = A.col_v1[i] == BondArrayCollision[i].GetVox1SInd() && A.col_v2[i] == BondArrayCollision[i].GetVox2SInd();
        if (!same) {
            A.col_v1.resize(n); A.col_v2.resize(n);
            for (int i = 0; i < n; i++) { A.col_v1[i] = BondArrayCollision[i].GetVox1SInd(); A.col_v2[i] = BondArrayCollision[i].GetVox2SInd(); }
            if (vxg_set_collision_pairs(A.batch, 0, n, A.col_v1.data(), A.col_v2.data(), nullptr)) die("engine error", vxg_last_error());
        }
    }

    uint32_t status = 0;
    if (vxg_step(A.batch, 1) || vxg_sync(A.batch, &status)) die("engine error", vxg_last_error());
    if (status & VXG_ST_COL_OVERFLOW) die("engine error", "contact-bond list overflow");
    if (status & VXG_ST_DIVERGED) return false;

    dt = DtFrac * OptimalDt;
    write_back(*this, A);
    CurTime += dt;
    CurStepCount++;
    return true;
}
