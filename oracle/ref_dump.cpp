// ref_dump -- test/baseline harness around the UNMODIFIED reference engine (oracle/_ref build).
//
// TEST INFRASTRUCTURE ONLY.  Links libvoxelyze_ref.a compiled straight from /root/reference (see
// oracle/Makefile); drives it exactly like the reference CLI does
// (evosoro/_voxcad/voxelyzeMain/main.cpp:50-110: LoadVXAFile -> Import -> UpdateCurTemp ->
// loop{TimeStep; UpdateCurTemp}) and dumps, through the PUBLIC members/getters only
// (VX_Sim.h:120-126 VoxArray/BondArrayInternal, VXS_Voxel.h:46-59, VXS_Bond.h:31-43),
//   * the imported model (per-voxel / per-bond descriptor) and
//   * per-step voxel state + bond outputs
// as a flat little-endian binary that tests/ and tools/make_golden.py read with numpy.
//
// Modes
//   ref_dump dump  <in.vxa> <out.bin> <n_steps> <every> [nocol]   trajectory dump
//   ref_dump time  <in.vxa> <warm> <steps> [nocol] [nostat]       prints "voxels steps seconds"
//   ref_dump full  <in.vxa> <out.bin> [nocol]                     run to StopConditionMet, dump final frame + fitness
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include "VX_Object.h"
#include "VX_Environment.h"
#include "VX_Sim.h"
#include "VX_SimGA.h"
#include "VX_MeshUtil.h"

static void wr(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(2); } }
static void wd(FILE* f, double v) { wr(f, &v, 8); }
static void wi(FILE* f, long long v) { wr(f, &v, 8); }

struct Ref {
    CVX_Object Object;
    CVX_Environment Environment;
    CVX_SimGA Sim;
    CVX_MeshUtil Mesh;
    std::string msg;
    bool load(const char* path, bool nocol, bool nostat) {
        Sim.pEnv = &Environment;
        Environment.pObj = &Object;
        Sim.setInternalMesh(&Mesh);
        if (!Sim.LoadVXAFile(path)) { fprintf(stderr, "ref_dump: cannot load %s\n", path); return false; }
        if (nocol) Sim.EnableFeature(VXSFEAT_COLLISIONS, false);
        if (!Sim.Import(&Environment, 0, &msg)) { fprintf(stderr, "ref_dump: import failed: %s\n", msg.c_str()); return false; }
        if (nostat) Sim.StatToCalc = CALCSTAT_NONE;
        Sim.pEnv->UpdateCurTemp(0.0);
#ifdef REF_LW
        Mesh.initializeDeformableMesh(&Sim);
#endif
        return true;
    }
    bool step() {
        bool ok = Sim.TimeStep(&msg);
        Sim.pEnv->UpdateCurTemp(Sim.CurTime);
        return ok;
    }
};

#define N_SIM 24
#define N_VD 42
#define N_BD 4
#define N_VS 15
#define N_BS 14

// The prescribed displacements of fixed DOFs have no public getter (VX_Voxel.h:121-123, protected): read them
// through pointers-to-member re-exported by a derived type (no reference source is changed).
struct VoxBcAccess : CVX_Voxel {
    using CVX_Voxel::ExternalDisp;
    using CVX_Voxel::ExternalTDisp;
};

static void dump_header(FILE* f, Ref& r, long long nframes) {
    CVX_SimGA& S = r.Sim;
    wr(f, "VXDUMP02", 8);
    wi(f, S.NumVox()); wi(f, S.NumBond()); wi(f, nframes);
    wi(f, N_SIM); wi(f, N_VD); wi(f, N_BD); wi(f, N_VS); wi(f, N_BS);
    double sp[N_SIM]; memset(sp, 0, sizeof sp);
    sp[0] = S.DtFrac;           sp[1] = S.OptimalDt;
    sp[2] = S.GetBondDampZ();   sp[3] = S.GetCollisionDampZ(); sp[4] = S.GetSlowDampZ();
    sp[5] = S.IsFeatureEnabled(VXSFEAT_GRAVITY) ? 1 : 0;
    sp[6] = S.pEnv->GetGravityAccel();
    sp[7] = S.IsFeatureEnabled(VXSFEAT_FLOOR) ? 1 : 0;
    sp[8] = S.IsFeatureEnabled(VXSFEAT_TEMPERATURE) ? 1 : 0;
    sp[9] = S.pEnv->GetTempAmplitude();
    sp[10] = S.pEnv->GetTempPeriod();
    sp[11] = S.GetInitCmTime();
    sp[12] = S.GetStopConditionValue();
    sp[13] = (double)S.GetStopConditionType();
    sp[14] = S.LocalVXC.GetLatticeDim();
    sp[15] = S.IsFeatureEnabled(VXSFEAT_COLLISIONS) ? 1 : 0;
    sp[16] = S.LocalVXC.GetVXDim(); sp[17] = S.LocalVXC.GetVYDim(); sp[18] = S.LocalVXC.GetVZDim();
#ifndef REF_LW
    sp[19] = S.getMinTempFact();
    sp[20] = S.pEnv->getGrowthAmplitude();
    sp[21] = S.GetAfterlifeTime();
    sp[22] = S.GetMidLifeFreezeTime();
#else
    sp[19] = 0.1; /* MIN_TEMP_FACTOR, LW/VX_Sim.h:31 */
    sp[22] = S.pEnv->GetFluidEnvironment() ? 1 : 0;
    sp[23] = S.pEnv->GetAggregateDragCoefficient();
#endif
    wr(f, sp, sizeof sp);
    for (int i = 0; i < S.NumVox(); i++) {
        CVXS_Voxel& v = S.VoxArray[i];
        double d[N_VD]; memset(d, 0, sizeof d);
        Vec3D<> np = v.GetNominalPosition();
        d[0] = np.x; d[1] = np.y; d[2] = np.z; d[3] = v.GetNominalSize();
        d[4] = v.GetMaterialIndex();
        d[5] = v.GetEMod(); d[6] = v.GetPoisson(); d[7] = v.GetCTE();
        d[8] = v.GetMass(); d[9] = v.GetInertia(); d[10] = v.GetFirstMoment();
        CVXC_Material* m = v.GetpMaterial();
        d[11] = m->GetuStatic(); d[12] = m->GetuDynamic(); d[13] = m->GetDensity();
        d[14] = m->GetElasticMod(); d[15] = (double)m->GetMatModel();
        d[16] = v.phaseOffset; d[17] = v.TempAmplitude; d[18] = v.TempPeriod;
        d[19] = (double)(unsigned char)v.GetDofFixed();
        d[20] = S.StoXIndexMap[i];
#ifndef REF_LW
        d[21] = v.TempAmpDamp; d[22] = v.initialVoxelSize;
#else
        d[21] = 1.0; d[22] = v.GetNominalSize();
#endif
        d[23] = v.GetLinearStiffness();
        Vec3D<> ef = v.GetExternalForce(), et = v.GetExternalTorque();
        Vec3D<> ed = v.*(&VoxBcAccess::ExternalDisp), ea = v.*(&VoxBcAccess::ExternalTDisp);
        d[24] = ef.x; d[25] = ef.y; d[26] = ef.z; d[27] = et.x; d[28] = et.y; d[29] = et.z;
        d[30] = ed.x; d[31] = ed.y; d[32] = ed.z; d[33] = ea.x; d[34] = ea.y; d[35] = ea.z;
#ifndef REF_LW   /* development parameters as ResetSimulation left them (VX_Sim.cpp:878-975) */
        d[36] = v.finalVoxelSize; d[37] = v.growthTime; d[38] = v.startGrowthTime; d[39] = v.finalPhaseOffset; d[40] = v.finalTempAmpDamp;
        CVX_Object* O = S.pEnv->pObj;
        d[41] = (O->GetUsingFinalVoxelSize() ? 1 : 0) | (O->GetUsingFinalPhaseOffset() ? 2 : 0) | (O->GetUsingFinalTempAmpDamp() ? 4 : 0) |
                ((O->GetUsingInitialVoxelSize() || O->GetUsingFinalVoxelSize()) ? 8 : 0);
#endif
        wr(f, d, sizeof d);
    }
    for (int j = 0; j < S.NumBond(); j++) {
        CVXS_BondInternal& b = S.BondArrayInternal[j];
        double d[N_BD] = { (double)b.GetVox1SInd(), (double)b.GetVox2SInd(), (double)b.GetBondAxis(), b.GetLinearStiffness() };
        wr(f, d, sizeof d);
    }
}

static void dump_frame(FILE* f, Ref& r) {
    CVX_SimGA& S = r.Sim;
    wd(f, S.CurTime); wd(f, S.dt); wd(f, (double)S.CurStepCount);
    Vec3D<> cm = S.GetCM(); wd(f, cm.x); wd(f, cm.y); wd(f, cm.z);
    wd(f, S.IniCM.x); wd(f, S.IniCM.y); wd(f, S.IniCM.z);
    wd(f, (double)S.NumColBond());
    for (int i = 0; i < S.NumVox(); i++) {
        CVXS_Voxel& v = S.VoxArray[i];
        Vec3D<double> p = v.GetCurPosHighAccuracy(); CQuat<double> q = v.GetCurAngleHighAccuracy();
        Vec3D<> vel = v.GetCurVel(), av = v.GetCurAngVel();
        double d[N_VS] = { p.x, p.y, p.z, q.w, q.x, q.y, q.z, vel.x, vel.y, vel.z, av.x, av.y, av.z,
                           v.GetCurScale(), v.GetCurStaticFric() ? 1.0 : 0.0 };
        wr(f, d, sizeof d);
    }
    for (int j = 0; j < S.NumBond(); j++) {
        CVXS_BondInternal& b = S.BondArrayInternal[j];
        Vec3D<> f1 = b.GetForce1(), f2 = b.GetForce2(), m1 = b.GetMoment1(), m2 = b.GetMoment2();
        double d[N_BS] = { f1.x, f1.y, f1.z, f2.x, f2.y, f2.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z,
                           b.GetEngStrain(), b.IsSmallAngle() ? 1.0 : 0.0 };
        wr(f, d, sizeof d);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_dump dump|time|full ...\n"); return 2; }
    std::string mode = argv[1];
    bool nocol = false, nostat = false;
    for (int i = 2; i < argc; i++) { if (!strcmp(argv[i], "nocol")) nocol = true; if (!strcmp(argv[i], "nostat")) nostat = true; }
    Ref* r = new Ref;
    if (!r->load(argv[2], nocol, nostat)) return 1;
    CVX_SimGA& S = r->Sim;

    if (mode == "dump") {
        if (argc < 6) return 2;
        long long n = atoll(argv[4]), every = atoll(argv[5]);
        if (every < 1) every = 1;
        FILE* f = fopen(argv[3], "wb"); if (!f) { perror(argv[3]); return 2; }
        long long nframes = 1 + n / every;
        dump_header(f, *r, nframes);
        dump_frame(f, *r);                       // frame 0 = imported state
        for (long long k = 1; k <= n; k++) {
            if (!r->step()) { fprintf(stderr, "ref_dump: diverged at step %lld\n", k); return 3; }
            if (k % every == 0) dump_frame(f, *r);
        }
        fclose(f);
        return 0;
    }
    if (mode == "time") {
        if (argc < 5) return 2;
        long long warm = atoll(argv[3]), n = atoll(argv[4]);
        for (long long k = 0; k < warm; k++) r->step();
        auto t0 = std::chrono::steady_clock::now();
        for (long long k = 0; k < n; k++) r->step();
        auto t1 = std::chrono::steady_clock::now();
        double sec = std::chrono::duration<double>(t1 - t0).count();
        Vec3D<> cm = S.GetCM();
        printf("%d %lld %.6f %.10g %.10g %.10g\n", S.NumVox(), n, sec, cm.x, cm.y, cm.z);
        return 0;
    }
    if (mode == "full") {
        if (argc < 4) return 2;
        FILE* f = fopen(argv[3], "wb"); if (!f) { perror(argv[3]); return 2; }
        dump_header(f, *r, 2);
        dump_frame(f, *r);
        long long steps = 0;
        while (!S.StopConditionMet()) { if (!r->step()) { fprintf(stderr, "ref_dump: diverged\n"); return 3; } steps++; }
        dump_frame(f, *r);
        // fitness probes exactly as the result writer derives them (VX_SimGA.cpp:36-47, 145-168)
        double lat = S.LocalVXC.GetLatticeDim();
        double dx = S.SS.CurCM.x - S.IniCM.x, dy = S.SS.CurCM.y - S.IniCM.y, dz = S.SS.CurCM.z - S.IniCM.z;
        wd(f, (double)steps);
        wd(f, sqrt(dx * dx + dy * dy) / lat);      // FinalDist / NormFinalDist
        wd(f, dy / lat);                            // finalDistY
        wd(f, sqrt(dx * dx + dy * dy + dz * dz) / lat); // LW normAbsoluteDisplacement
        wd(f, dx / lat); wd(f, dz / lat);
#ifndef REF_LW
        wd(f, S.SS.CurAnteriorDist); wd(f, S.SS.CurPosteriorDist); wd(f, S.SS.CurAnteriorY); wd(f, S.SS.CurPosteriorY);
        wd(f, S.SS.EndOfLifetimePosteriorY);
        wd(f, (double)S.SS.CurNumNonFeetTouchingFloor); wd(f, (double)S.SS.CurNumTouchingFloor);
        wd(f, S.CurTime - S.GetAfterlifeTime());
#else
        for (int i = 0; i < 8; i++) wd(f, 0.0);
#endif
        fclose(f);
        return 0;
    }
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
