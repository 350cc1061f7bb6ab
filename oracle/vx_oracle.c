/* vx_oracle.c -- plain-C fp64 restatement of the reference Voxelyze timestep.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, operation for operation (so that -ffp-contract=off builds are bit-comparable with the
 * unmodified reference built under oracle/_ref):
 *   CVX_Sim::Integrate                 evosoro/_voxcad/Voxelyze/VX_Sim.cpp:1763-1933
 *   CVXS_BondInternal::CalcLinForce    evosoro/_voxcad/Voxelyze/VXS_BondInternal.cpp:56-187
 *   CVXS_BondInternal::UpdateBondStrain                                            :189-307
 *   CVXS_BondInternal::AddDampForces                                               :310-346
 *   CVXS_BondCollision::CalcContactForce  evosoro/_voxcad/Voxelyze/VXS_BondCollision.cpp:41-59
 *   CVXS_Voxel::EulerStep              evosoro/_voxcad/Voxelyze/VXS_Voxel.cpp:169-427
 *   CVXS_Voxel::CalcTotalForce / CalcTotalMoment / CalcFloorEffect                 :482-758
 *   CQuat / Vec3D helpers              evosoro/_voxcad/Voxelyze/Utils/Vec3D.h:185-314
 * within the feature set SURVEY.md Appendix A lists (linear materials, no volume effects, no
 * plasticity, no max-velocity clamp, no equilibrium mode).  The land/water fork's scale law
 * (evosoro/_voxcad_land_water/Voxelyze/VXS_Voxel.cpp:211-235) is selected by VXG_F_LW_SCALE.
 *
 * Also restated (fp64, reference order of operations):
 *   CVX_Sim::UpdateCollisions / CalcL1Bonds, CVX_Voxel::CalcNearby   VX_Sim.cpp:1729-1755, :2357-2413; VX_Voxel.cpp:171-209
 *   land/water fluid drag: CVX_MeshUtil::LinkSimVoxels / GetCurVLoc, CMesh::CalcFaceNormals, the facet loop
 *       evosoro/_voxcad_land_water/Voxelyze/VX_MeshUtil.cpp:110-276, :388-428; Utils/Mesh.cpp:659-665; VX_Sim.cpp:1516-1597
 *
 * Not part of the product: nothing under evosoro_b200/ or include/ references this file.
 */
#include "vx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_REF 3.14159265358979                 /* Vec3D.h:21 */
#define HYST 1.1                                /* VEC3D_HYSTERESIS_FACTOR Vec3D.h:22 */
#define DISCARD_ANGLE_RAD 1e-7                  /* Vec3D.h:48 */
#define SMALL_ANGLE_RAD 1.732e-2                /* Vec3D.h:49 */
#define W_THRESH_ACOS2SQRT 0.9988               /* Vec3D.h:50 */
#define SA_BOND_BEND_RAD 0.05                   /* VXS_BondInternal.h:23 */
#define SA_BOND_EXT_PERC 1.30                   /* VXS_BondInternal.h:26 */

typedef struct { double x, y, z; } v3;
typedef struct { double w, x, y, z; } q4;

struct vxo_sim {
    int n_vox, n_bond, n_col;
    vxg_sim_params sim;
    vxg_vox_class* vc;  int n_vc;
    vxg_bond_class* bc; int n_bc;
    int *vcls, *bv1, *bv2, *bax, *bcls;
    int (*link)[6];            /* InternalBondIndices[6] in BD_PX,NX,PY,NY,PZ,NZ order; -1 none */
    v3* nom;
    /* voxel state */
    v3 *pos, *linmom, *angmom;
    q4* ang;
    double *scale, *last_scale;
    int* sfric;
    /* bond state + outputs */
    int* small;
    v3 *lp2, *la1, *la2;
    v3 *f1, *f2, *m1, *m2;
    double* strain;
    /* contact bonds */
    int *cv1, *cv2; double* ca1; v3 *cf1, *cf2;
    double cur_time, dt;
    int64_t step;
    int diverged;
    /* self-collision (UpdateCollisions state) */
    int col_auto;               /* the list is rebuilt here (no vxo_set_collision_pairs call) */
    int col_first;              /* ColEnableChanged */
    double col_acc, max_vel;    /* MaxDispSinceLastBondUpdate, SS.MaxVoxVel */
    int* ijk;                   /* lattice coordinates [n_vox][3] */
    int *near_off, *near_idx;   /* CalcNearby lists (CSR) */
    int* surf; int n_surf;      /* SurfVoxels */
    /* land/water fluid drag */
    v3 *sp, *sn;                /* StrainPosDirsCur / StrainNegDirsCur */
    v3 *cpos, *cneg;            /* CornerPosCur / CornerNegCur */
    v3* drag;                   /* DragForce */
    int n_vert, n_facet;
    int* vert_off; int* vert_vox; int* vert_corner;   /* CalcVerts[].ConVoxels (voxel, corner code) */
    int* facet_v; int* facet_owner;                   /* DefMesh.Facets[].vi[3], FacetToSIndex */
    v3* vpos;
    double small_w, smallish_w, slthresh;
};

static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 vscale(v3 a, double f) { return V(f * a.x, f * a.y, f * a.z); }          /* Vec3D::operator*(T)  Vec3D.h:83 */
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static v3 vdiv(v3 a, double f) { double inv = 1.0 / f; return V(inv * a.x, inv * a.y, inv * a.z); } /* Vec3D.h:85 */

/* CQuat::operator*  Vec3D.h:193 */
static q4 qmul(q4 a, q4 f) {
    q4 r;
    r.w = a.w * f.w - a.x * f.x - a.y * f.y - a.z * f.z;
    r.x = a.w * f.x + a.x * f.w + a.y * f.z - a.z * f.y;
    r.y = a.w * f.y - a.x * f.z + a.y * f.w + a.z * f.x;
    r.z = a.w * f.z + a.x * f.y - a.y * f.x + a.z * f.w;
    return r;
}
static q4 qconj(q4 a) { q4 r = {a.w, -a.x, -a.y, -a.z}; return r; }

/* CQuat::RotateVec3DInv  Vec3D.h:300-314 */
static v3 qrot_inv(q4 q, v3 f) {
    double tw = q.x * f.x + q.y * f.y + q.z * f.z;
    double tx = q.w * f.x - q.y * f.z + q.z * f.y;
    double ty = q.w * f.y + q.x * f.z - q.z * f.x;
    double tz = q.w * f.z - q.x * f.y + q.y * f.x;
    return V(tw * q.x + tx * q.w + ty * q.z - tz * q.y, tw * q.y - tx * q.z + ty * q.w + tz * q.x,
             tw * q.z + tx * q.y - ty * q.x + tz * q.w);
}

/* CQuat::ToRotationVector  Vec3D.h:270-285 */
static v3 q_to_rotvec(const vxo_sim* s, q4 q) {
    double sl = 1.0 - q.w * q.w;
    if (sl <= 0) return V(0, 0, 0);
    else if (sl < s->slthresh) {
        double f = sqrt((2 - 2 * (q.w > 1 ? 1 : q.w)) / sl);
        return vscale(vscale(V(q.x, q.y, q.z), 2.0), f);
    } else {
        double ac = acos(q.w > 1 ? 1 : q.w);
        return vdiv(vscale(vscale(V(q.x, q.y, q.z), 2.0), ac), sqrt(sl));
    }
}

/* CQuat::FromAngleToPosX  Vec3D.h:208-237 (quaternion starts as identity, Vec3D.h:161,175) */
static q4 q_from_angle_to_posx(v3 from) {
    q4 q = {1, 0, 0, 0};
    if (from.x == 0 && from.y == 0 && from.z == 0) return q;
    double yox = from.y / from.x, zox = from.z / from.x;
    if (yox < SMALL_ANGLE_RAD && yox > -SMALL_ANGLE_RAD && zox < SMALL_ANGLE_RAD && zox > -SMALL_ANGLE_RAD) {
        q.x = 0; q.y = 0.5 * zox; q.z = -0.5 * yox;
        q.w = 1 + 0.5 * (-q.y * q.y - q.z * q.z);
        return q;
    }
    v3 n = from;
    { /* Vec3D::NormalizeFast Vec3D.h:110 */
        double l = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
        if (l > 0) { double li = 1.0 / l; n.x *= li; n.y *= li; n.z *= li; }
    }
    double theta = acos(n.x);
    if (theta > PI_REF - DISCARD_ANGLE_RAD) { q.w = 0; q.x = 0; q.y = 1; q.z = 0; return q; }
    double ami = 1.0 / sqrt(n.z * n.z + n.y * n.y);
    double a = 0.5 * theta;
    double sn = sin(a);
    q.w = cos(a); q.x = 0; q.y = n.z * ami * sn; q.z = -n.y * ami * sn;
    return q;
}

/* CVX_Bond::ToXDirBond / ToOrigDirBond  VX_Bond.h:45-48 */
static v3 to_xdir(int ax, v3 v) {
    if (ax == 1) { double t = v.x; v.x = v.y; v.y = -t; }
    else if (ax == 2) { double t = v.x; v.x = v.z; v.z = -t; }
    return v;
}
static q4 to_xdir_q(int ax, q4 q) {
    if (ax == 1) { double t = q.x; q.x = q.y; q.y = -t; }
    else if (ax == 2) { double t = q.x; q.x = q.z; q.z = -t; }
    return q;
}
static v3 to_origdir(int ax, v3 v) {
    if (ax == 1) { double t = v.y; v.y = v.x; v.x = -t; }
    else if (ax == 2) { double t = v.z; v.z = v.x; v.x = -t; }
    return v;
}

#define AL(T, n) (T*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(T))
static void build_topology(vxo_sim* s);

vxo_sim* vxo_create(const vxg_robot_desc* d) {
    vxo_sim* s = (vxo_sim*)calloc(1, sizeof(vxo_sim));
    int nv = d->n_vox, nb = d->n_bond;
    s->n_vox = nv; s->n_bond = nb; s->sim = d->sim;
    s->n_vc = d->n_vclass; s->n_bc = d->n_bclass;
    s->vc = (vxg_vox_class*)malloc(sizeof(vxg_vox_class) * (size_t)(d->n_vclass > 0 ? d->n_vclass : 1));
    s->bc = (vxg_bond_class*)malloc(sizeof(vxg_bond_class) * (size_t)(d->n_bclass > 0 ? d->n_bclass : 1));
    memcpy(s->vc, d->vclasses, sizeof(vxg_vox_class) * (size_t)d->n_vclass);
    if (d->n_bclass) memcpy(s->bc, d->bclasses, sizeof(vxg_bond_class) * (size_t)d->n_bclass);
    s->vcls = AL(int, nv); s->bv1 = AL(int, nb); s->bv2 = AL(int, nb); s->bax = AL(int, nb); s->bcls = AL(int, nb);
    s->link = (int(*)[6])calloc((size_t)(nv > 0 ? nv : 1), sizeof(int[6]));
    s->nom = AL(v3, nv); s->pos = AL(v3, nv); s->linmom = AL(v3, nv); s->angmom = AL(v3, nv); s->ang = AL(q4, nv);
    s->scale = AL(double, nv); s->last_scale = AL(double, nv); s->sfric = AL(int, nv);
    s->small = AL(int, nb); s->lp2 = AL(v3, nb); s->la1 = AL(v3, nb); s->la2 = AL(v3, nb);
    s->f1 = AL(v3, nb); s->f2 = AL(v3, nb); s->m1 = AL(v3, nb); s->m2 = AL(v3, nb); s->strain = AL(double, nb);
    for (int i = 0; i < nv; i++) {
        s->vcls[i] = d->vox_class[i];
        s->nom[i] = V(d->nom_pos[3 * i], d->nom_pos[3 * i + 1], d->nom_pos[3 * i + 2]);
        for (int k = 0; k < 6; k++) s->link[i][k] = -1;
        /* CVXS_Voxel::ResetVoxel  VXS_Voxel.cpp:95-135 */
        s->pos[i] = s->nom[i];
        s->ang[i].w = 1.0;
        s->scale[i] = s->last_scale[i] = s->vc[s->vcls[i]].nom_size;
    }
    for (int j = 0; j < nb; j++) {
        s->bv1[j] = d->bond_v1[j]; s->bv2[j] = d->bond_v2[j]; s->bax[j] = d->bond_axis[j]; s->bcls[j] = d->bond_class[j];
        /* CreatePermBond VX_Sim.cpp:738-748: negative-side voxel gets BD_P*, positive-side BD_N* */
        s->link[s->bv1[j]][2 * s->bax[j]] = j;
        s->link[s->bv2[j]][2 * s->bax[j] + 1] = j;
        s->small[j] = 1; /* ResetBond VXS_BondInternal.cpp:43 */
    }
    /* Vec3D.h:57-60 */
    s->small_w = cos(SMALL_ANGLE_RAD * 0.5);
    s->smallish_w = cos(HYST * SMALL_ANGLE_RAD * 0.5);
    s->slthresh = 1.0 - W_THRESH_ACOS2SQRT * W_THRESH_ACOS2SQRT;
    s->cur_time = 0; s->dt = 0; s->step = 0; /* ClearAll VX_Sim.cpp:373, ResetSimulation :853-854 */
    s->col_auto = 1; s->col_first = 1;         /* ColEnableChanged = true  VX_Sim.cpp:50 */
    build_topology(s);
    return s;
}


/* ---- lattice topology: CalcNearby lists, surface voxels, the land/water surface mesh ------------------ */
static int find_vox(const vxo_sim* s, int x, int y, int z) {
    for (int i = 0; i < s->n_vox; i++) if (s->ijk[3 * i] == x && s->ijk[3 * i + 1] == y && s->ijk[3 * i + 2] == z) return i;
    return -1;
}
static void build_topology(vxo_sim* s) {
    int nv = s->n_vox;
    /* lattice coordinates from the nominal positions (cell centres at (i + 0.5) * pitch, VX_Object.cpp:543) */
    double pitch[3] = {0, 0, 0};
    for (int j = 0; j < s->n_bond; j++) { double p = s->bc[s->bcls[j]].nom_dist; int a = s->bax[j]; if (p > 0 && (pitch[a] == 0 || p < pitch[a])) pitch[a] = p; }
    for (int a = 0; a < 3; a++) if (!(pitch[a] > 0)) pitch[a] = s->vc[s->vcls[0]].nom_size;
    s->ijk = AL(int, 3 * nv);
    for (int i = 0; i < nv; i++) {
        s->ijk[3 * i] = (int)floor(s->nom[i].x / pitch[0]); s->ijk[3 * i + 1] = (int)floor(s->nom[i].y / pitch[1]); s->ijk[3 * i + 2] = (int)floor(s->nom[i].z / pitch[2]);
    }
    /* SurfVoxels (VX_Sim.cpp:675-680, IsSurfaceVoxel VX_Voxel.h:80) and CalcNearby(NumHops = (int)(CollisionHorizon*1.5)) */
    s->surf = AL(int, nv); s->n_surf = 0;
    for (int i = 0; i < nv; i++) { int n = 0; for (int k = 0; k < 6; k++) n += s->link[i][k] >= 0; if (n < 6) s->surf[s->n_surf++] = i; }
    int hops = (int)(s->sim.col_horizon * 1.5);
    s->near_off = AL(int, nv + 1);
    int cap = 64, used = 0;
    s->near_idx = AL(int, cap);
    int* list = AL(int, nv);
    for (int i = 0; i < nv; i++) {                                       /* VX_Voxel.cpp:171-209 */
        int n = 0, start = 0, stop = 1;
        list[n++] = i;
        for (int h = 0; h < hops; h++) {
            for (int j = start; j < stop; j++) {
                int v = list[j];
                for (int k = 0; k < 6; k++) {
                    int b = s->link[v][k];
                    if (b < 0) continue;
                    int o = (k % 2) ? s->bv1[b] : s->bv2[b];
                    int seen = 0; for (int q = 0; q < n; q++) if (list[q] == o) { seen = 1; break; }
                    if (!seen) list[n++] = o;
                }
            }
            start = stop; stop = n;
        }
        while (used + n > cap) { cap *= 2; s->near_idx = (int*)realloc(s->near_idx, sizeof(int) * (size_t)cap); }
        s->near_off[i] = used;
        memcpy(s->near_idx + used, list, sizeof(int) * (size_t)n);
        used += n;
    }
    s->near_off[nv] = used;
    free(list);
    /* fluid state */
    s->sp = AL(v3, nv); s->sn = AL(v3, nv); s->cpos = AL(v3, nv); s->cneg = AL(v3, nv); s->drag = AL(v3, nv);
    for (int i = 0; i < nv; i++) { double h = s->scale[i] / 2; s->cpos[i] = V(h, h, h); s->cneg[i] = V(-h, -h, -h); }   /* LW ResetVoxel :106-107 */
    if (!(s->sim.flags & VXG_F_FLUID)) return;
    /* CVX_MeshUtil::LinkSimVoxels (LW/VX_MeshUtil.cpp:110-276) */
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int i = 0; i < nv; i++) for (int a = 0; a < 3; a++) { int c = s->ijk[3 * i + a]; if (c < lo[a]) lo[a] = c; if (c > hi[a]) hi[a] = c; }
    int tx = hi[0] - lo[0] + 2, ty = hi[1] - lo[1] + 2, tz = hi[2] - lo[2] + 2;
    int ngrid = tx * ty * tz;
    int* cnt = AL(int, ngrid);
    for (int i = 0; i < nv; i++) for (int c = 0; c < 8; c++) {
        int gx = s->ijk[3 * i] - lo[0] + ((c >> 2) & 1), gy = s->ijk[3 * i + 1] - lo[1] + ((c >> 1) & 1), gz = s->ijk[3 * i + 2] - lo[2] + (c & 1);
        cnt[(gz * ty + gy) * tx + gx]++;
    }
    int* map = AL(int, ngrid);
    s->n_vert = 0;
    for (int g = 0; g < ngrid; g++) map[g] = (cnt[g] != 0 && cnt[g] != 8) ? s->n_vert++ : -1;   /* :249-262 */
    s->vert_off = AL(int, s->n_vert + 1);
    int tot = 0;
    for (int g = 0; g < ngrid; g++) if (map[g] >= 0) { s->vert_off[map[g]] = tot; tot += cnt[g]; }
    s->vert_off[s->n_vert] = tot;
    s->vert_vox = AL(int, tot); s->vert_corner = AL(int, tot);
    int* fill = AL(int, s->n_vert);
    for (int i = 0; i < nv; i++) for (int c = 0; c < 8; c++) {             /* ConVoxels in simulation-index order :131-144 */
        int gx = s->ijk[3 * i] - lo[0] + ((c >> 2) & 1), gy = s->ijk[3 * i + 1] - lo[1] + ((c >> 1) & 1), gz = s->ijk[3 * i + 2] - lo[2] + (c & 1);
        int u = map[(gz * ty + gy) * tx + gx];
        if (u < 0) continue;
        int at = s->vert_off[u] + fill[u]++;
        s->vert_vox[at] = i; s->vert_corner[at] = c;
    }
    /* facets: per voxel, open faces +X,-X,+Y,-Y,+Z,-Z, two triangles each, vertex order of :161-190 */
    static const int T1[6][3] = {{4, 6, 7}, {0, 3, 2}, {2, 3, 7}, {0, 5, 1}, {1, 5, 7}, {0, 6, 4}};
    static const int T2[6][3] = {{4, 7, 5}, {0, 1, 3}, {2, 7, 6}, {0, 4, 5}, {1, 7, 3}, {0, 2, 6}};
    static const int D[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    s->facet_v = AL(int, 3 * 12 * nv); s->facet_owner = AL(int, 12 * nv); s->n_facet = 0;
    for (int i = 0; i < nv; i++) for (int f = 0; f < 6; f++) {
        if (find_vox(s, s->ijk[3 * i] + D[f][0], s->ijk[3 * i + 1] + D[f][1], s->ijk[3 * i + 2] + D[f][2]) >= 0) continue;
        for (int t = 0; t < 2; t++) {
            const int* tri = t ? T2[f] : T1[f];
            for (int k = 0; k < 3; k++) {
                int c = tri[k];
                int gx = s->ijk[3 * i] - lo[0] + ((c >> 2) & 1), gy = s->ijk[3 * i + 1] - lo[1] + ((c >> 1) & 1), gz = s->ijk[3 * i + 2] - lo[2] + (c & 1);
                s->facet_v[3 * s->n_facet + k] = map[(gz * ty + gy) * tx + gx];
            }
            s->facet_owner[s->n_facet++] = i;
        }
    }
    s->vpos = AL(v3, s->n_vert);
    free(cnt); free(map); free(fill);
}

/* CQuat::RotateVec3D  Vec3D.h:293-299 */
static v3 qrot(q4 q, v3 f) {
    double fx = f.x, fy = f.y, fz = f.z;
    double tw = fx * q.x + fy * q.y + fz * q.z;
    double tx = fx * q.w - fy * q.z + fz * q.y;
    double ty = fx * q.z + fy * q.w - fz * q.x;
    double tz = -fx * q.y + fy * q.x + fz * q.w;
    return V(q.w * tx + q.x * tw + q.y * tz - q.z * ty, q.w * ty - q.x * tz + q.y * tw + q.z * tx, q.w * tz + q.x * ty - q.y * tx + q.z * tw);
}

/* the fluid block of LW CVX_Sim::Integrate (LW/VX_Sim.cpp:1516-1597) */
static void fluid_drag(vxo_sim* s) {
    for (int i = 0; i < s->n_vox; i++) s->drag[i] = V(0, 0, 0);
    for (int u = 0; u < s->n_vert; u++) {                                  /* UpdateMeshPhysicsOnlyNoColors -> GetCurVLoc :388-428 */
        v3 avg = V(0, 0, 0); double totw = 0;
        for (int k = s->vert_off[u]; k < s->vert_off[u + 1]; k++) {
            int i = s->vert_vox[k], c = s->vert_corner[k];
            v3 off = V((c & 4) ? s->cpos[i].x : s->cneg[i].x, (c & 2) ? s->cpos[i].y : s->cneg[i].y, (c & 1) ? s->cpos[i].z : s->cneg[i].z);
            v3 p = vadd(s->pos[i], qrot(s->ang[i], off));
            avg = vadd(avg, vscale(p, 1.0)); totw += 1.0;
        }
        s->vpos[u] = vdiv(avg, totw);
    }
    for (int f = 0; f < s->n_facet; f++) {
        v3 A = s->vpos[s->facet_v[3 * f]], B = s->vpos[s->facet_v[3 * f + 1]], C = s->vpos[s->facet_v[3 * f + 2]];
        v3 AB = vsub(B, A), AC = vsub(C, A);
        v3 cr = V(AB.y * AC.z - AB.z * AC.y, AB.z * AC.x - AB.x * AC.z, AB.x * AC.y - AB.y * AC.x);
        double crl = sqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
        double area = fabs(crl / 2.0);
        v3 n = crl > 0 ? vdiv(cr, crl) : cr;                               /* CalcFaceNormals: Normalized() */
        int i = s->facet_owner[f];
        v3 vel = vscale(s->linmom[i], s->vc[s->vcls[i]].mass_inv);         /* GetCurVel() */
        double vl = sqrt(vel.x * vel.x + vel.y * vel.y + vel.z * vel.z);
        v3 vn = vl > 0 ? vdiv(vel, vl) : vel;
        double nl = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
        v3 nn = nl > 0 ? vdiv(n, nl) : n;
        float ang = (float)acos(vn.x * nn.x + vn.y * nn.y + vn.z * nn.z);  /* `float speedNormalAngle` :1563 */
        v3 proj = V(0, 0, 0);
        if (fabs(ang) < PI_REF / 2) proj = vscale(nn, vel.x * n.x + vel.y * n.y + vel.z * n.z);   /* ProjectOnTo Vec3D.h:142 */
        double pl2 = proj.x * proj.x + proj.y * proj.y + proj.z * proj.z, pl = sqrt(pl2);
        v3 pn = pl > 0 ? vdiv(proj, pl) : proj;
        v3 d = vscale(pn, -(double)(float)s->sim.drag_coeff * area * pl2);  /* `float aggregateDragCoefficient` */
        s->drag[i] = vadd(s->drag[i], d);
    }
}

/* CVX_Sim::UpdateCollisions + CalcL1Bonds (VX_Sim.cpp:1729-1755, :2357-2413) */
static void update_collisions(vxo_sim* s) {
    double H = s->sim.col_horizon, lat = s->sim.lattice_dim;
    s->col_acc += fabs(s->max_vel * s->dt / lat);
    if (!(s->col_acc > (H - 1.0) / 2 || s->col_first)) return;
    s->col_first = 0; s->col_acc = 0.0;
    double filt = H * 1.5 * lat, filt2 = filt * filt;
    int cap = 64, n = 0;
    int* v1 = AL(int, cap); int* v2 = AL(int, cap); double* a1 = AL(double, cap);
    for (int i = 0; i < s->n_surf; i++) {
        int s1 = s->surf[i];
        for (int j = i + 1; j < s->n_surf; j++) {
            int s2 = s->surf[j];
            v3 d = vsub(s->pos[s1], s->pos[s2]);
            double d2 = d.x * d.x + d.y * d.y + d.z * d.z;
            if (!(d2 < filt2)) continue;
            int nearby = 0;
            for (int q = s->near_off[s1]; q < s->near_off[s1 + 1]; q++) if (s->near_idx[q] == s2) { nearby = 1; break; }
            if (nearby) continue;
            double act = H * (s->scale[s1] + s->scale[s1]) * 0.5;          /* voxel 1's scale twice :2382 */
            if (!(d2 < act * act)) continue;
            /* CreateColBond -> CVX_Bond::LinkVoxels / UpdateConstants: a1 = E A / L */
            const vxg_vox_class* c1 = &s->vc[s->vcls[s1]]; const vxg_vox_class* c2 = &s->vc[s->vcls[s2]];
            double E1 = c1->e_mod, E2 = c2->e_mod;
            if (E1 == 0 || E2 == 0) continue;
            double E = (E1 * E2 / (E1 + E2)) * 2, L = (c1->nom_size + c2->nom_size) * 0.5;
            if (n == cap) { cap *= 2; v1 = (int*)realloc(v1, sizeof(int) * (size_t)cap); v2 = (int*)realloc(v2, sizeof(int) * (size_t)cap); a1 = (double*)realloc(a1, sizeof(double) * (size_t)cap); }
            v1[n] = s1; v2[n] = s2; a1[n] = E * (L * L) / L; n++;
        }
    }
    free(s->cv1); free(s->cv2); free(s->ca1); free(s->cf1); free(s->cf2);
    s->n_col = n; s->cv1 = v1; s->cv2 = v2; s->ca1 = a1; s->cf1 = AL(v3, n); s->cf2 = AL(v3, n);
}

void vxo_destroy(vxo_sim* s) {
    if (!s) return;
    free(s->vc); free(s->bc); free(s->vcls); free(s->bv1); free(s->bv2); free(s->bax); free(s->bcls); free(s->link);
    free(s->nom); free(s->pos); free(s->linmom); free(s->angmom); free(s->ang); free(s->scale); free(s->last_scale);
    free(s->sfric); free(s->small); free(s->lp2); free(s->la1); free(s->la2); free(s->f1); free(s->f2); free(s->m1);
    free(s->m2); free(s->strain); free(s->cv1); free(s->cv2); free(s->ca1); free(s->cf1); free(s->cf2);
    free(s->ijk); free(s->near_off); free(s->near_idx); free(s->surf); free(s->sp); free(s->sn); free(s->cpos); free(s->cneg);
    free(s->drag); free(s->vert_off); free(s->vert_vox); free(s->vert_corner); free(s->facet_v); free(s->facet_owner); free(s->vpos);
    free(s);
}

void vxo_set_collision_pairs(vxo_sim* s, int n, const int32_t* v1, const int32_t* v2, const double* a1) {
    free(s->cv1); free(s->cv2); free(s->ca1); free(s->cf1); free(s->cf2);
    s->n_col = n; s->col_auto = 0;   /* host-managed list: no rebuild here */
    s->cv1 = AL(int, n); s->cv2 = AL(int, n); s->ca1 = AL(double, n); s->cf1 = AL(v3, n); s->cf2 = AL(v3, n);
    for (int i = 0; i < n; i++) { s->cv1[i] = v1[i]; s->cv2[i] = v2[i]; s->ca1[i] = a1[i]; }
}

/* CVXS_BondInternal::CalcLinForce (+UpdateBondStrain, AddDampForces) */
static void bond_update(vxo_sim* s, int j) {
    const vxg_bond_class* c = &s->bc[s->bcls[j]];
    int ax = s->bax[j], i1 = s->bv1[j], i2 = s->bv2[j];
    v3 rel = to_xdir(ax, vsub(s->pos[i2], s->pos[i1]));                 /* :58, :61 */
    q4 a1 = to_xdir_q(ax, s->ang[i1]), a2 = to_xdir_q(ax, s->ang[i2]);  /* :59-63 */
    v3 r = qrot_inv(a1, rel);                                           /* :65 */
    q4 na2 = qmul(qconj(a1), a2);                                       /* :66 */
    double nom = (s->scale[i1] + s->scale[i2]) * 0.5;                   /* :71 */

    int changed = 0;                                                    /* :74-78 */
    double small_turn = (fabs(r.z) + fabs(r.y)) / r.x;
    double ext = r.x / nom;
    if (!s->small[j] && na2.w > s->small_w && small_turn < SA_BOND_BEND_RAD && ext < SA_BOND_EXT_PERC) { s->small[j] = 1; changed = 1; }
    else if (s->small[j] && (!(na2.w > s->smallish_w) || small_turn > HYST * SA_BOND_BEND_RAD || ext > HYST * SA_BOND_EXT_PERC)) { s->small[j] = 0; changed = 1; }

    q4 tot; v3 p2, an1, an2;
    if (s->small[j]) {                                                  /* :87-95 */
        an1 = V(0, 0, 0);
        an2 = q_to_rotvec(s, na2);
        r.x -= nom;
        p2 = r;
        tot = qconj(a1);
    } else {                                                            /* :101-118 */
        q4 pa = q_from_angle_to_posx(r);
        tot = qmul(pa, qconj(a1));
        double len = sqrt(rel.x * rel.x + rel.y * rel.y + rel.z * rel.z);
        p2 = V(len - nom, 0, 0);
        an1 = q_to_rotvec(s, pa);
        an2 = q_to_rotvec(s, qmul(tot, a2));
    }

    /* UpdateBondStrain(_Pos2.x/L.x)  :120, :189-307 with MDL_LINEAR stress (VX_Object.cpp:1472-1482) */
    double strain = p2.x / c->len, stress;
    s->strain[j] = strain;
    double sv1 = strain, sv2 = strain;                                  /* CurStrainV1 = CurStrainV2 = CurStrainIn :208 */
    if (c->homogeneous) stress = c->e1 * strain;                        /* :206 */
    else {                                                              /* :226-258 */
        double st1 = c->e1 * sv1, st2 = c->e2 * sv2;
        int count = 0;
        double diff = (st1 >= st2) ? st1 - st2 : st2 - st1;
        double sum = st1 + st2; if (sum < 0) sum = -sum;
        while (diff > sum * .0005 && count < 3) {
            sv1 = 2 * st2 / (st1 + st2) * sv1;
            sv2 = 2 * st1 / (st1 + st2) * sv2;
            st1 = c->e1 * sv1; st2 = c->e2 * sv2;
            diff = (st1 >= st2) ? st1 - st2 : st2 - st1;
            sum = st1 + st2; if (sum < 0) sum = -sum;
            count++;
        }
        stress = (st1 + st2) / 2;
    }
    {   /* SetStrainDir(BD_P<axis>, CurStrainV1) / (BD_N<axis>, CurStrainV2)  :300-304 */
        int v1 = s->bv1[j], v2 = s->bv2[j], a = s->bax[j];
        if (a == 0) { s->sp[v1].x = sv1; s->sn[v2].x = sv2; } else if (a == 1) { s->sp[v1].y = sv1; s->sn[v2].y = sv2; } else { s->sp[v1].z = sv1; s->sn[v2].z = sv2; }
    }

    /* beam equations :127-128, :153-154  (CSArea1 == CSArea2 == s^2: VXS_Voxel.cpp:625-631, VXS_Bond.cpp:72) */
    v3 F1 = V(stress * (c->area + c->area) / 2, c->b1 * p2.y - c->b2 * (an1.z + an2.z), c->b1 * p2.z + c->b2 * (an1.y + an2.y));
    v3 F2 = vneg(F1);
    v3 M1 = V(c->a2 * (an1.x - an2.x), c->b2 * p2.z + c->b3 * (2 * an1.y + an2.y), -c->b2 * p2.y + c->b3 * (2 * an1.z + an2.z));
    v3 M2 = V(c->a2 * (an2.x - an1.x), c->b2 * p2.z + c->b3 * (an1.y + 2 * an2.y), -c->b2 * p2.y + c->b3 * (an1.z + 2 * an2.z));

    if (!changed) {                                                     /* :157, AddDampForces :310-346 */
        if (s->dt != 0) {
            double bz = 0.5 * s->sim.bond_damp_z;
            double dti = 1.0 / s->dt;
            v3 rv2 = vscale(vsub(p2, s->lp2[j]), dti);
            v3 ra1 = vscale(vsub(an1, s->la1[j]), dti);
            v3 ra2 = vscale(vsub(an2, s->la2[j]), dti);
            F1 = vadd(F1, vscale(V(c->c_a1m1 * rv2.x, c->c_b1m1 * rv2.y - c->c_b2fm1 * (ra1.z + ra2.z),
                                   c->c_b1m1 * rv2.z + c->c_b2fm1 * (ra1.y + ra2.y)), bz));
            if (!c->homogeneous)
                F2 = vadd(F2, vscale(V(-c->c_a1m2 * rv2.x, -c->c_b1m2 * rv2.y + c->c_b2fm2 * (ra1.z + ra2.z),
                                       -c->c_b1m2 * rv2.z - c->c_b2fm2 * (ra1.y + ra2.y)), bz));
            M1 = vadd(M1, vscale(V(-c->c_a2i1 * (ra2.x - ra1.x), c->c_b2fm1 * rv2.z + c->c_b3i1 * (2 * ra1.y + ra2.y),
                                   -c->c_b2fm1 * rv2.y + c->c_b3i1 * (2 * ra1.z + ra2.z)), 0.5 * bz));
            M2 = vadd(M2, vscale(V(c->c_a2i2 * (ra2.x - ra1.x), c->c_b2fm2 * rv2.z + c->c_b3i2 * (ra1.y + 2 * ra2.y),
                                   -c->c_b2fm2 * rv2.y + c->c_b3i2 * (ra1.z + 2 * ra2.z)), 0.5 * bz));
        }
        s->lp2[j] = p2; s->la1[j] = an1; s->la2[j] = an2;
    }

    /* back to world frame :161-170 */
    F1 = qrot_inv(tot, F1);
    if (!c->homogeneous) F2 = qrot_inv(tot, F2);
    M1 = qrot_inv(tot, M1);
    M2 = qrot_inv(tot, M2);
    F1 = to_origdir(ax, F1);
    if (c->homogeneous) F2 = vneg(F1); else F2 = to_origdir(ax, F2);
    M1 = to_origdir(ax, M1);
    M2 = to_origdir(ax, M2);
    s->f1[j] = F1; s->f2[j] = F2; s->m1[j] = M1; s->m2[j] = M2;
}

/* CVXS_BondCollision::CalcContactForce  VXS_BondCollision.cpp:41-59 */
static void col_update(vxo_sim* s, int k) {
    v3 p2 = vsub(s->pos[s->cv2[k]], s->pos[s->cv1[k]]);
    double nomd = (s->scale[s->cv1[k]] + s->scale[s->cv2[k]]) * 0.75;
    double len = sqrt(p2.x * p2.x + p2.y * p2.y + p2.z * p2.z);
    double rel = nomd - len;
    if (rel > 0) {
        s->cf2[k] = vscale(vscale(vdiv(p2, len), s->ca1[k]), rel);
        s->cf1[k] = vneg(s->cf2[k]);
    } else { s->cf2[k] = V(0, 0, 0); s->cf1[k] = V(0, 0, 0); }
}

/* CVXS_Voxel::EulerStep  VXS_Voxel.cpp:169-427 */
static void voxel_step(vxo_sim* s, int i) {
    const vxg_vox_class* c = &s->vc[s->vcls[i]];
    const vxg_sim_params* P = &s->sim;
    double dt = s->dt;
    int dof = c->dof_fixed;
    if ((dof & VXG_DOF_ALL) == VXG_DOF_ALL) {                           /* :174-179 */
        s->linmom[i] = V(0, 0, 0);
        s->pos[i] = vadd(s->nom[i], V(c->ext_disp[0], c->ext_disp[1], c->ext_disp[2]));
        s->angmom[i] = V(0, 0, 0);
        { /* FromRotationVector Vec3D.h:203-207 */
            v3 axv = V(c->ext_tdisp[0], c->ext_tdisp[1], c->ext_tdisp[2]);
            double l = sqrt(axv.x * axv.x + axv.y * axv.y + axv.z * axv.z);
            if (l > 0) { axv.x /= l; axv.y /= l; axv.z /= l; }
            double a = l * 0.5, sn = sin(a), cs = cos(a);
            s->ang[i].w = cs; s->ang[i].x = axv.x * sn; s->ang[i].y = axv.y * sn; s->ang[i].z = axv.z * sn;
        }
    } else {
        /* CalcTotalForce :482-651 */
        v3 vel = vscale(s->linmom[i], c->mass_inv);                     /* Vel of the previous step, :415 */
        v3 F = V(0, 0, 0);
        F = vadd(F, vscale(vscale(vel, -P->slow_damp_z), c->damp_lin)); /* :487 */
        for (int k = 0; k < 6; k++) {                                   /* :496-501 */
            int j = s->link[i][k];
            if (j < 0) continue;
            F = vadd(F, (k % 2) ? s->f2[j] : s->f1[j]);
        }
        if (P->flags & VXG_F_COLLISIONS)                                 /* :516-528 */
            for (int k = 0; k < s->n_col; k++) {
                if (s->cv2[k] == i) F = vadd(F, s->cf2[k]);
                else if (s->cv1[k] == i) F = vadd(F, s->cf1[k]);
            }
        F = vsub(F, V(0, 0, 0));                                        /* InputForce :531 */
        if (!(P->flags & VXG_F_FLUID)) F.z += c->mass * P->grav_acc;    /* :536 (grav_acc == 0 when disabled) */
        F = vadd(F, V(c->ext_force[0], c->ext_force[1], c->ext_force[2])); /* :540 */
        if (P->flags & VXG_F_FLUID) F = vadd(F, s->drag[i]);            /* LW/VXS_Voxel.cpp:370-371 */
        if ((P->flags & VXG_F_FLOOR) && !(P->flags & VXG_F_FLUID)) {    /* :635-639, CalcFloorEffect :708-758 */
            v3 FF = V(0, 0, 0);
            s->sfric[i] = 0;
            double pen = 0.5 * s->scale[i] - s->pos[i].z;               /* :698 */
            pen = pen <= 0 ? 0 : pen;
            if (pen > 0) {
                double normal = c->k_floor * pen;
                FF.z += normal;
                FF.z -= P->col_damp_z * c->damp_lin * vel.z;
                double sv = sqrt(vel.x * vel.x + vel.y * vel.y);
                double sva = atan2(vel.y, vel.x);
                double sf = sqrt(F.x * F.x + F.y * F.y);
                double dfr = c->mu_dynamic * normal;
                v3 fric = vneg(V(cos(sva) * dfr, sin(sva) * dfr, 0));
                if (P->flags & VXG_F_STICKY_FLOOR) { vel.x = 0; vel.y = 0; s->linmom[i].x = 0; s->linmom[i].y = 0; s->sfric[i] = 1; }
                if (vel.x == 0 && vel.y == 0) {
                    if (sf < c->mu_static * normal) s->sfric[i] = 1;
                } else {
                    if (dfr * dt < c->mass * sv) FF = vadd(FF, fric);
                    else { s->sfric[i] = 1; s->linmom[i].x = 0; s->linmom[i].y = 0; }
                }
            }
            F = vadd(F, FF);
            if (s->sfric[i]) { F.x = 0; F.y = 0; }
        }
        {   /* corner extents for the surface mesh, LW/VXS_Voxel.cpp:474-475 */
            double hs = c->nom_size / 2;
            s->cpos[i] = V((1 + s->sp[i].x) * c->nom_size / 2, (1 + s->sp[i].y) * c->nom_size / 2, (1 + s->sp[i].z) * c->nom_size / 2);
            s->cneg[i] = V(-(1 + s->sn[i].x) * c->nom_size / 2, -(1 + s->sn[i].y) * c->nom_size / 2, -(1 + s->sn[i].z) * c->nom_size / 2);
            (void)hs;
        }
        /* :184-197 */
        s->linmom[i] = vadd(s->linmom[i], vscale(F, dt));
        v3 disp = vscale(s->linmom[i], dt * c->mass_inv);
        s->pos[i] = vadd(s->pos[i], disp);
        if (dof & VXG_DOF_X) { s->pos[i].x = s->nom[i].x + c->ext_disp[0]; s->linmom[i].x = 0; }
        if (dof & VXG_DOF_Y) { s->pos[i].y = s->nom[i].y + c->ext_disp[1]; s->linmom[i].y = 0; }
        if (dof & VXG_DOF_Z) { s->pos[i].z = s->nom[i].z + c->ext_disp[2]; s->linmom[i].z = 0; }
        /* CalcTotalMoment :653-675 */
        v3 M = V(0, 0, 0);
        for (int k = 0; k < 6; k++) {
            int j = s->link[i][k];
            if (j < 0) continue;
            M = vsub(M, (k % 2) ? s->m2[j] : s->m1[j]);
        }
        M = vadd(M, V(c->ext_torque[0], c->ext_torque[1], c->ext_torque[2]));
        if (dof & VXG_DOF_TX) M.x = 0;
        if (dof & VXG_DOF_TY) M.y = 0;
        if (dof & VXG_DOF_TZ) M.z = 0;
        s->angmom[i] = vadd(s->angmom[i], vscale(M, dt));               /* :202 */
        {
            double fact = (1 - 10 * P->slow_damp_z * c->inertia_inv * c->damp_ang * dt); /* :206 */
            s->angmom[i].x *= fact; s->angmom[i].y *= fact; s->angmom[i].z *= fact;
        }
        v3 w = vscale(s->angmom[i], c->inertia_inv);                    /* :211 */
        q4 hq = {0 * 0.5, w.x * 0.5, w.y * 0.5, w.z * 0.5};           /* 0.5 * CQuat(0,w): operator*(T f, CQuat v) Vec3D.h:191 */
        q4 spin = qmul(hq, s->ang[i]);                                  /* ... * Angle :212 */
        q4* A = &s->ang[i];
        A->w += spin.w * dt; A->x += spin.x * dt; A->y += spin.y * dt; A->z += spin.z * dt; /* :214 */
        {   /* NormalizeFast Vec3D.h:243-246 */
            double l = sqrt(A->x * A->x + A->y * A->y + A->z * A->z + A->w * A->w);
            if (l != 0) { double li = 1.0 / l; A->w *= li; A->x *= li; A->y *= li; A->z *= li; }
            if (A->w >= 1.0) { A->w = 1.0; A->x = 0; A->y = 0; A->z = 0; }
        }
        if ((dof & VXG_DOF_TX) && (dof & VXG_DOF_TY) && (dof & VXG_DOF_TZ)) { /* :218-221 */
            v3 axv = V(c->ext_tdisp[0], c->ext_tdisp[1], c->ext_tdisp[2]);
            double l = sqrt(axv.x * axv.x + axv.y * axv.y + axv.z * axv.z);
            if (l > 0) { axv.x /= l; axv.y /= l; axv.z /= l; }
            double a = l * 0.5, sn = sin(a), cs = cos(a);
            A->w = cs; A->x = axv.x * sn; A->y = axv.y * sn; A->z = axv.z * sn;
            s->angmom[i] = V(0, 0, 0);
        }
    }
    /* SCALE */
    if (P->flags & VXG_F_LW_SCALE) {                                    /* LW/VXS_Voxel.cpp:211-235 */
        double tf = 1.0;
        if ((P->flags & VXG_F_TEMPERATURE) && s->cur_time >= P->init_cm_time)
            tf = (1 + (P->temp_amplitude * sin(2 * 3.1415926f * (s->cur_time / P->temp_period + c->phase_offset))) * c->cte);
        if (tf < P->min_scale_fact) tf = P->min_scale_fact;
        s->scale[i] = tf * c->nom_size;
    } else {                                                            /* VXS_Voxel.cpp:224-336 (without the velocity-adjusted windows :340-381) */
        double max_scale = P->max_scale_fact * c->nom_size;             /* :225 */
        double min_scale = P->min_scale_fact * c->nom_size;             /* :226 */
        double t = s->cur_time;
        double ctrl = 0, dev = 0, dphase = 0, dtad = 0, k = 0, frozen_adj = 0, freeze_init = 1;
        /* prenatal ramp :237-238 (initialVoxelSize is a float member, so the ratio is not exactly 1) */
        double cc = (t >= 0.5 * P->init_cm_time) ? 1.0 : 2 * t / P->init_cm_time;
        double prenatal = cc * ((c->init_size / c->nom_size) - 1);
        if (P->midlife_freeze_time > 0) {                               /* :241-259 */
            double middle = 0.5 * (P->stop_condition_value - P->init_cm_time);
            double fstart = middle - 0.5 * P->midlife_freeze_time, fend = middle + 0.5 * P->midlife_freeze_time;
            if (t > fstart && t < fend) {
                frozen_adj = t - fstart;
                if (t < fstart + P->init_cm_time) freeze_init = 0;
            }
            if (t > fend) frozen_adj = P->midlife_freeze_time;
        }
        {   /* postnatal linear development :262-291; the members are floats upstream, so float op float stays float */
            float start_f = (float)c->start_growth_time, growth_f = (float)c->growth_time;
            if (t >= start_f && growth_f > 0) {
                double eff = (t <= start_f + growth_f + P->midlife_freeze_time) ? t : start_f + growth_f + P->midlife_freeze_time;
                eff = eff - frozen_adj;
                k = (eff - start_f) / growth_f;
                if (c->dev_flags & VXG_DEV_FINAL_SIZE) dev = k * (((float)c->final_size / (float)c->init_size) - 1.0);
                if (c->dev_flags & VXG_DEV_FINAL_PHASE) dphase = k * ((float)c->final_phase_offset - (float)c->phase_offset);
                if (c->dev_flags & VXG_DEV_FINAL_TAD) dtad = k * ((float)c->final_temp_amp_damp - (float)c->temp_amp_damp);
            }
        }
        if ((P->flags & VXG_F_TEMPERATURE) && t >= P->init_cm_time) {  /* :294-305 */
            double phase = c->phase_offset + dphase, tad = c->temp_amp_damp + dtad;
            ctrl = tad * (P->temp_amplitude * sin(2 * 3.1415926f * (t / P->temp_period + phase))) * c->cte * freeze_init;
        }
        if (c->dev_flags & VXG_DEV_SIZE_ADJUST) {                       /* :308-318 */
            double cur_size = (1 + prenatal) * (1 + dev) * c->nom_size;
            double original = (cur_size / c->nom_size - 1) / P->growth_amplitude;
            double positive = (original + 1) * 0.5;
            double capped = (positive > 0.5) ? 0.5 : positive;
            ctrl = ctrl * capped * 2;
        }
        double cur = ctrl * c->nom_size + (1 + prenatal) * (1 + dev) * c->nom_size;  /* :321 */
        if (cur < s->last_scale[i] && cur < min_scale) cur = s->last_scale[i];
        if (cur > s->last_scale[i] && cur > max_scale) cur = s->last_scale[i];
        s->scale[i] = cur;
        s->last_scale[i] = cur;
    }
}

int64_t vxo_step(vxo_sim* s, int64_t n) {
    int64_t k;
    for (k = 0; k < n; k++) {
        if (s->diverged) break;
        if ((s->sim.flags & VXG_F_COLLISIONS) && s->col_auto) update_collisions(s);   /* TimeStep: VX_Sim.cpp:1131-1134 */
        int div = 0;
        for (int j = 0; j < s->n_bond; j++) {                           /* VX_Sim.cpp:1773-1776 */
            bond_update(s, j);
            if (s->strain[j] > 100) div = 1;
        }
        if (div) { s->diverged = 1; break; }
        if (s->sim.flags & VXG_F_COLLISIONS) for (int c = 0; c < s->n_col; c++) col_update(s, c); /* :1786-1788 */
        s->dt = s->sim.dt;                                              /* :1793 */
        if (s->sim.flags & VXG_F_FLUID) fluid_drag(s);                  /* LW/VX_Sim.cpp:1516-1597 */
        for (int i = 0; i < s->n_vox; i++) voxel_step(s, i);           /* :1913 */
        {   /* SS.MaxVoxVel (UpdateStats VX_Sim.cpp:1626-1649) */
            double m2 = 0;
            for (int i = 0; i < s->n_vox; i++) { v3 v = vscale(s->linmom[i], s->vc[s->vcls[i]].mass_inv); double l2 = v.x * v.x + v.y * v.y + v.z * v.z; if (l2 > m2) m2 = l2; }
            s->max_vel = sqrt(m2);
        }
        s->cur_time += s->dt;                                           /* :1929 */
        s->step++;
    }
    return k;
}

int vxo_diverged(const vxo_sim* s) { return s->diverged; }
double vxo_time(const vxo_sim* s) { return s->cur_time; }
int64_t vxo_step_count(const vxo_sim* s) { return s->step; }

void vxo_get_state(const vxo_sim* s, vxg_robot_state* o) {
    for (int i = 0; i < s->n_vox; i++) {
        if (o->pos) { o->pos[3 * i] = s->pos[i].x; o->pos[3 * i + 1] = s->pos[i].y; o->pos[3 * i + 2] = s->pos[i].z; }
        if (o->quat) { o->quat[4 * i] = s->ang[i].w; o->quat[4 * i + 1] = s->ang[i].x; o->quat[4 * i + 2] = s->ang[i].y; o->quat[4 * i + 3] = s->ang[i].z; }
        if (o->lin_mom) { o->lin_mom[3 * i] = s->linmom[i].x; o->lin_mom[3 * i + 1] = s->linmom[i].y; o->lin_mom[3 * i + 2] = s->linmom[i].z; }
        if (o->ang_mom) { o->ang_mom[3 * i] = s->angmom[i].x; o->ang_mom[3 * i + 1] = s->angmom[i].y; o->ang_mom[3 * i + 2] = s->angmom[i].z; }
        if (o->scale) o->scale[i] = s->scale[i];
        if (o->static_fric) o->static_fric[i] = s->sfric[i];
    }
    for (int j = 0; j < s->n_bond; j++) {
        if (o->bond_force1) { o->bond_force1[3 * j] = s->f1[j].x; o->bond_force1[3 * j + 1] = s->f1[j].y; o->bond_force1[3 * j + 2] = s->f1[j].z; }
        if (o->bond_force2) { o->bond_force2[3 * j] = s->f2[j].x; o->bond_force2[3 * j + 1] = s->f2[j].y; o->bond_force2[3 * j + 2] = s->f2[j].z; }
        if (o->bond_moment1) { o->bond_moment1[3 * j] = s->m1[j].x; o->bond_moment1[3 * j + 1] = s->m1[j].y; o->bond_moment1[3 * j + 2] = s->m1[j].z; }
        if (o->bond_moment2) { o->bond_moment2[3 * j] = s->m2[j].x; o->bond_moment2[3 * j + 1] = s->m2[j].y; o->bond_moment2[3 * j + 2] = s->m2[j].z; }
        if (o->bond_strain) o->bond_strain[j] = s->strain[j];
        if (o->bond_small_angle) o->bond_small_angle[j] = s->small[j];
    }
}

/* CVX_Sim::GetCM  VX_Sim.cpp:2415-2430 */
void vxo_get_com(const vxo_sim* s, double* com) {
    double tm = 0; v3 sum = V(0, 0, 0);
    for (int i = 0; i < s->n_vox; i++) {
        double m = s->vc[s->vcls[i]].mass;
        sum = vadd(sum, vscale(s->pos[i], m));
        tm += m;
    }
    sum = vdiv(sum, tm);
    com[0] = sum.x; com[1] = sum.y; com[2] = sum.z;
}

/* CVX_Sim::GetNumTouchingFloor  VX_Sim.cpp:2432-2441 */
int vxo_num_touching_floor(const vxo_sim* s) {
    int n = 0;
    for (int i = 0; i < s->n_vox; i++) if (0.5 * s->scale[i] - s->pos[i].z > 0) n++;
    return n;
}

double vxo_max_vel(const vxo_sim* s) {
    double mx = 0;
    for (int i = 0; i < s->n_vox; i++) {
        v3 v = vscale(s->linmom[i], s->vc[s->vcls[i]].mass_inv);
        double l = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        if (l > mx) mx = l;
    }
    return mx;
}
