/* vxgpu.h -- C-ABI of the B200-native Voxelyze timestep engine (libvxgpu.so).
 *
 * This is the drop-in boundary for the one hot path the engine replaces:
 *     CVX_Sim::Integrate()                       evosoro/_voxcad/Voxelyze/VX_Sim.cpp:1763-1933
 *       -> CVXS_BondInternal::UpdateBond()       evosoro/_voxcad/Voxelyze/VXS_BondInternal.cpp:35-346
 *       -> CVXS_Voxel::EulerStep()               evosoro/_voxcad/Voxelyze/VXS_Voxel.cpp:169-427, 482-758
 * batched over a population of independent robots.  The reference has no FFI of its own (SURVEY.md
 * section 8b); these are the entry points a `CVX_Sim` adapter binds: the host keeps
 * parsing the .vxa, building voxel/bond lists and computing the import-time constants
 * (CVX_Bond::UpdateConstants VX_Bond.cpp:123-173, CVX_Voxel::SetMaterial VX_Voxel.cpp:94-128) and hands
 * them over as a flat descriptor; the library owns all device memory behind an opaque handle.
 *
 * Plain C, no torch / CUDA types in any signature.  Return value: 0 = ok, negative = argument / CUDA
 * error, positive = the model uses a feature the kernels do not implement (the caller must abort --
 * there is NO CPU fallback).  vxg_last_error() gives a thread-local message for the last failure.
 * Any thread may call; one handle is not re-entrant; different handles may be driven concurrently.
 */
#ifndef VXGPU_H
#define VXGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXG_ABI_VERSION 5   /* 5: vxg_set_pooling, robots up to 32 767 voxels, negative vxg_pop_options.col_pairs_per_vox; 4: vxg_pop_* (ragged populations); 3: development fields, vxg_robot_state.vox_strain, vxg_robot_stats.mesh_volume */

/* ---- return codes ------------------------------------------------------------------------------ */
enum {
    VXG_OK = 0,
    VXG_ERR_ARG = -1,        /* bad argument / descriptor inconsistent                               */
    VXG_ERR_CUDA = -2,       /* CUDA runtime error (message has the cudaError string)                */
    VXG_ERR_NOMEM = -3,
    VXG_ERR_STATE = -4,      /* call order (e.g. step before all robots are uploaded)                */
    VXG_ERR_NO_DEVICE = -5,  /* no CUDA device: the engine never falls back to the CPU               */
    VXG_UNSUPPORTED_FEATURE = 1, /* descriptor asks for something outside the implemented path      */
    VXG_UNSUPPORTED_SIZE = 2     /* robot exceeds the limits given to vxg_create                     */
};

/* ---- feature flags (vxg_sim_params.flags) -------------------------------------------------------- */
#define VXG_F_FLOOR        0x0001u /* VXSFEAT_FLOOR        VXS_Voxel.cpp:635-639                       */
#define VXG_F_TEMPERATURE  0x0002u /* VXSFEAT_TEMPERATURE  VXS_Voxel.cpp:292-305 (thermal actuation)   */
#define VXG_F_LW_SCALE     0x0004u /* _voxcad_land_water scale law, LW/VXS_Voxel.cpp:211-235           */
#define VXG_F_STICKY_FLOOR 0x0008u /* getUsingStickyFloor() VXS_Voxel.cpp:736-739                      */
#define VXG_F_FLUID        0x0010u /* LW fluid environment: surface drag (LW/VX_Sim.cpp:1516-1597), gravity and floor inert (LW/VXS_Voxel.cpp:357,467) */
#define VXG_F_COLLISIONS   0x0020u /* VXSFEAT_COLLISIONS: self-collision contact bonds; the list is rebuilt on the device
                                      (UpdateCollisions / CalcL1Bonds, VX_Sim.cpp:1729-1755, :2357-2413) unless the host
                                      supplies it with vxg_set_collision_pairs; needs limits.max_col_pairs > 0          */

/* degrees of freedom, same bit values as the reference (VX_FRegion.h DOF_*) */
#define VXG_DOF_X  0x01
#define VXG_DOF_Y  0x02
#define VXG_DOF_Z  0x04
#define VXG_DOF_TX 0x08
#define VXG_DOF_TY 0x10
#define VXG_DOF_TZ 0x20
#define VXG_DOF_ALL 0x3F

/* ---- precision of the device arithmetic -------------------------------------------------------- */
#define VXG_PREC_F32 0  /* fp32 forces/moments/quaternions, fp64 positions and time (production)     */
#define VXG_PREC_F64 1  /* everything fp64 (logic-parity / high-accuracy mode; slower)               */

/* Per-voxel constants, de-duplicated into classes by the host (a class = one palette material in
 * evosoro-written files; per-voxel phase offsets or evolved stiffness simply make more classes).
 * All values are what the reference caches at import: VX_Voxel.cpp:109-125, VX_Voxel.h:64,
 * VX_Sim.cpp:878-991 (per-voxel actuation parameters). */
typedef struct vxg_vox_class {
    double mass;          /* rho s^3                                  VX_Voxel.cpp:110               */
    double mass_inv;      /* _massInv                                 VX_Voxel.cpp:121               */
    double inertia;       /* m s^2/6                                  VX_Voxel.cpp:111               */
    double inertia_inv;   /* _inertiaInv                              VX_Voxel.cpp:122               */
    double damp_lin;      /* _2xSqMxExS     = 2 sqrt(m E s)           VX_Voxel.cpp:123               */
    double damp_ang;      /* _2xSqIxExSxSxS = 2 sqrt(I E s^3)         VX_Voxel.cpp:124               */
    double k_floor;       /* GetLinearStiffness() = 2 E s             VX_Voxel.h:64                  */
    double mu_static;     /* uStatic                                  VXS_Voxel.cpp:717              */
    double mu_dynamic;    /* uDynamic                                 VXS_Voxel.cpp:716              */
    double nom_size;      /* NominalSize                              VX_Sim.cpp:527                 */
    double cte;           /* Vox_CTE                                  VXS_Voxel.cpp:295              */
    double temp_amp_damp; /* TempAmpDamp (1 unless evolved)           VX_Sim.cpp:885                 */
    double phase_offset;  /* phaseOffset (stored as float upstream)   VX_Sim.cpp:882                 */
    double e_mod;         /* Vox_E (information only)                                                */
    double init_size;     /* initialVoxelSize -- a `float` upstream (VXS_Voxel.h:100), so it differs  */
                          /* from nom_size by float rounding and feeds PreNatalTempFrac VXS_Voxel.cpp:237-238 */
    double ext_force[3];  /* ExternalForce                            VX_Sim.cpp:573                 */
    double ext_torque[3]; /* ExternalTorque                           VX_Sim.cpp:574                 */
    double ext_disp[3];   /* ExternalDisp  (fixed DOFs)               VX_Sim.cpp:576-578             */
    double ext_tdisp[3];  /* ExternalTDisp (fixed rot. DOFs)          VX_Sim.cpp:579-581             */
    int32_t dof_fixed;    /* VXG_DOF_* mask                           VX_Voxel.h:90                  */
    int32_t mat_index;    /* palette index (fitness probes key on 5/6/7, VX_Sim.cpp:2626)            */
    /* development (VX_Sim.cpp:878-975 sets them at reset, VXS_Voxel.cpp:240-318 reads them every step); all of them are
     * `float` members upstream (VXS_Voxel.h:92-100): the values here are the float-rounded ones.  Zeros + dev_flags 0 =
     * no development (every file evosoro's stock experiments write). */
    double final_size;            /* finalVoxelSize                    VX_Sim.cpp:908-924             */
    double growth_time;           /* growthTime                        VX_Sim.cpp:957-974             */
    double start_growth_time;     /* startGrowthTime                   VX_Sim.cpp:927-953             */
    double final_phase_offset;    /* finalPhaseOffset                  VX_Sim.cpp:883                 */
    double final_temp_amp_damp;   /* finalTempAmpDamp                  VX_Sim.cpp:886                 */
    int32_t dev_flags;            /* VXG_DEV_*                                                        */
    int32_t reserved;
} vxg_vox_class;

#define VXG_DEV_FINAL_SIZE  0x1  /* GetUsingFinalVoxelSize()   VXS_Voxel.cpp:272-275                          */
#define VXG_DEV_FINAL_PHASE 0x2  /* GetUsingFinalPhaseOffset() VXS_Voxel.cpp:278-284                          */
#define VXG_DEV_FINAL_TAD   0x4  /* GetUsingFinalTempAmpDamp() VXS_Voxel.cpp:286-289                          */
#define VXG_DEV_SIZE_ADJUST 0x8  /* GetUsingInitialVoxelSize() || GetUsingFinalVoxelSize(): actuation is scaled by the
                                    voxel's current size, VXS_Voxel.cpp:308-318                                 */

/* Per-bond constants, de-duplicated into classes (a class = an ordered material pair in
 * evosoro-written files).  VX_Bond.cpp:65-109 (LinkVoxels), :123-173 (UpdateConstants).
 * Voxels are cubes, so the y and z beam constants coincide (Iy == Iz). */
typedef struct vxg_bond_class {
    double a1;        /* E A / L        (axial; used by CalcMaxDt and contact bonds)                 */
    double a2;        /* G J / L        (torsion)                                                    */
    double b1;        /* 12 E I / L^3                                                                */
    double b2;        /* 6 E I / L^2                                                                 */
    double b3;        /* 2 E I / L                                                                   */
    double len;       /* L.x: nominal bond length used for strain      VXS_BondInternal.cpp:120      */
    double area;      /* (CSArea1 + CSArea2)/2 = s^2                   VXS_BondInternal.cpp:127      */
    double e1;        /* modulus the stress model uses for voxel 1     VXS_Voxel.cpp:831, VX_Object.cpp:1477 */
    double e2;        /* ... for voxel 2 (== e1 when homogeneous)                                    */
    double nom_dist;  /* nominal centre distance along the bond axis (|NominalPos2 - NominalPos1|)    */
    double c_a1m1, c_a1m2;   /* _2xSqA1xM1/2     VX_Bond.cpp:159                                     */
    double c_a2i1, c_a2i2;   /* _2xSqA2xI1/2     VX_Bond.cpp:160                                     */
    double c_b1m1, c_b1m2;   /* _2xSqB1{Y,Z}xM1/2  :161-162                                          */
    double c_b2fm1, c_b2fm2; /* _2xSqB2{Y,Z}xFM1/2 :163-164                                          */
    double c_b3i1, c_b3i2;   /* _2xSqB3{Y,Z}xI1/2  :169-170                                          */
    int32_t homogeneous;     /* HomogenousBond   VX_Bond.cpp:75                                      */
    int32_t reserved;
} vxg_bond_class;

/* Per-robot simulation constants: what CVX_Sim / CVX_Environment hold that the step reads. */
typedef struct vxg_sim_params {
    double dt;              /* DtFrac * OptimalDt, constant in scope   VX_Sim.cpp:1793, :1693-1727   */
    double bond_damp_z;     /* GetBondDampZ()       VXS_BondInternal.cpp:313                         */
    double col_damp_z;      /* GetCollisionDampZ()  VXS_Voxel.cpp:725                                */
    double slow_damp_z;     /* GetSlowDampZ()       VXS_Voxel.cpp:487, :206                          */
    double grav_acc;        /* GetGravityAccel(), 0 when VXSFEAT_GRAVITY is off  VXS_Voxel.cpp:536   */
    double init_cm_time;    /* GetInitCmTime()      VXS_Voxel.cpp:292                                */
    double temp_period;     /* TempPeriod (float upstream)  VXS_Voxel.cpp:302                        */
    double temp_amplitude;  /* TempAmplitude = TempAmp - TempBase (float upstream)                   */
    double min_scale_fact;  /* getMinTempFact()     VXS_Voxel.cpp:226                                */
    double max_scale_fact;  /* 1 + getGrowthAmplitude()  VXS_Voxel.cpp:225                           */
    double lattice_dim;     /* LocalVXC.GetLatticeDim()                                              */
    double drag_coeff;      /* LW AggregateDragCoefficient (VXG_F_FLUID)                             */
    double col_horizon;     /* CollisionHorizon (VXG_F_COLLISIONS), in lattice pitches  VX_Sim.cpp:1741, :2357-2385 */
    uint32_t flags;         /* VXG_F_*                                                               */
    uint32_t reserved;
    double growth_amplitude;     /* getGrowthAmplitude()   VXS_Voxel.cpp:313 (size-adjusted actuation)       */
    double midlife_freeze_time;  /* GetMidLifeFreezeTime() VXS_Voxel.cpp:241-259, :266                       */
    double stop_condition_value; /* GetStopConditionValue() (the freeze window is centred on the lifetime)   */
} vxg_sim_params;

/* One imported robot, flat.  Voxel i is the reference's simulation index i (rank among occupied
 * lattice cells, x fastest: VX_Sim.cpp:530-600).  Bond j joins voxel bond_v1[j] (negative side) to
 * bond_v2[j] (positive side) along bond_axis[j] (0 = X, 1 = Y, 2 = Z: VX_Sim.cpp:621-641,
 * :726-751).  A voxel has at most one bond per direction.  The host keeps ownership of every array;
 * vxg_upload copies. */
typedef struct vxg_robot_desc {
    int32_t n_vox, n_bond, n_vclass, n_bclass;
    const double* nom_pos;        /* [n_vox][3] NominalPosition, metres                               */
    const int32_t* vox_class;     /* [n_vox]                                                          */
    const int32_t* bond_v1;       /* [n_bond]                                                         */
    const int32_t* bond_v2;       /* [n_bond]                                                         */
    const int32_t* bond_axis;     /* [n_bond] 0/1/2                                                   */
    const int32_t* bond_class;    /* [n_bond]                                                         */
    const vxg_vox_class* vclasses;   /* [n_vclass]                                                    */
    const vxg_bond_class* bclasses;  /* [n_bclass]                                                    */
    vxg_sim_params sim;
} vxg_robot_desc;

/* Dynamic state of one robot in the reference's own terms (fp64 on the host side whatever the device
 * precision).  Any pointer may be NULL to skip that field.  Used by vxg_download_state (device ->
 * CVX_Sim::VoxArray / BondArrayInternal write-back for the GUI and for parity tests) and
 * vxg_upload_state (resume / perturbation tests). */
typedef struct vxg_robot_state {
    double* pos;        /* [n_vox][3]  Pos        VXS_Voxel.h:122                                     */
    double* quat;       /* [n_vox][4]  Angle (w,x,y,z)                                                */
    double* lin_mom;    /* [n_vox][3]  LinMom                                                         */
    double* ang_mom;    /* [n_vox][3]  AngMom                                                         */
    double* scale;      /* [n_vox]     Scale                                                          */
    int32_t* static_fric; /* [n_vox]   StaticFricFlag of the last step                                */
    double* bond_force1;  /* [n_bond][3] Force1 of the last step      VXS_Bond.h:54                   */
    double* bond_force2;  /* [n_bond][3]                                                              */
    double* bond_moment1; /* [n_bond][3]                                                              */
    double* bond_moment2; /* [n_bond][3]                                                              */
    double* bond_strain;  /* [n_bond]  CurStrainTot                                                   */
    int32_t* bond_small_angle; /* [n_bond] SmallAngle                                                 */
    double* vox_strain;   /* [n_vox][6] StrainPosDirsCur.xyz, StrainNegDirsCur.xyz of the last bond pass (LW/VXS_Voxel.cpp:258-268;
                             they set the corner extents of the surface mesh, :474-475); download only, batches with mesh capacity
                             (zeros otherwise)                                                                                  */
} vxg_robot_state;

/* Sizing of a batch: every robot slot gets the same stride so a slot is one contiguous blob. */
typedef struct vxg_limits {
    int32_t max_vox;      /* >= n_vox of every robot in the batch                                    */
    int32_t max_vclass;   /* >= n_vclass                                                             */
    int32_t max_bclass;   /* >= n_bclass                                                             */
    int32_t precision;    /* VXG_PREC_F32 / VXG_PREC_F64                                             */
    int32_t max_col_pairs;/* capacity for contact bonds per robot (0 = self-collision unused)        */
    int32_t max_mesh_verts;/* capacity for surface-mesh vertices per robot: needed (>= vxg_mesh_vertex_count of every
                              robot) to upload VXG_F_FLUID robots; 0 = no fluid-drag support in this batch      */
    int32_t reserved[2];  /* [0]: 0, or the voxel stride to lay the batch out with (one of the kernels' strides, >= what max_vox needs):
                             lets batches of differently sized small robots share one kernel image; [1]: 0                      */
} vxg_limits;

/* status bits written by vxg_sync */
#define VXG_ST_DIVERGED 0x1u /* a bond strain exceeded 100: Integrate() returned false, VX_Sim.cpp:1775-1777 */
#define VXG_ST_FINISHED 0x2u /* the robot has run the step budget set with vxg_set_step_budget      */
#define VXG_ST_COL_OVERFLOW 0x4u /* more contact bonds than limits.max_col_pairs: the robot stopped (never truncated silently) */

/* Per-robot reductions of the voxel state (what CVX_Sim::UpdateStats / the fitness probes read,
 * VX_Sim.cpp:1511-1690, :2415-2441, :2584-2712).  Filled by vxg_reduce_stats. */
typedef struct vxg_robot_stats {
    double com[3];        /* mass-weighted centre of mass  GetCM() VX_Sim.cpp:2415-2441              */
    double ini_com[3];    /* COM latched at the InitCmTime step   VX_Sim.cpp:1064-1075 (zeros before) */
    double max_vel;       /* max |Vel|  (SS.MaxVoxVel)                                               */
    double min_pos[3];    /* min over voxels of Pos                                                  */
    double max_pos[3];    /* max over voxels of Pos                                                  */
    double time;          /* CurTime                                                                 */
    int64_t steps;        /* CurStepCount                                                            */
    int32_t n_touching_floor; /* count of 0.5 Scale - z > 0     VX_Sim.cpp:2432-2441                 */
    int32_t ini_com_latched;
    /* fitness probes of the result writer (VX_Sim.cpp:2584-2712), in metres (the writer divides by the lattice pitch) */
    double anterior_dist;     /* max over voxels of the horizontal distance to IniCM   getAnteriorDist  */
    double posterior_dist;    /* min ...                                               getPosteriorDist */
    double anterior_y;        /* max Pos.y over voxels whose material index != 5, floor 0  getAnteriorY  */
    double posterior_y;       /* min Pos.y over voxels whose material index != 5, cap 1e5 lattice pitches  getPosteriorY */
    double eol_posterior_y;   /* posterior_y latched at the TimeStep() entry where CurTime >= StopConditionValue (VX_Sim.cpp:1078-1082); 0 = never */
    int32_t n_nonfeet_touching_floor; /* touching-floor count restricted to material index 6  GetNumNonFeetTouchingFloor */
    int32_t reserved;
    double mesh_volume;       /* volume enclosed by the deformed surface mesh, sum over its facets of (V1 x V2) . V3 / 6
                                 (CVX_MeshUtil::computeCurrentRobotVolume, LW/VX_MeshUtil.cpp:908-953 -> RobotVolumeEnd); -1 when the
                                 batch was created without mesh capacity (limits.max_mesh_verts = 0)                              */
} vxg_robot_stats;

typedef struct vxg_batch vxg_batch; /* opaque; owns all device memory */

/* Library / device discovery.  vxg_device_count returns 0 (not an error) when no GPU is visible so
 * host-only tools can still link the library; every compute entry point then fails with
 * VXG_ERR_NO_DEVICE. */
int vxg_abi_version(void);
int vxg_device_count(void);
const char* vxg_last_error(void);
/* on != 0: vxg_destroy keeps a batch's device and pinned blocks on a free list and vxg_create reuses them (a resident process that
 * builds a population per GA generation: `voxelyze --serve`); 0 (the default) releases the list.  Process-wide. */
int vxg_set_pooling(int on);

/* Replaces: construction of CVX_Sim::VoxArray / BondArrayInternal storage (VX_Sim.cpp:488-717). */
int vxg_create(vxg_batch** out, int device, int n_robots, const vxg_limits* lim);
void vxg_destroy(vxg_batch* b);

/* Replaces: CVX_Sim::Import's array fill + ResetSimulation (VX_Sim.cpp:530-641, :839-1007): copies the
 * descriptor of one robot into slot `robot` and resets its dynamic state (Pos = NominalPosition,
 * Angle = identity, momenta 0, Scale = NominalSize, CurTime = 0, dt = 0 for the first bond pass). */
int vxg_upload(vxg_batch* b, int robot, const vxg_robot_desc* desc);

/* Steps each robot stops after (CurStepCount limit; the host derives it from StopConditionMet,
 * VX_Sim.cpp:1398-1422, by replaying the fp64 `CurTime += dt` accumulation).  latch_step[r] is the
 * CurStepCount at whose TimeStep() entry IniCM is latched (VX_Sim.cpp:1064-1075); -1 = never.
 * Either pointer may be NULL (no limit / no latch). */
int vxg_set_step_budget(vxg_batch* b, const int64_t* total_steps, const int64_t* latch_step);
/* eol_step[r] = CurStepCount at whose TimeStep() entry EndOfLifetimePosteriorY is latched (CurTime >=
 * StopConditionValue, VX_Sim.cpp:1078-1082); -1 = never (the usual case without an afterlife). */
int vxg_set_eol_step(vxg_batch* b, const int64_t* eol_step);

/* Replaces: n_steps calls of CVX_Sim::Integrate() (VX_Sim.cpp:1763-1933) on every robot of the batch
 * that is neither diverged nor finished.  Asynchronous on the batch's stream. */
int vxg_step(vxg_batch* b, int n_steps);

/* Waits for the stream; status_per_robot (may be NULL) receives VXG_ST_* bits.  Maps to TimeStep()'s
 * `false` + "Simulation Diverged..." (VX_Sim.cpp:1147-1150) for bit 0. */
int vxg_sync(vxg_batch* b, uint32_t* status_per_robot);

/* Device time of the launches issued by the last vxg_step call (CUDA events on the batch stream);
 * valid after vxg_sync. */
int vxg_last_step_ms(vxg_batch* b, float* ms);
/* Number of kernels this library launched on the batch so far. */
int64_t vxg_launch_count(vxg_batch* b);
/* Benchmark markers: record CUDA event `slot` (0..VXG_N_MARKS-1) on the batch stream -- the stream the
 * kernels are launched on -- and read the device time between two recorded slots (waits for `slot1`). */
#define VXG_N_MARKS 8
int vxg_mark(vxg_batch* b, int slot);
int vxg_mark_elapsed_ms(vxg_batch* b, int slot0, int slot1, float* ms);

/* Replaces: the per-voxel / per-bond getters the GUI and CVX_MeshUtil read after a step
 * (VXS_Voxel.h:46-59, VXS_Bond.h:31-43). */
int vxg_download_state(vxg_batch* b, int robot, vxg_robot_state* out);
int vxg_upload_state(vxg_batch* b, int robot, const vxg_robot_state* in, double cur_time, int64_t cur_step);

/* Bond outputs (Force1/2, Moment1/2, strain) are only kept when recording is on: the last step of
 * every vxg_step call then also stores them for vxg_download_state. Off by default (costs HBM writes). */
int vxg_enable_bond_record(vxg_batch* b, int on);

/* Introspection for benchmarks/tests: launch geometry and per-robot byte counts of the state blob. */
int vxg_kernel_config(vxg_batch* b, int* threads, int* voxels_per_thread, int* smem_resident,
                      int64_t* smem_bytes, int64_t* blob_bytes, int64_t* dyn_bytes);
/* sizeof() of the ABI structs in the order vox_class, bond_class, sim_params, robot_desc, robot_state,
 * limits, robot_stats; returns how many there are. Lets a foreign-language mirror check its layout. */
int vxg_struct_sizes(int* out, int n);

/* Replaces: CVX_Sim::UpdateStats / GetCM (VX_Sim.cpp:1511-1690, :2415-2441) for the quantities the
 * result writer needs; out is [n_robots]. Synchronous. */
int vxg_reduce_stats(vxg_batch* b, vxg_robot_stats* out);

/* Host-managed contact bonds (seam B2, where CVX_Sim::UpdateCollisions stays on the host): replaces the list
 * CVX_Sim::CalcL1Bonds builds (VX_Sim.cpp:2357-2413) with n_pairs voxel index pairs (v1, v2) and switches the
 * device-side rebuild off for this robot.  The contact stiffness is derived from the voxels' classes
 * (CVX_Bond::LinkVoxels VX_Bond.cpp:90-102: a1 = composite E * mean size); `a1` is accepted for symmetry with
 * CVXS_BondCollision and may be NULL.  n_pairs < 0 hands the list back to the device. */
int vxg_set_collision_pairs(vxg_batch* b, int robot, int n_pairs, const int32_t* v1, const int32_t* v2,
                            const double* a1);

/* Host-buffer convenience used by the end-to-end benchmark: upload all descriptors, run n_steps,
 * download positions of every robot into pos_out ([n_robots][max_vox][3] doubles). */
int vxg_run_batch_host(vxg_batch* b, const vxg_robot_desc* descs, int n_descs, int n_steps, double* pos_out);
/* Surface-mesh vertices of a robot (lattice points shared by 1..7 voxels, LW/VX_MeshUtil.cpp:249-262): what
 * limits.max_mesh_verts must cover for the land/water fluid drag. */
int vxg_mesh_vertex_count(const vxg_robot_desc* desc);
/* Bytes one full population upload copies host -> device (for the benchmark's h2d accounting). */
int64_t vxg_upload_bytes(vxg_batch* b);

/* ---- ragged populations ---------------------------------------------------------------------------------------------------
 * Replaces evosoro's fan-out of one `voxelyze` process per individual (evosoro/tools/evaluation.py:89-90) for a population whose
 * members differ in size and features.  The robots are bucketed by what the kernels specialise on (voxel stride, self-collision,
 * surface mesh, production / wide kernel family, class-table size); every bucket is a vxg_batch of its own on its own stream, so
 * the buckets' grids run concurrently and no robot pays for the largest member's stride or the most general member's family.
 * Results are bit-identical to running each robot alone.  Robot indices below are positions in the `descs` array. */
typedef struct vxg_population vxg_population;
typedef struct vxg_pop_options {
    int32_t precision;          /* VXG_PREC_*                                                                                  */
    int32_t chunk_robots;       /* > 0: buckets are cut into batches of at most this many robots -- the granularity at which
                                   vxg_pop_run_host pipelines flatten / upload / simulate / download; 0 = one batch per bucket */
    int32_t want_mesh;          /* 1: surface-mesh tables for every VXG_F_LW_SCALE robot, not only the VXG_F_FLUID ones (the
                                   land/water result file reports RobotVolumeEnd for all of them)                               */
    int32_t col_pairs_per_vox;  /* contact-bond capacity of a self-colliding robot = this x n_vox (at least 512); 0 = 4;
                                   negative: exactly -this bonds (tests of the overflow path)                            */
    int32_t reserved[4];
} vxg_pop_options;
#define VXG_FAM_COMPACT 0x1  /* production family (compact bond table, gather form) */
#define VXG_FAM_COLLIDE 0x2
#define VXG_FAM_MESH    0x4
int vxg_pop_create(vxg_population** out, int device, const vxg_robot_desc* descs, int n_robots, const vxg_pop_options* opt);
void vxg_pop_destroy(vxg_population* p);
/* Reset every robot from its descriptor (same array shape as at create; ResetSimulation for the whole population). */
int vxg_pop_upload(vxg_population* p, const vxg_robot_desc* descs);
/* vxg_set_step_budget + vxg_set_eol_step with population indices; any pointer may be NULL. */
int vxg_pop_set_plan(vxg_population* p, const int64_t* total_steps, const int64_t* latch_step, const int64_t* eol_step);
int vxg_pop_step(vxg_population* p, int64_t n_steps);   /* asynchronous; batches whose robots have all finished are skipped */
int vxg_pop_run(vxg_population* p);                     /* to the end of every robot's budget (asynchronous)                */
int vxg_pop_sync(vxg_population* p, uint32_t* status_per_robot);
int vxg_pop_reduce_stats(vxg_population* p, vxg_robot_stats* out);
int vxg_pop_download_state(vxg_population* p, int robot, vxg_robot_state* out);
/* Host-buffer end-to-end call (what bench.py's `e2e` times): flatten + upload every descriptor, n_steps, positions of robot r into
 * pos_out[r][pos_stride_vox][3]; pipelined over the population's batches. */
int vxg_pop_run_host(vxg_population* p, const vxg_robot_desc* descs, int n_steps, double* pos_out, int pos_stride_vox);
/* introspection */
int vxg_pop_n_batches(const vxg_population* p);
int vxg_pop_batch_info(vxg_population* p, int batch, int32_t* n_robots, int32_t* stride, int32_t* threads, int32_t* family);
int vxg_pop_locate(const vxg_population* p, int robot, int32_t* batch, int32_t* slot);
int64_t vxg_pop_launch_count(const vxg_population* p);
/* vxg_mark / vxg_mark_elapsed_ms over all the batches' streams: device time from the earliest event of mark slot0 to the latest of slot1 */
int vxg_pop_mark(vxg_population* p, int slot);
int vxg_pop_mark_elapsed_ms(vxg_population* p, int slot0, int slot1, float* ms);
int64_t vxg_pop_upload_bytes(const vxg_population* p);    /* host -> device bytes of one whole-population upload */
int64_t vxg_pop_download_bytes(const vxg_population* p);  /* device -> host bytes of one vxg_pop_run_host        */

#ifdef __cplusplus
}
#endif
#endif /* VXGPU_H */
