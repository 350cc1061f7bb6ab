/* vxhost.h -- host side above the engine's C-ABI: .vxa import and the TimeStep bookkeeping the
 * reference keeps on the CPU.  Part of libvxgpu.so; plain C, no CUDA needed for any call here.
 *
 * Mirrors, for the hot path's callers only:
 *   CVX_Sim::LoadVXAFile / ReadVXA        evosoro/_voxcad/Voxelyze/VX_Sim.cpp:159-205, :263-356
 *   CVX_Environment::ReadXML              evosoro/_voxcad/Voxelyze/VX_Environment.cpp:123-293
 *   CVX_Object::ReadXML (lattice/palette/structure)  evosoro/_voxcad/Voxelyze/VX_Object.cpp:435-490, :1064-1074, :1344-1441, :1733-1930
 *   CVX_Sim::Import (voxel/bond arrays, constants)   evosoro/_voxcad/Voxelyze/VX_Sim.cpp:488-717, VX_Bond.cpp:65-173, VX_Voxel.cpp:94-128
 *   CVX_Sim::CalcMaxDt                    evosoro/_voxcad/Voxelyze/VX_Sim.cpp:1693-1727
 *   the CLI loop's stop test              evosoro/_voxcad/voxelyzeMain/main.cpp:89-110, VX_Sim.cpp:1398-1422
 * A .vxa that enables anything the kernels do not implement is REJECTED (VXG_UNSUPPORTED_FEATURE):
 * there is no CPU fallback to silently differ from (SURVEY.md section 7 "Unsupported-feature policy").
 */
#ifndef VXHOST_H
#define VXHOST_H

#include "vxgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vxh_model vxh_model; /* opaque: parsed .vxa + the flattened descriptor it owns */

/* engine variant whose semantics the file is imported with */
#define VXH_VARIANT_VOXCAD 0     /* evosoro/_voxcad            */
#define VXH_VARIANT_LAND_WATER 1 /* evosoro/_voxcad_land_water */

typedef struct vxh_info {
    int32_t n_vox, n_bond;
    int32_t x_voxels, y_voxels, z_voxels;
    int32_t stop_condition_type;   /* StopCondition enum, VX_Enums.h:55-65 */
    int32_t self_col_enabled;      /* <SelfColEnabled> as written in the file */
    int32_t write_fitness_file;
    double stop_condition_value, init_cm_time, afterlife_time;
    double dt_frac, optimal_dt, lattice_dim;
    double col_horizon;
    char fitness_file_name[512];
} vxh_info;

/* Replaces LoadVXAFile + Import.  `ignore_collisions` != 0 imports a SelfColEnabled=1 file with the
 * feature switched off (what `EnableFeature(VXSFEAT_COLLISIONS,false)` before Import does) instead of
 * rejecting it. Returns 0 / VXG_ERR_ARG (parse error) / VXG_UNSUPPORTED_FEATURE. */
int vxh_load_vxa_file(const char* path, int variant, int ignore_collisions, vxh_model** out);
int vxh_load_vxa_string(const char* text, int variant, int ignore_collisions, vxh_model** out);
void vxh_free(vxh_model* m);

const vxg_robot_desc* vxh_desc(const vxh_model* m);   /* valid until vxh_free */
int vxh_get_info(const vxh_model* m, vxh_info* out);
/* StoXIndexMap (VX_Sim.cpp:557): lattice cell index of every simulation voxel; out is [n_vox] */
int vxh_cell_index(const vxh_model* m, int32_t* out);

/* Replays the CLI's `while (!StopConditionMet()) TimeStep()` clock in fp64 (VX_Sim.cpp:1402-1403,
 * :1415-1422, :1929): total number of steps the reference takes, and the CurStepCount at whose
 * TimeStep() entry IniCM is latched (VX_Sim.cpp:1064).  Only SC_MAX_SIM_TIME / SC_MAX_TIME_STEPS /
 * SC_TEMP_CYCLES are time-driven; others return VXG_UNSUPPORTED_FEATURE. */
int vxh_step_plan(const vxh_model* m, int64_t* total_steps, int64_t* latch_step, double* final_time);

/* Same, plus the CurStepCount at whose TimeStep() entry EndOfLifetimePosteriorY is latched (VX_Sim.cpp:1078-1082; -1 = never). */
int vxh_step_plan_ex(const vxh_model* m, int64_t* total_steps, int64_t* latch_step, int64_t* eol_step, double* final_time);

/* Replaces CVX_SimGA::WriteResultFile / SaveResultFile (VX_SimGA.cpp:25-168; land/water fork: LW/VX_SimGA.cpp:36-75):
 * formats the <Voxelyze_Sim_Result><Fitness> block from the device reductions with the reference's tag set and
 * number formatting.  buf/cap receive the text (may be NULL/0); path (may be NULL) is written if given.
 * Returns the text length, or a negative error. */
int vxh_format_result(const vxh_model* m, const vxg_robot_stats* st, char* buf, int cap, const char* path);

/* thread-local message of the last failing vxh_* call */
const char* vxh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VXHOST_H */
