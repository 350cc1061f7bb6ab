/* vx_oracle.h -- CPU restatement of the reference Voxelyze step.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (evosoro_b200/, include/) links, imports or executes this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg do, and only as the checker.
 * Pinned against the unmodified reference built under oracle/_ref (ref_dump) -- see tests/test_oracle.py.
 *
 * It consumes the same flat descriptor as the engine (include/vxgpu.h) and follows the reference's
 * fp64 arithmetic operation for operation (file:line cited at each function in vx_oracle.c).
 */
#ifndef VX_ORACLE_H
#define VX_ORACLE_H
#include "../include/vxgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vxo_sim vxo_sim;

vxo_sim* vxo_create(const vxg_robot_desc* d);   /* Import + ResetSimulation */
void vxo_destroy(vxo_sim* s);
/* n x Integrate(); returns the number of steps actually taken (stops early when diverged) */
int64_t vxo_step(vxo_sim* s, int64_t n);
int vxo_diverged(const vxo_sim* s);
double vxo_time(const vxo_sim* s);
int64_t vxo_step_count(const vxo_sim* s);
void vxo_get_state(const vxo_sim* s, vxg_robot_state* out);
void vxo_get_com(const vxo_sim* s, double* com3);       /* GetCM() */
int vxo_num_touching_floor(const vxo_sim* s);
double vxo_max_vel(const vxo_sim* s);
void vxo_set_collision_pairs(vxo_sim* s, int n, const int32_t* v1, const int32_t* v2, const double* a1);

#ifdef __cplusplus
}
#endif
#endif
