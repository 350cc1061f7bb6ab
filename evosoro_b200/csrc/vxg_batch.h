// vxg_batch.h -- internals of libvxgpu.so shared by its translation units (vxg_engine.cu: one homogeneous batch;
// vxg_population.cu: a ragged population bucketed into such batches).  Not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vxgpu.h"
#include "vxg_launch.cuh"

int vxg_i_fail(int code, const char* fmt, ...);   // sets the calling thread's vxg_last_error() text, returns code
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return vxg_i_fail(VXG_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

using vx::BlobLayout; using vx::RobotCtl; using vx::ScaleK; using vx::StatsDev; using vx::ColPair;
using vx::BondK; using vx::BondKc; using vx::VoxK;

struct vxg_batch {
    int device = 0, n_robots = 0, prec = 0;
    vxg_limits lim{};
    BlobLayout L{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t marks[VXG_N_MARKS] = {};
    bool mark_set[VXG_N_MARKS] = {};
    bool timed = false;
    // device
    unsigned char* d_blobs = nullptr;
    RobotCtl* d_ctl = nullptr;
    void* d_bk = nullptr; double* d_bk_dist = nullptr; void* d_vk = nullptr; ScaleK* d_sk = nullptr;
    double* d_scale_c = nullptr; void* d_ext_disp = nullptr; void* d_ext_tdisp = nullptr; double* d_nom = nullptr;
    void* d_bond_rec = nullptr; StatsDev* d_stats = nullptr;
    uint32_t* d_bpair = nullptr; uint32_t* h_bpair = nullptr;   // voxel-class pair of every re-keyed bond class (compact family)
    // self-collision (allocated when limits.max_col_pairs > 0)
    ColPair* d_col_pairs = nullptr; int16_t* d_col_surf = nullptr; int8_t* d_col_lat4 = nullptr; unsigned long long* d_col_near = nullptr;
    ColPair* h_col_pairs = nullptr; int16_t* h_col_surf = nullptr; int8_t* h_col_lat4 = nullptr; unsigned long long* h_col_near = nullptr;
    // pinned host mirrors (whole batch, uploaded in one go before the first step after a change)
    unsigned char* h_blobs = nullptr;
    RobotCtl* h_ctl = nullptr;
    unsigned char* h_bkc = nullptr;   // compact bond classes of every robot (BondKc), used when the whole batch qualifies
    bool compact = false;             // mode of the tables currently on the device
    unsigned char* h_bk = nullptr; double* h_bk_dist = nullptr; unsigned char* h_vk = nullptr; ScaleK* h_sk = nullptr;
    double* h_scale_c = nullptr; unsigned char* h_ext_disp = nullptr; unsigned char* h_ext_tdisp = nullptr; double* h_nom = nullptr;
    StatsDev* h_stats = nullptr;
    bool dirty = false;          // host mirrors newer than device
    bool ctl_dirty = false;      // only h_ctl newer (step budget)
    bool tables_dirty = false;   // the class tables must travel again in the other form (bond recording switched on / off)
    std::vector<char> uploaded;
    std::vector<char> slot_dirty;    // the slot's host mirror (blob, Scale, control block) is newer than the device copy
    std::vector<char> reset_image;   // the slot's host mirror is the plain reset image vxg_upload packed (nothing pulled or edited since);
                                     // 2 = only its position planes and static section are filled in (pack_blob partial)
    std::vector<std::vector<int32_t>> bond_v1, bond_v2, bond_axis;   // for download mapping
    int64_t launches = 0;
    size_t last_upload_bytes = 0;    // blob + NominalPosition bytes the last population upload copied host -> device
    int vpt = 1, threads = 32;
    bool smem_mode = true;
    size_t smem_bytes = 0;
    size_t tb_wide = 0, tb_compact = 0;          // shared-memory table bytes with the wide / compact bond classes
    bool smem_wide_fits = false;                 // blob + wide tables fit in shared memory
    bool smem_mode_wide_or_compact = false;      // blob + compact tables fit
    size_t rs() const { return prec == VXG_PREC_F64 ? 8 : 4; }
    size_t bk_sz() const { return prec == VXG_PREC_F64 ? sizeof(BondK<double>) : sizeof(BondK<float>); }
    size_t bkc_sz() const { return prec == VXG_PREC_F64 ? sizeof(BondKc<double>) : sizeof(BondKc<float>); }
    size_t vk_sz() const { return prec == VXG_PREC_F64 ? sizeof(VoxK<double>) : sizeof(VoxK<float>); }
};


// make the device copy current (host mirrors -> device), asynchronous on the batch stream
int vxg_i_flush(vxg_batch* b);
// after a step: position planes + control blocks, device -> pinned mirrors, asynchronous on the batch stream
int vxg_i_enqueue_positions(vxg_batch* b);
// once the stream has been waited for: slot's position planes -> out[n_vox][3]
void vxg_i_unpack_positions(const vxg_batch* b, int slot, double* out);
// would this robot run the production (compact) family?  (rekey_bond_classes + pack_compact, vx_pack.h); *n_bclass receives the
// number of bond classes after re-keying by voxel-class pair (what limits.max_bclass has to cover for that family)
bool vxg_i_compact_ok(const vxg_robot_desc* d, int* n_bclass);
// host threads a call may use: VXG_HOST_THREADS, else hardware threads / LOCAL_WORLD_SIZE (one process per GPU shares the box)
int vxg_i_host_threads(int cap);
