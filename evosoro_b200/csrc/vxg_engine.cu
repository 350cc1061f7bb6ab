// vxg_engine.cu -- C-ABI of libvxgpu.so (include/vxgpu.h): batch management, descriptor flattening
// into the device blob layout, kernel dispatch, state download.  No CPU compute path exists here:
// every compute entry point fails with VXG_ERR_NO_DEVICE when no CUDA device is present.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "vxg_batch.h"
#include "vx_pack.h"

using namespace vx;

static thread_local std::string g_err;
int vxg_i_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define fail vxg_i_fail

static_assert(sizeof(StatsDev) == sizeof(vxg_robot_stats), "StatsDev must mirror vxg_robot_stats");

extern "C" int vxg_abi_version(void) { return VXG_ABI_VERSION; }
extern "C" const char* vxg_last_error(void) { return g_err.c_str(); }
extern "C" int vxg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
// sizes of the ABI structs as compiled, for the ctypes mirror's self-check
extern "C" int vxg_struct_sizes(int* out, int n) {
    int v[] = {(int)sizeof(vxg_vox_class), (int)sizeof(vxg_bond_class), (int)sizeof(vxg_sim_params), (int)sizeof(vxg_robot_desc),
               (int)sizeof(vxg_robot_state), (int)sizeof(vxg_limits), (int)sizeof(vxg_robot_stats)};
    int m = (int)(sizeof v / sizeof v[0]);
    for (int i = 0; i < n && i < m; i++) out[i] = v[i];
    return m;
}

template <typename Real> static BatchDev<Real> dev_view(const vxg_batch* b) {
    BatchDev<Real> B;
    B.blobs = b->d_blobs; B.L = b->L; B.ctl = b->d_ctl;
    B.bk = (unsigned char*)b->d_bk; B.bk_dist = b->d_bk_dist; B.vk = (VoxK<Real>*)b->d_vk; B.sk = b->d_sk;
    B.scale_c = b->d_scale_c; B.ext_disp = (Real*)b->d_ext_disp; B.ext_tdisp = (Real*)b->d_ext_tdisp; B.nom = b->d_nom;
    B.bond_rec = (Real*)b->d_bond_rec; B.bpair = b->d_bpair;
    B.max_vclass = b->lim.max_vclass; B.max_bclass = b->lim.max_bclass; B.n_robots = b->n_robots;
    B.col_pairs = b->d_col_pairs; B.col_surf = b->d_col_surf; B.col_lat4 = b->d_col_lat4; B.col_near = b->d_col_near;
    B.max_col = b->lim.max_col_pairs;
    return B;
}

// Picks the kernel family for the batch (each family lives in its own translation unit, vxg_k_*.cu).
// The compact bond table exists for the production path only (fp32, shared-memory resident).
static cudaError_t launch_step_p(vxg_batch* b, int n_steps, int record) {
    LaunchCfg c; c.n_robots = b->n_robots; c.nv = b->L.nv; c.smem_bytes = b->smem_bytes; c.stream = b->stream; c.n_steps = n_steps; c.record = record;
    c.vpt = b->vpt;
    if (b->prec == VXG_PREC_F64) {
        // fluid drag / self-collision in fp64 run on the blob in global memory (one code path for every robot size: this is the
        // logic-parity mode, not a production one)
        if (b->L.nvert_cap > 0 || b->lim.max_col_pairs > 0) return launch_f64_extras_global(dev_view<double>(b), c, b->L.nvert_cap > 0, b->lim.max_col_pairs > 0);
        return launch_f64(dev_view<double>(b), c, b->smem_mode);
    }
    const BatchDev<float> B = dev_view<float>(b);
    if ((b->L.nvert_cap > 0 || b->lim.max_col_pairs > 0) && !b->smem_mode) return launch_f32_extras_global(B, c, b->L.nvert_cap > 0, b->lim.max_col_pairs > 0);
    if (b->L.nvert_cap > 0)
        return (b->compact && b->smem_mode) ? launch_f32_fluid_compact(B, c, b->lim.max_col_pairs > 0) : launch_f32_fluid(B, c, b->lim.max_col_pairs > 0);
    if (b->lim.max_col_pairs > 0) return (b->compact && b->smem_mode) ? launch_f32_collide_compact(B, c) : launch_f32_collide(B, c);
    if (!b->smem_mode) return launch_f32_wide(B, c, false);
    return b->compact ? launch_f32_compact(B, c) : launch_f32_wide(B, c, true);
}

// ---- block cache behind vxg_create / vxg_destroy (vxg_set_pooling) --------------------------------------------------------------
// A resident engine process (voxelyze --serve) creates and destroys a population of batches per GA generation: ~35 device and pinned
// allocations each, 0.05 - 0.8 s of cudaMalloc / cudaMallocHost / cudaFree / cudaFreeHost per generation, all of which serialise with
// the next generation's launches inside the driver.  With pooling on, a destroyed batch's blocks go to a free list (sizes rounded up
// to 1/8 of the next power of two, so that a generation whose largest robot differs a little still hits) and the next vxg_create
// takes them from there; the one-shot CLI and the library's other users never switch it on.
namespace {
struct PoolBlock { void* p; size_t bytes; int where; };   // where: device ordinal, -1 = pinned host memory
std::mutex g_pool_mu;
std::unordered_map<void*, PoolBlock> g_pool_live;
std::vector<PoolBlock> g_pool_free;
size_t g_pool_free_bytes = 0;
bool g_pooling = false;
const size_t POOL_CAP_BYTES = (size_t)16 << 30;

size_t pool_round(size_t bytes) {
    size_t g = 4096;
    while (g * 8 < bytes) g <<= 1;
    return (bytes + g - 1) / g * g;
}
void pool_release(const PoolBlock& b) { if (b.where < 0) cudaFreeHost(b.p); else { cudaSetDevice(b.where); cudaFree(b.p); } }
}  // namespace

static cudaError_t vx_alloc(void** out, size_t bytes, int where) {
    *out = nullptr;
    if (bytes == 0) return cudaSuccess;
    bool pooling;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        pooling = g_pooling;
        if (pooling) {
            const size_t want = pool_round(bytes);
            for (size_t i = 0; i < g_pool_free.size(); i++)
                if (g_pool_free[i].where == where && g_pool_free[i].bytes == want) {
                    PoolBlock b = g_pool_free[i];
                    g_pool_free[i] = g_pool_free.back(); g_pool_free.pop_back(); g_pool_free_bytes -= b.bytes;
                    g_pool_live[b.p] = b; *out = b.p;
                    return cudaSuccess;
                }
        }
    }
    const size_t n = pooling ? pool_round(bytes) : bytes;
    cudaError_t e = where < 0 ? cudaMallocHost(out, n) : cudaMalloc(out, n);
    if (e == cudaSuccess && pooling) { std::lock_guard<std::mutex> g(g_pool_mu); g_pool_live[*out] = PoolBlock{*out, n, where}; }
    return e;
}
static void vx_free(void* p, int where) {
    if (!p) return;
    std::vector<PoolBlock> drop;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        auto it = g_pool_live.find(p);
        if (it == g_pool_live.end()) drop.push_back(PoolBlock{p, 0, where});          // allocated while pooling was off
        else {
            PoolBlock b = it->second; g_pool_live.erase(it);
            if (!g_pooling) drop.push_back(b);
            else {
                g_pool_free.push_back(b); g_pool_free_bytes += b.bytes;
                while (g_pool_free_bytes > POOL_CAP_BYTES && !g_pool_free.empty()) {   // oldest first
                    drop.push_back(g_pool_free.front()); g_pool_free_bytes -= g_pool_free.front().bytes; g_pool_free.erase(g_pool_free.begin());
                }
            }
        }
    }
    for (const PoolBlock& b : drop) pool_release(b);
}
extern "C" int vxg_set_pooling(int on) {
    std::vector<PoolBlock> drop;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        g_pooling = on != 0;
        if (!g_pooling) { drop.swap(g_pool_free); g_pool_free_bytes = 0; }
    }
    for (const PoolBlock& b : drop) pool_release(b);
    return VXG_OK;
}

static void free_all(vxg_batch* b) {
    vx_free(b->d_blobs, b->device); vx_free(b->d_ctl, b->device); vx_free(b->d_bk, b->device); vx_free(b->d_bk_dist, b->device); vx_free(b->d_vk, b->device); vx_free(b->d_sk, b->device);
    vx_free(b->d_scale_c, b->device); vx_free(b->d_ext_disp, b->device); vx_free(b->d_ext_tdisp, b->device); vx_free(b->d_nom, b->device); vx_free(b->d_bond_rec, b->device); vx_free(b->d_bpair, b->device); vx_free(b->h_bpair, -1);
    vx_free(b->d_stats, b->device);
    vx_free(b->d_col_pairs, b->device); vx_free(b->d_col_surf, b->device); vx_free(b->d_col_lat4, b->device); vx_free(b->d_col_near, b->device);
    vx_free(b->h_col_pairs, -1); vx_free(b->h_col_surf, -1); vx_free(b->h_col_lat4, -1); vx_free(b->h_col_near, -1);
    vx_free(b->h_bkc, -1);
    vx_free(b->h_blobs, -1); vx_free(b->h_ctl, -1); vx_free(b->h_bk, -1); vx_free(b->h_bk_dist, -1); vx_free(b->h_vk, -1);
    vx_free(b->h_sk, -1); vx_free(b->h_scale_c, -1); vx_free(b->h_ext_disp, -1); vx_free(b->h_ext_tdisp, -1); vx_free(b->h_nom, -1);
    vx_free(b->h_stats, -1);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    for (int i = 0; i < VXG_N_MARKS; i++) if (b->marks[i]) cudaEventDestroy(b->marks[i]);
    if (b->stream) cudaStreamDestroy(b->stream);
}

extern "C" int vxg_create(vxg_batch** out, int device, int n_robots, const vxg_limits* lim) {
    if (!out || !lim || n_robots <= 0) return fail(VXG_ERR_ARG, "vxg_create: bad argument");
    *out = nullptr;
    if (lim->max_vox <= 0 || lim->max_vclass <= 0 || lim->max_bclass <= 0) return fail(VXG_ERR_ARG, "vxg_create: limits must be positive");
    if (lim->precision != VXG_PREC_F32 && lim->precision != VXG_PREC_F64) return fail(VXG_ERR_ARG, "vxg_create: unknown precision");
    if (lim->max_vox > 32767) return fail(VXG_UNSUPPORTED_SIZE, "vxg_create: more than 32767 voxels per robot");
    if (lim->max_vclass > 65535 || lim->max_bclass > 65535) return fail(VXG_UNSUPPORTED_SIZE, "vxg_create: too many classes");
    int ndev = vxg_device_count();
    if (ndev <= 0) return fail(VXG_ERR_NO_DEVICE, "no CUDA device visible: the engine has no CPU path");
    if (device < 0 || device >= ndev) return fail(VXG_ERR_ARG, "vxg_create: device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    vxg_batch* b = new vxg_batch;
    b->device = device; b->n_robots = n_robots; b->lim = *lim; b->prec = lim->precision;
    if (lim->max_mesh_verts < 0 || lim->max_mesh_verts > 32767) { delete b; return fail(VXG_ERR_ARG, "vxg_create: max_mesh_verts out of range"); }
    int stride = stride_for(lim->max_vox);
    if (lim->reserved[0] > 0) {   // a wider stride asked for by the population layer (one kernel image for all small robots of a family)
        const bool warp_multiple = lim->reserved[0] >= 64 && lim->reserved[0] <= 256 && lim->reserved[0] % 32 == 0;   // every family has these
        if (lim->reserved[0] < lim->max_vox || !(warp_multiple || stride_for(lim->reserved[0]) == lim->reserved[0])) { delete b; return fail(VXG_ERR_ARG, "vxg_create: limits.reserved[0] = %d is not a stride that holds %d voxels", lim->reserved[0], lim->max_vox); }
        stride = lim->reserved[0];
    }
    b->L = make_layout(stride, (int)b->rs(), (lim->max_mesh_verts + 31) & ~31);
    // thread mapping: <= 512 threads per CTA, VPT voxels per thread
    b->vpt = b->L.nv <= 512 ? 1 : b->L.nv / 512;
    b->threads = b->L.nv / b->vpt;
    if (lim->max_col_pairs < 0 || lim->max_col_pairs > 16384) { delete b; return fail(VXG_ERR_ARG, "vxg_create: max_col_pairs out of range"); }
    const int mc = lim->max_col_pairs;
    size_t tb = b->prec == VXG_PREC_F64 ? tables_bytes<double>(lim->max_vclass, lim->max_bclass, false, mc) : tables_bytes<float>(lim->max_vclass, lim->max_bclass, false, mc);
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (tb > (size_t)max_smem) { delete b; return fail(VXG_UNSUPPORTED_SIZE, "vxg_create: class tables (%zu B) exceed shared memory", tb); }
    b->tb_wide = tb;
    b->tb_compact = tables_bytes<float>(lim->max_vclass, lim->max_bclass, true, mc);
    b->smem_wide_fits = b->L.total_bytes + tb <= (size_t)max_smem;
    b->smem_mode_wide_or_compact = b->prec == VXG_PREC_F32 && b->L.total_bytes + b->tb_compact <= (size_t)max_smem;
    b->smem_mode = b->smem_wide_fits || b->smem_mode_wide_or_compact;   // final choice is made at flush, when the classes are known
    b->smem_bytes = b->smem_mode ? b->L.total_bytes + tb : tb;

    size_t nr = (size_t)n_robots, mv = (size_t)lim->max_vclass, mb = (size_t)lim->max_bclass, nv = (size_t)b->L.nv;
    cudaError_t e = cudaSuccess;
#define DA(ptr, bytes) if (e == cudaSuccess) { e = vx_alloc((void**)&(ptr), (bytes), device); if (e == cudaSuccess) e = cudaMemset((ptr), 0, (bytes)); }
#define HA(ptr, bytes) if (e == cudaSuccess) { e = vx_alloc((void**)&(ptr), (bytes), -1); if (e == cudaSuccess) memset((ptr), 0, (bytes)); }
    DA(b->d_blobs, nr * b->L.total_bytes); DA(b->d_ctl, nr * sizeof(RobotCtl));
    DA(b->d_bk, nr * mb * b->bk_sz()); DA(b->d_bk_dist, nr * mb * sizeof(double)); DA(b->d_vk, nr * mv * b->vk_sz());
    DA(b->d_sk, nr * mv * sizeof(ScaleK)); DA(b->d_scale_c, nr * mv * sizeof(double)); DA(b->d_bpair, nr * mb * sizeof(uint32_t));
    DA(b->d_ext_disp, nr * mv * 3 * b->rs()); DA(b->d_ext_tdisp, nr * mv * 3 * b->rs()); DA(b->d_nom, nr * 3 * nv * sizeof(double));
    DA(b->d_stats, nr * sizeof(StatsDev));
    if (mc > 0) {
        DA(b->d_col_pairs, nr * (size_t)mc * sizeof(ColPair)); DA(b->d_col_surf, nr * nv * sizeof(int16_t));
        DA(b->d_col_lat4, nr * nv * 4); DA(b->d_col_near, nr * nv * sizeof(unsigned long long));
    }
    HA(b->h_blobs, nr * b->L.total_bytes); HA(b->h_ctl, nr * sizeof(RobotCtl));
    HA(b->h_bkc, nr * mb * b->bkc_sz());
    HA(b->h_bk, nr * mb * b->bk_sz()); HA(b->h_bk_dist, nr * mb * sizeof(double)); HA(b->h_vk, nr * mv * b->vk_sz());
    HA(b->h_sk, nr * mv * sizeof(ScaleK)); HA(b->h_scale_c, nr * mv * sizeof(double)); HA(b->h_bpair, nr * mb * sizeof(uint32_t));
    HA(b->h_ext_disp, nr * mv * 3 * b->rs()); HA(b->h_ext_tdisp, nr * mv * 3 * b->rs()); HA(b->h_nom, nr * 3 * nv * sizeof(double));
    HA(b->h_stats, nr * sizeof(StatsDev));
    if (mc > 0) {
        HA(b->h_col_pairs, nr * (size_t)mc * sizeof(ColPair)); HA(b->h_col_surf, nr * nv * sizeof(int16_t));
        HA(b->h_col_lat4, nr * nv * 4); HA(b->h_col_near, nr * nv * sizeof(unsigned long long));
    }
    // the zero fills above went to the default stream; the batch's own stream does not wait for that one (cudaStreamNonBlocking)
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
    for (int i = 0; i < VXG_N_MARKS; i++) if (e == cudaSuccess) e = cudaEventCreate(&b->marks[i]);
    if (e != cudaSuccess) {
        free_all(b); delete b; cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? VXG_ERR_NOMEM : VXG_ERR_CUDA, "vxg_create: %s", cudaGetErrorString(e));
    }
    b->uploaded.assign(nr, 0); b->reset_image.assign(nr, 0); b->slot_dirty.assign(nr, 1);
    b->bond_v1.resize(nr); b->bond_v2.resize(nr); b->bond_axis.resize(nr);
    for (size_t r = 0; r < nr; r++) { b->h_ctl[r].status = VXG_ST_FINISHED; b->h_ctl[r].total_steps = 0; }  // empty slots never run
    b->dirty = true;
    *out = b;
    return VXG_OK;
}

extern "C" void vxg_destroy(vxg_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    free_all(b);
    delete b;
}

template <typename Real> static void fill_classes(vxg_batch* b, int robot, const vxg_robot_desc* d) {
    size_t mb = (size_t)b->lim.max_bclass, mv = (size_t)b->lim.max_vclass, r = (size_t)robot;
    pack_classes<Real>(d, (BondK<Real>*)(b->h_bk) + r * mb, b->h_bk_dist + r * mb, (VoxK<Real>*)(b->h_vk) + r * mv,
                       b->h_sk + r * mv, b->h_scale_c + r * mv, (Real*)(b->h_ext_disp) + r * mv * 3, (Real*)(b->h_ext_tdisp) + r * mv * 3);
}

template <typename Real> static bool fill_compact(vxg_batch* b, int robot, const vxg_robot_desc* d, double* len) {
    return pack_compact<Real>(d, (BondKc<Real>*)(b->h_bkc) + (size_t)robot * b->lim.max_bclass, len);
}

template <typename Real> static void reset_blob(vxg_batch* b, int robot, const vxg_robot_desc* d) {
    // partial image: position planes + static section (incl. the fluid layouts' mesh tables, which pack_mesh fills in afterwards)
    pack_blob<Real>(d, b->L, b->h_blobs + (size_t)robot * b->L.total_bytes, b->h_nom + (size_t)robot * 3 * b->L.nv, true);
}
// the host mirror of a slot as a complete reset image (before anything reads or edits it on the host, or uploads it whole)
static void complete_mirror(vxg_batch* b, size_t r) {
    if (b->reset_image[r] != 2) return;
    unsigned char* blob = b->h_blobs + r * b->L.total_bytes;
    if (b->prec == VXG_PREC_F64) pack_blob_rest<double>(b->L, blob, b->h_ctl[r].n_vox); else pack_blob_rest<float>(b->L, blob, b->h_ctl[r].n_vox);
    b->reset_image[r] = 1;
}

extern "C" int vxg_upload(vxg_batch* b, int robot, const vxg_robot_desc* d) {
    if (!b || !d) return fail(VXG_ERR_ARG, "vxg_upload: null argument");
    if (robot < 0 || robot >= b->n_robots) return fail(VXG_ERR_ARG, "vxg_upload: robot %d out of range", robot);
    if (d->n_vox <= 0) return fail(VXG_ERR_ARG, "vxg_upload: robot has no voxels");
    if (d->n_vox > b->lim.max_vox) return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: %d voxels > max_vox %d", d->n_vox, b->lim.max_vox);
    if (d->n_vclass > b->lim.max_vclass || d->n_vclass <= 0) return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: %d voxel classes (max %d)", d->n_vclass, b->lim.max_vclass);
    if (d->n_bclass > b->lim.max_bclass) return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: %d bond classes (max %d)", d->n_bclass, b->lim.max_bclass);
    if (d->n_bond < 0 || (d->n_bond > 0 && (!d->bond_v1 || !d->bond_v2 || !d->bond_axis || !d->bond_class || !d->bclasses)))
        return fail(VXG_ERR_ARG, "vxg_upload: bond arrays missing");
    if (!d->nom_pos || !d->vox_class || !d->vclasses) return fail(VXG_ERR_ARG, "vxg_upload: voxel arrays missing");
    const vxg_sim_params& P = d->sim;
    if ((P.flags & VXG_F_COLLISIONS) && b->lim.max_col_pairs <= 0)
        return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: a robot with self-collision needs a batch created with limits.max_col_pairs > 0");
    if ((P.flags & VXG_F_FLUID) && b->L.nvert_cap <= 0)
        return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: a fluid-environment robot needs a batch created with limits.max_mesh_verts > 0");
    if (!(P.dt > 0) || !std::isfinite(P.dt)) return fail(VXG_ERR_ARG, "vxg_upload: dt must be positive");
    for (int i = 0; i < d->n_vox; i++)
        if (d->vox_class[i] < 0 || d->vox_class[i] >= d->n_vclass) return fail(VXG_ERR_ARG, "vxg_upload: voxel %d has class %d", i, d->vox_class[i]);
    {
        std::vector<char> used((size_t)d->n_vox * 6, 0);
        for (int j = 0; j < d->n_bond; j++) {
            int a = d->bond_axis[j], v1 = d->bond_v1[j], v2 = d->bond_v2[j];
            if (a < 0 || a > 2 || v1 < 0 || v2 < 0 || v1 >= d->n_vox || v2 >= d->n_vox || v1 == v2) return fail(VXG_ERR_ARG, "vxg_upload: bond %d malformed", j);
            if (d->bond_class[j] < 0 || d->bond_class[j] >= d->n_bclass) return fail(VXG_ERR_ARG, "vxg_upload: bond %d has class %d", j, d->bond_class[j]);
            if (used[(size_t)v1 * 6 + 2 * a] || used[(size_t)v2 * 6 + 2 * a + 1]) return fail(VXG_ERR_ARG, "vxg_upload: bond %d duplicates a voxel face", j);
            used[(size_t)v1 * 6 + 2 * a] = 1; used[(size_t)v2 * 6 + 2 * a + 1] = 1;
        }
    }
    for (int i = 0; i < d->n_vclass; i++) {
        const vxg_vox_class& c = d->vclasses[i];
        if (!(c.mass > 0) || !(c.inertia > 0) || !(c.nom_size > 0)) return fail(VXG_ERR_ARG, "vxg_upload: voxel class %d has non-positive mass/inertia/size", i);
    }
    // bond classes re-keyed by the voxel classes they join (one rest-length deviation per class and step in the compact family);
    // when that would outgrow the table the robot keeps its classes and the batch runs the wide family
    RekeyedBonds rk;
    rekey_bond_classes(d, b->lim.max_bclass, rk);
    const vxg_robot_desc* d_in = d;
    if (rk.ok) {
        d = &rk.desc;
        memset(b->h_bpair + (size_t)robot * b->lim.max_bclass, 0, (size_t)b->lim.max_bclass * sizeof(uint32_t));
        memcpy(b->h_bpair + (size_t)robot * b->lim.max_bclass, rk.pair.data(), rk.pair.size() * sizeof(uint32_t));
    }
    if (b->prec == VXG_PREC_F64) { fill_classes<double>(b, robot, d); reset_blob<double>(b, robot, d); }
    else { fill_classes<float>(b, robot, d); reset_blob<float>(b, robot, d); }
    init_ctl(d, b->h_ctl[robot]);
    if (b->lim.max_col_pairs > 0) {   // broad-phase tables of the self-collision list (CalcNearby, surface voxels)
        std::vector<int16_t> surf; std::vector<int8_t> lat4; std::vector<unsigned long long> near;
        const int hops = (int)(P.col_horizon * 1.5);
        if ((P.flags & VXG_F_COLLISIONS) && (!(P.col_horizon > 0) || !pack_collision_tables(d, hops, surf, lat4, near)))
            return fail(VXG_UNSUPPORTED_FEATURE, "vxg_upload: CollisionHorizon %g needs more than 3 bond hops (or the lattice is wider than 120 cells)", P.col_horizon);
        if (P.flags & VXG_F_COLLISIONS) {
            const size_t nv = (size_t)b->L.nv, r = (size_t)robot;
            memcpy(b->h_col_surf + r * nv, surf.data(), surf.size() * sizeof(int16_t));
            memcpy(b->h_col_lat4 + r * nv * 4, lat4.data(), lat4.size());
            memcpy(b->h_col_near + r * nv, near.data(), near.size() * sizeof(unsigned long long));
            b->h_ctl[robot].n_surf = (int)surf.size();
        }
    }
    if (b->L.nvert_cap > 0) {   // surface-mesh tables (also for non-fluid robots of a fluid batch: cheap, keeps one code path)
        int nfc = 0;
        int nvx = b->prec == VXG_PREC_F64 ? pack_mesh<double>(d, b->L, b->h_blobs + (size_t)robot * b->L.total_bytes, &nfc)
                                          : pack_mesh<float>(d, b->L, b->h_blobs + (size_t)robot * b->L.total_bytes, &nfc);
        if (nvx == -2) return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: robot %d has more surface facets than the batch's mesh capacity holds (raise limits.max_mesh_verts)", robot);
        b->h_ctl[robot].n_facet = nfc;
        if (nvx < 0) return fail(VXG_UNSUPPORTED_SIZE, "vxg_upload: robot %d has more than max_mesh_verts = %d surface-mesh vertices", robot, b->lim.max_mesh_verts);
        b->h_ctl[robot].n_vert = nvx;
    }
    {
        double len = 1.0;
        bool ok = b->prec == VXG_PREC_F64 ? fill_compact<double>(b, robot, d, &len) : fill_compact<float>(b, robot, d, &len);
        b->h_ctl[robot].compact_ok = (ok && rk.ok) ? 1 : 0;
        b->h_ctl[robot].bond_len = len;
    }
    d = d_in;
    b->bond_v1[robot].assign(d->bond_v1, d->bond_v1 + d->n_bond);
    b->bond_v2[robot].assign(d->bond_v2, d->bond_v2 + d->n_bond);
    b->bond_axis[robot].assign(d->bond_axis, d->bond_axis + d->n_bond);
    b->uploaded[robot] = 1; b->reset_image[robot] = 2; b->slot_dirty[robot] = 1;
    b->dirty = true;
    return VXG_OK;
}

// Device-side half of ResetSimulation (VX_Sim.cpp:839-877, VXS_Voxel.cpp:92-128, VXS_BondInternal.cpp:40-54) for slots whose
// host mirror is a plain reset image: only the words that differ per robot travel over PCIe (nominal positions and the
// static section); Angle = identity, zero momenta / damping history / accumulators, SmallAngle = true and the
// NominalPosition table are written here.  One CTA per robot slot.
template <typename Real>
__global__ void __launch_bounds__(256) vx_reset_kernel(unsigned char* blobs, BlobLayout L, const RobotCtl* ctl, double* nom) {
    const int robot = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    unsigned char* blob = blobs + (size_t)robot * L.total_bytes;
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* a = (uint4*)(blob + L.off_quat);
    for (size_t i = tid; i < ((L.dyn_bytes - L.off_quat) >> 4); i += T) a[i] = z;
    uint4* c = (uint4*)(blob + L.stat_bytes);
    for (size_t i = tid; i < ((L.total_bytes - L.stat_bytes) >> 4); i += T) c[i] = z;
    __syncthreads();
    RobotView<Real> R = make_view<Real>(blob, L, 0);
    int n_vox = ctl[robot].n_vox;
    if (n_vox > L.nv) n_vox = L.nv;
    for (int v = tid; v < n_vox; v += T) { R.quat[v].x = (Real)1; R.flags[v] = VX_FL_SMALL_ALL; }
    double* nm = nom + (size_t)robot * 3 * L.nv;
    for (int i = tid; i < 3 * L.nv; i += T) nm[i] = R.pos[i];
}

// Dynamic state, host mirrors -> device: the slots the host changed since the last flush only (slot_dirty); the others keep
// what the device computed.  Whole-population case (every slot a fresh reset image -- the GA's evaluate_all): positions +
// static section in two strided copies, the rest of the reset image is written by vx_reset_kernel.
static int flush_state(vxg_batch* b) {
    size_t nr = (size_t)b->n_robots, mv = (size_t)b->lim.max_vclass, nv = (size_t)b->L.nv;
    size_t n_dirty = 0; bool all_reset = true;
    for (size_t r = 0; r < nr; r++) if (b->slot_dirty[r]) { n_dirty++; if (b->uploaded[r] && !b->reset_image[r]) all_reset = false; }
    b->last_upload_bytes = 0;
    const size_t tb = b->L.total_bytes;
    if (n_dirty == nr && all_reset) {
        const size_t pos_bytes = 3 * nv * sizeof(double), st_bytes = b->L.stat_bytes - b->L.off_stat;
        CK(cudaMemcpy2DAsync(b->d_blobs + b->L.off_pos, tb, b->h_blobs + b->L.off_pos, tb, pos_bytes, nr, cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpy2DAsync(b->d_blobs + b->L.off_stat, tb, b->h_blobs + b->L.off_stat, tb, st_bytes, nr, cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_ctl, b->h_ctl, nr * sizeof(RobotCtl), cudaMemcpyHostToDevice, b->stream));
        if (b->prec == VXG_PREC_F64) vx_reset_kernel<double><<<(unsigned)nr, 256, 0, b->stream>>>(b->d_blobs, b->L, b->d_ctl, b->d_nom);
        else vx_reset_kernel<float><<<(unsigned)nr, 256, 0, b->stream>>>(b->d_blobs, b->L, b->d_ctl, b->d_nom);
        CK(cudaGetLastError());
        b->launches++;
        CK(cudaMemcpyAsync(b->d_scale_c, b->h_scale_c, nr * mv * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        b->last_upload_bytes += nr * (pos_bytes + st_bytes);
    } else {
        for (size_t r = 0; r < nr; r++) {
            if (!b->slot_dirty[r]) continue;
            if (b->uploaded[r]) complete_mirror(b, r);
            CK(cudaMemcpyAsync(b->d_blobs + r * tb, b->h_blobs + r * tb, tb, cudaMemcpyHostToDevice, b->stream));
            CK(cudaMemcpyAsync(b->d_nom + r * 3 * nv, b->h_nom + r * 3 * nv, 3 * nv * sizeof(double), cudaMemcpyHostToDevice, b->stream));
            CK(cudaMemcpyAsync(b->d_scale_c + r * mv, b->h_scale_c + r * mv, mv * sizeof(double), cudaMemcpyHostToDevice, b->stream));
            CK(cudaMemcpyAsync(b->d_ctl + r, b->h_ctl + r, sizeof(RobotCtl), cudaMemcpyHostToDevice, b->stream));
            b->last_upload_bytes += tb + 3 * nv * sizeof(double);
        }
    }
    return VXG_OK;
}

// make the device copy current (host mirrors -> device) on the batch stream
int vxg_i_flush(vxg_batch* b) {
    CK(cudaSetDevice(b->device));
    size_t nr = (size_t)b->n_robots, mv = (size_t)b->lim.max_vclass, mb = (size_t)b->lim.max_bclass, nv = (size_t)b->L.nv;
    if (b->dirty || b->tables_dirty) {
        // compact bond tables only if the production kernel runs and every robot of the batch qualifies
        bool compact = b->prec == VXG_PREC_F32 && b->smem_mode_wide_or_compact && !b->d_bond_rec;   // recording batches run the wide family
        for (size_t r = 0; r < nr && compact; r++) if (b->uploaded[r] && !b->h_ctl[r].compact_ok) compact = false;
        b->compact = compact;
        // stride 1024 (10^3 robots): the production family runs 1024 threads x 1 voxel (32 warps, 64 registers: measured 3 % faster
        // than 512 x 2 at 128 registers -- more warps hide the barrier and shared-memory latency); the other families keep 512 x 2
        if (b->L.nv == 1024) { b->vpt = (compact && b->lim.max_col_pairs == 0 && b->L.nvert_cap == 0) ? 1 : 2; b->threads = b->L.nv / b->vpt; }
        if (compact) { b->smem_mode = true; b->smem_bytes = b->L.total_bytes + b->tb_compact; }
        else {
            const bool f64_extras = b->prec == VXG_PREC_F64 && (b->L.nvert_cap > 0 || b->lim.max_col_pairs > 0);   // global-memory path only
            b->smem_mode = b->smem_wide_fits && !f64_extras;
            b->smem_bytes = b->smem_mode ? b->L.total_bytes + b->tb_wide : b->tb_wide;
        }
        if (b->dirty) { int rc = flush_state(b); if (rc) return rc; }
        if (compact) CK(cudaMemcpyAsync(b->d_bk, b->h_bkc, nr * mb * b->bkc_sz(), cudaMemcpyHostToDevice, b->stream));
        if (compact) CK(cudaMemcpyAsync(b->d_bpair, b->h_bpair, nr * mb * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
        else CK(cudaMemcpyAsync(b->d_bk, b->h_bk, nr * mb * b->bk_sz(), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_bk_dist, b->h_bk_dist, nr * mb * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_vk, b->h_vk, nr * mv * b->vk_sz(), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_sk, b->h_sk, nr * mv * sizeof(ScaleK), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_ext_disp, b->h_ext_disp, nr * mv * 3 * b->rs(), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->d_ext_tdisp, b->h_ext_tdisp, nr * mv * 3 * b->rs(), cudaMemcpyHostToDevice, b->stream));
        if (b->lim.max_col_pairs > 0 && b->dirty) {
            // Contact-bond lists are device-managed state (the rebuild writes them, only vxg_set_collision_pairs pulls them): only the
            // rows of the slots the host changed travel, the other robots' live lists stay as the device left them.
            const size_t mc = (size_t)b->lim.max_col_pairs;
            size_t n_dirty = 0;
            for (size_t r = 0; r < nr; r++) n_dirty += b->slot_dirty[r] ? 1 : 0;
            if (n_dirty == nr) {   // whole population (the GA's evaluate_all): four bulk copies
                CK(cudaMemcpyAsync(b->d_col_pairs, b->h_col_pairs, nr * mc * sizeof(ColPair), cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_surf, b->h_col_surf, nr * nv * sizeof(int16_t), cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_lat4, b->h_col_lat4, nr * nv * 4, cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_near, b->h_col_near, nr * nv * sizeof(unsigned long long), cudaMemcpyHostToDevice, b->stream));
            } else for (size_t r = 0; r < nr; r++) {
                if (!b->slot_dirty[r]) continue;
                CK(cudaMemcpyAsync(b->d_col_pairs + r * mc, b->h_col_pairs + r * mc, mc * sizeof(ColPair), cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_surf + r * nv, b->h_col_surf + r * nv, nv * sizeof(int16_t), cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_lat4 + r * nv * 4, b->h_col_lat4 + r * nv * 4, nv * 4, cudaMemcpyHostToDevice, b->stream));
                CK(cudaMemcpyAsync(b->d_col_near + r * nv, b->h_col_near + r * nv, nv * sizeof(unsigned long long), cudaMemcpyHostToDevice, b->stream));
            }
        }
    }
    if (b->ctl_dirty) CK(cudaMemcpyAsync(b->d_ctl, b->h_ctl, nr * sizeof(RobotCtl), cudaMemcpyHostToDevice, b->stream));   // the setters pulled it first
    if (b->dirty) std::fill(b->slot_dirty.begin(), b->slot_dirty.end(), 0);
    b->dirty = false; b->ctl_dirty = false; b->tables_dirty = false;
    return VXG_OK;
}

// device -> host mirrors (control blocks always; blobs/scales on request); synchronises the stream
static int pull(vxg_batch* b, bool blobs, int robot) {
    CK(cudaSetDevice(b->device));
    size_t nr = (size_t)b->n_robots, mv = (size_t)b->lim.max_vclass;
    // pending host-side changes go down first: dirtiness is per slot, so the device may still hold the newer state of the others
    // (and so do control-block edits of an earlier setter: two setters in a row must not lose the first one's values)
    if (b->dirty || b->ctl_dirty) { int rc = vxg_i_flush(b); if (rc) return rc; }
    CK(cudaMemcpyAsync(b->h_ctl, b->d_ctl, nr * sizeof(RobotCtl), cudaMemcpyDeviceToHost, b->stream));
    if (blobs) {
        size_t r0 = robot < 0 ? 0 : (size_t)robot, cnt = robot < 0 ? nr : 1;
        CK(cudaMemcpyAsync(b->h_blobs + r0 * b->L.total_bytes, b->d_blobs + r0 * b->L.total_bytes, cnt * b->L.total_bytes, cudaMemcpyDeviceToHost, b->stream));
        for (size_t r = r0; r < r0 + cnt; r++) b->reset_image[r] = 0;
        CK(cudaMemcpyAsync(b->h_scale_c + r0 * mv, b->d_scale_c + r0 * mv, cnt * mv * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    }
    CK(cudaStreamSynchronize(b->stream));
    return VXG_OK;
}

extern "C" int vxg_set_step_budget(vxg_batch* b, const int64_t* total_steps, const int64_t* latch_step) {
    if (!b) return fail(VXG_ERR_ARG, "vxg_set_step_budget: null batch");
    int rc = pull(b, false, -1);   // keep the device's time/step/status
    if (rc) return rc;
    for (int r = 0; r < b->n_robots; r++) {
        if (!b->uploaded[r]) continue;
        RobotCtl& c = b->h_ctl[r];
        if (total_steps) {
            c.total_steps = total_steps[r] < 0 ? LLONG_MAX : total_steps[r];
            if (c.step >= c.total_steps) c.status |= VXG_ST_FINISHED; else c.status &= ~VXG_ST_FINISHED;
        }
        if (latch_step) c.latch_step = latch_step[r];
    }
    b->ctl_dirty = true;
    return VXG_OK;
}

extern "C" int vxg_set_eol_step(vxg_batch* b, const int64_t* eol_step) {
    if (!b || !eol_step) return fail(VXG_ERR_ARG, "vxg_set_eol_step: null argument");
    int rc = pull(b, false, -1);
    if (rc) return rc;
    for (int r = 0; r < b->n_robots; r++) if (b->uploaded[r]) b->h_ctl[r].eol_step = eol_step[r];
    b->ctl_dirty = true;
    return VXG_OK;
}

extern "C" int vxg_enable_bond_record(vxg_batch* b, int on) {
    if (!b) return fail(VXG_ERR_ARG, "vxg_enable_bond_record: null batch");
    CK(cudaSetDevice(b->device));
    if (on && !b->d_bond_rec) {
        size_t bytes = (size_t)b->n_robots * 39 * b->L.nv * b->rs();
        CK(vx_alloc(&b->d_bond_rec, bytes, b->device));
        CK(cudaMemset(b->d_bond_rec, 0, bytes));
        b->tables_dirty = true;   // the production family does not record: the wide tables travel before the next step
    } else if (!on && b->d_bond_rec) {
        CK(cudaStreamSynchronize(b->stream));
        vx_free(b->d_bond_rec, b->device); b->d_bond_rec = nullptr;
        b->tables_dirty = true;
    }
    return VXG_OK;
}

extern "C" int vxg_step(vxg_batch* b, int n_steps) {
    if (!b) return fail(VXG_ERR_ARG, "vxg_step: null batch");
    if (n_steps < 0) return fail(VXG_ERR_ARG, "vxg_step: negative step count");
    for (int r = 0; r < b->n_robots; r++)
        if (!b->uploaded[r]) return fail(VXG_ERR_STATE, "vxg_step: robot slot %d was never uploaded", r);
    int rc = vxg_i_flush(b);
    if (rc) return rc;
    CK(cudaEventRecord(b->ev0, b->stream));
    // steps per launch: the state stays on chip inside a launch, and a launch is the granularity at which the CTAs of the other batches /
    // chunk streams of a population get the SMs.  Measured with 8 chunk streams: 1000 instead of 4096 steps per launch packs the waves of
    // small robots better (6^3 x 1024 end to end 1.69e10 -> 1.90e10) at no cost for large ones.
    const int chunk = 1000;
    int left = n_steps;
    while (left > 0) {
        int n = left < chunk ? left : chunk;
        int record = (b->d_bond_rec && n == left) ? 1 : 0;
        cudaError_t e = launch_step_p(b, n, record);
        if (e != cudaSuccess) return fail(VXG_ERR_CUDA, "vx_step_kernel launch: %s", cudaGetErrorString(e));
        b->launches++;
        left -= n;
    }
    CK(cudaEventRecord(b->ev1, b->stream));
    b->timed = true;
    return VXG_OK;
}

extern "C" int vxg_sync(vxg_batch* b, uint32_t* status) {
    if (!b) return fail(VXG_ERR_ARG, "vxg_sync: null batch");
    int rc = pull(b, false, -1);
    if (rc) return rc;
    if (status) for (int r = 0; r < b->n_robots; r++) status[r] = b->h_ctl[r].status;
    return VXG_OK;
}

extern "C" int vxg_last_step_ms(vxg_batch* b, float* ms) {
    if (!b || !ms) return fail(VXG_ERR_ARG, "vxg_last_step_ms: null argument");
    if (!b->timed) return fail(VXG_ERR_STATE, "vxg_last_step_ms: no step has been issued");
    CK(cudaSetDevice(b->device));
    CK(cudaEventSynchronize(b->ev1));
    CK(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return VXG_OK;
}

extern "C" int64_t vxg_launch_count(vxg_batch* b) { return b ? b->launches : 0; }

extern "C" int vxg_mark(vxg_batch* b, int slot) {
    if (!b || slot < 0 || slot >= VXG_N_MARKS) return fail(VXG_ERR_ARG, "vxg_mark: bad argument");
    CK(cudaSetDevice(b->device));
    CK(cudaEventRecord(b->marks[slot], b->stream));
    b->mark_set[slot] = true;
    return VXG_OK;
}

extern "C" int vxg_mark_elapsed_ms(vxg_batch* b, int s0, int s1, float* ms) {
    if (!b || !ms || s0 < 0 || s1 < 0 || s0 >= VXG_N_MARKS || s1 >= VXG_N_MARKS) return fail(VXG_ERR_ARG, "vxg_mark_elapsed_ms: bad argument");
    if (!b->mark_set[s0] || !b->mark_set[s1]) return fail(VXG_ERR_STATE, "vxg_mark_elapsed_ms: marker not recorded");
    CK(cudaSetDevice(b->device));
    CK(cudaEventSynchronize(b->marks[s1]));
    CK(cudaEventElapsedTime(ms, b->marks[s0], b->marks[s1]));
    return VXG_OK;
}

extern "C" int vxg_kernel_config(vxg_batch* b, int* threads, int* vpt, int* smem_mode, int64_t* smem_bytes, int64_t* blob_bytes, int64_t* dyn_bytes) {
    if (!b) return fail(VXG_ERR_ARG, "vxg_kernel_config: null batch");
    {   // the table form (wide / compact) and with it the shared-memory size is decided when the classes are flushed
        bool all = true;
        for (int r = 0; r < b->n_robots; r++) all = all && b->uploaded[r];
        if (all) { int rc = vxg_i_flush(b); if (rc) return rc; }
    }
    if (threads) *threads = b->threads;
    if (vpt) *vpt = b->vpt;
    if (smem_mode) *smem_mode = b->smem_mode ? 1 : 0;
    if (smem_bytes) *smem_bytes = (int64_t)b->smem_bytes;
    if (blob_bytes) *blob_bytes = (int64_t)b->L.total_bytes;
    if (dyn_bytes) *dyn_bytes = (int64_t)b->L.dyn_bytes;
    return VXG_OK;
}

template <typename Real> static void unpack_state(vxg_batch* b, int robot, vxg_robot_state* o) {
    const RobotCtl& c = b->h_ctl[robot];
    RobotView<Real> R = make_view<Real>(b->h_blobs + (size_t)robot * b->L.total_bytes, b->L, c.n_vox);
    const int nv = R.nv;
    const double* sc = b->h_scale_c + (size_t)robot * b->lim.max_vclass;
    for (int i = 0; i < c.n_vox; i++) {
        const R4<Real> q = R.quat[i], ma = R.momA[i]; const R2<Real> mb = R.momB[i];
        if (o->pos) for (int a = 0; a < 3; a++) o->pos[3 * i + a] = R.pos[a * nv + i];
        if (o->quat) { o->quat[4 * i] = (double)q.x; o->quat[4 * i + 1] = (double)q.y; o->quat[4 * i + 2] = (double)q.z; o->quat[4 * i + 3] = (double)q.w; }
        if (o->lin_mom) { o->lin_mom[3 * i] = (double)ma.x; o->lin_mom[3 * i + 1] = (double)ma.y; o->lin_mom[3 * i + 2] = (double)ma.z; }
        if (o->ang_mom) { o->ang_mom[3 * i] = (double)ma.w; o->ang_mom[3 * i + 1] = (double)mb.x; o->ang_mom[3 * i + 2] = (double)mb.y; }
        if (o->scale) o->scale[i] = sc[R.stat[i].vcls];
        if (o->static_fric) o->static_fric[i] = (R.flags[i] & VX_FL_STATIC_FRIC) ? 1 : 0;
        if (o->vox_strain) for (int a = 0; a < 6; a++) o->vox_strain[6 * i + a] = b->L.nvert_cap > 0 ? (double)R.strain[a * nv + i] : 0.0;
    }
}

extern "C" int vxg_download_state(vxg_batch* b, int robot, vxg_robot_state* o) {
    if (!b || !o) return fail(VXG_ERR_ARG, "vxg_download_state: null argument");
    if (robot < 0 || robot >= b->n_robots || !b->uploaded[robot]) return fail(VXG_ERR_ARG, "vxg_download_state: robot %d not present", robot);
    int rc = pull(b, true, robot);
    if (rc) return rc;
    if (b->prec == VXG_PREC_F64) unpack_state<double>(b, robot, o); else unpack_state<float>(b, robot, o);
    const int nb = (int)b->bond_v1[robot].size();
    const bool want_rec = o->bond_force1 || o->bond_force2 || o->bond_moment1 || o->bond_moment2 || o->bond_strain;
    if (want_rec && nb > 0) {
        if (!b->d_bond_rec || b->dirty) return fail(VXG_ERR_STATE, "vxg_download_state: bond outputs need vxg_enable_bond_record before the step");
        const int nv = b->L.nv;
        size_t bytes = (size_t)39 * nv * b->rs();
        std::vector<unsigned char> tmp(bytes);
        CK(cudaMemcpy(tmp.data(), (unsigned char*)b->d_bond_rec + (size_t)robot * bytes, bytes, cudaMemcpyDeviceToHost));
        for (int j = 0; j < nb; j++) {
            int a = b->bond_axis[robot][j], v = b->bond_v1[robot][j];
            double r[13];
            for (int cidx = 0; cidx < 13; cidx++) {
                size_t idx = (size_t)(a * 13 + cidx) * nv + v;
                r[cidx] = b->prec == VXG_PREC_F64 ? ((double*)tmp.data())[idx] : (double)((float*)tmp.data())[idx];
            }
            for (int x = 0; x < 3; x++) {
                if (o->bond_force1) o->bond_force1[3 * j + x] = r[x];
                if (o->bond_force2) o->bond_force2[3 * j + x] = r[3 + x];
                if (o->bond_moment1) o->bond_moment1[3 * j + x] = r[6 + x];
                if (o->bond_moment2) o->bond_moment2[3 * j + x] = r[9 + x];
            }
            if (o->bond_strain) o->bond_strain[j] = r[12];
        }
    }
    if (o->bond_small_angle) {
        const unsigned char* blob = b->h_blobs + (size_t)robot * b->L.total_bytes;
        const uint32_t* fl = (const uint32_t*)(blob + b->L.off_flags);   // flags plane is precision-independent
        for (int j = 0; j < nb; j++) o->bond_small_angle[j] = (fl[b->bond_v1[robot][j]] >> b->bond_axis[robot][j]) & 1u;
    }
    return VXG_OK;
}

template <typename Real> static void pack_state(vxg_batch* b, int robot, const vxg_robot_state* in) {
    const RobotCtl& c = b->h_ctl[robot];
    RobotView<Real> R = make_view<Real>(b->h_blobs + (size_t)robot * b->L.total_bytes, b->L, c.n_vox);
    const int nv = R.nv;
    for (int i = 0; i < c.n_vox; i++) {
        if (in->pos) for (int a = 0; a < 3; a++) R.pos[a * nv + i] = in->pos[3 * i + a];
        if (in->quat) { R.quat[i].x = (Real)in->quat[4 * i]; R.quat[i].y = (Real)in->quat[4 * i + 1]; R.quat[i].z = (Real)in->quat[4 * i + 2]; R.quat[i].w = (Real)in->quat[4 * i + 3]; }
        if (in->lin_mom) { R.momA[i].x = (Real)in->lin_mom[3 * i]; R.momA[i].y = (Real)in->lin_mom[3 * i + 1]; R.momA[i].z = (Real)in->lin_mom[3 * i + 2]; }
        if (in->ang_mom) { R.momA[i].w = (Real)in->ang_mom[3 * i]; R.momB[i].x = (Real)in->ang_mom[3 * i + 1]; R.momB[i].y = (Real)in->ang_mom[3 * i + 2]; }
    }
}

extern "C" int vxg_upload_state(vxg_batch* b, int robot, const vxg_robot_state* in, double cur_time, int64_t cur_step) {
    if (!b || !in) return fail(VXG_ERR_ARG, "vxg_upload_state: null argument");
    if (robot < 0 || robot >= b->n_robots || !b->uploaded[robot]) return fail(VXG_ERR_ARG, "vxg_upload_state: robot %d not present", robot);
    int rc = pull(b, true, robot);   // this slot's mirror becomes current; the others keep running from their device state
    if (rc) return rc;
    if (b->lim.max_col_pairs > 0) {   // its live contact-bond list travels back with it
        const size_t mc = (size_t)b->lim.max_col_pairs;
        CK(cudaMemcpy(b->h_col_pairs + (size_t)robot * mc, b->d_col_pairs + (size_t)robot * mc, mc * sizeof(ColPair), cudaMemcpyDeviceToHost));
    }
    if (b->prec == VXG_PREC_F64) pack_state<double>(b, robot, in); else pack_state<float>(b, robot, in);
    b->reset_image[robot] = 0;
    b->slot_dirty[robot] = 1;
    b->h_ctl[robot].cur_time = cur_time;
    b->h_ctl[robot].step = cur_step;
    b->dirty = true;
    return VXG_OK;
}

extern "C" int vxg_reduce_stats(vxg_batch* b, vxg_robot_stats* out) {
    if (!b || !out) return fail(VXG_ERR_ARG, "vxg_reduce_stats: null argument");
    int rc = vxg_i_flush(b);
    if (rc) return rc;
    if (b->prec == VXG_PREC_F64) launch_stats_f64(dev_view<double>(b), b->d_stats, b->n_robots, b->stream);
    else launch_stats_f32(dev_view<float>(b), b->d_stats, b->n_robots, b->stream);
    CK(cudaGetLastError());
    b->launches++;
    CK(cudaMemcpyAsync(b->h_stats, b->d_stats, (size_t)b->n_robots * sizeof(StatsDev), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    memcpy(out, b->h_stats, (size_t)b->n_robots * sizeof(StatsDev));
    return VXG_OK;
}

extern "C" int vxg_set_collision_pairs(vxg_batch* b, int robot, int n_pairs, const int32_t* v1, const int32_t* v2, const double* a1) {
    (void)a1;
    if (!b) return fail(VXG_ERR_ARG, "vxg_set_collision_pairs: null batch");
    if (robot < 0 || robot >= b->n_robots || !b->uploaded[robot]) return fail(VXG_ERR_ARG, "vxg_set_collision_pairs: robot %d not present", robot);
    if (b->lim.max_col_pairs <= 0) return fail(VXG_UNSUPPORTED_SIZE, "vxg_set_collision_pairs: the batch was created with limits.max_col_pairs = 0");
    if (n_pairs > b->lim.max_col_pairs) return fail(VXG_UNSUPPORTED_SIZE, "vxg_set_collision_pairs: %d pairs > max_col_pairs %d", n_pairs, b->lim.max_col_pairs);
    if (n_pairs > 0 && (!v1 || !v2)) return fail(VXG_ERR_ARG, "vxg_set_collision_pairs: null pair arrays");
    int rc = pull(b, true, robot);   // keep the device state of this slot; the others are not touched
    if (rc) return rc;
    RobotCtl& c = b->h_ctl[robot];
    if (n_pairs < 0) { c.col_host = 0; c.col_first = 1; c.n_col = 0; }
    else {
        ColPair* p = b->h_col_pairs + (size_t)robot * b->lim.max_col_pairs;
        for (int i = 0; i < n_pairs; i++) {
            if (v1[i] < 0 || v2[i] < 0 || v1[i] >= c.n_vox || v2[i] >= c.n_vox || v1[i] == v2[i]) return fail(VXG_ERR_ARG, "vxg_set_collision_pairs: pair %d malformed", i);
            p[i].v1 = (int16_t)v1[i]; p[i].v2 = (int16_t)v2[i];
        }
        c.col_host = 1; c.n_col = n_pairs; c.flags |= VXG_F_COLLISIONS;
    }
    b->slot_dirty[robot] = 1;   // its mirror was pulled above, so it is current; the pair lists travel whole
    b->dirty = true;
    return VXG_OK;
}

extern "C" int vxg_mesh_vertex_count(const vxg_robot_desc* d) {
    if (!d || d->n_vox <= 0 || !d->nom_pos || !d->vox_class || !d->vclasses) return fail(VXG_ERR_ARG, "vxg_mesh_vertex_count: bad descriptor");
    return mesh_vertex_count(d);
}

// bytes one full population upload moves host -> device (what flush() copies when everything is dirty)
extern "C" int64_t vxg_upload_bytes(vxg_batch* b) {
    if (!b) return 0;
    size_t nr = (size_t)b->n_robots, mv = (size_t)b->lim.max_vclass, mb = (size_t)b->lim.max_bclass, nv = (size_t)b->L.nv;
    // state part: what the last flush copied (positions + static section when every slot was a fresh reset image, else whole blobs
    // + NominalPosition); before any flush, the fresh-reset figure
    size_t state = b->last_upload_bytes ? b->last_upload_bytes : nr * (3 * nv * sizeof(double) + (b->L.stat_bytes - b->L.off_stat));
    size_t per = mb * (b->compact ? b->bkc_sz() : b->bk_sz()) + mb * sizeof(double) + mv * b->vk_sz() + mv * sizeof(ScaleK) + mv * sizeof(double) +
                 2 * mv * 3 * b->rs() + sizeof(RobotCtl);
    return (int64_t)(state + nr * per);
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// One process per GPU shares the box's cores with its sibling ranks: LOCAL_WORLD_SIZE (torchrun) divides them, VXG_HOST_THREADS
// overrides.  (8 ranks x 32 flatten threads on a 32-core box was what held the 1 -> 8 GPU end-to-end curve at 0.84.)
int vxg_i_host_threads(int cap) {
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (const char* e = getenv("VXG_HOST_THREADS")) { int v = atoi(e); if (v > 0) nt = v; }
    else {
        int ranks = 1;
        if (const char* w = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(w);
        if (ranks > 1) nt = std::max(1, nt / ranks);
    }
    if (nt > 64) nt = 64;
    if (cap > 0 && nt > cap) nt = cap;
    return nt;
}

bool vxg_i_compact_ok(const vxg_robot_desc* d, int* n_bclass) {
    RekeyedBonds rk;
    rekey_bond_classes(d, 65535, rk);
    if (!rk.ok) return false;
    std::vector<BondKc<float>> tmp((size_t)std::max(1, rk.desc.n_bclass));
    double len = 1.0;
    if (!pack_compact<float>(&rk.desc, tmp.data(), &len)) return false;
    if (n_bclass) *n_bclass = rk.desc.n_bclass;
    return true;
}

int vxg_i_enqueue_positions(vxg_batch* b) {
    // positions only: first 3*nv doubles of every blob (strided device -> host copy)
    const size_t w = (size_t)3 * b->L.nv * sizeof(double);
    CK(cudaMemcpy2DAsync(b->h_blobs, b->L.total_bytes, b->d_blobs, b->L.total_bytes, w, (size_t)b->n_robots, cudaMemcpyDeviceToHost, b->stream));
    std::fill(b->reset_image.begin(), b->reset_image.end(), 0);
    CK(cudaMemcpyAsync(b->h_ctl, b->d_ctl, (size_t)b->n_robots * sizeof(RobotCtl), cudaMemcpyDeviceToHost, b->stream));
    return VXG_OK;
}

void vxg_i_unpack_positions(const vxg_batch* b, int r, double* o) {   // plane layout -> the caller's [voxel][xyz] rows
    const int nv = b->L.nv;
    const double* p = (const double*)(b->h_blobs + (size_t)r * b->L.total_bytes);
    const int n = b->h_ctl[r].n_vox;
    for (int i = 0; i < n; i++) { o[3 * i] = p[i]; o[3 * i + 1] = p[nv + i]; o[3 * i + 2] = p[2 * nv + i]; }
}

// flatten descriptors into the pinned mirrors of slots [0, n) on the host cores (slots are independent)
static int upload_all(vxg_batch* b, const vxg_robot_desc* descs, int n_descs) {
    const int nt = vxg_i_host_threads(n_descs);
    std::vector<int> rcs((size_t)nt, 0);
    std::vector<std::string> msgs((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            for (int r = t; r < n_descs; r += nt) {
                int rc = vxg_upload(b, r, &descs[r]);
                if (rc) { rcs[t] = rc; msgs[t] = g_err; return; }
            }
        });
    for (auto& x : th) x.join();
    for (int t = 0; t < nt; t++) if (rcs[t]) { g_err = msgs[t]; return rcs[t]; }
    return VXG_OK;
}

extern "C" int vxg_run_batch_host(vxg_batch* b, const vxg_robot_desc* descs, int n_descs, int n_steps, double* pos_out) {
    if (!b || !descs || n_descs != b->n_robots) return fail(VXG_ERR_ARG, "vxg_run_batch_host: need one descriptor per robot slot");
    const bool timing = getenv("VXG_TIMING") != nullptr;
    const double t0 = now_ms();
    int rc = upload_all(b, descs, n_descs);
    if (rc) return rc;
    const double t1 = now_ms();
    rc = vxg_step(b, n_steps);
    if (rc) return rc;
    if (timing) { cudaStreamSynchronize(b->stream); }
    const double t2 = now_ms();
    if (pos_out) {
        rc = vxg_i_enqueue_positions(b);
        if (rc) return rc;
        CK(cudaStreamSynchronize(b->stream));
        const int mv = b->lim.max_vox;
        for (int r = 0; r < b->n_robots; r++)
            if (b->h_ctl[r].status & VXG_ST_DIVERGED) return fail(VXG_ERR_STATE, "vxg_run_batch_host: robot %d diverged (strain > 100)", r);
        const int nt = vxg_i_host_threads(std::min(16, b->n_robots));
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t]() { for (int r = t; r < b->n_robots; r += nt) vxg_i_unpack_positions(b, r, pos_out + (size_t)r * mv * 3); });
        for (auto& x : th) x.join();
    } else CK(cudaStreamSynchronize(b->stream));
    if (timing) fprintf(stderr, "vxg_run_batch_host: flatten %.2f ms, upload+step %.2f ms, download+unpack %.2f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
    return VXG_OK;
}
