// vxg_population.cu -- a ragged population on one GPU (include/vxgpu.h, vxg_pop_*).
//
// A vxg_batch is homogeneous by construction: one voxel stride, one kernel family, one table size -- that is what lets a CTA
// stream a robot in and out as one contiguous blob and lets the kernels be compiled per stride.  A population the GA hands over
// is not: CPPN morphologies of 30..216 voxels, some with self-collision lists, some outside the production family.  Putting it
// into ONE batch makes every robot pay for the largest one's stride and the most general member's kernel family.  Here the
// robots are bucketed by what the kernels specialise on -- (voxel stride, self-collision, surface mesh, production / wide
// family, class-table size) -- and every bucket becomes its own batch on its own stream, so the buckets' grids run concurrently
// on the device, each at its own family's speed.  Buckets are further cut into chunks so that the host-buffer path pipelines:
// while chunk c is being simulated, chunk c+1 is flattened and uploaded and chunk c-1 downloaded and unpacked.
//
// Host code only (the device work is the batches').  Replaces evosoro's per-robot process fan-out (evaluation.py:89-90).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

#include "vxg_batch.h"
#include "vx_pack.h"

using namespace vx;
#define fail vxg_i_fail

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Sub {                    // one homogeneous batch of the population
    vxg_batch* b = nullptr;
    std::vector<int> robots;    // population index of every slot
    int stride = 0, family = 0;
    int64_t max_total = -1;     // largest step budget among its robots (-1 = none set)
};

int pow2_bin(int n, int lo) { int b = lo; while (b < n) b *= 2; return b; }

template <typename F> void parallel_for(int n, int nt, F f) {
    nt = std::max(1, std::min(nt, n));
    if (nt == 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& x : th) x.join();
}

}  // namespace

struct vxg_population {
    int device = 0, n = 0, prec = VXG_PREC_F32;
    std::vector<Sub> subs;
    std::vector<int> sub_of, slot_of;   // per population robot
    std::vector<int> n_vox;
};

extern "C" void vxg_pop_destroy(vxg_population* p) {
    if (!p) return;
    for (auto& s : p->subs) vxg_destroy(s.b);
    delete p;
}

extern "C" int vxg_pop_create(vxg_population** out, int device, const vxg_robot_desc* descs, int n_robots, const vxg_pop_options* opt_in) {
    if (!out || !descs || n_robots <= 0) return fail(VXG_ERR_ARG, "vxg_pop_create: bad argument");
    *out = nullptr;
    vxg_pop_options opt; memset(&opt, 0, sizeof opt);
    if (opt_in) opt = *opt_in;
    if (opt.precision != VXG_PREC_F32 && opt.precision != VXG_PREC_F64) return fail(VXG_ERR_ARG, "vxg_pop_create: unknown precision");
    const int per_vox = opt.col_pairs_per_vox > 0 ? opt.col_pairs_per_vox : 4;
    const int nt = vxg_i_host_threads(n_robots);
    const bool timing = getenv("VXG_TIMING") != nullptr;
    const double t0 = now_s();
    if (timing) { cudaSetDevice(device); cudaFree(nullptr); }   // context creation, timed on its own
    const double t_ctx = now_s();

    // ---- classify ---------------------------------------------------------------------------------------------------------
    struct Cls { int stride, collide, mesh, compact, vbin, bbin, nbc, nverts; };
    std::vector<Cls> cls((size_t)n_robots);
    std::vector<int> bad((size_t)n_robots, 0);
    parallel_for(n_robots, nt, [&](int r) {
        const vxg_robot_desc& d = descs[r];
        if (d.n_vox <= 0 || d.n_vclass <= 0 || !d.nom_pos || !d.vox_class || !d.vclasses) { bad[r] = 1; return; }
        Cls c{};
        c.stride = stride_for(d.n_vox);
        c.collide = (d.sim.flags & VXG_F_COLLISIONS) ? 1 : 0;
        c.mesh = ((d.sim.flags & VXG_F_FLUID) || (opt.want_mesh && (d.sim.flags & VXG_F_LW_SCALE))) ? 1 : 0;
        int nbc = d.n_bclass;
        c.compact = (opt.precision == VXG_PREC_F32 && vxg_i_compact_ok(&d, &nbc)) ? 1 : 0;
        c.nbc = std::max(1, c.compact ? nbc : d.n_bclass);
        c.vbin = pow2_bin(d.n_vclass, 8); c.bbin = pow2_bin(c.nbc, 32);
        c.nverts = c.mesh ? vxg_mesh_vertex_count(&d) : 0;
        cls[r] = c;
    });
    for (int r = 0; r < n_robots; r++) if (bad[r]) return fail(VXG_ERR_ARG, "vxg_pop_create: robot %d has a malformed descriptor", r);
    // Small robots (strides up to 256) of one family share ONE stride, the largest any of them needs.  Measured (1024 CPPN-like
    // robots of 75 .. 161 voxels, SelfColEnabled=1): one bucket per 32-voxel stride 6.1e9 voxel-timesteps/s, strides 128 + 224
    // 7.7e9, everything at stride 192 9.5e9 -- every stride is its own kernel image, and several images resident on an SM at once
    // evict each other from the instruction cache (the step loop is ~25 KB of code; ncu: stall_no_instruction).  A CTA's idle
    // warps cost no issue slots, so the wider stride is the smaller tax.  Robots beyond 256 voxels keep a bucket per stride: there
    // the shared memory a robot occupies decides how many are resident.
    // (VXG_RT_STRIDE=1: the gather family's run-time-stride image instead -- one image, exact strides; measured slower, vxg_gather.cuh)
    const char* rt_env = getenv("VXG_RT_STRIDE");
    const bool rt = rt_env && rt_env[0] == '1';
    if (!getenv("VXG_POP_FINE_STRIDES")) {
        std::map<std::tuple<int, int, int>, int> widest;   // (collide, mesh, compact) -> widest small stride
        for (int r = 0; r < n_robots; r++) if (cls[r].stride <= 256) { int& w = widest[std::make_tuple(cls[r].collide, cls[r].mesh, cls[r].compact)]; w = std::max(w, cls[r].stride); }
        for (int r = 0; r < n_robots; r++) {
            if (cls[r].stride > 256) continue;
            if (rt && cls[r].compact) cls[r].stride = std::max(64, (descs[r].n_vox + 31) & ~31);
            else cls[r].stride = widest[std::make_tuple(cls[r].collide, cls[r].mesh, cls[r].compact)];
        }
    }

    const double t_cls = now_s();
    // ---- bucket, chunk ------------------------------------------------------------------------------------------------------
    typedef std::tuple<int, int, int, int, int, int> Key;
    std::map<Key, std::vector<int>> buckets;
    for (int r = 0; r < n_robots; r++) {
        const Cls& c = cls[r];
        buckets[Key(c.stride, c.collide, c.mesh, c.compact, c.vbin, c.bbin)].push_back(r);
    }
    vxg_population* p = new vxg_population;
    p->device = device; p->n = n_robots; p->prec = opt.precision;
    p->sub_of.assign(n_robots, -1); p->slot_of.assign(n_robots, -1); p->n_vox.resize(n_robots);
    for (int r = 0; r < n_robots; r++) p->n_vox[r] = descs[r].n_vox;
    for (auto& kv : buckets) {
        const std::vector<int>& rs = kv.second;
        const int want = opt.chunk_robots > 0 ? opt.chunk_robots : (int)rs.size();
        // even chunks (a last chunk of a handful of robots would idle most SMs at the tail)
        const int n_chunks = ((int)rs.size() + want - 1) / want;
        const int base = (int)rs.size() / n_chunks, rem = (int)rs.size() % n_chunks;
        size_t o = 0;
        for (int ci = 0; ci < n_chunks; ci++) {
            const size_t chunk = (size_t)(base + (ci < rem ? 1 : 0));
            Sub s;
            s.robots.assign(rs.begin() + o, rs.begin() + o + chunk);
            o += chunk;
            s.stride = std::get<0>(kv.first);
            s.family = std::get<3>(kv.first) | (std::get<1>(kv.first) << 1) | (std::get<2>(kv.first) << 2);
            vxg_limits lim; memset(&lim, 0, sizeof lim);
            lim.precision = opt.precision; lim.max_vox = lim.max_vclass = lim.max_bclass = 1;
            lim.reserved[0] = s.stride;   // the bucket's stride (may be wider than its own largest robot needs, see above)
            int col_cap = 0;
            for (int r : s.robots) {
                lim.max_vox = std::max(lim.max_vox, descs[r].n_vox);
                lim.max_vclass = std::max(lim.max_vclass, descs[r].n_vclass);
                lim.max_bclass = std::max(lim.max_bclass, cls[r].nbc);
                lim.max_mesh_verts = std::max(lim.max_mesh_verts, cls[r].nverts);
                if (cls[r].collide) col_cap = std::max(col_cap, opt.col_pairs_per_vox < 0 ? std::min(16384, -opt.col_pairs_per_vox) : std::min(16384, std::max(512, per_vox * descs[r].n_vox)));
            }
            int rc;
            for (;;) {   // the contact-bond list is shared-memory resident next to the robot: halve the capacity until both fit
                lim.max_col_pairs = col_cap;
                rc = vxg_create(&s.b, device, (int)s.robots.size(), &lim);
                if (rc) break;
                const bool resident = s.b->smem_wide_fits || (cls[s.robots[0]].compact && s.b->smem_mode_wide_or_compact);
                if (resident || col_cap <= 512) break;   // (at 512 and still too large: the batch runs from global memory)
                vxg_destroy(s.b); s.b = nullptr;
                col_cap /= 2;
            }
            if (rc) { vxg_pop_destroy(p); return rc; }
            for (size_t i = 0; i < s.robots.size(); i++) { p->sub_of[s.robots[i]] = (int)p->subs.size(); p->slot_of[s.robots[i]] = (int)i; }
            p->subs.push_back(std::move(s));
        }
    }
    // ---- upload ---------------------------------------------------------------------------------------------------------------
    const double t_alloc = now_s();
    int rc = vxg_pop_upload(p, descs);
    if (rc) { vxg_pop_destroy(p); return rc; }
    if (timing) fprintf(stderr, "vxg_pop_create: %d robots, %d batches: CUDA context %.3f s, classify + bucket %.3f s, batches (device + pinned allocations) %.3f s, flatten %.3f s (%d host threads)\n",
                        n_robots, (int)p->subs.size(), t_ctx - t0, t_cls - t_ctx, t_alloc - t_cls, now_s() - t_alloc, nt);
    *out = p;
    return VXG_OK;
}

static int upload_sub(vxg_population* p, Sub& s, const vxg_robot_desc* descs, int nt) {
    std::atomic<int> err(0);
    std::string msg;
    std::atomic_flag lock = ATOMIC_FLAG_INIT;
    parallel_for((int)s.robots.size(), nt, [&](int i) {
        if (err.load()) return;
        int rc = vxg_upload(s.b, i, &descs[s.robots[i]]);
        if (rc) { while (lock.test_and_set()) {} if (!err.load()) { err = rc; msg = vxg_last_error(); } lock.clear(); }
    });
    (void)p;
    if (err.load()) return fail(err.load(), "%s", msg.c_str());
    return VXG_OK;
}

extern "C" int vxg_pop_upload(vxg_population* p, const vxg_robot_desc* descs) {
    if (!p || !descs) return fail(VXG_ERR_ARG, "vxg_pop_upload: null argument");
    const int nt = vxg_i_host_threads(0);
    for (auto& s : p->subs) { int rc = upload_sub(p, s, descs, nt); if (rc) return rc; }
    return VXG_OK;
}

extern "C" int vxg_pop_set_plan(vxg_population* p, const int64_t* total, const int64_t* latch, const int64_t* eol) {
    if (!p) return fail(VXG_ERR_ARG, "vxg_pop_set_plan: null population");
    for (auto& s : p->subs) {
        const size_t m = s.robots.size();
        std::vector<int64_t> t(m), l(m), e(m);
        s.max_total = -1;
        for (size_t i = 0; i < m; i++) {
            const int r = s.robots[i];
            t[i] = total ? total[r] : -1; l[i] = latch ? latch[r] : -1; e[i] = eol ? eol[r] : -1;
            if (total) s.max_total = std::max(s.max_total, total[r] < 0 ? (int64_t)LLONG_MAX : total[r]);
        }
        int rc = vxg_set_step_budget(s.b, total ? t.data() : nullptr, latch ? l.data() : nullptr);
        if (!rc && eol) rc = vxg_set_eol_step(s.b, e.data());
        if (rc) return rc;
    }
    return VXG_OK;
}

extern "C" int vxg_pop_step(vxg_population* p, int64_t n_steps) {
    if (!p || n_steps < 0) return fail(VXG_ERR_ARG, "vxg_pop_step: bad argument");
    // round-robin over the batches in slices, so that their launches interleave on the device instead of one batch queueing
    // its whole lifetime before the next one starts
    const int64_t slice = 1 << 16;
    for (int64_t done = 0; done < n_steps; done += slice) {
        const int n = (int)std::min<int64_t>(slice, n_steps - done);
        for (auto& s : p->subs) {
            if (s.max_total >= 0 && done >= s.max_total) continue;   // every robot of this batch has finished
            int rc = vxg_step(s.b, n);
            if (rc) return rc;
        }
    }
    return VXG_OK;
}

extern "C" int vxg_pop_run(vxg_population* p) {
    if (!p) return fail(VXG_ERR_ARG, "vxg_pop_run: null population");
    int64_t mx = 0;
    for (auto& s : p->subs) {
        if (s.max_total < 0 || s.max_total == LLONG_MAX) return fail(VXG_ERR_STATE, "vxg_pop_run: every robot needs a step budget (vxg_pop_set_plan)");
        mx = std::max(mx, s.max_total);
    }
    return vxg_pop_step(p, mx);
}

extern "C" int vxg_pop_sync(vxg_population* p, uint32_t* status) {
    if (!p) return fail(VXG_ERR_ARG, "vxg_pop_sync: null population");
    for (auto& s : p->subs) {
        std::vector<uint32_t> st(s.robots.size());
        int rc = vxg_sync(s.b, st.data());
        if (rc) return rc;
        if (status) for (size_t i = 0; i < s.robots.size(); i++) status[s.robots[i]] = st[i];
    }
    return VXG_OK;
}

extern "C" int vxg_pop_reduce_stats(vxg_population* p, vxg_robot_stats* out) {
    if (!p || !out) return fail(VXG_ERR_ARG, "vxg_pop_reduce_stats: null argument");
    for (auto& s : p->subs) {
        std::vector<vxg_robot_stats> st(s.robots.size());
        int rc = vxg_reduce_stats(s.b, st.data());
        if (rc) return rc;
        for (size_t i = 0; i < s.robots.size(); i++) out[s.robots[i]] = st[i];
    }
    return VXG_OK;
}

extern "C" int vxg_pop_download_state(vxg_population* p, int robot, vxg_robot_state* out) {
    if (!p || robot < 0 || robot >= p->n) return fail(VXG_ERR_ARG, "vxg_pop_download_state: robot out of range");
    return vxg_download_state(p->subs[p->sub_of[robot]].b, p->slot_of[robot], out);
}

extern "C" int vxg_pop_n_batches(const vxg_population* p) { return p ? (int)p->subs.size() : 0; }

extern "C" int vxg_pop_batch_info(vxg_population* p, int i, int32_t* n_robots, int32_t* stride, int32_t* threads, int32_t* family) {
    if (!p || i < 0 || i >= (int)p->subs.size()) return fail(VXG_ERR_ARG, "vxg_pop_batch_info: batch out of range");
    Sub& s = p->subs[i];
    int th = 0, vpt = 0, sm = 0;
    int rc = vxg_kernel_config(s.b, &th, &vpt, &sm, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (n_robots) *n_robots = (int32_t)s.robots.size();
    if (stride) *stride = s.stride;
    if (threads) *threads = th;
    if (family) *family = (s.b->compact ? 1 : 0) | (s.family & ~1);   // what the batch actually runs
    return VXG_OK;
}

extern "C" int vxg_pop_locate(const vxg_population* p, int robot, int32_t* batch, int32_t* slot) {
    if (!p || robot < 0 || robot >= p->n) return fail(VXG_ERR_ARG, "vxg_pop_locate: robot out of range");
    if (batch) *batch = p->sub_of[robot];
    if (slot) *slot = p->slot_of[robot];
    return VXG_OK;
}

extern "C" int64_t vxg_pop_launch_count(const vxg_population* p) {
    int64_t n = 0;
    if (p) for (auto& s : p->subs) n += vxg_launch_count(s.b);
    return n;
}

// Benchmark markers over all the batches' streams.  Device time between two marks = (latest of the second mark's events) -
// (earliest of the first mark's), measured with CUDA events on the streams the kernels run on.
extern "C" int vxg_pop_mark(vxg_population* p, int slot) {
    if (!p) return fail(VXG_ERR_ARG, "vxg_pop_mark: null population");
    for (auto& s : p->subs) { int rc = vxg_mark(s.b, slot); if (rc) return rc; }
    return VXG_OK;
}
extern "C" int vxg_pop_mark_elapsed_ms(vxg_population* p, int s0, int s1, float* ms) {
    if (!p || !ms || s0 < 0 || s1 < 0 || s0 >= VXG_N_MARKS || s1 >= VXG_N_MARKS || p->subs.empty()) return fail(VXG_ERR_ARG, "vxg_pop_mark_elapsed_ms: bad argument");
    CK(cudaSetDevice(p->device));
    for (auto& s : p->subs) if (!s.b->mark_set[s0] || !s.b->mark_set[s1]) return fail(VXG_ERR_STATE, "vxg_pop_mark_elapsed_ms: marker not recorded");
    cudaEvent_t ref = p->subs[0].b->marks[s0];
    float first = 0.f, last = -1e30f;
    for (auto& s : p->subs) {
        CK(cudaEventSynchronize(s.b->marks[s1]));
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, ref, s.b->marks[s0]));
        CK(cudaEventElapsedTime(&b, ref, s.b->marks[s1]));
        first = std::min(first, a); last = std::max(last, b);
    }
    *ms = last - first;
    return VXG_OK;
}

extern "C" int64_t vxg_pop_upload_bytes(const vxg_population* p) {
    int64_t n = 0;
    if (p) for (auto& s : p->subs) n += vxg_upload_bytes(s.b);
    return n;
}

extern "C" int64_t vxg_pop_download_bytes(const vxg_population* p) {
    int64_t n = 0;
    if (p) for (auto& s : p->subs) n += (int64_t)s.robots.size() * ((int64_t)3 * s.b->L.nv * sizeof(double) + (int64_t)sizeof(RobotCtl));
    return n;
}

// The end-to-end call with host buffers, pipelined over the population's batches: flatten + upload + step + position download
// are issued chunk by chunk (every batch has its own stream), so the host flattens chunk c+1 while the device simulates chunk c,
// and unpacks chunk c while the device still simulates the later ones.
extern "C" int vxg_pop_run_host(vxg_population* p, const vxg_robot_desc* descs, int n_steps, double* pos_out, int pos_stride_vox) {
    if (!p || !descs || n_steps < 0) return fail(VXG_ERR_ARG, "vxg_pop_run_host: bad argument");
    if (pos_out) for (int r = 0; r < p->n; r++) if (p->n_vox[r] > pos_stride_vox) return fail(VXG_ERR_ARG, "vxg_pop_run_host: pos_stride_vox %d < %d voxels of robot %d", pos_stride_vox, p->n_vox[r], r);
    const int nt = vxg_i_host_threads(0);
    for (auto& s : p->subs) {
        int rc = upload_sub(p, s, descs, nt);
        if (!rc) rc = vxg_step(s.b, n_steps);
        if (!rc && pos_out) rc = vxg_i_enqueue_positions(s.b);
        if (rc) return rc;
    }
    for (auto& s : p->subs) {
        CK(cudaSetDevice(p->device));
        CK(cudaStreamSynchronize(s.b->stream));
        if (!pos_out) continue;
        for (size_t i = 0; i < s.robots.size(); i++)
            if (s.b->h_ctl[i].status & VXG_ST_DIVERGED) return fail(VXG_ERR_STATE, "vxg_pop_run_host: robot %d diverged (strain > 100)", s.robots[i]);
        parallel_for((int)s.robots.size(), std::min(nt, 16), [&](int i) { vxg_i_unpack_positions(s.b, i, pos_out + (size_t)s.robots[i] * pos_stride_vox * 3); });
    }
    return VXG_OK;
}
