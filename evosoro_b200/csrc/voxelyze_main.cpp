// voxelyze_main.cpp -- drop-in for the reference CLI `./voxelyze -f <file.vxa> [-p]`
// (evosoro/_voxcad/voxelyzeMain/main.cpp:9-133) on top of libvxgpu.so, plus the batched form the GPU wants:
//
//     voxelyze -f a.vxa [-f b.vxa ...] [--batch <dir>] [-p] [--device N | --devices 0,1,..] [--lw] [--ignore-collisions] [--plan]
//     voxelyze --serve [--devices ..] [--lw] [--ignore-collisions]
//
// --serve keeps the process (and with it the CUDA contexts: 1 - 1.5 s of a 2.6 s generation of 1024 robots) alive between
// generations: it answers "ready", then reads requests from stdin, one per line -- "run <list file>" (one .vxa path per line) --
// runs each as a batch exactly like the one-shot form, and answers "done <files> <failed>"; "quit" or end of input ends it.
//
// Every file is imported (LoadVXAFile + Import), the whole set is simulated as ONE batch on the GPU until
// each robot's stop condition, and each robot's <FitnessFileName> is written with the reference's tag set
// (CVX_SimGA::WriteResultFile).  Exit code: 1 = every robot completed and wrote its result, 0 = at least one did not -- the
// reference's own convention (main.cpp:28,57,132).  Failures are per robot: an unreadable / unsupported file or a diverged
// simulation costs that robot its result file (which is how evosoro recognises it, evaluation.py:107-119, 213-215) and nothing
// else.  There is no CPU path: a file the kernels cannot simulate (or a missing GPU) is an error with a message, never a silent
// fallback.
#include <dirent.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vxhost.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Options { bool print_scrn = false, lw = false, ignore_col = false, plan_only = false; std::vector<int> devices; };
static int run_files(const std::vector<std::string>& files, const Options& o, int* failed_out, std::vector<vxg_population*>* defer);

int main(int argc, char** argv) {
    std::vector<std::string> files;
    bool print_scrn = false, lw = false, ignore_col = false, plan_only = false, serve = false;
    std::vector<int> devices;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-f") && i + 1 < argc) files.push_back(argv[++i]);
        else if (!strcmp(argv[i], "--batch") && i + 1 < argc) {
            std::string dir = argv[++i];
            DIR* d = opendir(dir.c_str());
            if (!d) { fprintf(stderr, "voxelyze: cannot open directory %s\n", dir.c_str()); return 0; }
            std::vector<std::string> found;
            while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n.size() > 4 && n.substr(n.size() - 4) == ".vxa") found.push_back(dir + "/" + n); }
            closedir(d);
            std::sort(found.begin(), found.end());
            files.insert(files.end(), found.begin(), found.end());
        }
        else if (!strcmp(argv[i], "-p")) print_scrn = true;
        else if (!strcmp(argv[i], "--lw")) lw = true;
        else if (!strcmp(argv[i], "--ignore-collisions")) ignore_col = true;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) devices.push_back(atoi(argv[++i]));
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {   // --devices 0,1,2,3: the population is sharded over these GPUs
            for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t));
        }
        else if (!strcmp(argv[i], "--plan")) plan_only = true;     // import only: print "<file> <voxels> <bonds> <timesteps>" per input
        else if (!strcmp(argv[i], "--serve")) serve = true;
        else if (!strcmp(argv[i], "--computeShapeDescriptors")) { fprintf(stderr, "voxelyze: --computeShapeDescriptors (qhull) is not supported\n"); return 0; }
    }
    if (devices.empty()) devices.push_back(0);
    Options o; o.print_scrn = print_scrn; o.lw = lw; o.ignore_col = ignore_col; o.plan_only = plan_only; o.devices = devices;
    if (serve) {
        vxg_set_pooling(1);   // a population is built and torn down per request: its device / pinned blocks are reused, not reallocated
        // create the CUDA context of every device now (the first engine object does it), so that the first request does not pay for it
        for (int d : devices) {
            vxg_limits lim; memset(&lim, 0, sizeof lim);
            lim.max_vox = 1; lim.max_vclass = 1; lim.max_bclass = 1; lim.precision = VXG_PREC_F32;
            vxg_batch* b = nullptr;
            if (vxg_create(&b, d, 1, &lim)) { fprintf(stderr, "voxelyze: %s\n", vxg_last_error()); return 0; }
            vxg_destroy(b);
        }
        printf("ready\n"); fflush(stdout);
        std::thread reaper;
        char line[4096];
        while (fgets(line, sizeof line, stdin)) {
            std::string l = line;
            while (!l.empty() && (l.back() == '\n' || l.back() == '\r' || l.back() == ' ')) l.pop_back();
            if (l == "quit") break;
            if (l.compare(0, 4, "run ") != 0) { if (!l.empty()) { printf("error unknown request\n"); fflush(stdout); } continue; }
            std::vector<std::string> req;
            if (FILE* f = fopen(l.c_str() + 4, "r")) {
                char path[4096];
                while (fgets(path, sizeof path, f)) { std::string q = path; while (!q.empty() && (q.back() == '\n' || q.back() == '\r')) q.pop_back(); if (!q.empty()) req.push_back(q); }
                fclose(f);
            }
            int failed = (int)req.size();
            std::vector<vxg_population*> spent;
            if (!req.empty()) run_files(req, o, &failed, &spent);
            printf("done %d %d\n", (int)req.size(), failed); fflush(stdout);
            // freeing the device and pinned buffers takes 0.05 - 0.8 s: after the answer and on a thread of its own, so that neither this
            // generation's caller nor the next request waits for it
            if (reaper.joinable()) reaper.join();
            reaper = std::thread([spent]() { for (auto* p : spent) vxg_pop_destroy(p); });
        }
        if (reaper.joinable()) reaper.join();
        return 1;
    }
    if (files.empty()) { printf("\nInput file required. Quitting.\n"); return 0; }
    return run_files(files, o, nullptr, nullptr);
}

// One batch: import, simulate, write the result files.  Returns the exit code of the one-shot form (1 = every robot wrote its result).
static int run_files(const std::vector<std::string>& files, const Options& o, int* failed_out, std::vector<vxg_population*>* defer) {
    const bool timing = getenv("VXG_TIMING") != nullptr;   // phase times on stderr
    const double t_start = now_s();
    const bool print_scrn = o.print_scrn, lw = o.lw, ignore_col = o.ignore_col, plan_only = o.plan_only;
    const std::vector<int>& devices = o.devices;
    if (failed_out) *failed_out = (int)files.size();

    // Import every file on the host cores (LoadVXAFile + Import + the step plan are independent per file).  A file that cannot
    // be read or that enables something the kernels do not implement is dropped WITH a message and the others still run: the
    // reference runs one process per robot (evaluation.py:89-90), so one bad individual never costs its batch-mates their result.
    const int n_files = (int)files.size();
    std::vector<vxh_model*> loaded(n_files, nullptr);
    std::vector<int64_t> f_total(n_files), f_latch(n_files), f_eol(n_files);
    std::vector<std::string> why(n_files);
    {
        int nt = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("VXG_HOST_THREADS")) nt = atoi(e);
        nt = std::max(1, std::min(nt, std::min(n_files, 64)));
        std::atomic<int> next(0);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&]() {
                for (int r; (r = next.fetch_add(1)) < n_files;) {
                    int rc = vxh_load_vxa_file(files[r].c_str(), lw ? VXH_VARIANT_LAND_WATER : VXH_VARIANT_VOXCAD, ignore_col ? 1 : 0, &loaded[r]);
                    if (!rc) rc = vxh_step_plan_ex(loaded[r], &f_total[r], &f_latch[r], &f_eol[r], nullptr);
                    if (rc) { why[r] = vxh_last_error(); if (loaded[r]) { vxh_free(loaded[r]); loaded[r] = nullptr; } }
                }
            });
        for (auto& x : th) x.join();
    }
    const double t_imported = now_s();
    int n_failed = 0;
    std::vector<int> src;   // batch slot -> index into files
    for (int r = 0; r < n_files; r++) {
        if (loaded[r]) { src.push_back(r); continue; }
        n_failed++;
        fprintf(stderr, "\nProblem importing VXA file %s: %s. Skipped\n", files[r].c_str(), why[r].c_str());
    }
    const int n = (int)src.size();
    if (n == 0) { fprintf(stderr, "voxelyze: no input could be imported. Quitting\n"); return 0; }
    std::vector<vxh_model*> models(n, nullptr);
    std::vector<vxg_robot_desc> descs(n);
    std::vector<int64_t> total(n), latch(n), eol(n);
    for (int r = 0; r < n; r++) {
        models[r] = loaded[src[r]]; total[r] = f_total[src[r]]; latch[r] = f_latch[src[r]]; eol[r] = f_eol[src[r]];
        descs[r] = *vxh_desc(models[r]);
        if (print_scrn) printf("Imported %s: %d voxels, %d bonds, %lld steps\n", files[src[r]].c_str(), descs[r].n_vox, descs[r].n_bond, (long long)total[r]);
    }
    if (plan_only) {
        for (int r = 0; r < n; r++) printf("%s %d %d %lld\n", files[src[r]].c_str(), descs[r].n_vox, descs[r].n_bond, (long long)total[r]);
        for (int r = 0; r < n; r++) vxh_free(models[r]);
        if (failed_out) *failed_out = n_failed;
        return n_failed ? 0 : 1;
    }
    // The whole set is one ragged population per GPU: the library buckets it by voxel stride and kernel family (a 40-voxel CPPN
    // robot does not pay for a 216-voxel batch-mate's stride, nor a plain robot for a self-colliding one's kernel) and runs the
    // buckets concurrently; every robot stops at its own step budget.  Over several GPUs the robots are dealt out by cost --
    // timesteps x voxels, largest first onto the least loaded device (a robot with a bone-bone bond takes ten
    // times the steps of an all-tissue one) -- not by count.  No exchange between the devices: robots are independent.
    const int n_dev = (int)devices.size();
    std::vector<std::vector<int>> shard(n_dev);
    {
        std::vector<int> order(n);
        for (int r = 0; r < n; r++) order[r] = r;
        auto cost = [&](int r) { return (double)total[r] * (double)descs[r].n_vox; };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) != cost(b) ? cost(a) > cost(b) : a < b; });
        std::vector<double> load(n_dev, 0.0);
        for (int r : order) { int d = (int)(std::min_element(load.begin(), load.end()) - load.begin()); shard[d].push_back(r); load[d] += cost(r); }
        for (auto& sh : shard) std::sort(sh.begin(), sh.end());
    }
    vxg_pop_options opt; memset(&opt, 0, sizeof opt);
    opt.precision = VXG_PREC_F32;
    opt.want_mesh = 1;   // the land/water result file reports the volume of the deformed surface mesh for every robot of that fork
    // batches of 128 robots on their own streams: a launch ends when its last CTA does, so one 1024-robot batch pays for two whole
    // waves of CTAs in every launch where 1.7 would do, and for its stragglers; eight batches fill each other's gaps
    opt.chunk_robots = 128;
    if (const char* e = getenv("VXG_CHUNK_ROBOTS")) opt.chunk_robots = atoi(e);
    if (const char* e = getenv("VXG_COL_PAIRS_PER_VOX")) opt.col_pairs_per_vox = atoi(e);   // contact-bond capacity of the first pass (default 4 x voxels)
    std::vector<vxg_population*> pops(n_dev, nullptr);
    std::vector<uint32_t> status(n, 0);
    std::vector<vxg_robot_stats> stats(n);
    int rc = 0;
    for (int d = 0; d < n_dev && !rc; d++) {   // create + start every device's share (asynchronous), then collect
        const std::vector<int>& sh = shard[d];
        if (sh.empty()) continue;
        std::vector<vxg_robot_desc> dd; std::vector<int64_t> t, l, e;
        for (int r : sh) { dd.push_back(descs[r]); t.push_back(total[r]); l.push_back(latch[r]); e.push_back(eol[r]); }
        rc = vxg_pop_create(&pops[d], devices[d], dd.data(), (int)sh.size(), &opt);
        if (!rc) rc = vxg_pop_set_plan(pops[d], t.data(), l.data(), e.data());
        if (!rc) rc = vxg_pop_run(pops[d]);   // whole lifetimes on the device
    }
    const double t_launched = now_s();
    long long n_batches = 0, n_launches = 0;
    for (int d = 0; d < n_dev && !rc; d++) {
        if (!pops[d]) continue;
        const std::vector<int>& sh = shard[d];
        std::vector<uint32_t> st(sh.size()); std::vector<vxg_robot_stats> ss(sh.size());
        rc = vxg_pop_sync(pops[d], st.data());
        if (!rc) rc = vxg_pop_reduce_stats(pops[d], ss.data());
        for (size_t i = 0; i < sh.size() && !rc; i++) { status[sh[i]] = st[i]; stats[sh[i]] = ss[i]; }
        n_batches += vxg_pop_n_batches(pops[d]); n_launches += vxg_pop_launch_count(pops[d]);
    }
    // The reference's contact-bond list has no capacity (VX_Sim.cpp:2357-2413 grows a vector); ours is sized 4 x voxels, shared-memory
    // resident.  A robot that outgrows it stopped with VXG_ST_COL_OVERFLOW -- it is run again, from the start, in a population of
    // its own with sixteen times the room (the engine's limit is 16 384 bonds; beyond that the robot fails like a diverged one).
    if (!rc) {
        std::vector<int> again;
        for (int r = 0; r < n; r++) if (status[r] & VXG_ST_COL_OVERFLOW) again.push_back(r);
        if (!again.empty()) {
            fprintf(stderr, "voxelyze: %d robot(s) outgrew the contact-bond list; running them again with a larger one\n", (int)again.size());
            std::vector<vxg_robot_desc> dd; std::vector<int64_t> t, l, e;
            for (int r : again) { dd.push_back(descs[r]); t.push_back(total[r]); l.push_back(latch[r]); e.push_back(eol[r]); }
            vxg_pop_options big = opt;
            big.col_pairs_per_vox = 16 * (opt.col_pairs_per_vox > 0 ? opt.col_pairs_per_vox : 4);
            vxg_population* p2 = nullptr;
            std::vector<uint32_t> st(again.size()); std::vector<vxg_robot_stats> ss(again.size());
            int rc2 = vxg_pop_create(&p2, devices[0], dd.data(), (int)again.size(), &big);
            if (!rc2) rc2 = vxg_pop_set_plan(p2, t.data(), l.data(), e.data());
            if (!rc2) rc2 = vxg_pop_run(p2);
            if (!rc2) rc2 = vxg_pop_sync(p2, st.data());
            if (!rc2) rc2 = vxg_pop_reduce_stats(p2, ss.data());
            if (!rc2) for (size_t i = 0; i < again.size(); i++) { status[again[i]] = st[i]; stats[again[i]] = ss[i]; }
            else fprintf(stderr, "voxelyze: %s\n", vxg_last_error());
            if (p2) { n_batches += vxg_pop_n_batches(p2); n_launches += vxg_pop_launch_count(p2); vxg_pop_destroy(p2); }
        }
    }
    if (rc) {
        fprintf(stderr, "voxelyze: %s\n", vxg_last_error());
        for (auto* p : pops) if (p) vxg_pop_destroy(p);
        for (int r = 0; r < n; r++) vxh_free(models[r]);
        return 0;
    }
    const double t_simulated = now_s();
    if (print_scrn) printf("%d robots on %d device(s) in %lld batches, %lld kernel launches\n", n, n_dev, n_batches, n_launches);
    // result files, on the host cores like the import (1024 small files: 0.015 s, but up to 0.23 s on a file system that has just
    // had the previous generation's files deleted from the same directory)
    if (print_scrn) for (int r = 0; r < n; r++) printf("Ended at: %g\n", stats[r].time);
    {
        int nt = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("VXG_HOST_THREADS")) nt = atoi(e);
        nt = std::max(1, std::min(nt, std::min(n, 64)));
        std::atomic<int> next(0), bad(0);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&]() {
                for (int r; (r = next.fetch_add(1)) < n;) {
                    vxh_info info; vxh_get_info(models[r], &info);
                    const char* file = files[src[r]].c_str();
                    // A robot whose simulation diverged never reaches its stop condition in the reference (TimeStep keeps returning false,
                    // VX_Sim.cpp:1147-1150, so CurTime stands still and main.cpp:89 never leaves its loop): no result file is ever written and
                    // evosoro leaves the worst fitness on that individual (evaluation.py:107-110, 213-215).  Same here, without the hang:
                    // no file (a stale one is removed), a message, and the others are unaffected.
                    const bool failed = (status[r] & (VXG_ST_COL_OVERFLOW | VXG_ST_DIVERGED)) != 0;
                    if (status[r] & VXG_ST_COL_OVERFLOW) fprintf(stderr, "%s: the self-collision contact-bond list outgrew its capacity. No result written\n", file);
                    else if (status[r] & VXG_ST_DIVERGED) fprintf(stderr, "%s: Simulation Diverged. Please reduce forces or accelerations. No result written\n", file);
                    if (failed) { bad++; if (info.fitness_file_name[0]) remove(info.fitness_file_name); }
                    else if (info.fitness_file_name[0]) {   // SaveResultFile(FitnessFileName), main.cpp:130
                        if (vxh_format_result(models[r], &stats[r], nullptr, 0, info.fitness_file_name) < 0) { fprintf(stderr, "voxelyze: %s\n", vxh_last_error()); bad++; }
                    }
                    vxh_free(models[r]);
                }
            });
        for (auto& x : th) x.join();
        n_failed += bad.load();
    }
    const double t_written = now_s();
    for (auto* p : pops) { if (!p) continue; if (defer) defer->push_back(p); else vxg_pop_destroy(p); }
    if (timing) fprintf(stderr, "voxelyze: %d files: import %.3f s, populations (CUDA context, bucketing, flatten, upload, launches) %.3f s, simulation %.3f s, "
                                "result files %.3f s, teardown %.3f s\n", n_files, t_imported - t_start, t_launched - t_imported, t_simulated - t_launched,
                        t_written - t_simulated, now_s() - t_written);
    if (n_failed) fprintf(stderr, "voxelyze: %d of %d robots did not produce a result\n", n_failed, n_files);
    if (failed_out) *failed_out = n_failed;
    return n_failed ? 0 : 1;
}
