// voxelyze_main.cpp -- drop-in for the reference CLI `./voxelyze -f <file.vxa> [-p]`
// (evosoro/_voxcad/voxelyzeMain/main.cpp:9-133) on top of libvxgpu.so, plus the batched form the GPU wants:
//
//     voxelyze -f a.vxa [-f b.vxa ...] [--batch <dir>] [-p] [--device N] [--lw] [--ignore-collisions]
//
// Every file is imported (LoadVXAFile + Import), the whole set is simulated as ONE batch on the GPU until
// each robot's stop condition, and each robot's <FitnessFileName> is written with the reference's tag set
// (CVX_SimGA::WriteResultFile).  Exit code: 1 = completed, 0 = could not complete -- the reference's own
// convention (main.cpp:28,57,132) that evosoro's subprocess call relies on.  There is no CPU path: a file the
// kernels cannot simulate (or a missing GPU) is an error (exit 0 with a message), never a silent fallback.
#include <dirent.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vxhost.h"

int main(int argc, char** argv) {
    std::vector<std::string> files;
    bool print_scrn = false, lw = false, ignore_col = false;
    int device = 0;
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
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--computeShapeDescriptors")) { fprintf(stderr, "voxelyze: --computeShapeDescriptors (qhull) is not supported\n"); return 0; }
    }
    if (files.empty()) { printf("\nInput file required. Quitting.\n"); return 0; }

    const int n = (int)files.size();
    std::vector<vxh_model*> models(n, nullptr);
    std::vector<vxg_robot_desc> descs(n);
    std::vector<int64_t> total(n), latch(n), eol(n);
    vxg_limits lim; memset(&lim, 0, sizeof lim);
    lim.max_vox = lim.max_vclass = lim.max_bclass = 1; lim.precision = VXG_PREC_F32;
    int64_t max_total = 0;
    for (int r = 0; r < n; r++) {
        int rc = vxh_load_vxa_file(files[r].c_str(), lw ? VXH_VARIANT_LAND_WATER : VXH_VARIANT_VOXCAD, ignore_col ? 1 : 0, &models[r]);
        if (rc) {
            if (print_scrn || rc > 0) fprintf(stderr, "\nProblem importing VXA file %s: %s. Quitting\n", files[r].c_str(), vxh_last_error());
            return 0;
        }
        descs[r] = *vxh_desc(models[r]);
        rc = vxh_step_plan_ex(models[r], &total[r], &latch[r], &eol[r], nullptr);
        if (rc) { fprintf(stderr, "voxelyze: %s: %s\n", files[r].c_str(), vxh_last_error()); return 0; }
        lim.max_vox = std::max(lim.max_vox, descs[r].n_vox);
        lim.max_vclass = std::max(lim.max_vclass, descs[r].n_vclass);
        lim.max_bclass = std::max(lim.max_bclass, descs[r].n_bclass);
        if (descs[r].sim.flags & VXG_F_COLLISIONS) lim.max_col_pairs = 512;   // contact-bond list capacity (overflow is reported, never truncated)
        // the land/water result file reports the volume of the deformed surface mesh at the end: every robot of that fork needs the
        // mesh tables, not only those in a fluid environment
        if (descs[r].sim.flags & (VXG_F_FLUID | VXG_F_LW_SCALE)) lim.max_mesh_verts = std::max(lim.max_mesh_verts, vxg_mesh_vertex_count(&descs[r]));
        max_total = std::max(max_total, total[r]);
        if (print_scrn) printf("Imported %s: %d voxels, %d bonds, %lld steps\n", files[r].c_str(), descs[r].n_vox, descs[r].n_bond, (long long)total[r]);
    }
    vxg_batch* b = nullptr;
    int rc = vxg_create(&b, device, n, &lim);
    for (int r = 0; r < n && !rc; r++) rc = vxg_upload(b, r, &descs[r]);
    if (!rc) rc = vxg_set_step_budget(b, total.data(), latch.data());
    if (!rc) rc = vxg_set_eol_step(b, eol.data());
    std::vector<uint32_t> status(n, 0);
    for (int64_t done = 0; done < max_total && !rc;) {   // whole lifetimes on the device; robots stop at their own budgets
        int chunk = (int)std::min<int64_t>(max_total - done, 1 << 20);
        rc = vxg_step(b, chunk);
        done += chunk;
    }
    if (!rc) rc = vxg_sync(b, status.data());
    std::vector<vxg_robot_stats> stats(n);
    if (!rc) rc = vxg_reduce_stats(b, stats.data());
    if (rc) { fprintf(stderr, "voxelyze: %s\n", vxg_last_error()); return 0; }
    int ok = 1;
    for (int r = 0; r < n; r++) {
        vxh_info info; vxh_get_info(models[r], &info);
        if (status[r] & VXG_ST_COL_OVERFLOW) { fprintf(stderr, "%s: more than %d self-collision contact bonds\n", files[r].c_str(), lim.max_col_pairs); ok = 0; }
        else if (status[r] & VXG_ST_DIVERGED) { fprintf(stderr, "%s: Simulation Diverged. Please reduce forces or accelerations.\n", files[r].c_str()); }
        if (print_scrn) printf("Ended at: %g\n", stats[r].time);
        if (info.fitness_file_name[0]) {   // SaveResultFile(FitnessFileName), main.cpp:130
            if (vxh_format_result(models[r], &stats[r], nullptr, 0, info.fitness_file_name) < 0) { fprintf(stderr, "voxelyze: %s\n", vxh_last_error()); ok = 0; }
        }
        vxh_free(models[r]);
    }
    vxg_destroy(b);
    return ok ? 1 : 0;
}
