// hm_main_common.h — shared body of the three headless executables.
//
// Reference: main() of render_path_tracing.cu:817-841, render_nrc.cu:1080-1104,
// render_hair_msnn.cu:1146-1175 — `exe <config.json> [BETA]`, exit -1 when the scene does
// not load, BETA parsed with atoi (default 1).  The reference then opens a GLFW window and
// renders `spp` samples; this build renders the same number of samples headless and writes
// what the viewer's "Save PNG/EXR" buttons write (image_output, .exr next to it, and for
// HairMSNN the _pt / _nn components, render_hair_msnn.cu:1003-1022) plus stats_output.
#pragma once
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hairmsnn.h"

static std::string hm_with_suffix(const std::string& png, const char* suffix, const char* ext) {
    size_t dot = png.rfind('.');
    std::string stem = dot == std::string::npos ? png : png.substr(0, dot);
    return stem + suffix + ext;
}

// --gpus N: the launcher process starts N copies of itself, one per GPU (rank r on device r); rank 0 creates
// the NCCL id and hands it over through a file in /dev/shm, and builds the acceleration structure into a
// RAM-backed cache the other ranks restore.  No CUDA call is made in the launcher.
static int hm_launch_ranks(int argc, char** argv, int gpus) {
    char id_file[128], cache_dir[128];
    snprintf(id_file, sizeof(id_file), "/dev/shm/hm_comm_%d.id", (int)getpid());
    snprintf(cache_dir, sizeof(cache_dir), "/dev/shm/hm_bvh_%d", (int)getpid());
    const bool own_cache = getenv("HM_BVH_CACHE") == nullptr;
    if (own_cache) { mkdir(cache_dir, 0700); setenv("HM_BVH_CACHE", cache_dir, 1); }
    std::vector<pid_t> kids;
    for (int r = 0; r < gpus; ++r) {
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); return -1; }
        if (pid == 0) {
            std::vector<std::string> extra = {"--rank", std::to_string(r), "--world", std::to_string(gpus), "--comm-file", id_file};
            std::vector<char*> av(argv, argv + argc);
            for (auto& e : extra) av.push_back(const_cast<char*>(e.c_str()));
            av.push_back(nullptr);
            execv("/proc/self/exe", av.data());
            perror("execv");
            _exit(127);
        }
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = -1;
    }
    unlink(id_file);
    if (own_cache) { std::string cmd = std::string("rm -rf ") + cache_dir; if (system(cmd.c_str())) {} }
    return rc;
}

static bool hm_read_id_file(const std::string& path, unsigned char* id, int timeout_s) {
    for (int waited = 0; waited < timeout_s * 10; ++waited) {
        FILE* f = fopen(path.c_str(), "rb");
        if (f) {
            size_t n = fread(id, 1, HM_COMM_ID_BYTES, f);
            fclose(f);
            if (n == HM_COMM_ID_BYTES) return true;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    return false;
}

static int hm_main(int argc, char** argv, int kind, const char* name) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <config.json> [BETA] [--spp N] [--device D] [--gpus N [--shard spp|bands]] [--out image.png] [--stats stats.json] [--pretrain-steps K]\n", name);
        return -1;
    }
    std::string config = argv[1];
    int beta = 1, spp = -1, device = 0, pretrain = 200, gpus = 1, rank = -1, world = 1;
    std::string out_png, out_stats, comm_file, shard = "spp";
    int argi = 2;
    if (argi < argc && argv[argi][0] != '-') beta = atoi(argv[argi++]);
    for (; argi < argc; ++argi) {
        auto next = [&]() -> const char* { return argi + 1 < argc ? argv[++argi] : ""; };
        if (!strcmp(argv[argi], "--spp")) spp = atoi(next());
        else if (!strcmp(argv[argi], "--device")) device = atoi(next());
        else if (!strcmp(argv[argi], "--out")) out_png = next();
        else if (!strcmp(argv[argi], "--stats")) out_stats = next();
        else if (!strcmp(argv[argi], "--pretrain-steps")) pretrain = atoi(next());
        else if (!strcmp(argv[argi], "--gpus")) gpus = atoi(next());
        else if (!strcmp(argv[argi], "--shard")) shard = next();
        else if (!strcmp(argv[argi], "--rank")) rank = atoi(next());
        else if (!strcmp(argv[argi], "--world")) world = atoi(next());
        else if (!strcmp(argv[argi], "--comm-file")) comm_file = next();
        else { fprintf(stderr, "%s: unknown option %s\n", name, argv[argi]); return -1; }
    }
    if (shard != "spp" && shard != "bands") { fprintf(stderr, "%s: --shard is spp or bands\n", name); return -1; }
    if (kind == HM_RENDER_NRC && shard == "bands" && gpus > 1) { fprintf(stderr, "%s: render_nrc shards by samples only\n", name); return -1; }
    if (gpus > 1 && rank < 0) return hm_launch_ranks(argc, argv, gpus);
    const bool multi = rank >= 0 && world > 1;
    if (!multi) { rank = 0; world = 1; }
    else device = rank;
    hm_comm* comm = nullptr;
    unsigned char comm_id[HM_COMM_ID_BYTES];
    // ranks > 0 wait for rank 0's id file, which it writes AFTER building (and caching) the acceleration structure
    if (multi && rank > 0 && !hm_read_id_file(comm_file, comm_id, 600)) { fprintf(stderr, "%s: rank %d: no communicator id\n", name, rank); return -1; }
    if (rank == 0) printf("Loading scene %s\n", config.c_str());
    hm_scene* scene = nullptr;
    if (hm_scene_load(config.c_str(), &scene) != HM_OK) {
        fprintf(stderr, "Error loading scene: %s\n", hm_last_error());
        return -1;
    }
    hm_scene_info info;
    hm_scene_get_info(scene, &info);
    if (spp < 0) spp = info.spp;
    if (multi) {
        if (rank == 0) {
            if (hm_comm_get_unique_id(comm_id) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
            std::string tmp = comm_file + ".tmp";
            FILE* f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(comm_id, 1, HM_COMM_ID_BYTES, f) != HM_COMM_ID_BYTES) { fprintf(stderr, "%s: cannot write %s\n", name, tmp.c_str()); return -1; }
            fclose(f);
            rename(tmp.c_str(), comm_file.c_str());
        }
        if (hm_comm_create(comm_id, rank, world, device, &comm) != HM_OK) { fprintf(stderr, "%s: rank %d: %s\n", name, rank, hm_last_error()); return -1; }
    }
    const bool bands = multi && shard == "bands";
    // sample sharding: every rank renders ceil(spp / world) samples (ranks must issue the same collectives)
    const int my_spp = (multi && !bands) ? (spp + world - 1) / world : spp;
    if (rank == 0) printf("%d segments, %d triangles, %d BVH nodes, %dx%d, %d spp\n", info.num_segments, info.num_triangles, info.num_bvh_nodes,
           info.width, info.height, spp);
    hm_renderer* r = nullptr;
    if (hm_renderer_create(scene, kind, beta, device, bands ? rank : 0, bands ? world : 1, &r) != HM_OK) {
        fprintf(stderr, "%s: %s\n", name, hm_last_error());
        hm_scene_free(scene);
        return -1;
    }
    if (comm && hm_renderer_set_comm(r, comm) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
    hm_renderer_set_profiling(r, 1);
    auto t0 = std::chrono::steady_clock::now();
    if (kind == HM_RENDER_HAIR_MSNN && pretrain > 0) {
        if (hm_msnn_pretrain(r, pretrain) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rank == 0) printf("Initial training: %f sec\n", s);
        hm_renderer_reset_stats(r);
    }
    t0 = std::chrono::steady_clock::now();
    for (int done = 0; done < my_spp;) {
        int n = my_spp - done < 16 ? my_spp - done : 16;
        // enqueue only: the frames in flight overlap across the chunks; one synchronisation at the end
        if (hm_render_frames_async(r, n) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
        done += n;
    }
    if (hm_renderer_sync(r) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
    if (comm && hm_reduce_framebuffers(r) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int total_spp = (multi && !bands) ? my_spp * world : spp;
    if (rank == 0)
        printf("%d spp in %.3f s on %d GPU(s)%s: %.2f Mpaths/s\n", total_spp, secs, world, multi ? (bands ? " (row bands)" : " (sample groups)") : "",
               (double)info.width * info.height * total_spp / secs / 1e6);
    if (rank != 0) {   // outputs are rank 0's job
        hm_renderer_destroy(r);
        hm_comm_destroy(comm);
        hm_scene_free(scene);
        return 0;
    }

    // outputs: the scene's own paths when they are writable here, else next to the config
    std::string png = out_png;
    if (png.empty()) {
        size_t slash = config.rfind('/');
        png = (slash == std::string::npos ? std::string(".") : config.substr(0, slash)) + "/render.png";
    }
    int rc = 0;
    if (hm_save_png(r, png.c_str()) != HM_OK) { fprintf(stderr, "%s\n", hm_last_error()); rc = 1; }
    if (hm_save_exr(r, HM_BUF_FINAL_AVG, hm_with_suffix(png, "", ".exr").c_str()) != HM_OK) { fprintf(stderr, "%s\n", hm_last_error()); rc = 1; }
    if (kind == HM_RENDER_HAIR_MSNN) {
        hm_save_exr(r, HM_BUF_PT_AVG, hm_with_suffix(png, "_pt", ".exr").c_str());
        hm_save_exr(r, HM_BUF_NN_AVG, hm_with_suffix(png, "_nn", ".exr").c_str());
    }
    std::string stats = out_stats.empty() ? hm_with_suffix(png, "_stats", ".json") : out_stats;
    if (hm_write_stats(r, stats.c_str()) != HM_OK) { fprintf(stderr, "%s\n", hm_last_error()); rc = 1; }
    printf("wrote %s (+ .exr, stats %s)\n", png.c_str(), stats.c_str());
    hm_renderer_destroy(r);
    if (comm) hm_comm_destroy(comm);
    hm_scene_free(scene);
    return rc;
}
