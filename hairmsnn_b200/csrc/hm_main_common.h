// hm_main_common.h — shared body of the three headless executables.
//
// Reference: main() of render_path_tracing.cu:817-841, render_nrc.cu:1080-1104,
// render_hair_msnn.cu:1146-1175 — `exe <config.json> [BETA]`, exit -1 when the scene does
// not load, BETA parsed with atoi (default 1).  The reference then opens a GLFW window and
// renders `spp` samples; this build renders the same number of samples headless and writes
// what the viewer's "Save PNG/EXR" buttons write (image_output, .exr next to it, and for
// HairMSNN the _pt / _nn components, render_hair_msnn.cu:1003-1022) plus stats_output.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/hairmsnn.h"

static std::string hm_with_suffix(const std::string& png, const char* suffix, const char* ext) {
    size_t dot = png.rfind('.');
    std::string stem = dot == std::string::npos ? png : png.substr(0, dot);
    return stem + suffix + ext;
}

static int hm_main(int argc, char** argv, int kind, const char* name) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <config.json> [BETA] [--spp N] [--device D] [--out image.png] [--stats stats.json] [--pretrain-steps K]\n", name);
        return -1;
    }
    std::string config = argv[1];
    int beta = 1, spp = -1, device = 0, pretrain = 200;
    std::string out_png, out_stats;
    int argi = 2;
    if (argi < argc && argv[argi][0] != '-') beta = atoi(argv[argi++]);
    for (; argi < argc; ++argi) {
        auto next = [&]() -> const char* { return argi + 1 < argc ? argv[++argi] : ""; };
        if (!strcmp(argv[argi], "--spp")) spp = atoi(next());
        else if (!strcmp(argv[argi], "--device")) device = atoi(next());
        else if (!strcmp(argv[argi], "--out")) out_png = next();
        else if (!strcmp(argv[argi], "--stats")) out_stats = next();
        else if (!strcmp(argv[argi], "--pretrain-steps")) pretrain = atoi(next());
        else { fprintf(stderr, "%s: unknown option %s\n", name, argv[argi]); return -1; }
    }
    printf("Loading scene %s\n", config.c_str());
    hm_scene* scene = nullptr;
    if (hm_scene_load(config.c_str(), &scene) != HM_OK) {
        fprintf(stderr, "Error loading scene: %s\n", hm_last_error());
        return -1;
    }
    hm_scene_info info;
    hm_scene_get_info(scene, &info);
    if (spp < 0) spp = info.spp;
    printf("%d segments, %d triangles, %d BVH nodes, %dx%d, %d spp\n", info.num_segments, info.num_triangles, info.num_bvh_nodes,
           info.width, info.height, spp);
    hm_renderer* r = nullptr;
    if (hm_renderer_create(scene, kind, beta, device, 0, 1, &r) != HM_OK) {
        fprintf(stderr, "%s: %s\n", name, hm_last_error());
        hm_scene_free(scene);
        return -1;
    }
    hm_renderer_set_profiling(r, 1);
    auto t0 = std::chrono::steady_clock::now();
    if (kind == HM_RENDER_HAIR_MSNN && pretrain > 0) {
        if (hm_msnn_pretrain(r, pretrain) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("Initial training: %f sec\n", s);
        hm_renderer_reset_stats(r);
    }
    t0 = std::chrono::steady_clock::now();
    for (int done = 0; done < spp;) {
        int n = spp - done < 16 ? spp - done : 16;
        if (hm_render_frames(r, n) != HM_OK) { fprintf(stderr, "%s: %s\n", name, hm_last_error()); return -1; }
        done += n;
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%d spp in %.3f s: %.2f Mpaths/s\n", spp, secs, (double)info.width * info.height * spp / secs / 1e6);

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
    hm_scene_free(scene);
    return rc;
}
