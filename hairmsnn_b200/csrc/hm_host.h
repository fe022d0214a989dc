// hm_host.h — host-side scene containers shared by the loaders, the BVH builder and
// the C-ABI implementation.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "hm_bvh.h"

namespace hm {

// Fibre + head geometry in the layout the kernels consume.
struct HostGeometry {
    std::vector<F4> cps;          // Catmull-Rom control points incl. phantom endpoints; w = radius
    std::vector<int> seg_cp;      // per segment: index of the first of its 4 control points
    std::vector<int> seg_strand;  // per segment: strand id
    std::vector<F4> tri_verts;    // flattened triangle soup, 3 per triangle
    std::vector<F4> tri_normals;  // per-corner normals, 3 per triangle
    std::vector<float> tri_uv;    // per-corner texcoords, 6 per triangle
    int num_strands = 0;
    // bounds exactly as the reference accumulates them (hair bounds start at the
    // origin, headers/model.h:93-94; mesh bounds use the first corner of each
    // triangle only, model.cpp:309-310)
    float hair_min[3] = {0, 0, 0}, hair_max[3] = {0, 0, 0};
    float mesh_min[3] = {1e30f, 1e30f, 1e30f}, mesh_max[3] = {-1e30f, -1e30f, -1e30f};
    float hair_scale = 0.f;
    float scene_scale = 0.f;
    float kd[3] = {0, 0, 0};      // head diffuse colour (tinyobj default 0 when the .mtl has no Kd)
    float surf_alpha = 1.f;
};

struct HostBvh {
    std::vector<F4> nodes;
    std::vector<F4> leaf_data;   // 4 per leaf slot
    std::vector<int> leaf_code;
    std::vector<int> leaf_prim;
    // 8-wide quantised tree derived from the binary one (hm_bvh.h: WideNode)
    std::vector<F4> wnodes;      // 5 per wide node
    std::vector<F4> wleaf_data;  // 4 per wide leaf reference
    int wide_depth = 0;
};

void build_bvh(const HostGeometry& geo, HostBvh& out, int threads_hint = 0);
// Same through an on-disk cache of the GPU-side (8-wide) tree: directory from the environment variable
// HM_BVH_CACHE, file hm_bvh_<hash of geometry + build parameters>.bin, written atomically.  A scene
// restored from the cache carries no binary tree (only the builder's process has one).
void build_bvh_cached(const HostGeometry& geo, HostBvh& out, int threads_hint = 0);
uint64_t bvh_cache_key(const HostGeometry& geo);
bool load_bvh_cache(const std::string& path, uint64_t key, HostBvh& out);
void save_bvh_cache(const std::string& path, uint64_t key, const HostBvh& b);

// Everything parseScene produces (scene.cpp:119-339, headers/scene.h) plus the tables
// the frame drivers derive at start-up.
struct HostScene {
    HostGeometry geo;
    HostBvh bvh;
    // camera block
    float cam_from[3] = {0, 0, 1}, cam_to[3] = {0, 0, 0}, cam_up[3] = {0, 1, 0};
    float cos_fovy = 0.66f;
    // hair block (alpha already in radians: scene.cpp:207 multiplies by 3.14159f/180)
    float sigma_a[3] = {0.06f, 0.1f, 0.2f};
    float beta_m = 0.3f, beta_n = 0.3f, alpha = 0.f;
    float gains[4] = {1, 1, 1, 1};
    // lights block
    bool has_env = false;
    std::vector<float> env;      // RGBA32F
    int env_w = 0, env_h = 0;
    float env_scale = 1.f, env_rot = 0.f;
    std::vector<float> cpdf, ccdf, mpdf, mcdf;
    std::vector<float> dl_from;  // normalised, 3 per light
    std::vector<float> dl_emit;
    // integrator block
    int width = 0, height = 0, spp = 1, path_v1 = 1, path_v2 = 40;
    bool mis = true, env_pdf = true;
    std::string image_output, stats_output;
    // tcnn block
    std::string tcnn_config, tcnn_weights;
    bool tcnn_train = true;
    std::string base_dir;        // directory of config.json, for relative / foreign paths
};

// scene.cpp:349-425
void build_env_tables(HostScene& s);
// The host copy of the tables is built on first use (oracle hooks, HM_ENV_TABLES=host): the renderers build theirs on the
// device from the uploaded map (hm_wavefront.cu: launch_env_tables).  Thread-safe.
void ensure_env_tables(const HostScene& s);
// sin(theta) of every row's centre exactly as the host recipe computes it (scene.cpp:358): the device build takes
// these from the host so that libdevice's sinf cannot move a table entry by an ulp
void env_row_sines(int H, std::vector<float>& out);
// fetchSceneSamples (render_hair_msnn.cu:34-97): points on the strands for the TRAIN_DATA_GEN pass
void build_scene_samples(const HostGeometry& g, int num_samples, unsigned seed, std::vector<float>& points3);
// bounds / scales as the frame drivers compute them (render_hair_msnn.cu:414-430)
void finalize_geometry(HostScene& s);
// cameraChanged() (render_path_tracing.cu:767-797) through the viewer's camera
// round trip (owlViewer/Camera.cpp:94-120, Camera.h:55)
void camera_basis(const HostScene& s, int W, int H, float pos[3], float d00[3], float du[3], float dv[3]);

// Multi-GPU partition of one frame: `world` contiguous row bands; a band owns the
// training records whose pixel group (fbOfs / everyNth, cuda/hair_msnn.cu:199-203) lies wholly
// inside it (a group straddling a band edge is left out of that step's batch: its training
// pixel may fall on either side; with W % everyNth == 0, as in every shipped config, none do).  train_n = the band's record count rounded down to tcnn's batch granularity of 128
// (common.h:280) — what its backward pass consumes.  Pure host arithmetic: the renderer and the
// CPU-side multi-rank tests share it.
struct BandPartition {
    int row0, row1;          // rows [row0, row1)
    int slot0, slots;        // training records [slot0, slot0 + slots)
    int train_n;             // records fed to forward/backward
};
inline BandPartition band_partition(int W, int H, int records, int rank, int world) {
    BandPartition b;
    b.row0 = (int)((long long)H * rank / world);
    b.row1 = (int)((long long)H * (rank + 1) / world);
    const long long n = (long long)W * H;
    const int every_nth = records > 0 ? (int)(n / records) : 0;
    if (every_nth < 1) { b.slot0 = 0; b.slots = 0; b.train_n = 0; return b; }
    const long long px0 = (long long)b.row0 * W, px1 = (long long)b.row1 * W;
    b.slot0 = (int)((px0 + every_nth - 1) / every_nth);
    int slot1 = (int)(px1 / every_nth);
    if (slot1 > records) slot1 = records;
    if (b.slot0 > slot1) b.slot0 = slot1;
    b.slots = slot1 - b.slot0;
    b.train_n = b.slots - b.slots % 128;
    return b;
}

// Buffer sizes of render_nrc (RenderWindowNRC::initialize, render_nrc.cu:116-160; std::ceil of INTEGER
// divisions throughout): numTrainingRecords = 65536, MAX_BOUNCES = 40 (headers/render_nrc.h:126, nrc.cuh:13).
struct NrcLayout {
    int records;          // numTrainingRecords
    int train_pixels;     // numTrainingPixels = records / MAX_BOUNCES
    int every_nth;        // frameSize / numTrainingPixels (0: frame too small)
    int nn_frame_rows;    // nnFrameSize = frameSize + numTrainingPixels - numTrainingPixels % 128 + 128
};
inline NrcLayout nrc_layout(int W, int H) {
    NrcLayout l;
    const long long n = (long long)W * H;
    l.records = 65536;
    l.train_pixels = l.records / 40;
    l.every_nth = (int)(n / l.train_pixels);
    l.nn_frame_rows = (int)n + l.train_pixels - l.train_pixels % 128 + 128;
    return l;
}

inline GeomView make_view(const HostGeometry& g, const HostBvh& b) {
    GeomView v;
    v.nodes = b.nodes.data();
    v.leaf_data = b.leaf_data.data();
    v.leaf_code = b.leaf_code.data();
    v.leaf_prim = b.leaf_prim.data();
    v.cps = g.cps.data();
    v.tri_verts = g.tri_verts.data();
    v.num_segments = (int)g.seg_cp.size();
    v.num_tris = (int)(g.tri_verts.size() / 3);
    v.num_nodes = (int)(b.nodes.size() / 4);
    v.wnodes = b.wnodes.data();
    v.wleaf_data = b.wleaf_data.data();
    v.num_wnodes = (int)(b.wnodes.size() / 5);
    return v;
}

}  // namespace hm
