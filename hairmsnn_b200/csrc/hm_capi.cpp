// hm_capi.cpp — extern "C" boundary (include/hairmsnn.h).  Translates exceptions into
// status codes; owns the opaque handles.
#include "../../include/hairmsnn.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>

#include "hm_io.h"
#include "hm_renderer.h"

using namespace hm;

// The host scene is shared: a renderer keeps it alive, so hm_scene_free before hm_renderer_destroy is safe.
struct hm_scene {
    std::shared_ptr<HostScene> keep{new HostScene};
    HostScene& hs = *keep;
};
struct hm_mlp {
    std::unique_ptr<Mlp> owned;
    Mlp* m = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
    Renderer* owner = nullptr;   // a renderer's network: calls on it come after the frames the renderer holds back
};
static void flush_owner(hm_mlp* m) { if (m && m->owner) m->owner->flush(); }
struct hm_comm {
    std::unique_ptr<hm::Comm> c;
};
struct hm_renderer {
    std::shared_ptr<HostScene> scene_keep;   // declared first: destroyed after the renderer that references it
    std::unique_ptr<Renderer> r;
    hm_mlp mlp_view;
};

static thread_local std::string g_err;

namespace {
template <typename F>
int guarded(F&& f) {
    try {
        f();
        return HM_OK;
    } catch (const hm::IoError& e) {
        g_err = e.what();
        return HM_ERR_IO;
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        return HM_ERR_ARG;
    } catch (const std::logic_error& e) {
        g_err = e.what();
        return HM_ERR_STATE;
    } catch (const std::exception& e) {
        g_err = e.what();
        return (strncmp(e.what(), "CUDA", 4) == 0 || strncmp(e.what(), "NCCL", 4) == 0) ? HM_ERR_CUDA : HM_ERR_ARG;
    } catch (...) {
        g_err = "unknown error";
        return HM_ERR_ARG;
    }
}
void need(const void* p, const char* what) {
    if (!p) throw std::invalid_argument(std::string("null argument: ") + what);
}
void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e) + " (" + what + ")");
}
}  // namespace

extern "C" {

const char* hm_last_error(void) { return g_err.c_str(); }

int hm_frame_param_bytes(void) { return (int)sizeof(hm::FrameParams); }

int hm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- scene --------------------------------------------------------------------------
int hm_scene_load(const char* path, hm_scene** out) {
    return guarded([&] {
        need(path, "config_json_path"); need(out, "out");
        std::unique_ptr<hm_scene> s(new hm_scene);
        load_scene_file(path, s->hs);
        finalize_geometry(s->hs);
        build_bvh_cached(s->hs.geo, s->hs.bvh);
        *out = s.release();
    });
}

int hm_scene_create(const hm_scene_desc* d, hm_scene** out) {
    return guarded([&] {
        need(d, "desc"); need(out, "out");
        if (d->num_segments < 0 || d->num_triangles < 0 || (d->num_segments == 0 && d->num_triangles == 0))
            throw std::invalid_argument("Either hair or surface must be defined");
        if (d->num_segments > 0) { need(d->control_points, "control_points"); need(d->segment_first_cp, "segment_first_cp"); }
        if (d->num_triangles > 0) { need(d->tri_vertices, "tri_vertices"); need(d->tri_normals, "tri_normals"); }
        if (!d->env_rgba && d->num_dlights <= 0)
            throw std::invalid_argument("Either directional or environment light must be defined");
        if (d->width <= 0 || d->height <= 0 || ((int64_t)d->width * d->height) % 128 != 0)
            throw std::invalid_argument("width*height must be a positive multiple of 128 (scene.cpp:302-306)");
        std::unique_ptr<hm_scene> s(new hm_scene);
        HostScene& hs = s->hs;
        HostGeometry& g = hs.geo;
        g.cps.resize(d->num_control_points);
        if (d->num_control_points) memcpy(g.cps.data(), d->control_points, sizeof(float) * 4 * (size_t)d->num_control_points);
        g.seg_cp.assign(d->segment_first_cp, d->segment_first_cp + d->num_segments);
        for (int i = 0; i < d->num_segments; ++i)
            if (g.seg_cp[i] < 0 || g.seg_cp[i] + 3 >= d->num_control_points)
                throw std::invalid_argument("segment_first_cp out of range");
        g.num_strands = d->num_strands;
        const size_t nv = 3 * (size_t)d->num_triangles;
        g.tri_verts.resize(nv); g.tri_normals.resize(nv);
        for (size_t i = 0; i < nv; ++i) {
            g.tri_verts[i] = F4{d->tri_vertices[3 * i], d->tri_vertices[3 * i + 1], d->tri_vertices[3 * i + 2], 0.f};
            g.tri_normals[i] = F4{d->tri_normals[3 * i], d->tri_normals[3 * i + 1], d->tri_normals[3 * i + 2], 0.f};
        }
        for (int k = 0; k < 3; ++k) {
            g.hair_min[k] = d->hair_min[k]; g.hair_max[k] = d->hair_max[k]; g.kd[k] = d->surface_kd[k];
            hs.cam_from[k] = d->cam_from[k]; hs.cam_to[k] = d->cam_to[k]; hs.cam_up[k] = d->cam_up[k];
            hs.sigma_a[k] = d->sigma_a[k];
        }
        g.surf_alpha = d->surface_alpha;
        hs.cos_fovy = d->cos_fovy;
        hs.beta_m = d->beta_m; hs.beta_n = d->beta_n; hs.alpha = d->alpha;
        for (int k = 0; k < 4; ++k) hs.gains[k] = d->gains[k];
        hs.has_env = d->env_rgba != nullptr;
        if (hs.has_env) {
            if (d->env_w <= 0 || d->env_h <= 0) throw std::invalid_argument("bad environment size");
            hs.env_w = d->env_w; hs.env_h = d->env_h;
            hs.env.assign(d->env_rgba, d->env_rgba + 4 * (size_t)d->env_w * d->env_h);
        }
        hs.env_scale = d->env_scale; hs.env_rot = d->env_rotation;
        for (int i = 0; i < d->num_dlights; ++i) {
            float x = d->dl_from[3 * i], y = d->dl_from[3 * i + 1], z = d->dl_from[3 * i + 2];
            float r = 1.f / sqrtf(x * x + y * y + z * z);
            hs.dl_from.push_back(x * r); hs.dl_from.push_back(y * r); hs.dl_from.push_back(z * r);
            for (int k = 0; k < 3; ++k) hs.dl_emit.push_back(d->dl_emit[3 * i + k]);
        }
        hs.width = d->width; hs.height = d->height; hs.spp = d->spp;
        hs.path_v1 = d->path_v1; hs.path_v2 = d->path_v2;
        hs.mis = d->mis != 0; hs.env_pdf = d->env_pdf != 0;
        if (d->tcnn_config_path) hs.tcnn_config = d->tcnn_config_path;
        finalize_geometry(hs);
        build_bvh_cached(hs.geo, hs.bvh);
        *out = s.release();
    });
}

void hm_scene_free(hm_scene* s) { delete s; }

int hm_hair_file_load(const char* path, int* counts3, float* cps4, int* segment_first_cp, float* bounds6) {
    return guarded([&] {
        need(path, "path"); need(counts3, "counts3");
        hm::HostGeometry g;
        hm::load_hair_file(path, g);
        counts3[0] = (int)g.cps.size(); counts3[1] = (int)g.seg_cp.size(); counts3[2] = g.num_strands;
        if (cps4) memcpy(cps4, g.cps.data(), g.cps.size() * sizeof(hm::F4));
        if (segment_first_cp) memcpy(segment_first_cp, g.seg_cp.data(), g.seg_cp.size() * sizeof(int));
        if (bounds6) for (int k = 0; k < 3; ++k) { bounds6[k] = g.hair_min[k]; bounds6[3 + k] = g.hair_max[k]; }
    });
}

int hm_scene_save_bvh_cache(const hm_scene* s, const char* dir) {
    return guarded([&] {
        need(s, "scene"); need(dir, "dir");
        const uint64_t key = bvh_cache_key(s->hs.geo);
        char name[64];
        snprintf(name, sizeof(name), "/hm_bvh_%016llx.bin", (unsigned long long)key);
        save_bvh_cache(std::string(dir) + name, key, s->hs.bvh);
    });
}
int hm_scene_get_info(const hm_scene* s, hm_scene_info* info) {
    return guarded([&] {
        need(s, "scene"); need(info, "info");
        const HostScene& hs = s->hs;
        memset(info, 0, sizeof(*info));
        info->width = hs.width; info->height = hs.height; info->spp = hs.spp;
        info->path_v1 = hs.path_v1; info->path_v2 = hs.path_v2;
        info->num_segments = (int)hs.geo.seg_cp.size();
        info->num_control_points = (int)hs.geo.cps.size();
        info->num_triangles = (int)(hs.geo.tri_verts.size() / 3);
        info->num_strands = hs.geo.num_strands;
        info->num_bvh_nodes = (int)(hs.bvh.nodes.size() / 4);
        info->scene_scale = hs.geo.scene_scale;
        camera_basis(hs, hs.width, hs.height, info->cam_pos, info->cam_d00, info->cam_du, info->cam_dv);
        info->env_w = hs.env_w; info->env_h = hs.env_h;
        info->num_dlights = (int)(hs.dl_from.size() / 3);
        info->num_wide_nodes = (int)(hs.bvh.wnodes.size() / 5);
        info->num_wide_leaf_refs = (int)(hs.bvh.wleaf_data.size() / 4);
        info->wide_depth = hs.bvh.wide_depth;
    });
}

int hm_scene_get_arrays(const hm_scene* s, const float** nodes, const int** leaf_code, const int** leaf_prim,
                        const float** cps, const float** tri_v, const float** tri_n, const int** seg_cp,
                        const float** leaf_data) {
    return guarded([&] {
        need(s, "scene");
        const HostScene& hs = s->hs;
        if (nodes) *nodes = (const float*)hs.bvh.nodes.data();
        if (leaf_code) *leaf_code = hs.bvh.leaf_code.data();
        if (leaf_prim) *leaf_prim = hs.bvh.leaf_prim.data();
        if (cps) *cps = (const float*)hs.geo.cps.data();
        if (tri_v) *tri_v = (const float*)hs.geo.tri_verts.data();
        if (tri_n) *tri_n = (const float*)hs.geo.tri_normals.data();
        if (seg_cp) *seg_cp = hs.geo.seg_cp.data();
        if (leaf_data) *leaf_data = (const float*)hs.bvh.leaf_data.data();
    });
}

int hm_scene_get_env_tables(const hm_scene* s, const float** env, const float** cpdf, const float** ccdf,
                            const float** mpdf, const float** mcdf) {
    return guarded([&] {
        need(s, "scene");
        const HostScene& hs = s->hs;
        if (!hs.has_env) throw std::logic_error("scene has no environment light");
        hm::ensure_env_tables(hs);      // the host copy is built on first use; the renderers build theirs on the device
        if (env) *env = hs.env.data();
        if (cpdf) *cpdf = hs.cpdf.data();
        if (ccdf) *ccdf = hs.ccdf.data();
        if (mpdf) *mpdf = hs.mpdf.data();
        if (mcdf) *mcdf = hs.mcdf.data();
    });
}

int hm_image_load_exr(const char* path, float* rgba, size_t capacity_floats, int* width, int* height) {
    return guarded([&] {
        need(path, "path"); need(width, "width"); need(height, "height");
        std::vector<float> img;
        int w = 0, h = 0;
        load_exr_rgba(path, img, w, h);
        *width = w; *height = h;
        if (rgba) {
            if (capacity_floats < img.size()) throw std::invalid_argument("hm_image_load_exr: buffer too small");
            memcpy(rgba, img.data(), img.size() * sizeof(float));
        }
    });
}

static void load_tcnn_snapshot(hm::Mlp& mlp, const std::string& path);

// ---- renderer -----------------------------------------------------------------------
int hm_renderer_create(hm_scene* s, int kind, int beta_cli, int device, int rank, int world, hm_renderer** out) {
    return guarded([&] {
        need(s, "scene"); need(out, "out");
        if (kind != HM_RENDER_PATH_TRACING && kind != HM_RENDER_HAIR_MSNN && kind != HM_RENDER_NRC)
            throw std::invalid_argument("unknown renderer kind");
        std::unique_ptr<hm_renderer> h(new hm_renderer);
        h->scene_keep = s->keep;
        h->r.reset(new Renderer(s->hs, kind, beta_cli, device, rank, world));
        h->mlp_view.m = h->r->mlp();
        h->mlp_view.stream = h->r->stream();
        h->mlp_view.owner = h->r.get();
        h->mlp_view.device = device;
        // tcnn.init_weights (scene.cpp:318-322; loaded right after the TINY_MLP ctor, render_nrc.cu:148-150)
        if (h->r->mlp() && !s->hs.tcnn_weights.empty())
            load_tcnn_snapshot(*h->r->mlp(), hm::resolve_scene_path(s->hs.tcnn_weights, s->hs.base_dir));
        *out = h.release();
    });
}
void hm_renderer_destroy(hm_renderer* r) { delete r; }

int hm_render_frames(hm_renderer* r, int n) {
    return guarded([&] { need(r, "renderer"); r->r->render_frames(n); r->r->sync(); });
}
int hm_render_frames_async(hm_renderer* r, int n) {
    return guarded([&] { need(r, "renderer"); r->r->render_frames(n); });
}
int hm_render_flush(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->flush(); });
}
int hm_renderer_sync(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->sync(); });
}
int hm_renderer_reset_accumulation(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->reset_accumulation(); });
}
int hm_renderer_set_hair_params(hm_renderer* r, const float* sigma_a3, float beta_m, float beta_n, float alpha_radians,
                                const float* gains4) {
    return guarded([&] {
        need(r, "renderer"); need(sigma_a3, "sigma_a3"); need(gains4, "gains4");
        r->r->set_hair_params(sigma_a3, beta_m, beta_n, alpha_radians, gains4);
    });
}
int hm_renderer_set_environment(hm_renderer* r, float scale, float rotation) {
    return guarded([&] { need(r, "renderer"); r->r->set_environment(scale, rotation); });
}
int hm_renderer_set_sampling(hm_renderer* r, int mis, int env_pdf) {
    return guarded([&] { need(r, "renderer"); r->r->set_sampling(mis != 0, env_pdf != 0); });
}
int hm_renderer_accum_id(const hm_renderer* r) { return r ? r->r->accum_id() : -1; }
void* hm_renderer_stream(hm_renderer* r) { return r ? (void*)r->r->stream() : nullptr; }

int hm_comm_get_unique_id(void* out_id128) {
    return guarded([&] { need(out_id128, "out_id128"); hm::Comm::unique_id(out_id128); });
}
int hm_comm_create(const void* id128, int rank, int world, int device, hm_comm** out) {
    return guarded([&] {
        need(id128, "id128"); need(out, "out");
        std::unique_ptr<hm_comm> h(new hm_comm);
        h->c.reset(new hm::Comm(id128, rank, world, device));
        *out = h.release();
    });
}
void hm_comm_destroy(hm_comm* c) { delete c; }
int hm_comm_barrier(hm_comm* c) { return guarded([&] { need(c, "comm"); c->c->barrier(); }); }
int hm_comm_all_reduce_max(hm_comm* c, double* v) {
    return guarded([&] { need(c, "comm"); need(v, "inout"); *v = c->c->all_reduce_max(*v); });
}
int hm_comm_all_reduce_sum(hm_comm* c, double* v) {
    return guarded([&] { need(c, "comm"); need(v, "inout"); *v = c->c->all_reduce_sum(*v); });
}
int hm_renderer_set_comm(hm_renderer* r, hm_comm* c) {
    return guarded([&] { need(r, "renderer"); r->r->set_comm(c ? c->c.get() : nullptr); });
}
int hm_reduce_framebuffers(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->reduce_framebuffers(); });
}

namespace {
struct DevBuf {
    void* p = nullptr;
    DevBuf(const void* host, size_t bytes) {
        cuda_ok(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc");
        if (host) cuda_ok(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice), "cudaMemcpy");
    }
    ~DevBuf() { cudaFree(p); }
    float* f() const { return (float*)p; }
};
hm::HairLobes bsdf_lobes(int device, const float* sigma_a3, float beta_m, float beta_n, float alpha, const float* gains4) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("CUDA: no usable device (this library has no CPU path)");
    if (device < 0 || device >= ndev) throw std::runtime_error("CUDA: device index out of range");
    cuda_ok(cudaSetDevice(device), "cudaSetDevice");
    hm::HairLobes L;
    L.setup(beta_m, beta_n, alpha);
    L.sigma_a = hm::V3(sigma_a3[0], sigma_a3[1], sigma_a3[2]);
    for (int i = 0; i < 4; ++i) L.gain[i] = gains4[i];
    return L;
}
}  // namespace

int hm_bsdf_eval(int device, const float* sigma_a3, float beta_m, float beta_n, float alpha, const float* gains4,
                 const float* wo, const float* wi, const float* h, int n, float* out_f, float* out_pdf) {
    return guarded([&] {
        need(sigma_a3, "sigma_a3"); need(gains4, "gains4"); need(wo, "wo_local3"); need(wi, "wi_local3"); need(h, "h");
        need(out_f, "out_f3"); need(out_pdf, "out_pdf");
        if (n <= 0) throw std::invalid_argument("n must be positive");
        const hm::HairLobes L = bsdf_lobes(device, sigma_a3, beta_m, beta_n, alpha, gains4);
        DevBuf dwo(wo, (size_t)n * 12), dwi(wi, (size_t)n * 12), dh(h, (size_t)n * 4), df(nullptr, (size_t)n * 12), dp(nullptr, (size_t)n * 4);
        hm::launch_bsdf_eval(L, dwo.f(), dwi.f(), dh.f(), n, df.f(), dp.f(), nullptr);
        cuda_ok(cudaDeviceSynchronize(), "hm_bsdf_eval");
        cuda_ok(cudaMemcpy(out_f, df.p, (size_t)n * 12, cudaMemcpyDeviceToHost), "hm_bsdf_eval");
        cuda_ok(cudaMemcpy(out_pdf, dp.p, (size_t)n * 4, cudaMemcpyDeviceToHost), "hm_bsdf_eval");
    });
}
int hm_bsdf_sample(int device, const float* sigma_a3, float beta_m, float beta_n, float alpha, const float* gains4,
                   const float* wo, const float* h, const float* rand4, int n, float* out_wi, float* out_f, float* out_pdf) {
    return guarded([&] {
        need(sigma_a3, "sigma_a3"); need(gains4, "gains4"); need(wo, "wo_local3"); need(h, "h"); need(rand4, "rand4");
        need(out_wi, "out_wi_local3"); need(out_f, "out_f3"); need(out_pdf, "out_pdf");
        if (n <= 0) throw std::invalid_argument("n must be positive");
        const hm::HairLobes L = bsdf_lobes(device, sigma_a3, beta_m, beta_n, alpha, gains4);
        DevBuf dwo(wo, (size_t)n * 12), dh(h, (size_t)n * 4), du(rand4, (size_t)n * 16);
        DevBuf dw(nullptr, (size_t)n * 12), df(nullptr, (size_t)n * 12), dp(nullptr, (size_t)n * 4);
        hm::launch_bsdf_sample(L, dwo.f(), dh.f(), du.f(), n, dw.f(), df.f(), dp.f(), nullptr);
        cuda_ok(cudaDeviceSynchronize(), "hm_bsdf_sample");
        cuda_ok(cudaMemcpy(out_wi, dw.p, (size_t)n * 12, cudaMemcpyDeviceToHost), "hm_bsdf_sample");
        cuda_ok(cudaMemcpy(out_f, df.p, (size_t)n * 12, cudaMemcpyDeviceToHost), "hm_bsdf_sample");
        cuda_ok(cudaMemcpy(out_pdf, dp.p, (size_t)n * 4, cudaMemcpyDeviceToHost), "hm_bsdf_sample");
    });
}

int hm_msnn_trace(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->msnn_trace(); }); }
int hm_msnn_train_backward(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->msnn_train_backward(); }); }
int hm_msnn_train_apply(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->msnn_train_apply(); }); }
int hm_msnn_finish(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->msnn_finish(); }); }
int hm_nrc_trace(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->nrc_trace(); }); }
int hm_nrc_query(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->nrc_query(); }); }
int hm_nrc_train_backward(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->nrc_train_backward(); }); }
int hm_nrc_train_apply(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->nrc_train_apply(); }); }
int hm_nrc_end(hm_renderer* r) { return guarded([&] { need(r, "renderer"); r->r->nrc_end(); }); }
int hm_nrc_set_all_unbiased(hm_renderer* r, int on) {
    return guarded([&] { need(r, "renderer"); r->r->set_nrc_all_unbiased(on != 0); });
}
int hm_renderer_get_layout(const hm_renderer* r, int* out4) {
    return guarded([&] {
        need(r, "renderer"); need(out4, "out4");
        out4[0] = r->r->in_channels(); out4[1] = r->r->nn_frame_rows(); out4[2] = r->r->train_records(); out4[3] = r->r->every_nth();
    });
}
int hm_msnn_train_data_gen(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->msnn_train_data_gen(); r->r->sync(); });
}
int hm_msnn_pretrain(hm_renderer* r, int n) {
    return guarded([&] { need(r, "renderer"); r->r->msnn_pretrain(n); r->r->sync(); });
}
hm_mlp* hm_renderer_mlp(hm_renderer* r) { return (r && r->mlp_view.m) ? &r->mlp_view : nullptr; }

int hm_get_device_buffer(hm_renderer* r, int which, void** dev_ptr, size_t* bytes) {
    return guarded([&] {
        need(r, "renderer"); need(dev_ptr, "dev_ptr"); need(bytes, "bytes");
        void* p = r->r->device_buffer(which, bytes);
        if (!p) throw std::logic_error("buffer not available for this renderer kind");
        *dev_ptr = p;
    });
}

int hm_get_buffer(hm_renderer* r, int which, void* dst, size_t bytes) {
    return guarded([&] {
        need(r, "renderer"); need(dst, "host_dst");
        size_t have = 0;
        void* p = r->r->device_buffer(which, &have);
        if (!p) throw std::logic_error("buffer not available for this renderer kind");
        if (bytes > have) throw std::invalid_argument("requested more bytes than the buffer holds");
        r->r->sync();
        cuda_ok(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost), "hm_get_buffer");
    });
}

int hm_readback_async(hm_renderer* r, int which, void* dst, size_t bytes) {
    return guarded([&] { need(r, "renderer"); need(dst, "host_dst"); r->r->readback_async(which, dst, bytes); });
}

int hm_readback_rows_async(hm_renderer* r, int which, int row0, int rows, void* dst) {
    return guarded([&] {
        need(r, "renderer"); need(dst, "host_dst");
        if (which < 0 || which > HM_BUF_FB8) throw std::invalid_argument("hm_readback_rows_async: not an image buffer");
        const int W = r->r->width(), H = r->r->height();
        if (row0 < 0 || rows < 0 || row0 + rows > H) throw std::invalid_argument("row range outside the frame");
        r->r->readback_rows_async(which, row0, rows, dst);
        (void)W;
    });
}
int hm_renderer_get_rows(const hm_renderer* r, int* out2) {
    return guarded([&] { need(r, "renderer"); need(out2, "out2"); out2[0] = r->r->row0(); out2[1] = r->r->row1(); });
}

int hm_save_png(hm_renderer* r, const char* path) {
    return guarded([&] {
        need(r, "renderer"); need(path, "path");
        const int W = r->r->width(), H = r->r->height();
        std::vector<uint32_t> fb((size_t)W * H);
        size_t have = 0;
        void* p = r->r->device_buffer(HM_BUF_FB8, &have);
        r->r->sync();
        cuda_ok(cudaMemcpy(fb.data(), p, have, cudaMemcpyDeviceToHost), "hm_save_png");
        write_png_flipped(path, fb.data(), W, H);
    });
}

int hm_save_exr(hm_renderer* r, int which, const char* path) {
    return guarded([&] {
        need(r, "renderer"); need(path, "path");
        if (which < 0 || which > 5) throw std::invalid_argument("hm_save_exr: not a float4 image buffer");
        const int W = r->r->width(), H = r->r->height();
        std::vector<float> img((size_t)W * H * 4);
        size_t have = 0;
        void* p = r->r->device_buffer(which, &have);
        if (!p) throw std::logic_error("buffer not available for this renderer kind");
        r->r->sync();
        cuda_ok(cudaMemcpy(img.data(), p, have, cudaMemcpyDeviceToHost), "hm_save_exr");
        write_exr_flipped(path, img.data(), W, H);
    });
}

int hm_renderer_get_stats(hm_renderer* r, hm_stats* out) {
    return guarded([&] {
        need(r, "renderer"); need(out, "out");
        Stats s = r->r->stats();
        memset(out, 0, sizeof(*out));
        out->ms_primary = s.ms[0]; out->ms_shade = s.ms[1]; out->ms_extend = s.ms[2]; out->ms_shadow = s.ms[3];
        out->ms_finalize = s.ms[4]; out->ms_train = s.ms[5]; out->ms_infer = s.ms[6]; out->ms_composite = s.ms[7];
        out->ms_total = s.ms[8];
        out->rays_primary = s.rays_primary; out->rays_extend = s.rays_extend; out->rays_shadow = s.rays_shadow;
        out->shade_items = s.shade_items;
        out->kernel_launches = wavefront_launch_count() + (r->r->mlp() ? r->r->mlp()->launch_count() : 0);
        for (int i = 0; i < 8; ++i) out->stage_launches[i] = s.launches[i];
        for (int i = 0; i < 8; ++i) out->timed_launches[i] = s.timed_launches[i];
        out->trav_nodes_extend = s.trav[0]; out->trav_prims_extend = s.trav[1];
        out->trav_nodes_shadow = s.trav[2]; out->trav_prims_shadow = s.trav[3];
        out->trav_nodes_primary = s.trav[4]; out->trav_prims_primary = s.trav[5];
        out->trav_nodes_tail = s.tail_nodes; out->trav_prims_tail = s.tail_prims; out->rays_tail = s.tail_rays;
        out->last_loss = s.last_loss;
        out->frames = s.frames;
    });
}
int hm_renderer_set_collect_stats(hm_renderer* r, int on) {
    return guarded([&] { need(r, "renderer"); r->r->set_collect_stats(on != 0); });
}
int hm_renderer_set_profiling_stages(hm_renderer* r, unsigned mask) {
    return guarded([&] { need(r, "renderer"); r->r->set_profiling_stages(mask); });
}
int hm_renderer_set_profiling_period(hm_renderer* r, int n) {
    return guarded([&] { need(r, "renderer"); r->r->set_profiling_period(n); });
}
int hm_renderer_set_skip_unused_queries(hm_renderer* r, int on) {
    return guarded([&] { need(r, "renderer"); r->r->set_skip_unused_queries(on != 0); });
}
int hm_renderer_reset_stats(hm_renderer* r) {
    return guarded([&] { need(r, "renderer"); r->r->reset_stats(); });
}
int hm_renderer_set_frame_schedule(hm_renderer* r, int offset, int stride) {
    return guarded([&] {
        need(r, "renderer");
        if (stride < 1 || offset < 0) throw std::invalid_argument("bad frame schedule");
        r->r->set_frame_schedule(offset, stride);
    });
}
int hm_nrc_layout(int width, int height, int* out4) {
    return guarded([&] {
        need(out4, "out4");
        if (width <= 0 || height <= 0) throw std::invalid_argument("bad frame size");
        const NrcLayout l = nrc_layout(width, height);
        if (l.every_nth < 1) throw std::invalid_argument("render_nrc needs at least 1638 pixels (everyNth would be 0)");
        if (((long long)width * height) % 128) throw std::invalid_argument("render_nrc needs W*H to be a multiple of 128 (scene.cpp:302-306)");
        out4[0] = l.train_pixels; out4[1] = l.every_nth; out4[2] = l.nn_frame_rows; out4[3] = l.records;
    });
}
int hm_band_partition(int width, int height, int records, int rank, int world, int* out5) {
    return guarded([&] {
        if (!out5 || width <= 0 || height <= 0 || records < 0 || world < 1 || rank < 0 || rank >= world)
            throw std::invalid_argument("bad partition arguments");
        const hm::BandPartition b = hm::band_partition(width, height, records, rank, world);
        out5[0] = b.row0; out5[1] = b.row1; out5[2] = b.slot0; out5[3] = b.slots; out5[4] = b.train_n;
    });
}
int hm_renderer_set_profiling(hm_renderer* r, int on) {
    return guarded([&] { need(r, "renderer"); r->r->set_profiling(on != 0); });
}

// integrator.stats_output: the reference parses the key (scene.cpp:295) and never writes the file; this is
// the schema the headless executables write (SURVEY §5 "Metrics"): samples, seconds, Mpaths/s, per-stage
// milliseconds (event pairs, when profiling is on), ray counts and traversal work per ray (when the
// instrumented traversal is on), network queries per second, training loss.
static std::string stats_json(const Stats& s, int kind, int width, int rows, int nn_rows_per_frame) {
    std::ostringstream f;
    auto num = [](double v) -> std::string {
        if (!std::isfinite(v)) return "null";
        std::ostringstream o; o.precision(9); o << v; return o.str();
    };
    const double paths = (double)s.frames * rows * width;
    const char* kinds[3] = {"render_path_tracing", "render_nrc", "render_hair_msnn"};
    f << "{\n  \"renderer\": \"" << kinds[kind] << "\",\n";
    f << "  \"width\": " << width << ", \"rows\": " << rows << ", \"spp\": " << s.frames << ",\n";
    f << "  \"paths\": " << num(paths) << ",\n";
    f << "  \"seconds\": " << num(s.ms[8] * 1e-3) << ",\n";
    f << "  \"mpaths_per_s\": " << num(s.ms[8] > 0 ? paths / (s.ms[8] * 1e-3) / 1e6 : 0.0) << ",\n";
    f << "  \"ms\": {\"primary\": " << num(s.ms[0]) << ", \"shade_main\": " << num(s.ms[1]) << ", \"trace_main\": " << num(s.ms[2])
      << ", \"tail_piece\": " << num(s.ms[3]) << ", \"finalize\": " << num(s.ms[4]) << ", \"train\": " << num(s.ms[5])
      << ", \"infer\": " << num(s.ms[6]) << ", \"composite\": " << num(s.ms[7]) << "},\n";
    uint64_t launches = 0;
    for (int i = 0; i < 8; ++i) launches += s.launches[i];
    f << "  \"kernel_launches\": " << launches << ",\n";
    f << "  \"rays\": {\"primary\": " << s.rays_primary << ", \"extend\": " << s.rays_extend << ", \"shadow\": "
      << s.rays_shadow << "},\n";
    auto per = [&](uint64_t a, uint64_t b) { return b ? num((double)a / (double)b) : std::string("null"); };
    f << "  \"traversal_per_ray\": {\"primary\": {\"nodes\": " << per(s.trav[4], s.rays_primary) << ", \"primitives\": " << per(s.trav[5], s.rays_primary)
      << "}, \"extend\": {\"nodes\": " << per(s.trav[0], s.rays_extend) << ", \"primitives\": " << per(s.trav[1], s.rays_extend)
      << "}, \"shadow\": {\"nodes\": " << per(s.trav[2], s.rays_shadow) << ", \"primitives\": " << per(s.trav[3], s.rays_shadow) << "}},\n";
    f << "  \"mlp_rows_per_frame\": " << nn_rows_per_frame << ",\n";
    f << "  \"mlp_queries_per_s\": " << (s.ms[6] > 0 ? num((double)nn_rows_per_frame * s.frames / (s.ms[6] * 1e-3)) : std::string("null")) << ",\n";
    f << "  \"training_loss\": " << num(s.last_loss) << "\n}\n";
    return f.str();
}

int hm_write_stats(hm_renderer* r, const char* path) {
    return guarded([&] {
        need(r, "renderer"); need(path, "path");
        Stats s = r->r->stats();
        std::ofstream f(path);
        if (!f) throw hm::IoError(std::string("cannot write ") + path);
        f << stats_json(s, r->r->kind(), r->r->width(), r->r->row1() - r->r->row0(), r->r->mlp() ? r->r->nn_frame_rows() : 0);
        if (!f) throw hm::IoError(std::string("cannot write ") + path);
    });
}
// test hook (not part of the ABI header): the stats schema on made-up counters, incl. a non-finite loss
int hm_test_stats_json(char* buf, size_t capacity) {
    return guarded([&] {
        need(buf, "buf");
        Stats s;
        for (int i = 0; i < 9; ++i) { s.ms[i] = 1.5 * (i + 1); s.launches[i] = 10 + i; }
        s.frames = 4; s.rays_primary = 1000; s.rays_extend = 400; s.rays_shadow = 0;
        s.trav[0] = 12000; s.trav[1] = 1600; s.trav[4] = 15000; s.trav[5] = 1900;
        s.last_loss = std::numeric_limits<float>::quiet_NaN();
        const std::string j = stats_json(s, HM_RENDER_HAIR_MSNN, 256, 128, 32768);
        if (j.size() + 1 > capacity) throw std::invalid_argument("buffer too small");
        memcpy(buf, j.c_str(), j.size() + 1);
    });
}

// ---- stand-alone kernels ------------------------------------------------------------
int hm_trace_rays_device(hm_renderer* r, const float* d_org, const float* d_dir, int n, int any, float tmin, float tmax,
                         float* d_out) {
    return guarded([&] {
        need(r, "renderer"); need(d_org, "d_org3"); need(d_dir, "d_dir3"); need(d_out, "d_out_hit4");
        if (n > 0) r->r->trace_rays_device(d_org, d_dir, n, any, tmin, tmax, d_out, nullptr);
    });
}

int hm_trace_rays(hm_renderer* r, const float* org, const float* dir, int n, int any, float tmin, float tmax,
                  float* out_hit, int* out_stats) {
    return guarded([&] {
        need(r, "renderer"); need(org, "org3"); need(dir, "dir3"); need(out_hit, "out_hit4");
        if (n <= 0) return;
        cuda_ok(cudaSetDevice(r->r->device()), "set device");
        float *d_o = nullptr, *d_d = nullptr, *d_h = nullptr;
        int* d_s = nullptr;
        cuda_ok(cudaMalloc(&d_o, (size_t)n * 12), "malloc");
        cuda_ok(cudaMalloc(&d_d, (size_t)n * 12), "malloc");
        cuda_ok(cudaMalloc(&d_h, (size_t)n * 16), "malloc");
        if (out_stats) cuda_ok(cudaMalloc(&d_s, (size_t)n * 8), "malloc");
        cuda_ok(cudaMemcpy(d_o, org, (size_t)n * 12, cudaMemcpyHostToDevice), "h2d");
        cuda_ok(cudaMemcpy(d_d, dir, (size_t)n * 12, cudaMemcpyHostToDevice), "h2d");
        r->r->trace_rays_device(d_o, d_d, n, any, tmin, tmax, d_h, d_s);
        r->r->sync();
        cuda_ok(cudaMemcpy(out_hit, d_h, (size_t)n * 16, cudaMemcpyDeviceToHost), "d2h");
        if (out_stats) cuda_ok(cudaMemcpy(out_stats, d_s, (size_t)n * 8, cudaMemcpyDeviceToHost), "d2h");
        cudaFree(d_o); cudaFree(d_d); cudaFree(d_h); cudaFree(d_s);
    });
}

// ---- MLP ----------------------------------------------------------------------------
int hm_mlp_create(const char* config_path, int in_ch, int out_ch, int device, hm_mlp** out) {
    return guarded([&] {
        need(out, "out");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw std::runtime_error("CUDA: no usable device (this library has no CPU path)");
        if (device < 0 || device >= ndev) throw std::runtime_error("CUDA: device index out of range");
        cuda_ok(cudaSetDevice(device), "set device");
        MlpConfig cfg = config_path ? mlp_config_from_json(config_path, in_ch, out_ch) : MlpConfig();
        cfg.in_ch = in_ch; cfg.out_ch = out_ch;
        std::unique_ptr<hm_mlp> h(new hm_mlp);
        cuda_ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "stream");
        h->owned.reset(new Mlp(cfg, h->stream));
        h->m = h->owned.get();
        h->device = device;
        *out = h.release();
    });
}
void hm_mlp_destroy(hm_mlp* m) {
    if (!m || !m->owned) return;   // renderer-owned views are not freed here
    cudaSetDevice(m->device);
    m->owned.reset();
    cudaStreamDestroy(m->stream);
    delete m;
}

static void check_batch(int n) {
    if (n <= 0 || n % 128 != 0) throw std::invalid_argument("batch size must be a positive multiple of 128 (tcnn common.h:280)");
}

int hm_mlp_inference(hm_mlp* m, const float* d_in, float* d_out, int n) {
    return guarded([&] { need(m, "mlp"); flush_owner(m); need(d_in, "d_in"); need(d_out, "d_out"); check_batch(n); m->m->inference(d_in, d_out, n); });
}
int hm_mlp_inference_host(hm_mlp* m, const float* in, float* out, int n) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(in, "in"); need(out, "out"); check_batch(n);
        cuda_ok(cudaSetDevice(m->device), "set device");
        const int ic = m->m->config().in_ch, oc = m->m->config().out_ch;
        float *d_i = nullptr, *d_o = nullptr;
        cuda_ok(cudaMalloc(&d_i, (size_t)n * ic * 4), "malloc");
        cuda_ok(cudaMalloc(&d_o, (size_t)n * oc * 4), "malloc");
        cuda_ok(cudaMemcpyAsync(d_i, in, (size_t)n * ic * 4, cudaMemcpyHostToDevice, m->stream), "h2d");
        m->m->inference(d_i, d_o, n);
        cuda_ok(cudaMemcpyAsync(out, d_o, (size_t)n * oc * 4, cudaMemcpyDeviceToHost, m->stream), "d2h");
        cuda_ok(cudaStreamSynchronize(m->stream), "sync");
        cudaFree(d_i); cudaFree(d_o);
    });
}
int hm_mlp_forward_backward(hm_mlp* m, const float* d_in, const float* d_target, int n, int n_total) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(d_in, "d_in"); need(d_target, "d_target"); check_batch(n);
        m->m->forward_backward(d_in, d_target, n, n_total > 0 ? n_total : n);
    });
}
int hm_mlp_gradients(hm_mlp* m, float** d_grads, size_t* count) {
    return guarded([&] { need(m, "mlp"); flush_owner(m); need(d_grads, "d_grads"); need(count, "count"); *d_grads = m->m->gradients(); *count = m->m->n_params(); });
}
int hm_mlp_optimizer_step(hm_mlp* m) { return guarded([&] { need(m, "mlp"); flush_owner(m); m->m->optimizer_step(); }); }
int hm_mlp_loss(hm_mlp* m, float* loss) { return guarded([&] { need(m, "mlp"); flush_owner(m); need(loss, "loss"); *loss = m->m->loss(); }); }
int hm_mlp_train_step(hm_mlp* m, const float* d_in, const float* d_target, int n, float* loss) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(d_in, "d_in"); need(d_target, "d_target"); check_batch(n);
        m->m->forward_backward(d_in, d_target, n, n);
        m->m->optimizer_step();
        if (loss) *loss = m->m->loss();
    });
}
int hm_mlp_train_step_host(hm_mlp* m, const float* in, const float* target, int n, float* loss) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(in, "in"); need(target, "target"); check_batch(n);
        cuda_ok(cudaSetDevice(m->device), "set device");
        const int ic = m->m->config().in_ch, oc = m->m->config().out_ch;
        float *d_i = nullptr, *d_t = nullptr;
        cuda_ok(cudaMalloc(&d_i, (size_t)n * ic * 4), "malloc");
        cuda_ok(cudaMalloc(&d_t, (size_t)n * oc * 4), "malloc");
        cuda_ok(cudaMemcpyAsync(d_i, in, (size_t)n * ic * 4, cudaMemcpyHostToDevice, m->stream), "h2d");
        cuda_ok(cudaMemcpyAsync(d_t, target, (size_t)n * oc * 4, cudaMemcpyHostToDevice, m->stream), "h2d");
        m->m->forward_backward(d_i, d_t, n, n);
        m->m->optimizer_step();
        float l = m->m->loss();
        if (loss) *loss = l;
        cudaFree(d_i); cudaFree(d_t);
    });
}
int hm_mlp_reset(hm_mlp* m) { return guarded([&] { need(m, "mlp"); flush_owner(m); m->m->reset_weights(); }); }
int hm_mlp_reinitialize(hm_mlp* m) { return guarded([&] { need(m, "mlp"); flush_owner(m); m->m->reinitialize(); }); }
size_t hm_mlp_n_params(const hm_mlp* m) { return m ? m->m->n_params() : 0; }
int hm_mlp_get_params(hm_mlp* m, float* dst, size_t count) {
    return guarded([&] { need(m, "mlp"); flush_owner(m); need(dst, "host_dst"); m->m->get_params(dst, count); });
}
int hm_mlp_set_params(hm_mlp* m, const float* src, size_t count) {
    return guarded([&] { need(m, "mlp"); flush_owner(m); need(src, "host_src"); m->m->set_params(src, count); });
}
int hm_mlp_save(hm_mlp* m, const char* path) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(path, "path");
        std::vector<float> p(m->m->n_params());
        m->m->get_params(p.data(), p.size());
        std::ofstream f(path, std::ios::binary);
        if (!f) throw hm::IoError(std::string("cannot write ") + path);
        const char magic[8] = {'H', 'M', 'S', 'N', 'N', 'W', '1', 0};
        uint64_t n = p.size();
        f.write(magic, 8); f.write((const char*)&n, 8); f.write((const char*)p.data(), n * 4);
    });
}
// half bits -> float (snapshots of type "__half")
static float snapshot_half_to_float(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 0x1f, mnt = h & 0x3ff, bits;
    if (e == 0) {
        if (mnt == 0) bits = s << 31;
        else {
            int ee = -1;
            do { ee++; mnt <<= 1; } while (!(mnt & 0x400));
            bits = (s << 31) | ((uint32_t)(127 - 15 - ee) << 23) | ((mnt & 0x3ff) << 13);
        }
    } else if (e == 31) bits = (s << 31) | 0x7f800000u | (mnt << 13);
    else bits = (s << 31) | ((e + 112) << 23) | (mnt << 13);
    float f; memcpy(&f, &bits, 4);
    return f;
}

// Trainer::deserialize (trainer.h:285-310) over the TEXT form TINY_MLP::loadWeights reads
// (cuda/neural_network.cu:23-32): {"n_params": N, "params_type": "float" | "__half",
// "params_binary": {"bytes": [...], "subtype": null}} — nlohmann's JSON rendering of a binary value,
// accepted by tcnn's from_json (gpu_memory_json.h:58-67).  An "optimizer" member, if present, is ignored
// (the reference never writes one: serialize_optimizer defaults to false).
static void load_tcnn_snapshot(hm::Mlp& mlp, const std::string& path) {
    hm::Json j = hm::parse_json_file(path);
    std::string type = "__half";                       // type_to_string<precision_t>() of the reference build
    if (const hm::Json* t = j.find("params_type")) type = t->string();
    const hm::Json& pb = j.at("params_binary");
    const hm::Json* bytes = pb.type == hm::Json::Object ? pb.find("bytes") : nullptr;
    if (!bytes || bytes->type != hm::Json::Array) throw std::invalid_argument("snapshot: params_binary must be {\"bytes\": [...]}");
    const size_t nb = bytes->arr.size();
    std::vector<uint8_t> raw(nb);
    for (size_t i = 0; i < nb; ++i) raw[i] = (uint8_t)bytes->arr[i].number();
    std::vector<float> p;
    if (type == "float") {
        if (nb % 4) throw std::invalid_argument("snapshot: byte count is not a multiple of 4");
        p.resize(nb / 4);
        memcpy(p.data(), raw.data(), nb);
    } else if (type == "__half") {
        if (nb % 2) throw std::invalid_argument("snapshot: byte count is not a multiple of 2");
        p.resize(nb / 2);
        for (size_t i = 0; i < p.size(); ++i) { uint16_t h; memcpy(&h, raw.data() + 2 * i, 2); p[i] = snapshot_half_to_float(h); }
    } else {
        throw std::invalid_argument("Trainer: snapshot parameters must be of type float of __half");   // the reference's message
    }
    if (p.size() != mlp.n_params()) throw std::invalid_argument("snapshot has a different parameter count");
    mlp.set_params(p.data(), p.size());
}

int hm_mlp_load(hm_mlp* m, const char* path) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(path, "path");
        std::ifstream f(path, std::ios::binary);
        if (!f) throw hm::IoError(std::string("cannot read ") + path);
        char magic[8]; uint64_t n = 0;
        f.read(magic, 8);
        if (f && memcmp(magic, "HMSNNW1", 7) != 0) {   // not the raw blob: tiny-cuda-nn's JSON snapshot
            f.close();
            load_tcnn_snapshot(*m->m, path);
            return;
        }
        f.read((char*)&n, 8);
        if (!f) throw std::invalid_argument("not a weight file");
        if (n != m->m->n_params()) throw std::invalid_argument("weight file has a different parameter count");
        std::vector<float> p(n);
        f.read((char*)p.data(), n * 4);
        if (!f) throw hm::IoError("truncated weight file");
        m->m->set_params(p.data(), p.size());
    });
}
// Trainer::serialize (trainer.h:270-283) as text JSON, params_type "float"
int hm_mlp_save_snapshot(hm_mlp* m, const char* path) {
    return guarded([&] {
        need(m, "mlp"); flush_owner(m); need(path, "path");
        std::vector<float> p(m->m->n_params());
        m->m->get_params(p.data(), p.size());
        std::ofstream f(path, std::ios::binary);
        if (!f) throw hm::IoError(std::string("cannot write ") + path);
        std::string out;
        out.reserve(p.size() * 16 + 256);
        out += "{\"n_params\":" + std::to_string(p.size()) + ",\"params_binary\":{\"bytes\":[";
        const uint8_t* b = (const uint8_t*)p.data();
        char tmp[8];
        for (size_t i = 0; i < p.size() * 4; ++i) {
            int len = snprintf(tmp, sizeof(tmp), i ? ",%u" : "%u", (unsigned)b[i]);
            out.append(tmp, (size_t)len);
        }
        out += "],\"subtype\":null},\"params_type\":\"float\"}";
        f.write(out.data(), (std::streamsize)out.size());
        if (!f) throw hm::IoError(std::string("cannot write ") + path);
    });
}
void* hm_mlp_stream(hm_mlp* m) { return m ? (void*)m->stream : nullptr; }
uint64_t hm_mlp_launch_count(const hm_mlp* m) { return m ? m->m->launch_count() : 0; }

}  // extern "C"
