// hm_renderer.cu — see hm_renderer.h.
#include "hm_renderer.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/random.h>
#include <thrust/shuffle.h>

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace hm {

#define HM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));               \
    } while (0)

// ---------------------------------------------------------------------------------
template <typename T>
T* DeviceScene::upload(const T* src, size_t n) {
    if (n == 0) n = 1;
    void* p = nullptr;
    HM_CUDA(cudaMalloc(&p, n * sizeof(T)));
    if (src) HM_CUDA(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    else HM_CUDA(cudaMemset(p, 0, n * sizeof(T)));
    allocs_.push_back(p);
    bytes += n * sizeof(T);
    return (T*)p;
}

DeviceScene::DeviceScene(const HostScene& hs) {
    const HostGeometry& g = hs.geo;
    const HostBvh& b = hs.bvh;
    memset(&view, 0, sizeof(view));
    view.geom.nodes = upload(b.nodes.data(), b.nodes.size());
    view.geom.leaf_data = upload(b.leaf_data.data(), b.leaf_data.size());
    view.geom.leaf_code = nullptr;   // host-side only
    view.geom.leaf_prim = upload(b.leaf_prim.data(), b.leaf_prim.size());
    view.geom.cps = upload(g.cps.data(), g.cps.size());
    view.geom.tri_verts = upload(g.tri_verts.data(), g.tri_verts.size());
    view.geom.num_segments = (int)g.seg_cp.size();
    view.geom.num_tris = (int)(g.tri_verts.size() / 3);
    view.geom.num_nodes = (int)(b.nodes.size() / 4);
    view.seg_cp = upload(g.seg_cp.data(), g.seg_cp.size());
    view.tri_normals = upload(g.tri_normals.data(), g.tri_normals.size());

    LightSet& L = view.lights;
    L.env.has_env = hs.has_env ? 1 : 0;
    L.env.pdf_sampling = hs.env_pdf ? 1 : 0;
    L.env.W = hs.env_w; L.env.H = hs.env_h;
    L.env.scale = hs.env_scale; L.env.rot_phi = hs.env_rot;
    if (hs.has_env) {
        L.env.env = upload(hs.env.data(), hs.env.size());
        L.env.cpdf = upload(hs.cpdf.data(), hs.cpdf.size());
        L.env.ccdf = upload(hs.ccdf.data(), hs.ccdf.size());
        L.env.mpdf = upload(hs.mpdf.data(), hs.mpdf.size());
        L.env.mcdf = upload(hs.mcdf.data(), hs.mcdf.size());
    }
    L.num_dlights = (int)(hs.dl_from.size() / 3);
    if (L.num_dlights > kMaxDirLights) throw std::runtime_error("too many directional lights (max 8)");
    for (int i = 0; i < L.num_dlights; ++i)
        for (int k = 0; k < 3; ++k) { L.dl_from[i][k] = hs.dl_from[3 * i + k]; L.dl_emit[i][k] = hs.dl_emit[3 * i + k]; }
    L.num_total = L.num_dlights + (hs.has_env ? 1 : 0);

    view.lobes.setup(hs.beta_m, hs.beta_n, hs.alpha);
    view.lobes.sigma_a = V3(hs.sigma_a[0], hs.sigma_a[1], hs.sigma_a[2]);
    for (int i = 0; i < 4; ++i) view.lobes.gain[i] = hs.gains[i];
    for (int k = 0; k < 3; ++k) view.kd[k] = g.kd[k];
    view.surf_alpha = g.surf_alpha;
    view.scene_scale = g.scene_scale;
    view.mis = hs.mis ? 1 : 0;
}

DeviceScene::~DeviceScene() {
    for (void* p : allocs_) cudaFree(p);
}

// ---------------------------------------------------------------------------------
Renderer::Renderer(const HostScene& hs, int kind, int beta_cli, int device, int rank, int world)
    : hs_(hs), kind_(kind), beta_(beta_cli - 1), device_(device), rank_(rank), world_(world) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("CUDA: no usable device (this library has no CPU path)");
    if (device < 0 || device >= ndev) throw std::runtime_error("CUDA: device index out of range");
    HM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw std::runtime_error("CUDA: device is not sm_100-class (kernels are built for sm_100a only)");
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad rank/world");
    W_ = hs.width; H_ = hs.height;
    if (W_ <= 0 || H_ <= 0) throw std::invalid_argument("bad frame size");
    // contiguous row bands; band edges are multiples of 8 rows where possible so that
    // every band holds whole training-record groups and a multiple of 128 pixels
    row0_ = (int)((int64_t)H_ * rank / world);
    row1_ = (int)((int64_t)H_ * (rank + 1) / world);
    HM_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    scene_.reset(new DeviceScene(hs));
    camera_basis(hs, W_, H_, cam_.pos, cam_.d00, cam_.du, cam_.dv);

    const size_t n = (size_t)W_ * H_;
    auto alloc = [&](size_t bytes) {
        void* p = nullptr;
        HM_CUDA(cudaMalloc(&p, bytes));
        HM_CUDA(cudaMemset(p, 0, bytes));
        allocs_.push_back(p);
        return p;
    };
    paths_.rng = (uint32_t*)alloc(n * 4);
    paths_.ray_o = (float4*)alloc(n * 16);
    paths_.ray_d = (float4*)alloc(n * 16);
    paths_.hit = (float4*)alloc(n * 16);
    paths_.beta = (float4*)alloc(n * 16);
    paths_.color = (float4*)alloc(n * 16);
    paths_.dl_beta = (float4*)alloc(n * 16);
    paths_.dl_light = (float4*)alloc(n * 16);
    paths_.dl_bsdf = (float4*)alloc(n * 16);
    paths_.vis = (uint32_t*)alloc(n * 4);
    if (kind_ == HM_KIND_MSNN) {
        paths_.beta_short = (float4*)alloc(n * 16);
        paths_.color_short = (float4*)alloc(n * 16);
        paths_.dl_beta_short = (float4*)alloc(n * 16);
    }
    q_.shade[0] = (int*)alloc(n * 4);
    q_.shade[1] = (int*)alloc(n * 4);
    q_.extend = (int*)alloc(n * 4);
    q_.shadow = (float4*)alloc(n * 2 * 32);
    q_.counts = (int*)alloc(16 * 4);
    d_trav_ = (unsigned long long*)alloc(8 * 8);
    q_.trav = d_trav_;
    HM_CUDA(cudaMallocHost((void**)&h_counts_, 16 * 4));
    memset(h_counts_, 0, 16 * 4);

    for (int i = 0; i < 6; ++i) bufs_[i] = (float4*)alloc(n * 16);
    fb_ = (uint32_t*)alloc(n * 4);

    if (kind_ == HM_KIND_MSNN) {
        in_ch_ = 12;
        records_ = 128 * 128;   // numTrainRecordsX * numTrainRecordsY (headers/render_hair_msnn.h:127-130)
        // everyNth = std::ceil(frameSize / numTrainRecords) with INTEGER division (render_hair_msnn.cu:106)
        every_nth_ = (int)(n / (size_t)records_);
        if (every_nth_ < 1)
            throw std::invalid_argument("render_hair_msnn needs at least 16384 pixels (everyNth would be 0)");
        MlpConfig cfg = hs.tcnn_config.empty() ? MlpConfig() : mlp_config_from_json(hs.tcnn_config, in_ch_, 3);
        cfg.in_ch = in_ch_; cfg.out_ch = 3;
        mlp_.reset(new Mlp(cfg, stream_));
        h_train_idxs_.resize(records_);
        for (int i = 0; i < records_; ++i) h_train_idxs_[i] = i;   // thrust::sequence
        d_train_idxs_ = (int*)alloc((size_t)records_ * 4);
        HM_CUDA(cudaMemcpy(d_train_idxs_, h_train_idxs_.data(), (size_t)records_ * 4, cudaMemcpyHostToDevice));
        nn_frame_in_ = (float*)alloc(n * in_ch_ * 4);
        nn_frame_out_ = (float*)alloc(n * 3 * 4);
        nn_train_in_ = (float*)alloc((size_t)records_ * in_ch_ * 4);
        nn_train_out_ = (float*)alloc((size_t)records_ * 3 * 4);
        gbuffer_ = (float4*)alloc(n * 16);
    }
    HM_CUDA(cudaDeviceSynchronize());
}

Renderer::~Renderer() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    mlp_.reset();
    for (void* p : allocs_) cudaFree(p);
    if (h_counts_) cudaFreeHost(h_counts_);
    scene_.reset();
    for (auto& p : pending_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : event_pool_) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
}

void Renderer::sync() {
    HM_CUDA(cudaSetDevice(device_));
    HM_CUDA(cudaStreamSynchronize(stream_));
}

cudaEvent_t Renderer::take_event() {
    if (!event_pool_.empty()) { cudaEvent_t e = event_pool_.back(); event_pool_.pop_back(); return e; }
    cudaEvent_t e;
    HM_CUDA(cudaEventCreate(&e));
    return e;
}

// Event pairs are recorded around each stage launch and resolved later, so profiling does
// not add host synchronisation inside the frame.
template <typename F>
void Renderer::timed(int stage, F&& f) {
    stats_.launches[stage]++;
    if (!profiling_) { f(); return; }
    Pending p{stage, take_event(), take_event()};
    HM_CUDA(cudaEventRecord(p.a, stream_));
    f();
    HM_CUDA(cudaEventRecord(p.b, stream_));
    pending_.push_back(p);
}

void Renderer::resolve_events() {
    if (pending_.empty()) return;
    HM_CUDA(cudaStreamSynchronize(stream_));
    for (auto& p : pending_) {
        float ms = 0.f;
        HM_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
        stats_.ms[p.stage] += ms;
        event_pool_.push_back(p.a); event_pool_.push_back(p.b);
    }
    pending_.clear();
}

void Renderer::reset_stats() {
    resolve_events();
    stats_ = Stats();
    cudaMemsetAsync(d_trav_, 0, 8 * 8, stream_);
}

FrameParams Renderer::base_params() {
    FrameParams P;
    memset(&P, 0, sizeof(P));
    P.scene = scene_->view;
    P.cam = cam_;
    P.paths = paths_;
    P.q = q_;
    P.W = W_; P.H = H_;
    P.row0 = row0_; P.row1 = row1_;
    P.accum_id = accum_id_;
    P.frame_id = frame_offset_ + accum_id_ * frame_stride_;
    P.collect_stats = collect_stats_ ? 1 : 0;
    P.v1_stop = hs_.path_v1 - 1;
    P.v2_stop = hs_.path_v2 - 1;
    P.accum = bufs_[1]; P.average = bufs_[0]; P.fb = fb_;
    P.in_ch = in_ch_;
    return P;
}

// One wavefront loop: primary, then (shade -> shadow + extend) per path vertex.
void Renderer::trace_bounces(FrameParams& P, int max_vertices) {
    HM_CUDA(cudaMemsetAsync(q_.counts, 0, 16 * 4, stream_));
    timed(0, [&] { launch_primary(P, stream_); });
    stats_.rays_primary += (uint64_t)(row1_ - row0_) * W_;
    int src = 0;
    for (int vertex = 0; vertex < max_vertices; ++vertex) {
        timed(1, [&] { launch_shade(P, src, stream_); });
        timed(3, [&] { launch_shadow(P, stream_); });
        const int dst = src ^ 1;
        HM_CUDA(cudaMemsetAsync(q_.counts + dst, 0, 4, stream_));
        timed(2, [&] { launch_extend(P, dst, stream_); });
        const bool last = vertex + 1 >= max_vertices;
        // Long paths (PT, training paths): look at the queue sizes every few vertices so
        // the loop ends once every path died.  Short HairMSNN paths never synchronise.
        const bool poll = collect_stats_ || (!last && max_vertices > 4 && (vertex & 1) == 1);
        if (poll) {
            HM_CUDA(cudaMemcpyAsync(h_counts_, q_.counts, 16 * 4, cudaMemcpyDeviceToHost, stream_));
            HM_CUDA(cudaStreamSynchronize(stream_));
            stats_.shade_items += (uint64_t)h_counts_[src];
            stats_.rays_extend += (uint64_t)h_counts_[2];
            stats_.rays_shadow += (uint64_t)h_counts_[3];
            if (h_counts_[dst] == 0) break;
        }
        HM_CUDA(cudaMemsetAsync(q_.counts + 2, 0, 8, stream_));   // extend + shadow counters
        src = dst;
    }
    timed(4, [&] { launch_finalize(P, stream_); });
}

void Renderer::frame_pt() {
    FrameParams P = base_params();
    P.mode = MODE_PT;
    // vertices shaded: the primary hit plus up to v2_stop bounces
    int max_vertices = P.v2_stop + 1;
    if (max_vertices < 1) max_vertices = 1;
    trace_bounces(P, max_vertices);
}

void Renderer::shuffle_train_idxs() {
    // thrust::shuffle(trainIdxs, default_random_engine()) — a freshly constructed engine
    // every frame, so the permutation applied is the same each time and the sequence of
    // compositions is deterministic (render_hair_msnn.cu:711-714)
    thrust::device_ptr<int> p = thrust::device_pointer_cast(d_train_idxs_);
    thrust::shuffle(thrust::cuda::par.on(stream_), p, p + records_, thrust::default_random_engine());
}

void Renderer::msnn_trace() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    HM_CUDA(cudaSetDevice(device_));
    shuffle_train_idxs();
    FrameParams P = base_params();
    P.mode = MODE_MSNN;
    P.msnn_beta = beta_;
    P.every_nth = every_nth_;
    P.train_idxs = d_train_idxs_;
    // this band's training records: fbOfs / everyNth over its pixel range
    const int64_t px0 = (int64_t)row0_ * W_, px1 = (int64_t)row1_ * W_;
    P.train_slot0 = (int)(px0 / every_nth_);
    int slot1 = (int)((px1 + every_nth_ - 1) / every_nth_);
    if (slot1 > records_) slot1 = records_;
    P.train_slots = slot1 - P.train_slot0;
    P.nn_frame_in = nn_frame_in_;
    P.nn_train_in = nn_train_in_;
    P.nn_train_out = nn_train_out_;
    P.gbuffer = gbuffer_;
    // Training paths run to full length (pathV2 vertices); with the queue polling in
    // trace_bounces the loop ends as soon as they have all terminated.
    int max_vertices = P.v2_stop + 1;
    if (max_vertices < 1) max_vertices = 1;
    trace_bounces(P, max_vertices);
}

void Renderer::msnn_train_backward() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    const int64_t px0 = (int64_t)row0_ * W_, px1 = (int64_t)row1_ * W_;
    int s0 = (int)(px0 / every_nth_);
    int s1 = (int)((px1 + every_nth_ - 1) / every_nth_);
    if (s1 > records_) s1 = records_;
    int n = s1 - s0;
    n -= n % 128;
    timed(5, [&] {
        mlp_->forward_backward(nn_train_in_ + (size_t)s0 * in_ch_, nn_train_out_ + (size_t)s0 * 3, n, world_ == 1 ? n : records_);
    });
}

void Renderer::msnn_train_apply() {
    timed(5, [&] { mlp_->optimizer_step(); });
}

void Renderer::msnn_finish() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    const size_t first = (size_t)row0_ * W_;
    const int count = (row1_ - row0_) * W_;
    timed(6, [&] { mlp_->inference(nn_frame_in_ + first * in_ch_, nn_frame_out_ + first * 3, count); });
    MsnnComposite C;
    C.final_avg = bufs_[0]; C.final_accum = bufs_[1];
    C.pt_avg = bufs_[2]; C.pt_accum = bufs_[3];
    C.nn_avg = bufs_[4]; C.nn_accum = bufs_[5];
    C.fb = fb_; C.gbuffer = gbuffer_; C.nn_out = nn_frame_out_;
    C.accum_id = accum_id_;
    C.first = (int)first; C.count = count;
    timed(7, [&] { launch_msnn_composite(C, stream_); });
    accum_id_++;
    stats_.frames++;
}

void Renderer::msnn_pretrain(int steps) {
    // The reference pre-trains for one wall-clock second on rays towards random strand
    // points from an UNSET camera.  Deterministic stand-in: `steps`
    // G_BUFFER passes from the real camera, each followed by a training step; the
    // accumulation counter advances as genTrainingData() does and is reset afterwards.
    for (int i = 0; i < steps; ++i) {
        msnn_trace();
        msnn_train_backward();
        msnn_train_apply();
        accum_id_++;
    }
    accum_id_ = 0;
}

void Renderer::render_frames(int n) {
    HM_CUDA(cudaSetDevice(device_));
    for (int i = 0; i < n; ++i) {
        Pending whole{8, nullptr, nullptr};
        if (profiling_) {
            whole.a = take_event(); whole.b = take_event();
            HM_CUDA(cudaEventRecord(whole.a, stream_));
        }
        if (kind_ == HM_KIND_PT) {
            frame_pt();
            accum_id_++;
            stats_.frames++;
        } else if (kind_ == HM_KIND_MSNN) {
            msnn_trace();
            if (hs_.tcnn_train) {
                msnn_train_backward();
                msnn_train_apply();
            }
            msnn_finish();
        } else {
            throw std::logic_error("render_nrc is not implemented in this build");
        }
        if (whole.a) {
            HM_CUDA(cudaEventRecord(whole.b, stream_));
            pending_.push_back(whole);
        }
    }
}

Stats Renderer::stats() {
    resolve_events();
    if (mlp_) stats_.last_loss = mlp_->loss();
    HM_CUDA(cudaStreamSynchronize(stream_));
    unsigned long long t[8];
    HM_CUDA(cudaMemcpy(t, d_trav_, sizeof(t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) stats_.trav[i] = t[i];
    return stats_;
}

void* Renderer::device_buffer(int which, size_t* bytes) {
    const size_t n = (size_t)W_ * H_;
    switch (which) {
        case 0: case 1: case 2: case 3: case 4: case 5: *bytes = n * 16; return bufs_[which];
        case 6: *bytes = n * 4; return fb_;
        case 7: *bytes = n * in_ch_ * 4; return nn_frame_in_;
        case 8: *bytes = n * 3 * 4; return nn_frame_out_;
        case 9: *bytes = (size_t)records_ * in_ch_ * 4; return nn_train_in_;
        case 10: *bytes = (size_t)records_ * 3 * 4; return nn_train_out_;
        case 11: *bytes = n * 16; return gbuffer_;
        case 12: *bytes = (size_t)records_ * 4; return d_train_idxs_;
        default: *bytes = 0; return nullptr;
    }
}

void Renderer::trace_rays_device(const float* d_org, const float* d_dir, int n, int any, float tmin, float tmax,
                                 float* d_out_hit, int* d_out_stats) {
    HM_CUDA(cudaSetDevice(device_));
    launch_trace_rays(scene_->view, d_org, d_dir, n, any, tmin, tmax, (float4*)d_out_hit, d_out_stats, stream_);
}

}  // namespace hm
