// hm_renderer.cu — see hm_renderer.h.
#include "hm_renderer.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/random.h>
#include <thrust/shuffle.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

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
    // The pooled traversal (hm_trace_dev.cuh) packs (leaf reference << 5 | lane) and (primitive id << 5 | lane) into 32-bit
    // words: 2^26 references / primitives at most (the curly scene has 39.0 M references, 3.5 M primitives).
    if (b.wleaf_data.size() / 4 >= ((size_t)1 << 26))
        throw std::invalid_argument("scene too large: the traversal supports up to 67 108 863 leaf references");
    // the kernels traverse the 8-wide quantised tree only; the binary tree stays on the host (oracle hook)
    view.geom.nodes = nullptr;
    view.geom.leaf_data = nullptr;
    view.geom.wnodes = upload(b.wnodes.data(), b.wnodes.size());
    view.geom.wleaf_data = upload(b.wleaf_data.data(), b.wleaf_data.size());
    view.geom.num_wnodes = (int)(b.wnodes.size() / 5);
    view.geom.k47 = 0x47000000u;
    view.geom.leaf_code = nullptr;   // host-side only
    view.geom.leaf_prim = nullptr;
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
        const char* where = getenv("HM_ENV_TABLES");
        if (where && std::string(where) == "host") {
            // A/B and test path: the host recipe's tables, uploaded
            ensure_env_tables(hs);
            L.env.cpdf = upload(hs.cpdf.data(), hs.cpdf.size());
            L.env.ccdf = upload(hs.ccdf.data(), hs.ccdf.size());
            L.env.mpdf = upload(hs.mpdf.data(), hs.mpdf.size());
            L.env.mcdf = upload(hs.mcdf.data(), hs.mcdf.size());
        } else {
            // default: built on the device from the uploaded map (bit-identical, tests/test_gpu_pt.py)
            const size_t cw = (size_t)hs.env_w + 1;
            L.env.cpdf = upload<float>(nullptr, cw * hs.env_h);
            L.env.ccdf = upload<float>(nullptr, cw * hs.env_h);
            L.env.mpdf = upload<float>(nullptr, (size_t)hs.env_h + 1);
            L.env.mcdf = upload<float>(nullptr, (size_t)hs.env_h + 1);
            std::vector<float> sines;
            env_row_sines(hs.env_h, sines);
            float* d_sin = upload(sines.data(), sines.size());
            launch_env_tables(L.env.env, d_sin, hs.env_w, hs.env_h, const_cast<float*>(L.env.cpdf), const_cast<float*>(L.env.ccdf),
                              const_cast<float*>(L.env.mpdf), const_cast<float*>(L.env.mcdf), nullptr);
            HM_CUDA(cudaDeviceSynchronize());
        }
        L.env.ccoarse = nullptr;
        if (cdf_direct_ok(hs.env_w) && !getenv("HM_ENV_ONE_LEVEL")) {
            float* coarse = upload<float>(nullptr, (size_t)(hs.env_w >> 6) * hs.env_h);
            launch_env_coarse(L.env.ccdf, hs.env_w, hs.env_h, coarse, nullptr);
            HM_CUDA(cudaDeviceSynchronize());
            L.env.ccoarse = coarse;
        }
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
    row0_ = band_partition(W_, H_, 0, rank, world).row0;
    row1_ = band_partition(W_, H_, 0, rank, world).row1;
    int prio_lo = 0, prio_hi = 0;
    HM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    HM_CUDA(cudaStreamCreateWithPriority(&main_stream_, cudaStreamNonBlocking, prio_lo));
    // HM_MAIN_STREAMS > 1 spreads the main pieces of consecutive frames over several streams.  Measured on
    // B200 (profiles/r1k_sweep_main_streams.txt): 1 -> 194-201, 2 -> 178, 3 -> 151 Mpaths/s — concurrent
    // persistent traversal grids fight over the SMs and the L2, so the default stays 1.
    if (const char* e = getenv("HM_MAIN_STREAMS")) n_work_ = std::max(1, std::min((int)kMaxWorkStreams, atoi(e)));
    for (int i = 0; i < n_work_ && n_work_ > 1; ++i) HM_CUDA(cudaStreamCreateWithPriority(&work_streams_[i], cudaStreamNonBlocking, prio_lo));
    HM_CUDA(cudaStreamCreateWithPriority(&order_stream_, cudaStreamNonBlocking, getenv("HM_ORDER_LOW_PRIO") ? prio_lo : prio_hi));
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
    d_trav_ = (unsigned long long*)alloc(16 * 8);
    if (kind_ == HM_KIND_MSNN) {
        in_ch_ = 12;
        records_ = 128 * 128;   // numTrainRecordsX * numTrainRecordsY (headers/render_hair_msnn.h:127-130)
        // everyNth = std::ceil(frameSize / numTrainRecords) with INTEGER division (render_hair_msnn.cu:106)
        every_nth_ = (int)(n / (size_t)records_);
        if (every_nth_ < 1)
            throw std::invalid_argument("render_hair_msnn needs at least 16384 pixels (everyNth would be 0)");
    }
    if (kind_ == HM_KIND_NRC) {
        // RenderWindowNRC::initialize (render_nrc.cu:116-160); std::ceil of INTEGER divisions throughout
        if (world != 1) throw std::invalid_argument("render_nrc shards by samples (hm_renderer_set_frame_schedule), not by row bands");
        in_ch_ = 9;
        const NrcLayout nl = nrc_layout(W_, H_);     // shared with hm_nrc_layout (pure host arithmetic)
        static_assert(kNrcTrainRecords == 65536 && kNrcMaxBounces == 40, "nrc_layout() holds the same constants");
        records_ = nl.records;
        nrc_train_pixels_ = nl.train_pixels;
        every_nth_ = nl.every_nth;
        if (every_nth_ < 1) throw std::invalid_argument("render_nrc needs at least 1638 pixels (everyNth would be 0)");
        if (n % 128) throw std::invalid_argument("render_nrc needs W*H to be a multiple of 128 (scene.cpp:302-306)");
        nn_frame_rows_ = nl.nn_frame_rows;
    }
    if (const char* e = getenv("HM_TAIL_BOUND")) tail_bound_items_ = atoi(e);
    if (const char* e = getenv("HM_FRAMES_IN_FLIGHT")) frames_in_flight_ = std::max(1, std::min((int)kFramesInFlight, atoi(e)));
    if (kind_ == HM_KIND_MSNN) {
        tail_group_ = 4;
        if (const char* e = getenv("HM_TAIL_GROUP")) tail_group_ = std::max(1, std::min((int)kTailGroupMax, atoi(e)));
        if (tail_group_ > 1) {
            // a group being filled + two groups' tails and order-stream work in flight, unless told otherwise
            if (!getenv("HM_FRAMES_IN_FLIGHT")) frames_in_flight_ = std::min((int)kFramesInFlight, 3 * tail_group_);
            frames_in_flight_ = std::max(tail_group_, frames_in_flight_ / tail_group_ * tail_group_);
        }
    }
    // Per-pixel state is allocated for this rank's row band only.  Slots stay FULL-frame pixel indices (RNG keys,
    // training-pixel rule), so each array's base pointer is moved back by the band's first pixel:
    // kernels index it with the global slot and touch [first, first + nb) only.
    const size_t nb = (size_t)(row1_ - row0_) * W_;        // pixels of the band
    const size_t first = (size_t)row0_ * W_;
    if (nb > (size_t)2048 * 2048)
        throw std::invalid_argument("a renderer's band holds more than 2048 x 2048 pixels (scene.cpp:302-306 limit, applied per GPU band): use more bands");
    // the TRAIN_DATA_GEN pass uses slots 0 .. records-1 of the same arrays (band offset undone, params_for): a band
    // smaller than the record count still allocates that many elements
    const size_t n_alloc = kind_ == HM_KIND_MSNN ? std::max(nb, (size_t)(128 * 128)) : nb;
    auto alloc_px = [&](size_t bytes_per_px) { return (char*)alloc(n_alloc * bytes_per_px) - first * bytes_per_px; };
    // Path state the tail kernels touch: the contexts of one tail group are slices, n_alloc elements apart, of one
    // block per array, so a merged tail addresses frame k's slot s as element s + k * n_alloc of frame 0's arrays.
    std::vector<char*> group_blocks;
    size_t group_next = 0;
    auto alloc_state = [&](int ci, size_t bytes_per_px) {
        const int K = tail_group_;
        if (K == 1) return alloc_px(bytes_per_px);
        if (ci % K == 0) group_blocks.push_back((char*)alloc((size_t)K * n_alloc * bytes_per_px));
        char* base = group_blocks[group_next++];
        return base + (size_t)(ci % K) * n_alloc * bytes_per_px - first * bytes_per_px;
    };
    for (int ci = 0; ci < frames_in_flight_; ++ci) {
        FrameCtx& c = ctx_[ci];
        if (ci % tail_group_ == 0) group_blocks.clear();
        group_next = 0;
        c.paths.rng = (uint32_t*)alloc_state(ci, 4);
        c.paths.ray_o = (float4*)alloc_state(ci, 16);
        c.paths.ray_d = (float4*)alloc_state(ci, 16);
        c.paths.hit = (float4*)alloc_state(ci, 16);
        c.paths.beta = (float4*)alloc_state(ci, 16);
        c.paths.color = (float4*)alloc_state(ci, 16);
        c.paths.dl_beta = (float4*)alloc_state(ci, 16);
        c.paths.dl_light = (float4*)alloc_state(ci, 16);
        c.paths.dl_bsdf = (float4*)alloc_state(ci, 16);
        c.paths.vis = (uint32_t*)alloc_state(ci, 4);
        if (kind_ == HM_KIND_MSNN) {
            c.paths.beta_short = (float4*)alloc_state(ci, 16);
            c.paths.color_short = (float4*)alloc_state(ci, 16);
            c.paths.dl_beta_short = (float4*)alloc_state(ci, 16);
            if (tail_group_ > 1 && ci % tail_group_ == 0) {
                Queues& gq = group_q_[ci / tail_group_];
                const size_t cap = (size_t)tail_group_ * records_;   // a frame has at most `records` training paths
                gq.shade[0] = (int*)alloc(cap * 4);
                gq.shade[1] = (int*)alloc(cap * 4);
                gq.extend = (int*)alloc(cap * 4);
                gq.shadow = (float4*)alloc(cap * 2 * 32);
                gq.counts = (int*)alloc(16 * 4);
                gq.trav = d_trav_;
            }
            c.train_idxs = (int*)alloc((size_t)records_ * 4);
            c.nn_frame_in = (float*)alloc_px((size_t)in_ch_ * 4);
            c.nn_train_in = (float*)alloc((size_t)records_ * in_ch_ * 4);
            c.nn_train_out = (float*)alloc((size_t)records_ * 3 * 4);
            c.gbuffer = (float4*)alloc_px(16);
            // one flag per 128-pixel tile of the FULL frame (tiny): bands need not start on a tile boundary
            c.query_tiles = (int*)alloc((n / 128 + 1) * 4);
        }
        if (kind_ == HM_KIND_NRC) {
            c.paths.nrc_state = (float4*)alloc(n * 16);
            c.paths.nrc_prev = (float4*)alloc(n * 16);
            c.train_idxs = (int*)alloc((size_t)nrc_train_pixels_ * 4);
            c.nn_frame_in = (float*)alloc((size_t)nn_frame_rows_ * in_ch_ * 4);
            c.nn_train_in = (float*)alloc((size_t)records_ * in_ch_ * 4);
            c.nn_train_out = (float*)alloc((size_t)records_ * 3 * 4);
            c.gbuffer = (float4*)alloc(n * 16);
            c.gbuffer_b = (float4*)alloc(n * 16);
            c.tbuffer = (NrcTrainRec*)alloc((size_t)nrc_train_pixels_ * sizeof(NrcTrainRec));
        }
        c.q.shade[0] = (int*)alloc(n_alloc * 4);
        c.q.shade[1] = (int*)alloc(n_alloc * 4);
        c.q.extend = (int*)alloc(n_alloc * 4);
        c.q.shadow = (float4*)alloc(n_alloc * 2 * 32);
        c.q.counts = (int*)alloc(16 * 4);
        c.q.trav = d_trav_;
        HM_CUDA(cudaStreamCreateWithPriority(&c.tail_stream, cudaStreamNonBlocking, getenv("HM_TAIL_LOW_PRIO") ? prio_lo : prio_hi));
        HM_CUDA(cudaEventCreateWithFlags(&c.ev_main_done, cudaEventDisableTiming));
        HM_CUDA(cudaEventCreateWithFlags(&c.ev_traced, cudaEventDisableTiming));
        HM_CUDA(cudaEventCreateWithFlags(&c.ev_free, cudaEventDisableTiming));
        HM_CUDA(cudaEventCreateWithFlags(&c.ev_shuffled, cudaEventDisableTiming));
    }
    for (int i = 0; i < 6; ++i) bufs_[i] = (float4*)alloc(n * 16);
    fb_ = (uint32_t*)alloc(n * 4);

    if (kind_ == HM_KIND_MSNN || kind_ == HM_KIND_NRC) {
        MlpConfig cfg = hs.tcnn_config.empty() ? MlpConfig() : mlp_config_from_json(hs.tcnn_config, in_ch_, 3);
        cfg.in_ch = in_ch_; cfg.out_ch = 3;
        mlp_.reset(new Mlp(cfg, order_stream_));
        n_idxs_ = kind_ == HM_KIND_NRC ? nrc_train_pixels_ : records_;
        if (kind_ == HM_KIND_MSNN) nn_frame_rows_ = (int)n;
        std::vector<int> seq(n_idxs_);
        for (int i = 0; i < n_idxs_; ++i) seq[i] = i;   // thrust::sequence
        d_train_idxs_ = (int*)alloc((size_t)n_idxs_ * 4);
        HM_CUDA(cudaMemcpy(d_train_idxs_, seq.data(), (size_t)n_idxs_ * 4, cudaMemcpyHostToDevice));
        nn_frame_out_ = kind_ == HM_KIND_MSNN ? (float*)alloc_px(3 * 4) : (float*)alloc((size_t)nn_frame_rows_ * 3 * 4);
    }
    last_ctx_ = &ctx_[0];
    HM_CUDA(cudaDeviceSynchronize());
}

Renderer::~Renderer() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    mlp_.reset();
    scratch_.release();
    for (void* p : allocs_) cudaFree(p);
    scene_.reset();
    for (auto& p : pending_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : event_pool_) cudaEventDestroy(e);
    for (FrameCtx& c : ctx_) {
        if (c.tail_stream) cudaStreamDestroy(c.tail_stream);
        if (c.ev_main_done) cudaEventDestroy(c.ev_main_done);
        if (c.ev_traced) cudaEventDestroy(c.ev_traced);
        if (c.ev_free) cudaEventDestroy(c.ev_free);
        if (c.ev_shuffled) cudaEventDestroy(c.ev_shuffled);
    }
    if (main_stream_) cudaStreamDestroy(main_stream_);
    for (auto w : work_streams_) if (w) cudaStreamDestroy(w);
    if (order_stream_) cudaStreamDestroy(order_stream_);
}

void Renderer::set_hair_params(const float sigma_a[3], float beta_m, float beta_n, float alpha_rad, const float gains[4]) {
    sync();
    SceneView& v = scene_->view;
    v.lobes.setup(beta_m, beta_n, alpha_rad);
    v.lobes.sigma_a = V3(sigma_a[0], sigma_a[1], sigma_a[2]);
    for (int i = 0; i < 4; ++i) v.lobes.gain[i] = gains[i];
    accum_id_ = 0;
}

void Renderer::set_environment(float scale, float rotation) {
    sync();
    scene_->view.lights.env.scale = scale;
    scene_->view.lights.env.rot_phi = rotation;
    accum_id_ = 0;
}

void Renderer::set_sampling(bool mis, bool env_pdf) {
    sync();
    scene_->view.mis = mis ? 1 : 0;
    scene_->view.lights.env.pdf_sampling = env_pdf ? 1 : 0;
    accum_id_ = 0;
}

void Renderer::sync() {
    HM_CUDA(cudaSetDevice(device_));
    flush_deferred();
    HM_CUDA(cudaStreamSynchronize(main_stream_));
    for (int i = 0; i < n_work_; ++i) if (work_streams_[i]) HM_CUDA(cudaStreamSynchronize(work_streams_[i]));
    for (FrameCtx& c : ctx_) if (c.tail_stream) HM_CUDA(cudaStreamSynchronize(c.tail_stream));
    HM_CUDA(cudaStreamSynchronize(order_stream_));
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
void Renderer::timed(int stage, cudaStream_t s, F&& f) {
    stats_.launches[stage]++;
    if (!profiling_ || !((profile_mask_ >> stage) & 1u) || (frames_issued_ % (uint64_t)profile_period_) != 0) { f(); return; }
    stats_.timed_launches[stage]++;
    Pending p{stage, take_event(), take_event()};
    HM_CUDA(cudaEventRecord(p.a, s));
    f();
    HM_CUDA(cudaEventRecord(p.b, s));
    pending_.push_back(p);
}

void Renderer::resolve_events() {
    if (pending_.empty()) return;
    sync();
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
    sync();
    stats_ = Stats();
    HM_CUDA(cudaMemset(d_trav_, 0, 16 * 8));
}

FrameParams Renderer::params_for(const FrameCtx& c) {
    FrameParams P;
    memset(&P, 0, sizeof(P));
    P.scene = scene_->view;
    P.cam = cam_;
    P.paths = c.paths;
    P.q = c.q;
    P.W = W_; P.H = H_;
    P.row0 = row0_; P.row1 = row1_;
    P.accum_id = c.accum_id;
    P.frame_id = c.frame_id;
    P.collect_stats = collect_stats_ ? 1 : 0;
    P.v1_stop = hs_.path_v1 - 1;
    P.v2_stop = hs_.path_v2 - 1;
    P.accum = bufs_[1]; P.average = bufs_[0]; P.fb = fb_;
    P.in_ch = in_ch_;
    P.n_primary = (row1_ - row0_) * W_;
    if (c.pretrain) {
        // TRAIN_DATA_GEN: slot == record index 0..records-1, whatever the band: undo the band offset of the
        // per-pixel arrays (the constructor moved their base pointers back by the band's first pixel)
        const ptrdiff_t first = (ptrdiff_t)row0_ * W_;
        if (first) {
            PathBuffers& b = P.paths;
            b.rng += first; b.ray_o += first; b.ray_d += first; b.hit += first; b.beta += first; b.color += first;
            b.dl_beta += first; b.dl_light += first; b.dl_bsdf += first; b.vis += first;
            if (b.beta_short) { b.beta_short += first; b.color_short += first; b.dl_beta_short += first; }
        }
        P.pretrain = 1;
        P.n_primary = records_;
        P.sampled_points = d_scene_points_;
        P.scene_indices = d_scene_indices_;
    }
    if (kind_ == HM_KIND_MSNN) {
        P.mode = MODE_MSNN;
        P.msnn_beta = beta_;
        P.every_nth = every_nth_;
        P.train_idxs = c.train_idxs;
        const BandPartition bp = band_partition(W_, H_, records_, rank_, world_);
        P.train_slot0 = c.pretrain ? 0 : bp.slot0;
        P.train_slots = c.pretrain ? records_ : bp.slots;
        P.train_records = records_;
        const size_t unshift = c.pretrain ? (size_t)row0_ * W_ : 0;
        P.nn_frame_in = c.nn_frame_in + unshift * in_ch_;
        P.nn_train_in = c.nn_train_in;
        P.nn_train_out = c.nn_train_out;
        P.gbuffer = c.gbuffer + unshift;
        P.query_tiles = c.pretrain ? nullptr : c.query_tiles;
    } else if (kind_ == HM_KIND_NRC) {
        P.mode = MODE_NRC;
        P.every_nth = every_nth_;
        P.train_idxs = c.train_idxs;
        P.nn_frame_in = c.nn_frame_in;
        P.gbuffer = c.gbuffer;
        P.gbuffer_b = c.gbuffer_b;
        P.tbuffer = c.tbuffer;
        P.nrc_train_pixels = nrc_train_pixels_;
        P.nrc_all_unbiased = nrc_all_unbiased_ ? 1 : 0;
        P.nrc_c = nrc_c_;
    } else {
        P.mode = MODE_PT;
    }
    return P;
}

void Renderer::shuffle_train_idxs(FrameCtx& c) {
    // thrust::shuffle(trainIdxs, default_random_engine()) — a freshly constructed engine
    // every frame, so the permutation applied is the same each time and the sequence of
    // compositions is deterministic (render_hair_msnn.cu:711-714).  The frame keeps a copy:
    // later frames re-shuffle the persistent array while this one is still in flight.
    thrust::device_ptr<int> p = thrust::device_pointer_cast(d_train_idxs_);
    // on main_stream_ (frame order); the frame's own stream picks the copy up through ev_shuffled
    if (c.main != main_stream_) HM_CUDA(cudaStreamWaitEvent(main_stream_, c.ev_free, 0));
    thrust::shuffle(thrust::cuda::par_nosync(scratch_).on(main_stream_), p, p + n_idxs_, thrust::default_random_engine());
    HM_CUDA(cudaMemcpyAsync(c.train_idxs, d_train_idxs_, (size_t)n_idxs_ * 4, cudaMemcpyDeviceToDevice, main_stream_));
    if (c.main != main_stream_) {
        HM_CUDA(cudaEventRecord(c.ev_shuffled, main_stream_));
        HM_CUDA(cudaStreamWaitEvent(c.main, c.ev_shuffled, 0));
    }
}

FrameCtx& Renderer::begin_frame(bool pretrain) {
    HM_CUDA(cudaSetDevice(device_));
    FrameCtx& c = ctx_[frames_issued_ % frames_in_flight_];
    c.main = (pretrain || n_work_ <= 1) ? main_stream_ : work_streams_[frames_issued_ % n_work_];
    frames_issued_++;
    c.accum_id = accum_id_;
    c.frame_id = frame_offset_ + accum_id_ * frame_stride_;
    // the context is reusable once the frame that last used it has been composited
    HM_CUDA(cudaStreamWaitEvent(c.main, c.ev_free, 0));
    c.pretrain = pretrain;
    return c;
}

// Wavefront loop: primary, then (shade -> shadow + extend) per path vertex.  The first
// `main_vertices` vertices run on the main stream (trace_main), the rest on a tail stream (trace_tail).
// No host synchronisation: every stage reads its queue length from device memory, and the
// tail always issues the full vertex budget (empty launches cost a few microseconds).
void Renderer::trace_main(FrameCtx& c) {
    if ((kind_ == HM_KIND_MSNN || kind_ == HM_KIND_NRC) && !c.pretrain) shuffle_train_idxs(c);
    FrameParams P = params_for(c);
    int max_vertices = P.v2_stop + 1;   // the primary hit plus up to v2_stop bounces
    if (kind_ == HM_KIND_NRC) {
        max_vertices = kNrcMaxBounces;  // vertices 0..39; path_v1/path_v2 do not apply (cuda/nrc.cu:169)
        // what render() clears after every frame (render_nrc.cu:683-689: owlBufferClear x4 + RESET pass)
        HM_CUDA(cudaMemsetAsync(c.nn_frame_in, 0, (size_t)nn_frame_rows_ * in_ch_ * 4, c.main));
        HM_CUDA(cudaMemsetAsync(c.nn_train_in, 0, (size_t)records_ * in_ch_ * 4, c.main));
        HM_CUDA(cudaMemsetAsync(c.nn_train_out, 0, (size_t)records_ * 3 * 4, c.main));
        HM_CUDA(cudaMemsetAsync(c.tbuffer, 0, (size_t)nrc_train_pixels_ * sizeof(NrcTrainRec), c.main));
    }
    if (max_vertices < 1) max_vertices = 1;
    int main_vertices = kind_ == HM_KIND_MSNN ? beta_ + 2 : 6;
    if (main_vertices < 2) main_vertices = 2;
    if (main_vertices > max_vertices) main_vertices = max_vertices;

    cudaStream_t s = c.main;
    HM_CUDA(cudaMemsetAsync(c.q.counts, 0, 16 * 4, s));
    if (c.query_tiles && !c.pretrain) HM_CUDA(cudaMemsetAsync(c.query_tiles, 0, ((size_t)W_ * H_ / 128 + 1) * 4, s));
    timed(0, s, [&] { launch_primary(P, s); });
    int src = 0;
    for (int vertex = 0; vertex < main_vertices; ++vertex) {
        timed(1, s, [&] { launch_shade(P, src, s, 0); });
        const int dst = src ^ 1;
        HM_CUDA(cudaMemsetAsync(c.q.counts + dst, 0, 4, s));
        timed(2, s, [&] { launch_trace(P, dst, s, 0); });
        HM_CUDA(cudaMemsetAsync(c.q.counts + 2, 0, 16, s));   // extend + shadow counters and their work cursors
        src = dst;
    }
    c.mp_src = src; c.mp_vertex = main_vertices; c.mp_max = max_vertices;
    if (main_vertices < max_vertices) HM_CUDA(cudaEventRecord(c.ev_main_done, c.main));
}

// Tail piece of m frames whose main pieces stopped at the same vertex (m > 1: HairMSNN frames of one tail group,
// consecutive contexts): the remaining vertices, then each frame's finalize pass; the order stream waits for it.
void Renderer::trace_tail(FrameCtx* const* cs, int m) {
    FrameCtx& c0 = *cs[0];
    cudaStream_t s = c0.main;
    FrameParams P = params_for(c0);
    int src = c0.mp_src;
    const bool has_tail = c0.mp_vertex < c0.mp_max;
    if (has_tail) {
        // a merged tail runs on its group's stream: the group's queues have one user at a time
        s = m > 1 ? ctx_[(int)(cs[0] - ctx_) / tail_group_ * tail_group_].tail_stream : c0.tail_stream;
        for (int k = 0; k < m; ++k) HM_CUDA(cudaStreamWaitEvent(s, cs[k]->ev_main_done, 0));
        P.tail = 1;
        // HairMSNN: past the main piece only training paths are alive (16384 per frame at most; all records in a pre-training pass)
        long long tail_bound = 0;
        if (kind_ == HM_KIND_MSNN) tail_bound = (tail_bound_items_ > 0 ? tail_bound_items_ : records_) * (long long)m;
        if (m > 1) {
            const int gi = (int)(cs[0] - ctx_) / tail_group_;
            const size_t nb = (size_t)(row1_ - row0_) * W_;
            TailMerge M;
            memset(&M, 0, sizeof(M));
            for (int k = 0; k < m; ++k) { M.counts[k] = cs[k]->q.counts; M.queue[k] = cs[k]->q.shade[src]; }
            M.n = m; M.src = src; M.stride = (int)std::max(nb, (size_t)(128 * 128));
            M.cap = tail_group_ * records_;
            M.out = group_q_[gi].shade[src]; M.out_counts = group_q_[gi].counts;
            HM_CUDA(cudaMemsetAsync(group_q_[gi].counts, 0, 16 * 4, s));
            timed(3, s, [&] { launch_merge_tail(M, records_, s); });
            P.q = group_q_[gi];
            P.tail_merged = 1;
        }
        for (int vertex = c0.mp_vertex; vertex < c0.mp_max; ++vertex) {
            timed(3, s, [&] { launch_shade(P, src, s, tail_bound); });
            const int dst = src ^ 1;
            HM_CUDA(cudaMemsetAsync(P.q.counts + dst, 0, 4, s));
            timed(3, s, [&] { launch_trace(P, dst, s, tail_bound); });
            HM_CUDA(cudaMemsetAsync(P.q.counts + 2, 0, 16, s));
            src = dst;
        }
    }
    for (int k = 0; k < m; ++k) {
        FrameCtx& c = *cs[k];
        if (!has_tail) s = c.main;
        if (kind_ != HM_KIND_PT) {
            FrameParams Pk = params_for(c);
            timed(4, s, [&] { launch_finalize(Pk, s); });   // frame-local outputs only
        }
        HM_CUDA(cudaEventRecord(c.ev_traced, s));
        HM_CUDA(cudaStreamWaitEvent(order_stream_, c.ev_traced, 0));
        last_ctx_ = &c;
    }
}

void Renderer::trace_frame(FrameCtx& c) {
    trace_main(c);
    FrameCtx* cs[1] = {&c};
    trace_tail(cs, 1);
}

void Renderer::finish_pt(FrameCtx& c) {
    FrameParams P = params_for(c);
    timed(4, order_stream_, [&] { launch_finalize(P, order_stream_); });   // accumulates: frame order
}

void Renderer::end_frame(FrameCtx& c) {
    HM_CUDA(cudaEventRecord(c.ev_free, order_stream_));
    // deferred frames (render_frames with tail groups) advanced the counters when their main piece was enqueued
    if (c.counted) { c.counted = false; return; }
    accum_id_++;
    stats_.frames++;
}

void Renderer::msnn_trace() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    if (current_) throw std::logic_error("msnn_trace: the previous frame was not finished (call msnn_finish)");
    flush_deferred();
    FrameCtx& c = begin_frame();
    trace_frame(c);
    current_ = &c;
}

void Renderer::msnn_train_backward() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    if (!current_) throw std::logic_error("msnn_train_backward: call msnn_trace first");
    const BandPartition bp = band_partition(W_, H_, records_, rank_, world_);
    const int s0 = bp.slot0, n = bp.train_n;
    FrameCtx& c = *current_;
    timed(5, order_stream_, [&] {
        mlp_->forward_backward(c.nn_train_in + (size_t)s0 * in_ch_, c.nn_train_out + (size_t)s0 * 3, n,
                               (world_ == 1 ? n : records_) * groups_);
    });
}

// One NCCL all-reduce of the fp32 gradient buffer (1 000 448 floats, 4 MB) over NVLink, in frame order on the
// order stream: backward -> all-reduce -> Adam.  Every rank applies the same sum, so the replicas' weights
// stay bit-identical.
void Renderer::all_reduce_gradients() {
    if (!comm_ || comm_->world() == 1) return;
    timed(5, order_stream_, [&] { comm_->all_reduce_sum(mlp_->gradients(), mlp_->n_params(), order_stream_); });
}

void Renderer::set_comm(Comm* comm) {
    sync();
    if (!comm) { comm_ = nullptr; groups_ = 1; frame_offset_ = 0; frame_stride_ = 1; return; }
    if (comm->device() != device_) throw std::invalid_argument("communicator and renderer are on different devices");
    if (comm->world() % world_ != 0 || comm->rank() % world_ != rank_)
        throw std::invalid_argument("communicator rank must be group * bands + band (comm world a multiple of the band count)");
    comm_ = comm;
    groups_ = comm->world() / world_;
    frame_offset_ = comm->rank() / world_;
    frame_stride_ = groups_;
}

void Renderer::reduce_framebuffers() {
    if (!comm_) throw std::logic_error("reduce_framebuffers: no communicator attached (hm_renderer_set_comm)");
    if (current_) throw std::logic_error("reduce_framebuffers: a frame is in flight");
    HM_CUDA(cudaSetDevice(device_));
    flush_deferred();
    const size_t n = (size_t)W_ * H_;
    const double total = comm_->all_reduce_sum((double)accum_id_) / world_;   // samples per pixel over all groups
    if (total <= 0) throw std::logic_error("reduce_framebuffers: nothing rendered yet");
    const int pairs = kind_ == HM_KIND_MSNN ? 3 : 1;
    for (int i = 0; i < pairs; ++i) {
        float4* avg = bufs_[2 * i];
        float4* accum = bufs_[2 * i + 1];
        HM_CUDA(cudaMemcpyAsync(avg, accum, n * 16, cudaMemcpyDeviceToDevice, order_stream_));
        comm_->all_reduce_sum((float*)avg, n * 4, order_stream_);
        launch_resolve_sum(avg, i == 0 ? fb_ : nullptr, 1.f / (float)total, (int)n, order_stream_);
    }
    HM_CUDA(cudaStreamSynchronize(order_stream_));
}

void Renderer::msnn_train_apply() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    all_reduce_gradients();
    timed(5, order_stream_, [&] { mlp_->optimizer_step(); });
}

// Everything of a traced HairMSNN frame that depends on the frames before it, in frame order on the order stream.
void Renderer::msnn_order_work(FrameCtx& c) {
    current_ = &c;
    if (hs_.tcnn_train) {
        msnn_train_backward();
        msnn_train_apply();
    }
    msnn_finish();
}

// Enqueue what render_frames() held back: the merged tail of the deferred frames, then — in the order the calls were
// made — each frame's order-stream work and the read-backs queued behind it.
void Renderer::flush_deferred() {
    if (deferred_.empty()) return;
    std::vector<Deferred> ops;
    ops.swap(deferred_);
    deferred_frames_ = 0;
    FrameCtx* cs[kTailGroupMax];
    int m = 0;
    for (const Deferred& d : ops) if (d.kind == 0) cs[m++] = d.c;
    if (m > 0) trace_tail(cs, m);
    for (const Deferred& d : ops) {
        if (d.kind == 0) {
            // frame span on the order stream (stage 8), as render_frames() records it for frames that are not held back
            Pending whole{8, nullptr, nullptr};
            if (profiling_ && ((profile_mask_ >> 8) & 1u)) {
                whole.a = take_event(); whole.b = take_event();
                HM_CUDA(cudaEventRecord(whole.a, order_stream_));
            }
            msnn_order_work(*d.c);
            if (whole.a) {
                HM_CUDA(cudaEventRecord(whole.b, order_stream_));
                pending_.push_back(whole);
            }
        } else {
            HM_CUDA(cudaMemcpyAsync(d.dst, d.src, d.bytes, cudaMemcpyDeviceToHost, order_stream_));
        }
    }
}

void Renderer::msnn_finish() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    if (!current_) throw std::logic_error("msnn_finish: call msnn_trace first");
    FrameCtx& c = *current_;
    const size_t first = (size_t)row0_ * W_;
    const int count = (row1_ - row0_) * W_;
    // rows whose output nobody reads (background, head) are skipped tile-wise; results are identical
    const int* mask = (skip_unused_queries_ && first % 128 == 0) ? c.query_tiles + first / 128 : nullptr;
    timed(6, order_stream_, [&] { mlp_->inference(c.nn_frame_in + first * in_ch_, nn_frame_out_ + first * 3, count, mask); });
    MsnnComposite C;
    C.final_avg = bufs_[0]; C.final_accum = bufs_[1];
    C.pt_avg = bufs_[2]; C.pt_accum = bufs_[3];
    C.nn_avg = bufs_[4]; C.nn_accum = bufs_[5];
    C.fb = fb_; C.gbuffer = c.gbuffer; C.nn_out = nn_frame_out_;
    C.accum_id = c.accum_id;
    C.first = (int)first; C.count = count;
    timed(7, order_stream_, [&] { launch_msnn_composite(C, order_stream_); });
    end_frame(c);
    current_ = nullptr;
}

// ---- render_nrc (render_nrc.cu:640-700) ------------------------------------------------
void Renderer::nrc_trace() {
    if (kind_ != HM_KIND_NRC) throw std::logic_error("not an NRC renderer");
    if (current_) throw std::logic_error("nrc_trace: the previous frame was not finished (call nrc_end)");
    FrameCtx& c = begin_frame();
    trace_frame(c);
    current_ = &c;
}

void Renderer::nrc_query() {
    if (kind_ != HM_KIND_NRC) throw std::logic_error("not an NRC renderer");
    if (!current_) throw std::logic_error("nrc_query: call nrc_trace first");
    FrameCtx& c = *current_;
    timed(6, order_stream_, [&] { mlp_->inference(c.nn_frame_in, nn_frame_out_, nn_frame_rows_); });
    NrcRender R;
    R.accum = bufs_[1]; R.average = bufs_[0]; R.fb = fb_;
    R.gbuffer = c.gbuffer; R.gbuffer_b = c.gbuffer_b; R.tbuffer = c.tbuffer;
    R.train_idxs = c.train_idxs;
    R.nn_out = nn_frame_out_;
    R.train_in = c.nn_train_in; R.train_gt = c.nn_train_out;
    R.W = W_; R.H = H_; R.in_ch = in_ch_; R.every_nth = every_nth_;
    R.train_pixels = nrc_train_pixels_; R.all_unbiased = nrc_all_unbiased_ ? 1 : 0;
    R.accum_id = c.accum_id;
    timed(7, order_stream_, [&] { launch_nrc_render(R, order_stream_); });
}

void Renderer::nrc_train_backward() {
    if (kind_ != HM_KIND_NRC) throw std::logic_error("not an NRC renderer");
    if (!current_) throw std::logic_error("nrc_train_backward: call nrc_trace first");
    FrameCtx& c = *current_;
    timed(5, order_stream_, [&] { mlp_->forward_backward(c.nn_train_in, c.nn_train_out, records_, records_ * groups_); });
}

void Renderer::nrc_train_apply() {
    if (kind_ != HM_KIND_NRC) throw std::logic_error("not an NRC renderer");
    all_reduce_gradients();
    timed(5, order_stream_, [&] { mlp_->optimizer_step(); });
}

void Renderer::nrc_end() {
    if (kind_ != HM_KIND_NRC) throw std::logic_error("not an NRC renderer");
    if (!current_) throw std::logic_error("nrc_end: call nrc_trace first");
    end_frame(*current_);
    current_ = nullptr;
}

void Renderer::ensure_scene_samples() {
    if (d_scene_points_) return;
    std::vector<float> pts;
    build_scene_samples(hs_.geo, 1000000, 1337u, pts);      // numSamples = 1e6 (headers/render_hair_msnn.h:127-130)
    n_scene_samples_ = (int)(pts.size() / 3);
    if (n_scene_samples_ < records_) throw std::invalid_argument("TRAIN_DATA_GEN needs at least 16384 strand samples");
    HM_CUDA(cudaMalloc((void**)&d_scene_points_, pts.size() * 4));
    allocs_.push_back(d_scene_points_);
    HM_CUDA(cudaMemcpy(d_scene_points_, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice));
    std::vector<int> seq(n_scene_samples_);
    for (int i = 0; i < n_scene_samples_; ++i) seq[i] = i;   // thrust::sequence
    HM_CUDA(cudaMalloc((void**)&d_scene_indices_, (size_t)n_scene_samples_ * 4));
    allocs_.push_back(d_scene_indices_);
    HM_CUDA(cudaMemcpy(d_scene_indices_, seq.data(), (size_t)n_scene_samples_ * 4, cudaMemcpyHostToDevice));
    // the initial shuffle of sceneIndices (render_hair_msnn.cu:395-400)
    thrust::device_ptr<int> p = thrust::device_pointer_cast(d_scene_indices_);
    thrust::shuffle(thrust::cuda::par_nosync(scratch_).on(main_stream_), p, p + n_scene_samples_, thrust::default_random_engine());
}

void Renderer::msnn_train_data_gen() {
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    if (current_) throw std::logic_error("msnn_train_data_gen: a frame is in flight");
    HM_CUDA(cudaSetDevice(device_));
    sync();
    ensure_scene_samples();
    FrameCtx& c = begin_frame(true);
    trace_frame(c);
    end_frame(c);
}

void Renderer::msnn_pretrain(int steps) {
    // The reference's initial training (render_hair_msnn.cu:633-641): genTrainingData() = the
    // TRAIN_DATA_GEN pass — 128 x 128 full-length training paths whose first ray runs from the camera
    // position to a random point on the strands — followed by train() = shuffle(sceneIndices) +
    // training_step, repeated for one WALL-CLOCK second, before cameraChanged() has set the launch
    // parameters' camera (the rays start wherever the uninitialised camera.pos points).
    // Here: the scene file's camera position, a fixed number of steps, and strand samples from a fixed
    // seed instead of the wall clock.  accumId advances per pass and is reset afterwards (cameraChanged).
    if (kind_ != HM_KIND_MSNN) throw std::logic_error("not a HairMSNN renderer");
    if (current_) throw std::logic_error("msnn_pretrain: a frame is in flight");
    HM_CUDA(cudaSetDevice(device_));
    sync();
    ensure_scene_samples();
    for (int i = 0; i < steps; ++i) {
        FrameCtx& c = begin_frame(true);
        trace_frame(c);
        current_ = &c;
        thrust::device_ptr<int> p = thrust::device_pointer_cast(d_scene_indices_);
        thrust::shuffle(thrust::cuda::par_nosync(scratch_).on(main_stream_), p, p + n_scene_samples_, thrust::default_random_engine());
        // with a communicator every rank contributes a full batch (groups render different sample ids, the bands
        // of a group the same one): global batch = records x comm world
        timed(5, order_stream_, [&] { mlp_->forward_backward(c.nn_train_in, c.nn_train_out, records_, records_ * (comm_ ? comm_->world() : 1)); });
        all_reduce_gradients();
        timed(5, order_stream_, [&] { mlp_->optimizer_step(); });
        end_frame(c);
        current_ = nullptr;
    }
    sync();
    stats_.frames -= steps;
    accum_id_ = 0;
}

void Renderer::render_frames(int n) {
    HM_CUDA(cudaSetDevice(device_));
    for (int i = 0; i < n; ++i) {
        Pending whole{8, nullptr, nullptr};
        const bool held_back = kind_ == HM_KIND_MSNN && tail_group_ > 1;   // flush_deferred() records the span
        if (!held_back && profiling_ && ((profile_mask_ >> 8) & 1u)) {
            // frame span on the order stream: previous frame's completion -> this frame's completion
            whole.a = take_event(); whole.b = take_event();
            HM_CUDA(cudaEventRecord(whole.a, order_stream_));
        }
        if (kind_ == HM_KIND_PT) {
            FrameCtx& c = begin_frame();
            trace_frame(c);
            finish_pt(c);
            end_frame(c);
        } else if (kind_ == HM_KIND_NRC) {
            nrc_trace();
            nrc_query();
            if (hs_.tcnn_train) {
                nrc_train_backward();
                nrc_train_apply();
            }
            nrc_end();
        } else if (tail_group_ > 1) {
            if (current_) throw std::logic_error("render_frames: a split frame is in flight (call msnn_finish)");
            // main piece now; tail, training step, inference and composite when the tail group is complete
            FrameCtx& c = begin_frame();
            trace_main(c);
            c.counted = true;
            accum_id_++;
            stats_.frames++;
            deferred_.push_back(Deferred{0, &c, nullptr, nullptr, 0});
            deferred_frames_++;
            const int ci = (int)(&c - ctx_);
            if ((ci + 1) % tail_group_ == 0) flush_deferred();
        } else {
            msnn_trace();
            if (hs_.tcnn_train) {
                msnn_train_backward();
                msnn_train_apply();
            }
            msnn_finish();
        }
        if (whole.a) {
            HM_CUDA(cudaEventRecord(whole.b, order_stream_));
            pending_.push_back(whole);
        }
    }
}

Stats Renderer::stats() {
    resolve_events();
    sync();
    if (mlp_) stats_.last_loss = mlp_->loss();
    unsigned long long t[16];
    HM_CUDA(cudaMemcpy(t, d_trav_, sizeof(t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) stats_.trav[i] = t[i];
    stats_.tail_nodes = t[10]; stats_.tail_prims = t[11]; stats_.tail_rays = t[12];
    stats_.rays_extend = t[6]; stats_.rays_shadow = t[7]; stats_.shade_items = t[8]; stats_.rays_primary = t[9];
    return stats_;
}

void* Renderer::device_buffer(int which, size_t* bytes) {
    flush_deferred();   // per-frame buffers are those of the last frame enqueued in full
    return buffer_ptr(which, bytes);
}

void* Renderer::buffer_ptr(int which, size_t* bytes) {
    const size_t n = (size_t)W_ * H_;
    const size_t nb = (size_t)(row1_ - row0_) * W_, first = (size_t)row0_ * W_;
    FrameCtx& c = *last_ctx_;
    switch (which) {
        case 0: case 1: case 2: case 3: case 4: case 5: *bytes = n * 16; return bufs_[which];
        case 6: *bytes = n * 4; return fb_;
        // per-pixel working buffers hold this renderer's row band only: [rows of the band][W] (the whole frame when world = 1)
        case 7: *bytes = (kind_ == HM_KIND_NRC ? (size_t)nn_frame_rows_ : nb) * in_ch_ * 4; return c.nn_frame_in + first * in_ch_;
        case 8: *bytes = (kind_ == HM_KIND_NRC ? (size_t)nn_frame_rows_ : nb) * 3 * 4; return nn_frame_out_ + first * 3;
        case 9: *bytes = (size_t)records_ * in_ch_ * 4; return c.nn_train_in;
        case 10: *bytes = (size_t)records_ * 3 * 4; return c.nn_train_out;
        case 11: *bytes = nb * 16; return c.gbuffer + first;
        case 12: *bytes = (size_t)n_idxs_ * 4; return c.train_idxs;
        case 13: *bytes = n * 16; return c.gbuffer_b;
        case 14: *bytes = (size_t)nrc_train_pixels_ * sizeof(NrcTrainRec); return c.tbuffer;
        case 15: *bytes = (size_t)n_scene_samples_ * 4; return d_scene_indices_;
        case 16: *bytes = (size_t)n_scene_samples_ * 12; return d_scene_points_;
        case 17: *bytes = ((size_t)hs_.env_w + 1) * hs_.env_h * 4; return (void*)scene_->view.lights.env.cpdf;
        case 18: *bytes = ((size_t)hs_.env_w + 1) * hs_.env_h * 4; return (void*)scene_->view.lights.env.ccdf;
        case 19: *bytes = ((size_t)hs_.env_h + 1) * 4; return (void*)scene_->view.lights.env.mpdf;
        case 20: *bytes = ((size_t)hs_.env_h + 1) * 4; return (void*)scene_->view.lights.env.mcdf;
        default: *bytes = 0; return nullptr;
    }
}

// Image buffers do not belong to a frame context: a read-back asked for while frames are held back is queued
// behind them (same stream order as an immediate copy would have had).
static bool is_image_buffer(int which) { return which >= 0 && which <= 6; }

void Renderer::readback_async(int which, void* host_dst, size_t bytes) {
    const bool defer = is_image_buffer(which) && !deferred_.empty();
    size_t have = 0;
    void* p = defer ? buffer_ptr(which, &have) : device_buffer(which, &have);
    if (!p) throw std::logic_error("buffer not available for this renderer kind");
    if (bytes > have) throw std::invalid_argument("requested more bytes than the buffer holds");
    if (defer) { deferred_.push_back(Deferred{1, nullptr, p, host_dst, bytes}); return; }
    HM_CUDA(cudaMemcpyAsync(host_dst, p, bytes, cudaMemcpyDeviceToHost, order_stream_));
}

void Renderer::readback_rows_async(int which, int row0, int rows, void* host_dst) {
    const bool defer = is_image_buffer(which) && !deferred_.empty();
    size_t have = 0;
    char* p = (char*)(defer ? buffer_ptr(which, &have) : device_buffer(which, &have));
    if (!p) throw std::logic_error("buffer not available for this renderer kind");
    const size_t px = which == 6 ? 4 : 16;
    char* src = p + (size_t)row0 * W_ * px;
    const size_t bytes = (size_t)rows * W_ * px;
    if (defer) { deferred_.push_back(Deferred{1, nullptr, src, host_dst, bytes}); return; }
    HM_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, order_stream_));
}

void Renderer::trace_rays_device(const float* d_org, const float* d_dir, int n, int any, float tmin, float tmax,
                                 float* d_out_hit, int* d_out_stats) {
    HM_CUDA(cudaSetDevice(device_));
    launch_trace_rays(scene_->view, d_org, d_dir, n, any, tmin, tmax, (float4*)d_out_hit, d_out_stats, (int*)(d_trav_ + 15), order_stream_);
}

}  // namespace hm
