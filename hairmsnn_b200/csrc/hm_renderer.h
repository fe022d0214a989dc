// hm_renderer.h — host frame drivers: device scene, path-state buffers, per-frame pass
// sequencing for the three renderer kinds.
//
// Stands in for RenderWindowPT / RenderWindowNRC / RenderWindow_HairMSNN
// initialize()/render()/train() (render_path_tracing.cu, render_nrc.cu,
// render_hair_msnn.cu) minus the GLFW/ImGui shell.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hm_comm.h"
#include "hm_host.h"
#include "hm_mlp.h"
#include "hm_wavefront.h"

namespace hm {

struct Stats {
    double ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // primary shade(main piece) trace(main piece) shade+trace(tail piece) finalize train infer composite total
    uint64_t launches[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t timed_launches[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // launches that carried an event pair (ms[] sums over these)
    uint64_t rays_primary = 0, rays_extend = 0, rays_shadow = 0, shade_items = 0;
    uint64_t trav[6] = {0, 0, 0, 0, 0, 0};        // extend nodes/prims, shadow nodes/prims, primary nodes/prims
    uint64_t tail_nodes = 0, tail_prims = 0, tail_rays = 0;   // share of extend+shadow traced by tail-piece launches
    float last_loss = 0.f;
    int frames = 0;
};

// Scratch allocator for thrust: its default allocator calls cudaMalloc/cudaFree per algorithm invocation, and
// cudaFree synchronises the whole device — which would serialise the frames in flight.  Blocks are kept and
// reused.  One instance per Renderer: every use is ordered on that renderer's main stream, on its device.
struct ThrustScratch {
    typedef char value_type;
    struct Block { char* p; size_t n; bool busy; };
    std::vector<Block> blocks;
    char* allocate(std::ptrdiff_t n) {
        for (auto& b : blocks)
            if (!b.busy && b.n >= (size_t)n) { b.busy = true; return b.p; }
        char* p = nullptr;
        if (cudaMalloc((void**)&p, (size_t)n) != cudaSuccess) throw std::runtime_error("CUDA: scratch allocation failed");
        blocks.push_back(Block{p, (size_t)n, true});
        return p;
    }
    void deallocate(char* p, size_t) {
        for (auto& b : blocks)
            if (b.p == p) { b.busy = false; return; }
    }
    void release() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); }   // called by ~Renderer after a device sync
};

class DeviceScene {
public:
    explicit DeviceScene(const HostScene& hs);
    ~DeviceScene();
    SceneView view;
    size_t bytes = 0;
private:
    std::vector<void*> allocs_;
    template <typename T> T* upload(const T* src, size_t n);
};

// Per-frame working set.  Several frames are in flight at once (see Renderer), each with
// its own path state, queues and network I/O buffers.
struct FrameCtx {
    PathBuffers paths{};
    Queues q{};
    int* train_idxs = nullptr;        // this frame's copy of the shuffled trainIdxs
    float* nn_frame_in = nullptr;
    float* nn_train_in = nullptr;
    float* nn_train_out = nullptr;
    float4* gbuffer = nullptr;
    float4* gbuffer_b = nullptr;      // NRC: throughput + bounce count at the cache query
    int* query_tiles = nullptr;       // HairMSNN: 128-pixel tiles holding at least one hair hit (the only rows whose
                                      // network output the RENDER pass reads)
    NrcTrainRec* tbuffer = nullptr;   // NRC: per-training-pixel path records
    cudaStream_t main = nullptr;      // the stream this frame's main piece runs on
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t ev_main_done = nullptr, ev_traced = nullptr, ev_free = nullptr, ev_shuffled = nullptr;
    int frame_id = 0, accum_id = 0;
    bool pretrain = false;            // this frame is a TRAIN_DATA_GEN pass
    bool counted = false;             // accum_id / frame count already advanced for this frame (deferred frames)
    // where the main piece stopped: shade queue holding the survivors, next vertex, vertex budget
    int mp_src = 0, mp_vertex = 0, mp_max = 0;
};

// Frame scheduling.  The reference renders one frame at a time on the null stream.  Here a
// frame is three chained pieces on different streams:
//   main  stream : shuffle, primary rays and the first vertices of every path (the bulk of
//                  the rays; saturates the GPU)
//   tail  stream : the remaining vertices of the few long paths (HairMSNN training paths,
//                  deep path-tracing paths): dozens of tiny latency-bound launches
//   order stream : everything with a cross-frame dependency, strictly in frame order:
//                  training step, inference, composite / accumulation
// Frame N+1's main part does not depend on frame N's tail, training or inference, so up to
// kFramesInFlight frames overlap: tails and the MLP run in the shadow of the next frames'
// main parts.  Results are identical to one-at-a-time execution.
class Renderer {
public:
    static constexpr int kFramesInFlight = 24;   // contexts allocated lazily: frames_in_flight_ of them are used
    int frames_in_flight_ = 8;                   // HM_FRAMES_IN_FLIGHT overrides
    // render_hair_msnn: the tail pieces of up to tail_group_ consecutive frames run as ONE launch sequence
    // (HM_TAIL_GROUP overrides; 1 = a tail per frame).  Frames handed to render_frames() are enqueued up to the end
    // of their main piece at once; their tail, training step, inference and composite are enqueued when the group
    // is full or when anything that observes results is called (sync, read-backs of per-frame buffers, stats, ...).
    // Read-backs of image buffers asked for in between are queued behind the frame they follow.
    int tail_group_ = 1;

    Renderer(const HostScene& hs, int kind, int beta_cli, int device, int rank, int world);
    ~Renderer();

    void render_frames(int n);          // enqueue n frames (async)
    void sync();
    void flush() { flush_deferred(); }  // enqueue everything render_frames() held back (no host wait)
    void reset_accumulation() { sync(); accum_id_ = 0; }
    // Live edits of the viewer's panels (render_hair_msnn.cu:780-1000 drawUI: hair colour / roughness / tilt /
    // lobe gains, environment scale and rotation, MIS and ENV_PDF switches).  The values travel in the kernel
    // parameter block of every launch, so an edit is a host-side store; accumulation restarts as in the viewer.
    void set_hair_params(const float sigma_a[3], float beta_m, float beta_n, float alpha_rad, const float gains[4]);
    void set_environment(float scale, float rotation);
    void set_sampling(bool mis, bool env_pdf);
    int accum_id() const { return accum_id_; }
    cudaStream_t stream() const { return order_stream_; }

    // HairMSNN split frame
    void msnn_trace();
    void msnn_train_backward();
    void msnn_train_apply();
    void msnn_finish();
    void msnn_pretrain(int steps);
    void msnn_train_data_gen();        // one TRAIN_DATA_GEN pass (genTrainingData), no training step
    Mlp* mlp() { return mlp_.get(); }

    // NRC split frame (render_nrc.cu:640-700): trace = G_BUFFER pass; query = inference over the
    // frame + training suffixes, then the RENDER pass (training records + composite);
    // train_backward / train_apply = training_step; end = buffer clears + RESET pass + accumId++
    void nrc_trace();
    void nrc_query();
    void nrc_train_backward();
    void nrc_train_apply();
    void nrc_end();
    void set_nrc_all_unbiased(bool on) { nrc_all_unbiased_ = on; }
    int nn_frame_rows() const { return nn_frame_rows_; }
    int train_records() const { return records_; }
    int in_channels() const { return in_ch_; }
    int every_nth() const { return every_nth_; }

    void* device_buffer(int which, size_t* bytes);
    void* buffer_ptr(int which, size_t* bytes);   // the same without enqueuing held-back frames first
    void trace_rays_device(const float* d_org, const float* d_dir, int n, int any, float tmin, float tmax,
                           float* d_out_hit, int* d_out_stats);
    // enqueue a device->host copy of an output buffer behind the last enqueued frame
    void readback_async(int which, void* host_dst, size_t bytes);
    void readback_rows_async(int which, int row0, int rows, void* host_dst);   // image buffers 0..6 only
    void set_profiling(bool on) { profiling_ = on; }
    // bit s set: stage s (Stats::ms index) gets event pairs while profiling is on; default all
    void set_profiling_stages(unsigned mask) { profile_mask_ = mask; }
    // event pairs only around the launches of every n-th frame (default 1 = every frame): an event record between two
    // kernels of a stream costs a few microseconds of launch gap, ~5 % of the frame rate when every frame is timed
    void set_profiling_period(int n) { profile_period_ = n < 1 ? 1 : n; }
    void set_collect_stats(bool on) { collect_stats_ = on; }
    // off: evaluate the cache for every pixel as the reference does (default: skip 128-pixel tiles without a hair hit)
    void set_skip_unused_queries(bool on) { skip_unused_queries_ = on; }
    // Multi-GPU.  The communicator's ranks form `groups` sample groups of `world` row bands each
    // (comm world = groups * band world, comm rank = group * band world + band rank).  Attaching it
    //  * sets the sample schedule: group g renders sample ids g, g + groups, ... (weak scaling by samples);
    //  * makes every training step all-reduce the network's gradients over ALL ranks on the order stream,
    //    between backward and Adam, with the loss normalised by the global batch — replicas stay bit-identical;
    //  * enables reduce_framebuffers().
    void set_comm(Comm* comm);
    // Sum of all ranks' accumulation buffers -> average over the global sample count + 8-bit frame, left in the
    // average / fb8 buffers of EVERY rank (rows a rank does not own are zero in its accumulation buffers, so one
    // all-reduce serves sample groups and row bands alike).  The accumulation buffers stay local: rendering can
    // continue afterwards.
    void reduce_framebuffers();
    int sample_groups() const { return groups_; }
    // spp sharding: RNG frame id = offset + accum_id * stride (accum_id counts this renderer's own samples)
    void set_frame_schedule(int offset, int stride) { frame_offset_ = offset; frame_stride_ = stride; }
    void reset_stats();
    Stats stats();
    int width() const { return W_; }
    int height() const { return H_; }
    int kind() const { return kind_; }
    int row0() const { return row0_; }
    int row1() const { return row1_; }
    int device() const { return device_; }
    const HostScene& host_scene() const { return hs_; }

private:
    FrameCtx& begin_frame(bool pretrain = false);
    void trace_frame(FrameCtx& c);        // main + tail pieces, ends with order_stream waiting on ev_traced
    void trace_main(FrameCtx& c);         // main piece only (fills c.mp_*)
    void trace_tail(FrameCtx* const* cs, int m);   // tail piece of m frames (m > 1: merged), finalize, ev_traced
    void msnn_order_work(FrameCtx& c);    // training step, inference, composite of one traced frame (order stream)
    void flush_deferred();
    struct Deferred { int kind; FrameCtx* c; void* src; void* dst; size_t bytes; };   // kind 0: frame, 1: read-back
    std::vector<Deferred> deferred_;
    int deferred_frames_ = 0;
    Queues group_q_[kFramesInFlight] = {};   // merged-tail queues, one per tail group of contexts
    void finish_pt(FrameCtx& c);
    void end_frame(FrameCtx& c);
    FrameParams params_for(const FrameCtx& c);
    void shuffle_train_idxs(FrameCtx& c);
    template <typename F> void timed(int stage, cudaStream_t s, F&& f);

    const HostScene& hs_;
    int kind_, beta_, device_, rank_, world_;
    int W_, H_, row0_, row1_;
    int accum_id_ = 0;
    uint64_t frames_issued_ = 0;
    FrameCtx* current_ = nullptr;   // frame between msnn_trace() and msnn_finish()
    // main_stream_: shuffles (strictly in frame order), pre-training frames and — by default — every frame's main
    // piece; work_streams_ (HM_MAIN_STREAMS > 1, an experiment that lost): main pieces round-robin
    static constexpr int kMaxWorkStreams = 4;
    cudaStream_t main_stream_ = nullptr, order_stream_ = nullptr;
    cudaStream_t work_streams_[kMaxWorkStreams] = {nullptr, nullptr, nullptr, nullptr};
    int n_work_ = 1;
    std::unique_ptr<DeviceScene> scene_;
    Camera cam_;
    FrameCtx ctx_[kFramesInFlight];
    std::vector<void*> allocs_;
    ThrustScratch scratch_;
    // outputs
    float4* bufs_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // final avg/accum, pt avg/accum, nn avg/accum
    uint32_t* fb_ = nullptr;
    // msnn
    std::unique_ptr<Mlp> mlp_;
    int in_ch_ = 12, records_ = 16384, every_nth_ = 1;
    int* d_train_idxs_ = nullptr;   // persistent permutation, re-shuffled every frame
    int n_idxs_ = 0;                // its length (training records for HairMSNN, training pixels for NRC)
    float* nn_frame_out_ = nullptr;
    // TRAIN_DATA_GEN pass: points on the strands + their shuffled index list (fetchSceneSamples)
    float* d_scene_points_ = nullptr;
    int* d_scene_indices_ = nullptr;
    int n_scene_samples_ = 0;
    void ensure_scene_samples();
    // nrc
    int nrc_train_pixels_ = 0;      // numTrainingPixels = numTrainingRecords / MAX_BOUNCES
    int nn_frame_rows_ = 0;         // rows fed to inference (nnFrameSize for NRC, W*H for HairMSNN)
    bool nrc_all_unbiased_ = false;
    float nrc_c_ = 0.01f;           // headers/render_nrc.h:132
    FrameCtx* last_ctx_ = nullptr;
    Comm* comm_ = nullptr;          // not owned
    int groups_ = 1;                // sample groups of the communicator (comm world / band world)
    void all_reduce_gradients();
    bool profiling_ = false, collect_stats_ = false, skip_unused_queries_ = true;
    int frame_offset_ = 0, frame_stride_ = 1;
    unsigned profile_mask_ = 0xffffffffu;
    int profile_period_ = 1;
    int tail_bound_items_ = 0;      // HM_TAIL_BOUND: grid bound (in queue items) of tail-piece launches; 0 = training records
    Stats stats_;
    struct Pending { int stage; cudaEvent_t a, b; };
    std::vector<Pending> pending_;
    std::vector<cudaEvent_t> event_pool_;
    cudaEvent_t take_event();
    void resolve_events();
    unsigned long long* d_trav_ = nullptr;
};

}  // namespace hm
