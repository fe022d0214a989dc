/* hairmsnn.h — C ABI of the B200-native HairMSNN per-path rendering loop.
 *
 * The reference (facebookresearch/HairMSNN) has no plugin/FFI surface: the path is
 * compiled into three executables.  This header is the boundary a maintainer would
 * bind instead of linking OWL/OptiX + tiny-cuda-nn; each entry point cites the
 * reference code it stands in for (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success
 * and a negative hm_status on failure; hm_last_error() gives the message of the
 * calling thread's last failure.  No exceptions cross the boundary.  Handles are
 * not thread-safe; independent handles may be used from different threads.
 * There is NO CPU fallback: every compute entry point fails with HM_ERR_CUDA when
 * no sm_100-class device is usable.
 */
#ifndef HAIRMSNN_H
#define HAIRMSNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hm_scene hm_scene;
typedef struct hm_renderer hm_renderer;
typedef struct hm_mlp hm_mlp;
typedef struct hm_comm hm_comm;

typedef enum {
    HM_OK = 0,
    HM_ERR_ARG = -1,      /* bad argument / malformed scene                          */
    HM_ERR_IO = -2,       /* file missing or unreadable                               */
    HM_ERR_CUDA = -3,     /* CUDA / NCCL runtime error, or no usable device           */
    HM_ERR_STATE = -4,    /* call not valid for this handle (e.g. wrong renderer kind) */
    HM_ERR_UNSUPPORTED = -5
} hm_status;

/* renderer kinds == the reference's three executables */
enum { HM_RENDER_PATH_TRACING = 0, HM_RENDER_NRC = 1, HM_RENDER_HAIR_MSNN = 2 };

/* hm_get_buffer selectors.  float4 buffers are W*H*16 bytes, FB8 is W*H*4 bytes.
 * (viewer members accumBuffer/averageBuffer/fbPointer, OWLViewer.h:232-249;
 *  HairMSNN's pt/nn/final triplets, render_hair_msnn.cu:139-145) */
enum {
    HM_BUF_FINAL_AVG = 0, HM_BUF_FINAL_ACCUM = 1,
    HM_BUF_PT_AVG = 2, HM_BUF_PT_ACCUM = 3,
    HM_BUF_NN_AVG = 4, HM_BUF_NN_ACCUM = 5,
    HM_BUF_FB8 = 6,
    HM_BUF_NN_FRAME_INPUT = 7,   /* float[W*H][in_ch]  (nnFrameInput)  */
    HM_BUF_NN_FRAME_OUTPUT = 8,  /* float[W*H][3]      (nnFrameOutput) */
    HM_BUF_NN_TRAIN_INPUT = 9,   /* float[records][in_ch]              */
    HM_BUF_NN_TRAIN_OUTPUT = 10, /* float[records][3]                  */
    HM_BUF_GBUFFER = 11,         /* float4[W*H]: rgb short-path colour, w = flags */
    HM_BUF_TRAIN_IDXS = 12,      /* int[records] (trainIdxs after this frame's shuffle; NRC: int[training pixels]) */
    HM_BUF_GBUFFER_B = 13,       /* render_nrc: float4[W*H]: rgb GBuffer::beta, w = GBuffer::bounces (int bits) */
    HM_BUF_NRC_TRAIN_RECORDS = 14,/* render_nrc: TrainBuffer[training pixels] as 5 x float[40][3] (vert wo n
                                    vertRadiance vertBeta) + int bounces + int hit = 2408 bytes each */
    HM_BUF_SCENE_INDICES = 15,   /* render_hair_msnn: int[numSamples] sceneIndices (after the first pre-training call) */
    HM_BUF_SCENE_POINTS = 16,    /* render_hair_msnn: float[numSamples][3] sampledPoints */
    /* the environment importance tables as the device holds them (built there from the uploaded map):
     * cPdf / cCdf float[env_h][env_w + 1], mPdf / mCdf float[env_h + 1] (generateEnvSamplingTables, scene.cpp:349-425) */
    HM_BUF_ENV_CPDF = 17, HM_BUF_ENV_CCDF = 18, HM_BUF_ENV_MPDF = 19, HM_BUF_ENV_MCDF = 20
};

const char* hm_last_error(void);
int hm_device_count(void);
/* bytes of the kernel parameter block a stage launch sends host->device (accounting only) */
int hm_frame_param_bytes(void);

/* ---- scene -------------------------------------------------------------------- */

/* parseScene(path, Scene&)  (scene.cpp:119-339): same config.json schema and
 * defaults.  Paths inside the file that do not exist as written (the shipped scenes
 * hold Windows absolute paths) are retried relative to the config's directory,
 * case-insensitively. */
int hm_scene_load(const char* config_json_path, hm_scene** out);

/* A scene from arrays (synthetic benchmarks, tests).  All pointers are HOST memory
 * and are copied.  Geometry is what Scene::extractHairData (scene.cpp:10-73) and
 * loadOBJ (model.cpp:233-330) produce. */
typedef struct {
    const float* control_points;   /* [num_control_points][4] xyz + radius (phantom endpoints included) */
    int num_control_points;
    const int* segment_first_cp;   /* [num_segments] index of the first of 4 control points */
    int num_segments;
    int num_strands;
    float hair_min[3], hair_max[3];/* bounds of the real points, grown from the origin (headers/model.h:93-94) */
    const float* tri_vertices;     /* [num_triangles*3][3] flattened soup */
    const float* tri_normals;      /* [num_triangles*3][3] */
    int num_triangles;
    float surface_kd[3];
    float surface_alpha;
    /* camera{} */
    float cam_from[3], cam_to[3], cam_up[3], cos_fovy;
    /* hair{} — alpha in RADIANS (the loader applies scene.cpp:207's degree conversion) */
    float sigma_a[3], beta_m, beta_n, alpha, gains[4];
    /* lights{} */
    const float* env_rgba;         /* [env_h][env_w][4] or NULL */
    int env_w, env_h;
    float env_scale, env_rotation;
    const float* dl_from;          /* [num_dlights][3], normalised here as scene.cpp:263 does */
    const float* dl_emit;          /* [num_dlights][3] */
    int num_dlights;
    /* integrator{} */
    int width, height, spp, path_v1, path_v2, mis, env_pdf;
    /* tcnn{}: path of the tiny-cuda-nn JSON, or NULL for the shipped tcnn_hairmsnn.json values */
    const char* tcnn_config_path;
} hm_scene_desc;

/* Both constructors build the acceleration structure (owlGroupBuildAccel's job in the reference,
 * render_hair_msnn.cu:401,412).  With the environment variable HM_BVH_CACHE naming a directory, the built
 * tree is stored there (file name = hash of geometry + build parameters, written atomically) and later
 * constructions of the same geometry — other ranks of a multi-GPU job, later runs — load it instead. */
int hm_scene_create(const hm_scene_desc* desc, hm_scene** out);
/* loadEnvTexture's file reader (model.cpp:158-231, tinyexr LoadEXR): single-part scanline OpenEXR,
 * HALF/FLOAT/UINT channels, NONE/ZIPS/ZIP/PIZ compression -> RGBA32F, top row first, alpha 1 when
 * the file has none.  Call with rgba = NULL to get the size, then with a buffer of
 * capacity_floats >= 4*w*h. */
int hm_image_load_exr(const char* path, float* rgba, size_t capacity_floats, int* width, int* height);
/* LoadCemYuksel + Scene::extractHairData (scene.cpp:75-117, 10-73) alone — the .hair reader without the rest of the
 * scene or the acceleration structure.  counts3 = control points (phantom end points included), segments, strands.
 * Call with NULL arrays for the counts, then with cps4 [control points][4] (xyz + radius = 0.2 x file thickness),
 * segment_first_cp [segments] and bounds6 (min xyz, max xyz of the file's points, grown from the origin as
 * headers/model.h:93-94 does); any of the three may stay NULL. */
int hm_hair_file_load(const char* path, int* counts3, float* cps4, int* segment_first_cp, float* bounds6);
/* stores the scene's wide tree in `dir` under the name HM_BVH_CACHE lookups use (no-op if it came from there) */
int hm_scene_save_bvh_cache(const hm_scene* scene, const char* dir);
void hm_scene_free(hm_scene* scene);

typedef struct {
    int width, height, spp, path_v1, path_v2;
    int num_segments, num_control_points, num_triangles, num_strands, num_bvh_nodes;
    float scene_scale;
    float cam_pos[3], cam_d00[3], cam_du[3], cam_dv[3];  /* cameraChanged(), render_path_tracing.cu:767-797 */
    int env_w, env_h, num_dlights;
    /* the 8-wide quantised tree the kernels traverse (80 B nodes), its leaf references (64 B each) and
     * depth.  num_bvh_nodes counts the binary SAH tree it is derived from: 0 when the scene came from the
     * HM_BVH_CACHE directory (the cache stores the wide tree only). */
    int num_wide_nodes, num_wide_leaf_refs, wide_depth;
} hm_scene_info;
int hm_scene_get_info(const hm_scene* scene, hm_scene_info* info);
/* host copies of the acceleration structure + geometry (test hook: lets the oracle
 * traverse the SAME tree).  Pointers stay valid until hm_scene_free. */
int hm_scene_get_arrays(const hm_scene* scene, const float** bvh_nodes, const int** leaf_code, const int** leaf_prim,
                        const float** control_points, const float** tri_vertices4, const float** tri_normals4,
                        const int** segment_first_cp, const float** leaf_data16);
/* generateEnvSamplingTables (scene.cpp:349-425): cPdf/cCdf [(w+1)*h], mPdf/mCdf [h+1] */
int hm_scene_get_env_tables(const hm_scene* scene, const float** env_rgba, const float** cpdf, const float** ccdf,
                            const float** mpdf, const float** mcdf);

/* ---- renderer ------------------------------------------------------------------ */

/* RenderWindowPT / RenderWindowNRC / RenderWindow_HairMSNN ::initialize()
 * (render_path_tracing.cu:98-402, render_nrc.cu:116-630, render_hair_msnn.cu:98-645).
 * beta_cli is the [BETA] argument (render_hair_msnn.cu:1146-1170: internal beta = BETA-1).
 * The frame is split into `world` contiguous row bands; this handle renders band
 * `rank` on CUDA device `device`.  world = 1 renders everything. */
/* The renderer shares ownership of the host scene: hm_scene_free may be called before hm_renderer_destroy. */
int hm_renderer_create(hm_scene* scene, int kind, int beta_cli, int device, int rank, int world, hm_renderer** out);
void hm_renderer_destroy(hm_renderer* r);

/* n calls of RenderWindow*::render() (render_path_tracing.cu:404-423,
 * render_hair_msnn.cu:699-771): one sample per pixel each, accumId advances. */
int hm_render_frames(hm_renderer* r, int n_frames);
/* same, but returns after enqueueing on the renderer's stream.
 * render_hair_msnn renderers run the tail pieces (the long training paths) of up to 4 consecutive frames as one
 * launch sequence: a frame's main piece is enqueued by this call, its tail, training step, inference and composite
 * when its group of frames is complete or when a call that observes results is made (hm_renderer_sync, hm_render_frames,
 * hm_get_buffer, hm_get_device_buffer, hm_renderer_get_stats, hm_reduce_framebuffers, hm_renderer_mlp, the split-frame
 * and live-edit calls ...).
 * hm_readback_async / hm_readback_rows_async of the image buffers issued in between are queued behind the frame they
 * follow, so a per-frame read-back keeps its place in the stream order without breaking the group.  Results are
 * bit-identical to frame-at-a-time execution.  hm_render_flush enqueues whatever is held back without waiting
 * (call it before recording an event on hm_renderer_stream that is meant to follow the frames). */
int hm_render_frames_async(hm_renderer* r, int n_frames);
int hm_render_flush(hm_renderer* r);
int hm_renderer_sync(hm_renderer* r);
/* restart accumulation (cameraChanged(): accumId = 0) */
int hm_renderer_reset_accumulation(hm_renderer* r);
int hm_renderer_accum_id(const hm_renderer* r);
/* Live edits, as the viewer's ImGui panels make them (render_hair_msnn.cu drawUI / render_path_tracing.cu drawUI:
 * hair sigma_a, beta_m, beta_n, alpha, the four lobe gains; environment scale and rotation; MIS and ENV_PDF).
 * Each call waits for the frames in flight, stores the values (they travel in every launch's parameter block)
 * and restarts accumulation (accumId = 0), which is what the panels do.  alpha in radians. */
int hm_renderer_set_hair_params(hm_renderer* r, const float* sigma_a3, float beta_m, float beta_n, float alpha_radians,
                                const float* gains4);
int hm_renderer_set_environment(hm_renderer* r, float scale, float rotation);
int hm_renderer_set_sampling(hm_renderer* r, int mis, int env_pdf);
/* CUDA stream the renderer launches on (cudaStream_t as void*), for event timing */
void* hm_renderer_stream(hm_renderer* r);

/* HairMSNN only: the split halves of render(), for multi-GPU gradient exchange
 * (SURVEY §8e).  trace = G_BUFFER pass; then hm_renderer_mlp() forward/backward,
 * caller all-reduces hm_mlp_gradients(), hm_mlp_optimizer_step(); then finish =
 * inference + RENDER pass. */
int hm_msnn_trace(hm_renderer* r);
int hm_msnn_train_backward(hm_renderer* r);
int hm_msnn_train_apply(hm_renderer* r);
int hm_msnn_finish(hm_renderer* r);
/* The initial training of RenderWindow_HairMSNN::initialize (render_hair_msnn.cu:633-641): n_steps of
 * genTrainingData() [the TRAIN_DATA_GEN pass, cuda/hair_msnn.cu:222-233: 128 x 128 training paths towards
 * random points on the strands (fetchSceneSamples, render_hair_msnn.cu:34-97)] + train().  The reference
 * runs it for one wall-clock second from an unset camera with time-seeded samples; here the step count is
 * the caller's, the rays start at the scene's camera position and the samples come from a fixed seed.
 * hm_msnn_train_data_gen runs one pass without the training step (records: HM_BUF_NN_TRAIN_INPUT/OUTPUT). */
int hm_msnn_pretrain(hm_renderer* r, int n_steps);
int hm_msnn_train_data_gen(hm_renderer* r);
hm_mlp* hm_renderer_mlp(hm_renderer* r);

/* render_nrc only: the pieces of RenderWindowNRC::render() (render_nrc.cu:640-700) in the reference's
 * order.  trace = shuffle + G_BUFFER pass (nrcTracePaths, cuda/nrc.cu:135-310); query = inference over
 * the frame's cache queries and the training suffixes (nnFrameSize rows) + RENDER pass
 * (nrcGenerateTrainingData + composite, cuda/nrc.cu:69-133,367-381); train_backward / train_apply =
 * trainer->training_step over the 65536 records (a multi-GPU caller all-reduces hm_mlp_gradients()
 * in between); end = accumId++ (the buffer clears and the RESET pass happen at the next trace).
 * For render_nrc BUF_GBUFFER holds rgb = GBuffer::pathRadiance, w = hit.  With the 9 input channels the
 * tcnn composite encoding has no identity part and OneBlob pads: as tiny-cuda-nn does (oneblob.h:221-225,
 * pinned by tests/golden/tcnn_9.npz), network inputs 38..45 are overwritten with 1 and the padding inputs
 * 56..63 — left unwritten by the reference — are 0. */
int hm_nrc_trace(hm_renderer* r);
int hm_nrc_query(hm_renderer* r);
int hm_nrc_train_backward(hm_renderer* r);
int hm_nrc_train_apply(hm_renderer* r);
int hm_nrc_end(hm_renderer* r);
/* LaunchParams::allUnbiased (cuda_headers/nrc.cuh:77): every training pixel traces its full path */
int hm_nrc_set_all_unbiased(hm_renderer* r, int on);
/* out4 = MLP input channels, rows fed to inference per frame (W*H, or nnFrameSize for render_nrc,
 * render_nrc.cu:140-145), training records per step, everyNth */
int hm_renderer_get_layout(const hm_renderer* r, int* out4);

int hm_get_buffer(hm_renderer* r, int which, void* host_dst, size_t bytes);
/* asynchronous variant: the copy is enqueued behind the last enqueued frame; host_dst (ideally
 * pinned) is valid after hm_renderer_sync().  Lets a caller stream every frame's result to the
 * host without stalling the frames in flight. */
int hm_readback_async(hm_renderer* r, int which, void* host_dst, size_t bytes);
/* same for rows [row0, row0 + rows) of an image buffer (HM_BUF_FINAL_AVG .. HM_BUF_FB8): what a row-band rank owns.
 * host_dst receives rows * W pixels. */
int hm_readback_rows_async(hm_renderer* r, int which, int row0, int rows, void* host_dst);
/* the rows this renderer renders: out2 = row0, row1 (hm_band_partition of its rank / world) */
int hm_renderer_get_rows(const hm_renderer* r, int* out2);
/* device pointer of the same buffers (for in-place NCCL gathers) */
int hm_get_device_buffer(hm_renderer* r, int which, void** dev_ptr, size_t* bytes);

/* OWLViewer::screenShot (OWLViewer.cpp:109-124) / saveEXR (model.cpp:366-383): rows flipped */
int hm_save_png(hm_renderer* r, const char* path);
int hm_save_exr(hm_renderer* r, int which, const char* path);
/* integrator.stats_output — parsed by the reference (scene.cpp:295) but never written */
int hm_write_stats(hm_renderer* r, const char* path);

/* ms_shade / ms_extend: k_shade / k_trace launches of a frame's main piece (a k_trace launch traces a vertex's
 * occlusion probes AND continuation rays); ms_shadow: all launches of its tail piece (vertices past beta+1,
 * long paths only). */
typedef struct {
    double ms_primary, ms_shade, ms_extend, ms_shadow, ms_finalize, ms_train, ms_infer, ms_composite, ms_total;
    uint64_t rays_primary, rays_extend, rays_shadow, shade_items;
    uint64_t kernel_launches;
    /* launches per stage: primary shade extend shadow finalize train infer composite */
    uint64_t stage_launches[8];
    /* instrumented traversal (hm_renderer_set_collect_stats): BVH nodes visited / primitives tested */
    uint64_t trav_nodes_extend, trav_prims_extend, trav_nodes_shadow, trav_prims_shadow, trav_nodes_primary, trav_prims_primary;
    /* share of the extend+shadow totals traced by tail-piece launches */
    uint64_t trav_nodes_tail, trav_prims_tail, rays_tail;
    float last_loss;
    int frames;
    /* launches per stage that carried an event pair: the ms_* sums cover these (hm_renderer_set_profiling_period) */
    uint64_t timed_launches[8];
} hm_stats;
int hm_renderer_get_stats(hm_renderer* r, hm_stats* out);
/* per-stage CUDA-event timing on/off (event pairs around each launch, resolved in hm_renderer_get_stats) */
int hm_renderer_set_profiling(hm_renderer* r, int on);
/* which stages get event pairs while profiling is on: bit 0 primary, 1 shade (main piece), 2 trace (main piece), 3 shade + trace (tail
 * piece), 4 finalize, 5 train, 6 infer, 7 composite, 8 whole frame; default all.  A frame has ~320 launches:
 * timing all of them costs ~3 % of the frame rate. */
int hm_renderer_set_profiling_stages(hm_renderer* r, unsigned mask);
/* event pairs around the launches of every n-th frame only (default 1).  An event record between two kernels of a stream
 * costs a few microseconds of launch gap: timing every frame's traversal and network launches costs ~5 % of the frame rate. */
int hm_renderer_set_profiling_period(hm_renderer* r, int every_nth_frame);
/* instrumented traversal + queue-size accounting (polls the queue counters every bounce) */
int hm_renderer_set_collect_stats(hm_renderer* r, int on);
/* render_hair_msnn: the RENDER pass reads the network output of hair-hit pixels only (cuda/hair_msnn.cu:325-340).
 * on (default): inference skips 128-pixel tiles without a hair hit (their HM_BUF_NN_FRAME_OUTPUT rows keep
 * stale values; images are identical).  off: all W*H rows are evaluated, as the reference does. */
int hm_renderer_set_skip_unused_queries(hm_renderer* r, int on);
int hm_renderer_reset_stats(hm_renderer* r);
/* Sample schedule for spp-sharded rendering: the RNG frame id of this renderer's
 * k-th sample is offset + k * stride (default 0, 1 == the reference's accumId). */
int hm_renderer_set_frame_schedule(hm_renderer* r, int offset, int stride);
/* Row-band partition used by hm_renderer_create(rank, world) (pure host arithmetic, no
 * device needed): out5 = row0, row1, first training record, training records owned, records fed
 * to the backward pass (rounded down to tcnn's 128-row granularity, common.h:280).  `records` is
 * numTrainRecordsX*Y = 16384 for render_hair_msnn (headers/render_hair_msnn.h:127-130). */
int hm_band_partition(int width, int height, int records, int rank, int world, int* out5);
/* Buffer sizes of render_nrc for a frame (RenderWindowNRC::initialize, render_nrc.cu:116-160; pure host
 * arithmetic): out4 = numTrainingPixels (65536 / MAX_BOUNCES), everyNth, nnFrameSize (rows fed to inference:
 * frame + training suffixes, rounded up to tcnn's 128-row granularity), numTrainingRecords. */
int hm_nrc_layout(int width, int height, int* out4);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch ----------
 * The reference is single-GPU (render_hair_msnn.cu:1056: owlContextCreate(nullptr, 1)); these calls are the
 * new partitioning of its render()/train() loop.  Rank 0 obtains a 128-byte id and hands it to the other
 * processes by any means (file, pipe, TCP store); every process then creates its communicator on its device. */
#define HM_COMM_ID_BYTES 128
int hm_comm_get_unique_id(void* out_id128);
int hm_comm_create(const void* id128, int rank, int world, int device, hm_comm** out);
void hm_comm_destroy(hm_comm* c);
int hm_comm_barrier(hm_comm* c);
/* max / sum of one double over all ranks (timing: max over ranks of a device-measured duration) */
int hm_comm_all_reduce_max(hm_comm* c, double* inout);
int hm_comm_all_reduce_sum(hm_comm* c, double* inout);
/* Attaches the communicator to a renderer (NULL detaches).  comm world = G sample groups x B row bands, where
 * B is the renderer's own `world` (hm_renderer_create) and comm rank = group * B + band.  From then on
 *  - group g renders sample ids g, g + G, g + 2G, ... (hm_renderer_set_frame_schedule is set accordingly);
 *  - every training step inside hm_render_frames / hm_msnn_train_apply / hm_nrc_train_apply / hm_msnn_pretrain
 *    all-reduces the network's gradients over all ranks between backward and Adam (ncclAllReduce on the
 *    renderer's stream, loss normalised by the global batch), so the replicas' weights stay bit-identical;
 *  - hm_reduce_framebuffers is available.
 * Every rank must make the same sequence of rendering calls.  The communicator must outlive the attachment. */
int hm_renderer_set_comm(hm_renderer* r, hm_comm* c);
/* Sums the ranks' accumulation buffers (row bands a rank does not own are zero in its buffers, so one
 * all-reduce serves sample groups and bands alike) and leaves the average over the GLOBAL sample count and
 * the 8-bit frame in HM_BUF_*_AVG / HM_BUF_FB8 of every rank.  Accumulation buffers stay local. */
int hm_reduce_framebuffers(hm_renderer* r);

/* ---- stand-alone kernels (parity tests, micro-benchmarks) ----------------------- */

/* owl::traceRay (owl_device.h:153-177) for a batch of HOST rays.  any_hit = 0: closest
 * hit (ray type 0), 1: occlusion (ray type 1).  out_hit [n][4] = t, prim (int bits),
 * u, v; prim = segment id, or num_segments + triangle id, or -1.
 * out_stats [n][2] = nodes visited, primitives tested (may be NULL). */
int hm_trace_rays(hm_renderer* r, const float* org3, const float* dir3, int n, int any_hit, float tmin, float tmax,
                  float* out_hit4, int* out_stats2);
/* same with DEVICE pointers, asynchronous on the renderer's stream */
int hm_trace_rays_device(hm_renderer* r, const float* d_org3, const float* d_dir3, int n, int any_hit, float tmin,
                         float tmax, float* d_out_hit4);

/* disney_hair(si, pdf) (cuda_headers/disney_hair.cuh:185-266): f * cos and pdf of the fibre scattering model for n
 * HOST (wo_local, wi_local, h) triples, evaluated by the sm_100a build on `device`.  Local frame: x = fibre tangent;
 * h = dot(to_local[1], n).  alpha in radians; gains4 = R, TT, TRT, TRRT.  NaNs come back as NaNs (the renderer's
 * callers scrub them as the reference does). */
int hm_bsdf_eval(int device, const float* sigma_a3, float beta_m, float beta_n, float alpha_radians, const float* gains4,
                 const float* wo_local3, const float* wi_local3, const float* h, int n, float* out_f3, float* out_pdf);
/* sample_disney_hair (disney_hair.cuh:276-386): rand4 = the four draws in the reference's order (lobe, theta,
 * phi-of-theta, dphi); returns the sampled LOCAL direction and disney_hair evaluated on it. */
int hm_bsdf_sample(int device, const float* sigma_a3, float beta_m, float beta_n, float alpha_radians, const float* gains4,
                   const float* wo_local3, const float* h, const float* rand4, int n, float* out_wi_local3, float* out_f3, float* out_pdf);

/* ---- MLP: TINY_MLP (cuda_headers/neural_network.cuh:33-50, cuda/neural_network.cu) ---- */

/* TINY_MLP(configPath, inCh, outCh): config_path = tiny-cuda-nn JSON (NULL = the shipped
 * tcnn_hairmsnn.json values).  Weights are initialised as tcnn does (pcg32 seeded from
 * std::seed_seq{1337}; trainer.h:53-99). */
int hm_mlp_create(const char* config_path, int in_ch, int out_ch, int device, hm_mlp** out);
void hm_mlp_destroy(hm_mlp* m);
/* TINY_MLP::inference(float* in, float* out, int n): DEVICE pointers, AoS [n][in_ch] -> [n][out_ch];
 * n must be a multiple of 128 (tcnn batch granularity, common.h:280) */
int hm_mlp_inference(hm_mlp* m, const float* d_in, float* d_out, int n);
/* host-pointer convenience wrapper (copies in and out) */
int hm_mlp_inference_host(hm_mlp* m, const float* in, float* out, int n);
/* trainer->training_step(input, target) + trainer->loss()  (trainer.h:168-190):
 * forward + loss + backward (+ optimizer step).  DEVICE pointers.  loss may be NULL. */
int hm_mlp_train_step(hm_mlp* m, const float* d_in, const float* d_target, int n, float* loss);
int hm_mlp_train_step_host(hm_mlp* m, const float* in, const float* target, int n, float* loss);
/* split form: forward+backward with loss normalised by n_total_records (global batch), then step */
int hm_mlp_forward_backward(hm_mlp* m, const float* d_in, const float* d_target, int n, int n_total_records);
int hm_mlp_gradients(hm_mlp* m, float** d_grads, size_t* count);   /* fp32, loss-scaled by 128 */
int hm_mlp_optimizer_step(hm_mlp* m);
int hm_mlp_loss(hm_mlp* m, float* loss);
/* TINY_MLP::reset() zeroes the weights (cuda/neural_network.cu:17-21) */
int hm_mlp_reset(hm_mlp* m);
int hm_mlp_reinitialize(hm_mlp* m);   /* back to the seeded initial weights + fresh optimizer */
size_t hm_mlp_n_params(const hm_mlp* m);
/* fp32 master parameters, tcnn order: MLP matrices first, then the hash grid
 * (network_with_input_encoding.h:113-130) */
int hm_mlp_get_params(hm_mlp* m, float* host_dst, size_t count);
int hm_mlp_set_params(hm_mlp* m, const float* host_src, size_t count);
/* hm_mlp_save: raw little-endian fp32 blob.  hm_mlp_save_snapshot: Trainer::serialize (trainer.h:270-283)
 * as the text JSON TINY_MLP::loadWeights reads (cuda/neural_network.cu:23-32): n_params, params_type
 * "float", params_binary {"bytes": [...]}.  hm_mlp_load accepts both, and snapshots of params_type
 * "__half" (Trainer::deserialize, trainer.h:285-310).  scene.tcnn.init_weights is loaded this way by
 * hm_renderer_create (render_nrc.cu:148-150). */
int hm_mlp_save(hm_mlp* m, const char* path);
int hm_mlp_save_snapshot(hm_mlp* m, const char* path);
int hm_mlp_load(hm_mlp* m, const char* path);
void* hm_mlp_stream(hm_mlp* m);
uint64_t hm_mlp_launch_count(const hm_mlp* m);

#ifdef __cplusplus
}
#endif
#endif /* HAIRMSNN_H */
