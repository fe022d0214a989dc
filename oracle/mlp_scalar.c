/* oracle/mlp_scalar.c — TEST INFRASTRUCTURE / CPU BASELINE (never linked into or called by the product).
 *
 * Scalar C restatement of the tiny-cuda-nn network the reference instantiates (scenes/.../tcnn_hairmsnn.json):
 * Composite[HashGrid 16x2 | OneBlob 6x4 | Identity] -> 64 -> ReLU 64 -> ReLU 64 -> 16 (3 used), RelativeL2Luminance
 * loss, Adam.  Paths below are relative to /root/reference/extern/tiny-cuda-nn.
 *   grid:     include/tiny-cuda-nn/encodings/grid.h:171-204 (index, scale), :221-351 (forward), :395-516 (backward)
 *   pos_fract include/tiny-cuda-nn/common_device.h:434-445; quartic_cdf :492-497
 *   oneblob   include/tiny-cuda-nn/encodings/oneblob.h:99-127; identity encodings/identity.h:46-66
 *   mlp       src/fully_fused_mlp.cu:499-557 (forward), :150-259 + :784-842 (backward, weight gradients)
 *   loss      include/tiny-cuda-nn/losses/relative_l2_luminance.h:40-87
 *   adam      include/tiny-cuda-nn/optimizers/adam.h:48-120
 * Arithmetic is fp32 throughout on the parameter values it is given (pass fp16-rounded values to follow tcnn's
 * storage); it is the "scalar MLP with the same weights" of BASELINE.md §3 and is pinned against the golden
 * vectors of tiny-cuda-nn itself in tests/test_cpu_mlp_oracle.py (fp16-accumulation tolerance stated there).
 * Rows are distributed over `threads` pthreads.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEVELS 16
#define FEATS 2
#define WIDTH 64
#define OUT_PAD 16
#define BLOB_DIMS 6
#define BLOB_BINS 4
#define N_MATRIX (WIDTH * WIDTH * 2 + OUT_PAD * WIDTH)

typedef struct {
    uint32_t offset[LEVELS + 1];
    float scale[LEVELS];
    uint32_t res[LEVELS];
} Layout;

static void make_layout(Layout* L) {
    uint32_t off = 0;
    for (int l = 0; l < LEVELS; ++l) {
        float scale = exp2f((float)l * log2f(2.0f)) * 16.0f - 1.0f;   /* grid_scale, grid.h:195-199 */
        uint32_t res = (uint32_t)ceilf(scale) + 1;                  /* grid_resolution, grid.h:201-204 */
        uint64_t n = (uint64_t)res * res * res;
        if (n > 0x7fffffffu) n = 0x7fffffffu;
        n = (n + 7) / 8 * 8;
        if (n > (1u << 15)) n = 1u << 15;
        L->offset[l] = off; L->scale[l] = scale; L->res[l] = res;
        off += (uint32_t)n;
    }
    L->offset[LEVELS] = off;
}

static inline uint32_t grid_index(const uint32_t p[3], uint32_t res, uint32_t size) {
    uint64_t stride = 1;
    uint32_t index = 0;
    int dim = 0;
    for (; dim < 3 && stride <= size; ++dim) { index += p[dim] * (uint32_t)stride; stride *= res; }
    if (size < stride) index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    return index % size;
}

static inline float quartic_cdf(float x, float inv_radius) {
    float u = x * inv_radius, u2 = u * u, u4 = u2 * u2;
    float v = (15.0f / 16.0f) * u * (1.0f - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f;
    return v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
}

typedef struct { uint32_t idx[LEVELS][8]; float w[LEVELS][8]; } Corners;

static void encode_row(const Layout* L, const float* params, const float* x, int in_ch, float* e, Corners* keep) {
    const float* table = params + N_MATRIX;
    for (int l = 0; l < LEVELS; ++l) {
        uint32_t pg[3]; float fr[3];
        for (int d = 0; d < 3; ++d) {
            float pos = x[d] * L->scale[l] + 0.5f;
            float fl = floorf(pos);
            pg[d] = (uint32_t)(int32_t)fl;
            fr[d] = pos - fl;
        }
        const uint32_t size = L->offset[l + 1] - L->offset[l];
        float a0 = 0.f, a1 = 0.f;
        for (int c = 0; c < 8; ++c) {
            float w = 1.f; uint32_t loc[3];
            for (int d = 0; d < 3; ++d) {
                if ((c >> d) & 1) { w *= fr[d]; loc[d] = pg[d] + 1u; }
                else { w *= 1.f - fr[d]; loc[d] = pg[d]; }
            }
            uint32_t i = grid_index(loc, L->res[l], size) + L->offset[l];
            a0 += w * table[2 * (size_t)i]; a1 += w * table[2 * (size_t)i + 1];
            if (keep) { keep->idx[l][c] = i; keep->w[l][c] = w; }
        }
        e[2 * l] = a0; e[2 * l + 1] = a1;
    }
    int c0 = LEVELS * FEATS;
    for (int j = 0; j < BLOB_DIMS; ++j) {
        float xv = x[3 + j];
        float left = quartic_cdf(-xv, BLOB_BINS) + quartic_cdf(-xv - 1.f, BLOB_BINS) + quartic_cdf(-xv + 1.f, BLOB_BINS);
        for (int k = 0; k < BLOB_BINS; ++k) {
            float rb = (float)(k + 1) / BLOB_BINS;
            float right = quartic_cdf(rb - xv, BLOB_BINS) + quartic_cdf(rb - xv - 1.f, BLOB_BINS) + quartic_cdf(rb - xv + 1.f, BLOB_BINS);
            e[c0 + j * BLOB_BINS + k] = right - left;
            left = right;
        }
    }
    int c1 = c0 + BLOB_DIMS * BLOB_BINS, ident = in_ch - 3 - BLOB_DIMS;
    for (int j = 0; j < ident; ++j) e[c1 + j] = x[3 + BLOB_DIMS + j];
    for (int j = c1 + ident; j < WIDTH; ++j) e[j] = 1.f;
}

static void layers(const float* params, const float* e, float* h1, float* h2, float* y) {
    const float* W0 = params; const float* W1 = params + WIDTH * WIDTH; const float* Wo = params + 2 * WIDTH * WIDTH;
    for (int o = 0; o < WIDTH; ++o) { float s = 0.f; for (int i = 0; i < WIDTH; ++i) s += W0[o * WIDTH + i] * e[i]; h1[o] = s > 0.f ? s : 0.f; }
    for (int o = 0; o < WIDTH; ++o) { float s = 0.f; for (int i = 0; i < WIDTH; ++i) s += W1[o * WIDTH + i] * h1[i]; h2[o] = s > 0.f ? s : 0.f; }
    for (int o = 0; o < OUT_PAD; ++o) { float s = 0.f; for (int i = 0; i < WIDTH; ++i) s += Wo[o * WIDTH + i] * h2[i]; y[o] = s; }
}

typedef struct {
    const Layout* L; const float* params; const float* x; const float* target; float* out;
    float* grads;       /* private [N_MATRIX] accumulator of the matrix gradients */
    float* grid_grads;  /* shared grid gradient buffer, updated with atomic adds */
    int in_ch, r0, r1, n_total; double loss;
} Job;

static inline void atomic_addf(float* p, float v) {
    uint32_t* u = (uint32_t*)p;
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
    do {
        float f; memcpy(&f, &old, 4);
        f += v;
        memcpy(&neu, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static void* infer_job(void* p) {
    Job* j = (Job*)p;
    float e[WIDTH], h1[WIDTH], h2[WIDTH], y[OUT_PAD];
    for (int r = j->r0; r < j->r1; ++r) {
        encode_row(j->L, j->params, j->x + (size_t)r * j->in_ch, j->in_ch, e, NULL);
        layers(j->params, e, h1, h2, y);
        for (int k = 0; k < 3; ++k) j->out[3 * (size_t)r + k] = y[k];
    }
    return NULL;
}

/* forward + loss + backward of rows [r0,r1): matrix gradients into this job's private buffer, grid gradients
 * scattered into the shared buffer (all loss-scaled by 128) */
static void* train_job(void* p) {
    Job* j = (Job*)p;
    const float* W0 = j->params; const float* W1 = j->params + WIDTH * WIDTH; const float* Wo = j->params + 2 * WIDTH * WIDTH;
    float* g = j->grads;
    float e[WIDTH], h1[WIDTH], h2[WIDTH], y[OUT_PAD], dy[OUT_PAD], dh2[WIDTH], dh1[WIDTH], de[WIDTH];
    Corners cn;
    const float n_total = (float)j->n_total * 3.f;
    double loss = 0.0;
    for (int r = j->r0; r < j->r1; ++r) {
        encode_row(j->L, j->params, j->x + (size_t)r * j->in_ch, j->in_ch, e, &cn);
        layers(j->params, e, h1, h2, y);
        float lum = 0.299f * y[0] + 0.587f * y[1] + 0.114f * y[2];
        float denom = lum * lum + 0.01f;
        memset(dy, 0, sizeof(dy));
        for (int k = 0; k < 3; ++k) {
            float diff = y[k] - j->target[3 * (size_t)r + k];
            loss += (double)(diff * diff / denom / n_total);
            dy[k] = 128.f * (2.f * diff / denom) / n_total;
        }
        for (int i = 0; i < WIDTH; ++i) { float s = 0.f; for (int o = 0; o < OUT_PAD; ++o) s += dy[o] * Wo[o * WIDTH + i]; dh2[i] = h2[i] > 0.f ? s : 0.f; }
        for (int i = 0; i < WIDTH; ++i) { float s = 0.f; for (int o = 0; o < WIDTH; ++o) s += dh2[o] * W1[o * WIDTH + i]; dh1[i] = h1[i] > 0.f ? s : 0.f; }
        for (int i = 0; i < LEVELS * FEATS; ++i) { float s = 0.f; for (int o = 0; o < WIDTH; ++o) s += dh1[o] * W0[o * WIDTH + i]; de[i] = s; }
        for (int o = 0; o < WIDTH; ++o) for (int i = 0; i < WIDTH; ++i) g[o * WIDTH + i] += dh1[o] * e[i];
        for (int o = 0; o < WIDTH; ++o) for (int i = 0; i < WIDTH; ++i) g[WIDTH * WIDTH + o * WIDTH + i] += dh2[o] * h1[i];
        for (int o = 0; o < OUT_PAD; ++o) for (int i = 0; i < WIDTH; ++i) g[2 * WIDTH * WIDTH + o * WIDTH + i] += dy[o] * h2[i];
        for (int l = 0; l < LEVELS; ++l)
            for (int c = 0; c < 8; ++c) {
                float* t = j->grid_grads + 2 * (size_t)cn.idx[l][c];
                atomic_addf(t, cn.w[l][c] * de[2 * l]); atomic_addf(t + 1, cn.w[l][c] * de[2 * l + 1]);
            }
    }
    j->loss = loss;
    return NULL;
}

static int run_jobs(void* (*fn)(void*), Job* jobs, int nt) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
    for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

size_t mlps_n_params(void) { Layout L; make_layout(&L); return (size_t)N_MATRIX + (size_t)L.offset[LEVELS] * FEATS; }

/* TINY_MLP::inference: x AoS [n][in_ch] -> out AoS [n][3] */
void mlps_inference(const float* params, const float* x, int n, int in_ch, float* out, int threads) {
    Layout L; make_layout(&L);
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    Job* jobs = (Job*)calloc(threads, sizeof(Job));
    for (int t = 0; t < threads; ++t) {
        jobs[t].L = &L; jobs[t].params = params; jobs[t].x = x; jobs[t].out = out; jobs[t].in_ch = in_ch;
        jobs[t].r0 = (int)((long long)n * t / threads); jobs[t].r1 = (int)((long long)n * (t + 1) / threads);
    }
    run_jobs(infer_job, jobs, threads);
    free(jobs);
}

/* forward + loss + backward: ACCUMULATES the gradients (loss-scaled by 128) into grads [n_params] — the caller
 * zeroes it (mlps_adam clears what it consumes, as the optimiser kernel does); returns the loss */
double mlps_gradients(const float* params, const float* x, const float* target, int n, int in_ch, int n_total_records, float* grads, int threads) {
    Layout L; make_layout(&L);
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    Job* jobs = (Job*)calloc(threads, sizeof(Job));
    for (int t = 0; t < threads; ++t) {
        jobs[t].L = &L; jobs[t].params = params; jobs[t].x = x; jobs[t].target = target; jobs[t].in_ch = in_ch;
        jobs[t].n_total = n_total_records > 0 ? n_total_records : n;
        jobs[t].grads = (float*)calloc(N_MATRIX, sizeof(float));
        jobs[t].grid_grads = grads + N_MATRIX;
        jobs[t].r0 = (int)((long long)n * t / threads); jobs[t].r1 = (int)((long long)n * (t + 1) / threads);
    }
    run_jobs(train_job, jobs, threads);
    double loss = 0.0;
    for (int t = 0; t < threads; ++t) {
        for (size_t i = 0; i < (size_t)N_MATRIX; ++i) grads[i] += jobs[t].grads[i];
        loss += jobs[t].loss;
        free(jobs[t].grads);
    }
    free(jobs);
    return loss;
}

/* adam_step (adam.h:48-120) on parameters [first, first+count): fp32 master/m1/m2 + per-parameter step counts;
 * L2 regularisation on the matrices only, grid entries with a zero gradient are skipped; consumed gradients are cleared. */
void mlps_adam(float* master, float* m1, float* m2, uint32_t* steps, float* grads_scaled, size_t first, size_t count,
               float lr, float beta1, float beta2, float eps, float l2_reg) {
    for (size_t i = first; i < first + count; ++i) {
        float g = grads_scaled[i] / 128.f;
        grads_scaled[i] = 0.f;
        const int is_matrix = i < (size_t)N_MATRIX;
        if (!is_matrix && g == 0.f) continue;
        if (is_matrix) g += l2_reg * master[i];
        const float a = m1[i] = beta1 * m1[i] + (1.f - beta1) * g;
        const float b = m2[i] = beta2 * m2[i] + (1.f - beta2) * g * g;
        const uint32_t s = ++steps[i];
        const float lr_t = lr * sqrtf(1.f - powf(beta2, (float)s)) / (1.f - powf(beta1, (float)s));
        master[i] -= lr_t / (sqrtf(b) + eps) * a;
    }
}
