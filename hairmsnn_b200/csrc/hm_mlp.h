// hm_mlp.h — the radiance-cache network: composite encoding (hash grid | OneBlob |
// identity) -> 64-wide fused MLP, RelativeL2Luminance loss, Adam + exponential decay.
//
// Stands in for TINY_MLP over tiny-cuda-nn (cuda/neural_network.cu:34-64,
// cuda_headers/neural_network.cuh:33-50) for the one network configuration the
// shipped scenes instantiate.  Kernels live in hm_mlp.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace hm {

struct MlpConfig {
    int in_ch = 12, out_ch = 3;
    // HashGrid
    int grid_dims = 3, n_levels = 16, feats = 2, log2_hashmap = 15, base_res = 16;
    float per_level_scale = 2.0f;
    // OneBlob
    int blob_dims = 6, blob_bins = 4;
    // network
    int width = 64, hidden_layers = 2;
    // optimizer
    float lr = 1e-2f, beta1 = 0.9f, beta2 = 0.99f, eps = 1e-15f, l2_reg = 1e-6f;
    int decay_start = 4000, decay_interval = 4000;
    float decay_base = 0.33f;
    uint32_t seed = 1337;
};

// Parses the tiny-cuda-nn JSON subset; throws std::runtime_error on unknown otypes.
MlpConfig mlp_config_from_json(const std::string& path, int in_ch, int out_ch);

static constexpr int kMlpMaxLevels = 16;
struct GridLayout {
    uint32_t offset[kMlpMaxLevels + 1];   // in entries (x feats for parameters)
    float scale[kMlpMaxLevels];
    uint32_t resolution[kMlpMaxLevels];
    int n_levels;
};

class Mlp {
public:
    Mlp(const MlpConfig& cfg, cudaStream_t stream);
    ~Mlp();
    Mlp(const Mlp&) = delete;
    Mlp& operator=(const Mlp&) = delete;

    const MlpConfig& config() const { return cfg_; }
    size_t n_params() const { return n_params_; }
    size_t n_matrix_params() const { return n_matrix_; }
    cudaStream_t stream() const { return stream_; }

    // AoS fp32 [n][in_ch] -> [n][out_ch]; n % 128 == 0
    // d_tile_mask (optional, [n / 128] ints): 128-row tiles whose flag is 0 are skipped and their outputs
    // left untouched — for callers that know which rows' results are never read
    void inference(const float* d_in, float* d_out, int n, const int* d_tile_mask = nullptr);
    // forward + loss + backward into the fp32 gradient buffer (loss-scaled by 128).
    // n_total_records normalises the loss (global batch for data-parallel training).
    void forward_backward(const float* d_in, const float* d_target, int n, int n_total_records);
    void optimizer_step();
    float loss();                       // sum of per-element losses of the last forward (syncs)
    float* gradients() { return d_grads_; }
    void reset_weights();               // TINY_MLP::reset
    void reinitialize();
    void get_params(float* host, size_t count);
    void set_params(const float* host, size_t count);
    uint64_t launch_count() const { return launches_; }
    int step_count() const { return step_; }

    // host-side initial parameters in tcnn's order and RNG stream
    static void initial_params(const MlpConfig& cfg, std::vector<float>& out, size_t& n_matrix);
    static GridLayout grid_layout(const MlpConfig& cfg);

private:
    void ensure_train_buffers(int n);
    void sync_half_params(bool all);

    MlpConfig cfg_;
    cudaStream_t stream_;
    GridLayout layout_;
    size_t n_params_ = 0, n_matrix_ = 0;
    float* d_master_ = nullptr;     // fp32 weights
    void* d_half_ = nullptr;        // fp16 weights (same order)
    void* d_wpack_ = nullptr;       // fp16 MLP matrices re-laid-out for the tensor-core kernels
    float* d_grads_ = nullptr;      // fp32
    float* d_m1_ = nullptr; float* d_m2_ = nullptr; uint32_t* d_steps_ = nullptr;
    float* d_loss_ = nullptr;       // [1] accumulated loss
    // training activations
    int train_cap_ = 0;
    void* d_x_ = nullptr; void* d_h1_ = nullptr; void* d_h2_ = nullptr; void* d_dy_ = nullptr;
    float lr_factor_ = 1.f;
    int step_ = 0;
    uint64_t launches_ = 0;
    bool use_tc_ = true;
    bool fused_train_ = true;
};

}  // namespace hm
