// temporary stub
#include "hm_mlp.h"
#include <stdexcept>
namespace hm {
MlpConfig mlp_config_from_json(const std::string&, int in_ch, int out_ch) { MlpConfig c; c.in_ch=in_ch; c.out_ch=out_ch; return c; }
Mlp::Mlp(const MlpConfig& cfg, cudaStream_t s) : cfg_(cfg), stream_(s) {}
Mlp::~Mlp() {}
void Mlp::inference(const float*, float*, int) { throw std::logic_error("mlp stub"); }
void Mlp::forward_backward(const float*, const float*, int, int) { throw std::logic_error("mlp stub"); }
void Mlp::optimizer_step() {}
float Mlp::loss() { return 0.f; }
void Mlp::reset_weights() {}
void Mlp::reinitialize() {}
void Mlp::get_params(float*, size_t) {}
void Mlp::set_params(const float*, size_t) {}
}
