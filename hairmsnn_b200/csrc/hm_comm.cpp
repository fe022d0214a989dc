// hm_comm.cpp — see hm_comm.h.
#include "hm_comm.h"

#include <nccl.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace hm {

static_assert(sizeof(ncclUniqueId) == kCommIdBytes, "ncclUniqueId is 128 bytes");

#define HM_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess)                                                                          \
            throw std::runtime_error(std::string("NCCL: ") + ncclGetErrorString(r_) + " at " + __FILE__ + \
                                     ":" + std::to_string(__LINE__));                                   \
    } while (0)
#define HM_CUDA_C(call)                                                                                 \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " + __FILE__ + \
                                     ":" + std::to_string(__LINE__));                                   \
    } while (0)

void Comm::unique_id(void* out128) {
    ncclUniqueId id;
    HM_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
}

Comm::Comm(const void* id128, int rank, int world, int device) : rank_(rank), world_(world), device_(device) {
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad rank/world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("CUDA: no usable device (this library has no CPU path)");
    if (device < 0 || device >= ndev) throw std::runtime_error("CUDA: device index out of range");
    HM_CUDA_C(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    HM_NCCL(ncclCommInitRank(&c, world, id, rank));
    comm_ = c;
    HM_CUDA_C(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    HM_CUDA_C(cudaMalloc((void**)&d_scalar_, sizeof(double)));
}

Comm::~Comm() {
    cudaSetDevice(device_);
    if (d_scalar_) cudaFree(d_scalar_);
    if (comm_) ncclCommDestroy((ncclComm_t)comm_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Comm::all_reduce_sum(float* d_buf, size_t count, cudaStream_t s) {
    HM_NCCL(ncclAllReduce(d_buf, d_buf, count, ncclFloat, ncclSum, (ncclComm_t)comm_, s));
}

static double reduce_scalar(void* comm, cudaStream_t s, double* d, double v, ncclRedOp_t op, int device) {
    HM_CUDA_C(cudaSetDevice(device));
    HM_CUDA_C(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, s));
    HM_NCCL(ncclAllReduce(d, d, 1, ncclDouble, op, (ncclComm_t)comm, s));
    double out = 0;
    HM_CUDA_C(cudaMemcpyAsync(&out, d, sizeof(double), cudaMemcpyDeviceToHost, s));
    HM_CUDA_C(cudaStreamSynchronize(s));
    return out;
}

void Comm::barrier() { (void)reduce_scalar(comm_, stream_, d_scalar_, 0.0, ncclSum, device_); }
double Comm::all_reduce_max(double v) { return reduce_scalar(comm_, stream_, d_scalar_, v, ncclMax, device_); }
double Comm::all_reduce_sum(double v) { return reduce_scalar(comm_, stream_, d_scalar_, v, ncclSum, device_); }

}  // namespace hm
