// hm_comm.h — multi-GPU plumbing of the rendering loop: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference is single-GPU (render_hair_msnn.cu:1056, owlContextCreate(nullptr, 1)); this is the new
// work of SURVEY §8(e): pixels and samples are independent, so ranks shard samples (spp groups) and/or row
// bands; the only exchange steps are (1) the per-step all-reduce of the network's gradients between backward
// and Adam, so every replica holds bit-identical weights, and (2) one reduction of the framebuffers per output.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace hm {

constexpr size_t kCommIdBytes = 128;   // sizeof(ncclUniqueId)

class Comm {
public:
    // fills out128 with a fresh NCCL unique id (rank 0 calls this and hands the bytes to the other ranks)
    static void unique_id(void* out128);
    Comm(const void* id128, int rank, int world, int device);
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;

    int rank() const { return rank_; }
    int world() const { return world_; }
    int device() const { return device_; }

    // in-place sum over all ranks, enqueued on `s` (every rank must enqueue the same sequence of collectives)
    void all_reduce_sum(float* d_buf, size_t count, cudaStream_t s);
    // host-side helpers on the communicator's own stream (synchronise before returning)
    void barrier();
    double all_reduce_max(double v);
    double all_reduce_sum(double v);

private:
    void* comm_ = nullptr;   // ncclComm_t
    int rank_, world_, device_;
    cudaStream_t stream_ = nullptr;
    double* d_scalar_ = nullptr;
};

}  // namespace hm
