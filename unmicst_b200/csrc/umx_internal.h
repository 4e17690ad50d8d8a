// Host-side structures of the engine (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/unmicst_b200.h"
#include "umx_kernels.cuh"

namespace umx {

void set_error(const char* fmt, ...);

#define UMX_CUDA_TRY(expr)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ::umx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return UMX_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

#define UMX_TRY(expr)              \
    do {                           \
        int r__ = (expr);          \
        if (r__ != UMX_OK) return r__; \
    } while (0)

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
};

// A per-tile activation buffer living in the workspace (fp32 NHWC).
struct Buffer {
    std::string name;
    int h = 0, w = 0, c = 0;
    float* d = nullptr;             // [max_batch][h][w][c]
    int64_t per_tile() const { return (int64_t)h * w * c; }
};

enum OpKind { OP_CONV = 0, OP_TOP = 1 };

struct Op {
    OpKind kind = OP_CONV;
    std::string name;
    // OP_CONV (conv and conv-transpose, fp32 CUDA-core implementation)
    ConvParams cp{};                 // device pointers filled at plan time; n_tiles patched per launch
    // OP_TOP
    TopParams tp{};
    int top_src = -1;
    int out_buf = -1;
    double flops_per_tile = 0;       // algorithmic (2*MAC)
    double bytes_per_tile = 0;       // activations read + written once, fp32
    double weight_bytes = 0;         // per launch
    int prof_slot = -1;
};

struct ProfSlot {
    std::string name;
    int64_t launches = 0;
    double ms = 0, flops = 0, bytes = 0;
};

struct PendingEvent {
    int slot;
    cudaEvent_t a, b;
    double flops, bytes;
};

}  // namespace umx

struct umx_handle {
    int device = 0;
    umx_model_desc desc{};
    int S = 0, C = 0, K = 0, L = 0, margin = 0, sub = 0;
    int max_batch = 0;
    std::vector<int> chan;                       // nOutX
    std::map<std::string, umx::HostTensor> tensors;
    std::vector<float*> dev_allocs;              // weights etc.
    std::vector<umx::Buffer> bufs;
    std::vector<umx::Op> ops;
    int in_buf = -1;                             // network input buffer
    float* probs = nullptr;                      // [max_batch][S][S][K] (forward_tiles)
    // image path
    void* d_img = nullptr; size_t d_img_bytes = 0;
    float* d_probs_rows = nullptr; size_t d_probs_rows_bytes = 0;
    uint8_t* d_stage_u8[2] = {nullptr, nullptr}; size_t d_stage_u8_bytes = 0;
    float* d_stage_f32[2] = {nullptr, nullptr}; size_t d_stage_f32_bytes = 0;
    cudaEvent_t ev_stitch[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    // profiling
    bool profiling = false;
    std::vector<umx::ProfSlot> prof;
    std::vector<umx::PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    int64_t launches = 0;
};
