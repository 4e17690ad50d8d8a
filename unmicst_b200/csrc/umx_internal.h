// Host-side structures of the engine (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/unmicst_b200.h"
#include "umx_kernels.cuh"
#include "umx_tc.cuh"

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

// A per-tile activation buffer living in the workspace (NHWC); fp32 and/or fp16 hi[/lo] planes,
// depending on which kernels consume it.
struct Buffer {
    std::string name;
    int h = 0, w = 0, c = 0;
    bool need_f = false, need_h = false, need_lo = false;     // need_lo: some consumer multiplies with the hi/lo split
    float* d = nullptr;             // [cap][h][w][c] fp32
    __half* dh = nullptr;           // [planes][cap][h][w][c] fp16
    int planes = 0;
    int64_t plane_elems = 0;
    int cs() const { return (c + 7) & ~7; }          // channel stride of the fp16 planes (TMA needs 16-byte strides)
    int64_t per_tile() const { return (int64_t)h * w * c; }
    int64_t per_tile_h() const { return (int64_t)h * w * cs(); }
};

struct TermHost {
    int src0 = -1, src1 = -1;         // buffer ids (src1: second concat source)
    std::vector<float> w;             // fp32 [tap][c0+c1][cout], BN scale folded where the graph allows
    int k = 3;
};

struct ConvSpec {
    std::vector<TermHost> terms;
    int cout = 0;
    bool transpose = false;
    bool has_bias = false, has_post = false, pool = false;
    std::vector<float> bias, post_scale, post_shift;
    int act = 0;
};

enum OpKind { OP_CONV = 0, OP_TOP = 1, OP_TAPS = 2 };

struct Op {
    OpKind kind = OP_CONV;
    std::string name;
    ConvSpec spec;
    bool use_tc = false;
    bool use_first = false;          // dedicated first-layer kernel (1-2 input channels, 3x3, pooled)
    FirstParams fp{};
    // how the tensor path maps the op: 1 plain (one or two wide concat sources), 3 two terms with a one-channel 1x1
    // shortcut in the epilogue, 4 two terms with a wide 1x1 term joining the K loop at the centre tap
    int tc_mode = 0;
    int fuse_top = -1;               // index of the OP_TOP fused into this conv's epilogue
    bool fused_away = false;         // OP_TOP executed inside the preceding conv
    ConvParams cp{};                 // fp32 CUDA-core implementation
    TcConvParams tcp{};              // tcgen05 implementation
    alignas(64) CUtensorMap mapA0, mapA1, mapB, mapB1;
    TopParams tp{};
    TapsParams taps{};               // OP_TAPS: k x k tap expansion of a narrow buffer
    int taps_src = -1, taps_k = 0;
    std::vector<float> top_w, top_b;
    int top_src = -1;
    int out_buf = -1;
    double flops_per_tile = 0;       // algorithmic (2*MAC)
    double bytes_per_tile = 0;       // activations read + written once at 4 B/element
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
    int cap_tiles = 0;                           // workspace capacity in tiles (max_batch rounded up to 8)
    int num_sms = 148;
    int precision = UMX_PREC_SPLIT3;
    uint64_t single_mask = 0;                    // UMX_PREC_MIXED: ops (by index) that run with one MMA per product
    bool fuse_gather = false;                    // the first layer is the only reader of the gathered tiles: it can read the image itself
    bool full_split = false;                     // every tensor-path op carries both planes of everything (umx_set_op_terms allowed)
    std::vector<int8_t> op_terms;                // umx_create_ex: per op, t0 | t1 << 2 (-1: what the precision implies)
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
    uint8_t* d_band_u8 = nullptr; size_t d_band_u8_bytes = 0;     // UMX_F_CLI_QUANT with resizing: the band's maps at inference size
    uint8_t* d_out_u8 = nullptr; size_t d_out_u8_bytes = 0;       // ... and the resized pages before the D2H copy
    int carry_row = -1; int carry_h = 0, carry_w = 0;             // tile row whose probabilities sit in slot 0 of d_probs_rows (UMX_F_CONTINUE)
    float* d_lut = nullptr;                      // 65537-entry normalisation table for integer samples
    unsigned long long* d_minmax = nullptr;      // umx_resample_minmax result
    unsigned long long* d_dbg = nullptr;         // UMX_TC_EXP=64 cycle counters
    // profiling
    bool profiling = false;
    std::vector<umx::ProfSlot> prof;
    std::vector<umx::PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    int64_t launches = 0;
};
