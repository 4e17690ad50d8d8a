// Kernel-facing parameter blocks and launch wrappers shared between the kernel
// translation units and the C-ABI (umx_api.cu).  Device code only sees PODs.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umx {

constexpr int kMaxTaps = 25;     // 5x5
constexpr int kMaxPhases = 4;    // stride-2 conv-transpose = 4 sub-pixel phases

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// One linear term of a fused convolution: out += conv(concat(src0, src1), w, taps).
// Tap tables give, per output phase, the input offset (dy,dx) and the weight tap
// index; a plain SAME conv has one phase with k*k taps, a stride-2 conv-transpose
// has four phases with the taps of matching parity.
struct ConvTerm {
    const float* src0;
    const float* src1;          // second concat source (skip ‖ up) or nullptr
    const float* w;             // [tap][c0+c1][cout] fp32
    int32_t c0, c1;
    int32_t hy0, hy1, hx0, hx1; // halo: max(-dy), max(dy), max(-dx), max(dx) over all phases
    int8_t  ntaps[kMaxPhases];
    int8_t  dy[kMaxPhases][kMaxTaps];
    int8_t  dx[kMaxPhases][kMaxTaps];
    int8_t  wi[kMaxPhases][kMaxTaps];
};

struct ConvParams {
    ConvTerm term[2];
    int32_t nterms;
    int32_t n_tiles;            // image tiles in this launch
    int32_t in_h, in_w;         // input grid (per image tile)
    int32_t cout;
    int32_t os;                 // output stride: 1 = conv, 2 = conv-transpose
    int32_t nphase;             // 1 or 4
    int32_t ph, pw, nt;         // CTA pixel patch: nt image tiles x ph x pw = 128 pixels
    int32_t act;                // Act
    int32_t pool;               // 1 = fused 2x2 max-pool (os must be 1)
    float   leaky;
    const float* bias;          // [cout] or nullptr
    const float* post_scale;    // [cout] affine applied after the activation (legacy BN) or nullptr
    const float* post_shift;
    float* out;                 // NHWC fp32 or nullptr
    __half* out_h;              // fp16 hi[/lo] planes [out_planes][n][oh][ow][cout] for tensor-path consumers, or nullptr
    int64_t out_plane_elems;
    int32_t out_planes;
    int32_t out_cs;             // channel stride of out_h (cout rounded up to 8)
};

// First layer of the v2 graphs: 3x3 SAME conv from 1 or 2 input channels + bias + activation +
// 2x2 max-pool, straight from the normalised tile to the pooled fp16/fp32 feature map.
struct FirstParams {
    const float* src;           // [n][S][S][cin]
    const float* w;             // [9][cin][cout] fp32 (BN scale folded)
    const float* bias;          // [cout]
    int32_t n_tiles, S, cin, cout;
    int32_t ks, pool;           // kernel size (3 or 5), fused 2x2 max-pool
    int32_t act;
    float leaky;
    float* out;                 // [n][S/2][S/2][cout] fp32 or nullptr
    __half* out_h;              // fp16 hi[/lo] planes or nullptr
    int64_t out_plane_elems;
    int32_t out_planes;
    int32_t out_cs;             // channel stride of out_h (cout rounded up to 8)
    // optional second output: the k x k tap expansion of the input (see TapsParams), written from the same patch
    __half* taps_out;           // fp16 hi[/lo] planes [planes][n][S][S][taps_cs] or nullptr
    int64_t taps_plane_elems;
    int32_t taps_planes, taps_cs;
};

struct TapsParams {             // k x k tap expansion (im2col of the SAME-padded tile) of a 1-2 channel fp32 buffer
    const float* src;           // [n][S][S][cin]
    __half* out;                // fp16 hi[/lo] planes [planes][n][S][S][cs], channel = tap*cin + c, pad channels 0
    int64_t out_plane_elems;
    int32_t out_planes;
    int32_t n_tiles, S, cin, ks, cs;
};

struct TopParams {              // lt 1x1 conv (+ folded BN) + softmax over K classes
    const float* src;           // [n_pix][cin]
    const float* w;             // [cin][K]
    const float* bias;          // [K] or nullptr
    float* probs;               // [n_pix][K]
    int64_t n_pix;
    int32_t cin, k;
};

struct PreMap {                 // mirrors umx_premap
    double in_scale;
    int32_t rescale;
    double imin, imax, omin, omax;
};

struct GatherParams {           // PI2D.getPatch + (x-mean)/std for a run of tiles
    const void* img;            // [C][img_rows][W] samples, device
    int32_t dtype;              // UMX_U8/U16/F32/F64
    int32_t n_planes;           // planes present in img (1 may be broadcast to C)
    int64_t plane_stride;       // elements
    int32_t img_row0;           // image row held at img row 0 (band uploads)
    int32_t img_rows;           // rows present in the buffer
    int32_t H, W;               // full image size
    int32_t S, margin, sub, npc;
    int32_t C;                  // network input channels
    int32_t tile0, n_tiles;     // global tile index of the first tile, count
    double mean, std_dev;
    PreMap pre;
    int32_t has_pre;
    float* out;                 // [n_tiles][S][S][C]
};

struct StitchParams {           // PI2D.patchOutput/getValidOutput as a gather + quantise
    const float* probs;         // tile rows [tr_lo, tr_hi): [tr - tr_lo][npc][S][S][K]
    int32_t tr_lo, tr_hi;
    int32_t S, margin, sub, npc, npr, K;
    int32_t H, W;
    int32_t row0, row1;         // image rows [row0,row1) to emit
    uint8_t* out_u8;            // [K][*][W]; row r lands at (r - out_row_base); nullptr = skip
    float*   out_f32;
    int64_t  out_plane_stride;  // elements between class planes
    int32_t  out_row_base;
};

// launchers (kernels_simt.cu)
cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t s);
cudaError_t launch_top_softmax(const TopParams& p, cudaStream_t s);
cudaError_t launch_first_conv(const FirstParams& p, cudaStream_t s);
cudaError_t launch_taps(const TapsParams& p, cudaStream_t s);
size_t first_conv_smem_bytes(int cin, int ks, int cout);
cudaError_t launch_gather_tiles(const GatherParams& p, cudaStream_t s);
cudaError_t launch_stitch(const StitchParams& p, cudaStream_t s);
cudaError_t conv_simt_configure();   // opt in to > 48 KB dynamic shared memory
size_t conv_simt_smem_bytes(const ConvParams& p);

}  // namespace umx
