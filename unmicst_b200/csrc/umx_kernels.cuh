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
// Tile gather fused into the first layer: when lut != nullptr the kernel reads the 8/16-bit samples of the
// image itself through the normalisation look-up table (GatherParams.lut) instead of a gathered fp32 tile buffer.
struct FirstImage {
    const void* img;            // [C][img_rows][W] samples (device)
    const float* lut;           // [ch * 65537 + code]; entry n_codes = the zero padding outside the image
    int64_t plane_stride;
    int32_t dtype, n_planes;    // UMX_U8 / UMX_U16
    int32_t img_row0, H, W;
    int32_t margin, sub, npc, tile0;
};

struct FirstParams {
    FirstImage im;
    const float* src;           // [n][S][S][cin]   (unused when im.lut is set)
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
    int32_t rpb;                // 32x32 regions per block (set by the launcher)
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

// skimage.transform.resize(img, (dst_h, dst_w)) with its defaults (UnMicst1-5.py:813-815, :850), i.e.
// scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True), preceded when an axis shrinks by
// gaussian_filter(sigma=(in/out-1)/2, truncate=4, mode='mirror'); evaluated per output sample in float64.
constexpr int kMaxResampleRadius = 16;
constexpr int kMaxImagePlanes = 4;     // network input channels with their own PreMap / look-up table
struct Resample {
    int32_t on;                 // 0: identity (the source grid is the destination grid)
    int32_t src_h, src_w;       // source grid
    int32_t ry, rx;             // Gaussian radius per axis (0: no anti-aliasing on that axis)
    double zoom_y, zoom_x;      // src / dst per axis (ndimage.zoom's coordinate scale with grid_mode=True)
    double gy[2 * kMaxResampleRadius + 1], gx[2 * kMaxResampleRadius + 1];   // normalised Gaussian taps
};

struct GatherParams {           // PI2D.getPatch + (x-mean)/std for a run of tiles
    const void* img;            // [C][img_rows][W] samples, device (the SOURCE grid when rs.on)
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
    PreMap pre[kMaxImagePlanes]; // per network channel (the same map in every entry unless UMX_F_PREMAP_PER_PLANE)
    int32_t has_pre;
    Resample rs;                // rs.on: H x W is a resized view of the src_h x src_w samples in img (img_row0/img_rows/W-stride refer to the source)
    const float* lut;           // optional: normalised value per 8/16-bit sample code (no resample): out = lut[ch * 65537 + sample]
    float* out;                 // [n_tiles][S][S][C]
};

struct NormLutParams {          // builds GatherParams.lut: every possible 8/16-bit sample through the float64 map
    int32_t n;                  // 256 or 65536
    double mean, std_dev;
    PreMap pre;
    int32_t has_pre;
    float* out;
};

struct ResizeU8Params {         // uint8 page -> resize -> np.uint8(255 * x)   (UnMicst1-5.py:850-853), K planes
    const uint8_t* src;         // [K][src_rows][src_w]: rows [src_row0, src_row0 + src_rows) of the source grid
    int64_t src_plane_stride;
    int32_t src_row0, src_rows;
    int32_t K;
    int32_t dst_h, dst_w;       // destination grid
    int32_t row0, row1;         // destination rows to produce
    Resample rs;                // src grid = the probability maps at inference size
    uint8_t* out;               // [K][*][dst_w]; row r lands at r - out_row_base
    int64_t out_plane_stride;
    int32_t out_row_base;
};

struct MinMaxParams {           // min / max of img_as_float(img) resized to dst_h x dst_w (rescale_intensity's in_range)
    const void* img;
    int32_t dtype;
    int32_t dst_h, dst_w;
    double in_scale;
    Resample rs;
    unsigned long long* out;    // [2]: order-preserving encodings of min and max (see minmax_decode)
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
    int32_t  replace;           // 1: PI2D 'replace' mode (PartitionOfImage.py:99-100): the last tile written over a pixel wins, no weights
    int32_t  fp16_quant;        // 1: first quantisation as the reference evaluates it, np.uint8(255 * <float16 array>) (UnMicst1-5.py:848)
    int32_t  requant;           // 1: out_u8 = uint8(255 * (uint8(255 p) * (1/255))), the reference's second quantisation at equal size (UnMicst1-5.py:850-853)
};

// launchers (kernels_simt.cu)
cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t s);
cudaError_t launch_top_softmax(const TopParams& p, cudaStream_t s);
cudaError_t launch_first_conv(const FirstParams& p, cudaStream_t s);
cudaError_t launch_taps(const TapsParams& p, cudaStream_t s);
size_t first_conv_smem_bytes(int cin, int ks, int cout);
cudaError_t launch_gather_tiles(const GatherParams& p, cudaStream_t s);
cudaError_t launch_stitch(const StitchParams& p, cudaStream_t s);
cudaError_t launch_norm_lut(const NormLutParams& p, cudaStream_t s);
cudaError_t launch_resize_u8(const ResizeU8Params& p, cudaStream_t s);
cudaError_t launch_resample_minmax(const MinMaxParams& p, cudaStream_t s);
double minmax_decode(unsigned long long code);
// host: fill a Resample for src -> dst (returns false when the Gaussian radius exceeds kMaxResampleRadius)
bool make_resample(Resample* rs, int src_h, int src_w, int dst_h, int dst_w);
cudaError_t conv_simt_configure();   // opt in to > 48 KB dynamic shared memory
size_t conv_simt_smem_bytes(const ConvParams& p);

}  // namespace umx
