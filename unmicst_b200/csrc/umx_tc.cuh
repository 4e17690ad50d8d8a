// Parameter block of the tcgen05 implicit-GEMM convolution (kernels_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umx {

// The taps of one output phase form an ny x nx grid with constant steps, so the kernels walk them with adds instead
// of table look-ups: tap (iy, ix) reads the input at (dy0 + dstep*iy, dx0 + dstep*ix) and the weight tap
// wi0 + iy*wiy + ix*wix.  conv k x k: one phase, dstep +1; stride-2 conv-transpose: four sub-pixel phases, dstep -1.
constexpr int kTcMaxCols = 1280;     // widest layer the tensor path takes (columns = n_ntiles * n_t)

struct TcPhaseGrid { int32_t ntaps, nx, dy0, dx0, dstep, wi0, wiy, wix; };

struct TcConvParams {
    int32_t n_tiles;             // image tiles in this launch
    int32_t in_h, in_w;          // input grid per image tile
    int32_t bw, bh, bn;          // TMA box over (w, h, tile): bn*bh*bw = 128 GEMM rows
    int32_t c0, c1;              // channels of the two concat sources (c1 = 0: single source)
    int32_t cout;
    int32_t n_t;                 // GEMM N per CTA tile (multiple of 16, <= 256)
    int32_t n_ntiles;            // ceil(cout / n_t)
    int32_t nphase, os;          // conv: 1,1   conv-transpose: 4,2
    int32_t merge_px;            // conv-transpose in halo mode: a work item covers both px phases of a row parity (shared patches, two n_t-column
                                 // accumulator halves): half the patch loads, barrier round trips and epilogue hand-overs
    TcPhaseGrid grid[4];         // per phase tap grid (conv 3x3/5x5: 1 phase; conv-transpose: 4 phases, <= 9 taps)
    int32_t planes;              // 1: fp16 operands, 1 MMA/product; 2: hi/lo split, up to 3 MMAs/product
    int32_t planes_a, planes_b;  // planes a slot holds (layout): A hi [+ lo], B hi [+ lo]; fixed when the op is lowered
    int32_t terms0, terms1;      // split layers, per concat source: bit 0 = a_hi*w_lo correction, bit 1 = a_lo*w_hi correction (3 = full split)
    int32_t stages;              // smem pipeline depth (halo mode: patch slots)
    int32_t halo;                // 1: one (bh+halo) x (bw+halo) pixel patch per 64-channel slab serves every tap
    int32_t pw, ph, hx0, hy0;    // patch size and left/top halo (halo mode)
    int32_t halo_nh;             // halo mode on 8x8 grids: the patch holds bn tiles row-interleaved, [h][tile][w], so that
                                 // consecutive 8-pixel row groups stay one patch row apart (uniform SBO); GEMM row m = (y*bn + tile)*bw + x
    int32_t b_stages, gb;        // halo mode: weight ring depth, taps per weight slot
    int32_t b_res_bytes;         // bytes of the resident weight region (per CTA)
    int32_t b1_hi_only;          // streamed (non-resident) halo mode with hi + lo weights: mapB1 is the same box with the hi plane only,
                                 // used for the slabs of a source that has no a_hi*w_lo term
    int32_t ncat;                // resident mode, CTA pairs: a_hi x [w_hi | w_lo] as one MMA of width 2*n_t, halves added in the epilogue
    int32_t res_m_planes;        // resident mode: weight tiles per tap of a full slab: 1 = hi, 2 = hi + lo ([plane][tap]), 3 = ncat ([tap][X, X, Y])
    int32_t res_c_planes;        // resident mode: weight planes a centre-only (1x1 term) slab keeps: 1 = hi, 2 = hi + lo (its a_hi*w_lo term is on)
    int32_t b_resident;          // halo mode: every (slab, tap) weight tile of the layer stays in shared memory (one slot of gb = all taps per slab)
    int32_t kslab;               // plain mode: 64-channel slabs per ring slot (more MMAs per barrier round trip)
    uint32_t epi_nap_ns;         // epilogue warps nap this long between polls of the accumulator-ready barrier (0: hinted try_wait only)
    int32_t exp_flags;           // timing experiments only (results invalid): 1 no TMA loads, 4 one MMA per slab, 8 first epilogue chunk only, 16 no global stores, 32 no epilogue work
    unsigned long long* dbg;     // exp_flags & 64: 16 cycle counters (producer / MMA / epilogue waits and work), else nullptr
    int32_t pair;                // 1: CTA pairs, tcgen05.mma.cta_group::2 (M = 256 across two SMs)
    int32_t act;                 // umx::Act
    float   leaky;
    int32_t pool;                // fused 2x2 max-pool (conv only)
    int32_t a1_center;           // 1: the second source is a 1x1 shortcut: its slabs join the K loop at the centre tap only
    int32_t center_tap;          // linear index (= device weight tap) of the (0,0) tap of a k x k conv
    const float* post_scale;     // [cout] affine applied after the activation (legacy: batch-norm follows the ReLU) or nullptr
    const float* post_shift;
    __half* out_h;               // fp16 plane(s) [planes][n][oh][ow][cout] or nullptr
    int64_t out_plane_elems;     // elements between the hi and lo plane
    int32_t out_planes;          // planes to write into out_h (1 or 2)
    int32_t out_cs;              // channel stride of out_h (cout rounded up to 8; pad channels are written as 0)
    float*  out_f;               // fp32 [n][oh][ow][cout] or nullptr
    // one-channel fp32 source entering as a 1x1 shortcut in the epilogue (legacy down layers)
    const float* skip_src;       // [n][in_h][in_w] or nullptr
    const float* skip_w;         // [cout] fp32
    int32_t skip_c;              // 1 when skip_src is set
    // fused lt 1x1 conv + softmax (replaces the activation store when top_w != nullptr)
    const float* top_w;          // non-null marks the fusion (values live in tab_topw / tab_topb)
    float* top_probs;            // [n][oh][ow][K]
    int32_t top_k;
    // Epilogue tables in the kernel-parameter constant bank (padded with zeros to n_ntiles * n_t entries): warp-uniform
    // operands come from the constant cache instead of shared memory, whose bandwidth belongs to the tensor core
    alignas(16) float tab_bias[kTcMaxCols];  // bias (0 when the layer has none)
    alignas(16) float tab_topw[256 * 4];     // fused lt weights [channel][4]
    float tab_topb[4];
};

size_t tc_conv_smem_bytes(const TcConvParams& p);
size_t tc_conv_a_bytes(const TcConvParams& p);
size_t tc_conv_b_bytes(const TcConvParams& p);
size_t tc_conv_fixed_bytes(const TcConvParams& p);
cudaError_t tc_conv_configure();
// b1: the weight map with a one-tap box (resident mode, centre-only slabs); equal to b when unused
cudaError_t launch_tc_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b1,
                           const TcConvParams& p, int num_sms, cudaStream_t s);

// Host helpers: build TMA descriptors through the driver entry point (no libcuda link dependency).
// activations: fp16 [planes][n][h][w][c]; box = {64 ch, bw, bh, bn, planes}
// nh_order: dimension order {C, W, tiles, H, plane} (box rows interleave the tiles) instead of {C, W, H, tiles, plane}
int make_act_tensor_map(CUtensorMap* out, const __half* base, int planes, int64_t plane_elems, int n, int h, int w, int c,
                        int bw, int bh, int bn, int box_planes, int nh_order = 0);
// weights: fp16 [planes][tap][cout][cin]; box = {64 cin, n_t, box_taps, planes}
int make_weight_tensor_map(CUtensorMap* out, const __half* base, int planes, int taps, int cout, int cin, int n_t,
                           int box_planes, int box_taps);

}  // namespace umx
