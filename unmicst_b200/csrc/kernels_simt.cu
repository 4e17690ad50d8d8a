// fp32 CUDA-core kernels of the UnMicst probability-map path (sm_100a).
//
//  conv_simt_kernel   tf.nn.conv2d SAME / tf.nn.conv2d_transpose s2 SAME, NHWC fp32, with the
//                     channel concat (UnMicst1-5.py:196), shortcut add (:106-114), bias /
//                     batch-norm affine, ReLU / leaky-ReLU and 2x2 max-pool (:117) fused.
//                     Exact reference arithmetic (fp32 FMA); used for the narrow / first
//                     layers and as the UMX_PREC_FP32 path of every layer.
//  top_softmax_kernel lt 1x1 conv + folded BN + softmax (UnMicst1-5.py:212-237)
//  gather_tiles       PI2D.getPatch + (patch-mean)/std (PartitionOfImage.py:77-82, UnMicst1-5.py:700)
//  stitch_kernel      PI2D.patchOutput/getValidOutput as an atomic-free gather
//                     (PartitionOfImage.py:92-122) + np.uint8(255*p) (UnMicst1-5.py:848)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "umx_kernels.cuh"
#include "../../include/unmicst_b200.h"

namespace umx {

namespace {

constexpr int CK = 8;        // input channels staged per step
constexpr int COT = 64;      // output channels per CTA
constexpr int NTHR = 256;    // 16 pixel groups x 16 channel groups
constexpr int PIX_PER_CTA = 128;

__device__ __forceinline__ float apply_act(float v, int act, float leaky) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LEAKY) return v > 0.f ? v : v * leaky;
    return v;
}

// 4 consecutive channels -> fp16 hi (and lo = fp16(v - hi)) planes
__device__ __forceinline__ void store_h2x4(__half* o, int64_t plane_elems, int planes, const float* v) {
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = __float2half_rn(v[j]); l[j] = __float2half_rn(v[j] - __half2float(h[j])); }
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(h);
    if (planes == 2) *reinterpret_cast<uint2*>(o + plane_elems) = *reinterpret_cast<uint2*>(l);
}

// 16 consecutive channels (32-byte aligned destination) -> one 256-bit store per plane
__device__ __forceinline__ void store_h16(__half* o, int64_t plane_elems, int planes, const float* v) {
    __half2 hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hi[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const uint32_t* h = reinterpret_cast<const uint32_t*>(hi);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]),
                 "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
    if (planes == 2) {
        __half2 lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f = __half22float2(hi[j]);
            lo[j] = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
        }
        const uint32_t* l = reinterpret_cast<const uint32_t*>(lo);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + plane_elems), "r"(l[0]), "r"(l[1]), "r"(l[2]),
                     "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
    }
}

// 8 consecutive channels: one 16-byte store per plane
__device__ __forceinline__ void store_h8(__half* o, int64_t plane_elems, int planes, const float* v) {
    __half2 hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hi[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
    if (planes == 2) {
        __half2 lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(hi[j]);
            lo[j] = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
        }
        *reinterpret_cast<uint4*>(o + plane_elems) = *reinterpret_cast<const uint4*>(lo);
    }
}

// Each CTA: 128 output pixels (nt image tiles x ph x pw) x 64 output channels of one phase.
// Each thread: a 2x4 pixel patch x 4 output channels (32 fp32 accumulators).
__global__ void __launch_bounds__(NTHR) conv_simt_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int cg = tid & 15;
    const int pg = tid >> 4;
    const int phase = blockIdx.z;
    const int co0 = blockIdx.y * COT;

    int n0, y0, x0;
    if (p.nt > 1) {
        n0 = blockIdx.x * p.nt; y0 = 0; x0 = 0;
    } else {
        const int bx = p.in_w / p.pw, by = p.in_h / p.ph;
        n0 = blockIdx.x / (bx * by);
        const int r = blockIdx.x % (bx * by);
        y0 = (r / bx) * p.ph; x0 = (r % bx) * p.pw;
    }
    const int gpt = (p.ph >> 1) * (p.pw >> 2);      // thread groups per image tile
    const int nt_l = pg / gpt;
    const int rem = pg % gpt;
    const int py0 = (rem / (p.pw >> 2)) * 2;
    const int px0 = (rem % (p.pw >> 2)) * 4;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < p.nterms; ++t) {
        const ConvTerm& T = p.term[t];
        const int ntap = T.ntaps[phase];
        if (ntap == 0) continue;
        const int prow = p.ph + T.hy0 + T.hy1;
        const int pcol = p.pw + T.hx0 + T.hx1;
        const int npix = p.nt * prow * pcol;
        float* xs = smem;
        float* ws = smem + ((npix * CK + 3) & ~3);
        const int ctot = T.c0 + T.c1;
        const int nch0 = (T.c0 + CK - 1) / CK;
        const int nch1 = (T.c1 + CK - 1) / CK;
        for (int cb = 0; cb < nch0 + nch1; ++cb) {
            const float* src; int C, cbase, wrow;
            if (cb < nch0) { src = T.src0; C = T.c0; cbase = cb * CK; wrow = cbase; }
            else { src = T.src1; C = T.c1; cbase = (cb - nch0) * CK; wrow = T.c0 + cbase; }
            const int cvalid = min(CK, C - cbase);
            __syncthreads();
            // ---- stage the input patch (zero outside the image tile: SAME padding per tile)
            const bool vec = ((C & 3) == 0);
            for (int idx = tid; idx < npix * 2; idx += NTHR) {
                const int pix = idx >> 1, half = idx & 1;
                const int col = pix % pcol;
                const int rr = pix / pcol;
                const int row = rr % prow;
                const int nl = rr / prow;
                const int gy = y0 + row - T.hy0, gx = x0 + col - T.hx0, gn = n0 + nl;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gn < p.n_tiles && gy >= 0 && gy < p.in_h && gx >= 0 && gx < p.in_w) {
                    const float* g = src + (((int64_t)gn * p.in_h + gy) * p.in_w + gx) * C + cbase + half * 4;
                    const int rem_c = cvalid - half * 4;
                    if (vec && rem_c >= 4) {
                        v = __ldg(reinterpret_cast<const float4*>(g));
                    } else {
                        if (rem_c > 0) v.x = __ldg(g);
                        if (rem_c > 1) v.y = __ldg(g + 1);
                        if (rem_c > 2) v.z = __ldg(g + 2);
                        if (rem_c > 3) v.w = __ldg(g + 3);
                    }
                }
                *reinterpret_cast<float4*>(xs + pix * CK + half * 4) = v;
            }
            // ---- stage the weights [tap][ci][64 co]
            const bool wvec = ((p.cout & 3) == 0);
            for (int idx = tid; idx < ntap * CK * (COT / 4); idx += NTHR) {
                const int c4 = idx & 15;
                const int ci = (idx >> 4) & (CK - 1);
                const int tap = idx >> 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int co = co0 + c4 * 4;
                if (ci < cvalid && co < p.cout) {
                    const float* g = T.w + ((int64_t)T.wi[phase][tap] * ctot + wrow + ci) * p.cout + co;
                    if (wvec) {
                        v = __ldg(reinterpret_cast<const float4*>(g));
                    } else {
                        v.x = __ldg(g);
                        if (co + 1 < p.cout) v.y = __ldg(g + 1);
                        if (co + 2 < p.cout) v.z = __ldg(g + 2);
                        if (co + 3 < p.cout) v.w = __ldg(g + 3);
                    }
                }
                *reinterpret_cast<float4*>(ws + (tap * CK + ci) * COT + c4 * 4) = v;
            }
            __syncthreads();
            // ---- accumulate
            for (int tap = 0; tap < ntap; ++tap) {
                const int dy = T.dy[phase][tap] + T.hy0;
                const int dx = T.dx[phase][tap] + T.hx0;
                const float* xb = xs + ((nt_l * prow + py0 + dy) * pcol + px0 + dx) * CK;
                const float* wb = ws + tap * CK * COT + cg * 4;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 wv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(wb + (h * 4 + j) * COT);
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) {
                        const float4 a = *reinterpret_cast<const float4*>(xb + ((pp >> 2) * pcol + (pp & 3)) * CK + h * 4);
                        acc[pp][0] = fmaf(a.x, wv[0].x, acc[pp][0]); acc[pp][1] = fmaf(a.x, wv[0].y, acc[pp][1]);
                        acc[pp][2] = fmaf(a.x, wv[0].z, acc[pp][2]); acc[pp][3] = fmaf(a.x, wv[0].w, acc[pp][3]);
                        acc[pp][0] = fmaf(a.y, wv[1].x, acc[pp][0]); acc[pp][1] = fmaf(a.y, wv[1].y, acc[pp][1]);
                        acc[pp][2] = fmaf(a.y, wv[1].z, acc[pp][2]); acc[pp][3] = fmaf(a.y, wv[1].w, acc[pp][3]);
                        acc[pp][0] = fmaf(a.z, wv[2].x, acc[pp][0]); acc[pp][1] = fmaf(a.z, wv[2].y, acc[pp][1]);
                        acc[pp][2] = fmaf(a.z, wv[2].z, acc[pp][2]); acc[pp][3] = fmaf(a.z, wv[2].w, acc[pp][3]);
                        acc[pp][0] = fmaf(a.w, wv[3].x, acc[pp][0]); acc[pp][1] = fmaf(a.w, wv[3].y, acc[pp][1]);
                        acc[pp][2] = fmaf(a.w, wv[3].z, acc[pp][2]); acc[pp][3] = fmaf(a.w, wv[3].w, acc[pp][3]);
                    }
                }
            }
        }
    }

    // ---- epilogue: bias -> activation -> post affine -> (pool) -> store
    const int gn = n0 + nt_l;
    const int co = co0 + cg * 4;
    if (gn >= p.n_tiles || co >= p.cout) return;
    float b[4] = {0.f, 0.f, 0.f, 0.f}, ps[4] = {1.f, 1.f, 1.f, 1.f}, pt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (co + j < p.cout) {
            if (p.bias) b[j] = __ldg(p.bias + co + j);
            if (p.post_scale) { ps[j] = __ldg(p.post_scale + co + j); pt[j] = __ldg(p.post_shift + co + j); }
        }
    }
#pragma unroll
    for (int pp = 0; pp < 8; ++pp)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = apply_act(acc[pp][j] + b[j], p.act, p.leaky);
            if (p.post_scale) v = fmaf(v, ps[j], pt[j]);
            acc[pp][j] = v;
        }
    const bool ovec = ((p.cout & 3) == 0);
    if (p.pool) {
        const int oh = p.in_h >> 1, ow = p.in_w >> 1;
        const int oy = (y0 + py0) >> 1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int ox = ((x0 + px0) >> 1) + q;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = fmaxf(fmaxf(acc[2 * q][j], acc[2 * q + 1][j]), fmaxf(acc[4 + 2 * q][j], acc[4 + 2 * q + 1][j]));
            const int64_t oo = (((int64_t)gn * oh + oy) * ow + ox) * p.cout + co;
            if (p.out) {
                float* o = p.out + oo;
                if (ovec) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                else
                    for (int j = 0; j < 4; ++j) if (co + j < p.cout) o[j] = v[j];
            }
            if (p.out_h) store_h2x4(p.out_h + (((int64_t)gn * oh + oy) * ow + ox) * p.out_cs + co, p.out_plane_elems, p.out_planes, v);
        }
    } else {
        const int oh = p.in_h * p.os, ow = p.in_w * p.os;
        const int phy = (p.os == 2) ? (phase >> 1) : 0, phx = (p.os == 2) ? (phase & 1) : 0;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
            const int oy = (y0 + py0 + (pp >> 2)) * p.os + phy;
            const int ox = (x0 + px0 + (pp & 3)) * p.os + phx;
            const int64_t oo = (((int64_t)gn * oh + oy) * ow + ox) * p.cout + co;
            if (p.out) {
                float* o = p.out + oo;
                if (ovec) *reinterpret_cast<float4*>(o) = make_float4(acc[pp][0], acc[pp][1], acc[pp][2], acc[pp][3]);
                else
                    for (int j = 0; j < 4; ++j) if (co + j < p.cout) o[j] = acc[pp][j];
            }
            if (p.out_h) store_h2x4(p.out_h + (((int64_t)gn * oh + oy) * ow + ox) * p.out_cs + co, p.out_plane_elems, p.out_planes, acc[pp]);
        }
    }
}

// First layer: k x k SAME conv from 1 or 2 input channels (+ bias, activation, optional 2x2 max-pool).
// Block: p.rpb consecutive 32x32-pixel regions (of one tile when rpb divides the regions per tile); thread: one 2x2 pixel
// block, all output channels in groups of 16.  Weights are read from shared memory as warp-wide broadcasts (padded to a
// multiple of 16 channels with zeros) and staged once per block.  The input patch is double buffered: the samples of
// the next region are requested before the FMAs of the current one, looked up (fused gather) after its first channel
// group and stored at its end, so the two dependent global-load latencies of the fused gather hide behind arithmetic.
// FAST = the common configuration resolved at compile time (2x2 max-pool, leaky ReLU, fp16 planes out, no fp32 copy):
// the per-group flag tests of the general path are a fifth of its instructions.
template <int CIN, int KS, bool FAST>
__global__ void __launch_bounds__(256, (KS == 3 && (FAST || CIN == 1)) ? 2 : 0) first_conv_kernel(const FirstParams p) {
    constexpr int R = KS / 2, PW = 32 + 2 * R, NB = 2 + 2 * R;
    constexpr bool DUP = NB * NB * CIN <= 16;            // keep the patch as (x, x) pairs: FFMA2 operands without moves
    const bool pool = FAST ? true : p.pool != 0;
    const int act = FAST ? (int)ACT_LEAKY : p.act;
    const bool wide_out = (p.out_cs & 15) == 0 && (p.out_plane_elems & 15) == 0;
    const bool wide_taps = (p.taps_cs & 15) == 0 && (p.taps_plane_elems & 15) == 0;
    constexpr int NL = (PW * PW + 255) / 256;            // patch pixels per thread
    constexpr int XN = (PW * PW * CIN + 3) & ~3;
    extern __shared__ __align__(16) float smem[];
    const int cpad = (p.cout + 15) & ~15;
    float* xin0 = smem;                                  // 2 x [PW][PW][CIN]
    float* ws = smem + 2 * XN;                           // [KS*KS][CIN][cpad]
    float* bs = ws + KS * KS * CIN * cpad;               // [cpad]
    const int nb = p.S / 32;
    const bool fused = p.im.lut != nullptr;
    const int n_codes = p.im.dtype == UMX_U8 ? 256 : 65536;

    // stage 1 of a patch load: the raw sample (fused gather: its integer code; else the fp32 value) of each patch pixel.
    // code -1 = outside the tile (the conv's zero padding); n_codes = inside the tile, outside the image (PI2D's frame).
    int32_t raw[NL * CIN];
    auto fetch = [&](int region) {
        const int n = region / (nb * nb), rb = region % (nb * nb);
        const int y0 = (rb / nb) * 32, x0 = (rb % nb) * 32;
        // PI2D.getPatch + (x - mean)/std straight from the image: tile t = (ti, tj) covers frame rows ti*sub .. +S, the
        // frame is the image at offset (margin, margin) and 0 (-> lut[n_codes]) elsewhere
        const int t = p.im.tile0 + n, ti = fused ? t / p.im.npc : 0, tj = fused ? t - ti * p.im.npc : 0;
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int i = threadIdx.x + l * 256;
            const int r = i / PW, c = i % PW;
            const int gy = y0 + r - R, gx = x0 + c - R;
            const bool inb = i < PW * PW && gy >= 0 && gy < p.S && gx >= 0 && gx < p.S;
            if (fused) {
                const int ir = ti * p.im.sub + gy - p.im.margin, ic = tj * p.im.sub + gx - p.im.margin;
                const bool inside = inb && ir >= 0 && ir < p.im.H && ic >= 0 && ic < p.im.W;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    int code = inb ? n_codes : -1;
                    if (inside) {
                        const int64_t idx = (int64_t)(p.im.n_planes == 1 ? 0 : ci) * p.im.plane_stride + (int64_t)(ir - p.im.img_row0) * p.im.W + ic;
                        code = p.im.dtype == UMX_U8 ? (int)__ldg(reinterpret_cast<const uint8_t*>(p.im.img) + idx)
                                                    : (int)__ldg(reinterpret_cast<const uint16_t*>(p.im.img) + idx);
                    }
                    raw[l * CIN + ci] = code;
                }
            } else {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
                    raw[l * CIN + ci] = inb ? __float_as_int(__ldg(p.src + (((int64_t)n * p.S + gy) * p.S + gx) * CIN + ci)) : 0;
            }
        }
    };
    // stage 2: code -> normalised fp32 through the look-up table (in place)
    auto lookup = [&]() {
        if (!fused) return;
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const int code = raw[l * CIN + ci];
                raw[l * CIN + ci] = code >= 0 ? __float_as_int(__ldg(p.im.lut + ci * 65537 + code)) : 0;
            }
    };
    auto stash = [&](float* xin) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int i = threadIdx.x + l * 256;
            if (i < PW * PW) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) xin[i * CIN + ci] = __int_as_float(raw[l * CIN + ci]);
            }
        }
    };

    const int region0 = blockIdx.x * p.rpb;
    fetch(region0);
    for (int i = threadIdx.x; i < KS * KS * CIN * cpad; i += 256) {
        const int c = i % cpad;
        ws[i] = c < p.cout ? __ldg(p.w + (size_t)(i / cpad) * p.cout + c) : 0.f;
    }
    for (int i = threadIdx.x; i < cpad; i += 256) bs[i] = (p.bias && i < p.cout) ? __ldg(p.bias + i) : 0.f;
    lookup();
    stash(xin0);
    __syncthreads();
    const int py = threadIdx.x >> 4, px = threadIdx.x & 15;
    const int oh = pool ? p.S / 2 : p.S;
    for (int rr = 0; rr < p.rpb; ++rr) {
        const int region = region0 + rr;
        const int n = region / (nb * nb), rb = region % (nb * nb);
        const int y0 = (rb / nb) * 32, x0 = (rb % nb) * 32;
        const float* xin = xin0 + (rr & 1) * XN;
        const bool more = rr + 1 < p.rpb;
        if (more) fetch(region + 1);
        float in[NB][NB][CIN];
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) in[a][b][ci] = xin[((2 * py + a) * PW + 2 * px + b) * CIN + ci];
        float2 in2[DUP ? NB : 1][DUP ? NB : 1][DUP ? CIN : 1];
        if (DUP) {
#pragma unroll
            for (int a = 0; a < NB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) in2[DUP ? a : 0][DUP ? b : 0][DUP ? ci : 0] = make_float2(in[a][b][ci], in[a][b][ci]);
        }
        if (p.taps_out) {
            // tap expansion of this thread's 2x2 pixels for the tensor-path consumer (lu0.conv2): channel = tap*CIN + c
            constexpr int NCH = KS * KS * CIN;
            const bool wide = wide_taps;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t pix = ((int64_t)n * p.S + y0 + 2 * py + (q >> 1)) * p.S + x0 + 2 * px + (q & 1);
                __half* o = p.taps_out + pix * p.taps_cs;
#pragma unroll
                for (int g = 0; g < (NCH + 15) / 16; ++g) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ch = g * 16 + j;
                        v[j] = ch < NCH ? in[(q >> 1) + (ch / CIN) / KS][(q & 1) + (ch / CIN) % KS][ch % CIN] : 0.f;
                    }
                    if (wide) store_h16(o + g * 16, p.taps_plane_elems, p.taps_planes, v);
                    else {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            if (g * 16 + j4 * 4 < p.taps_cs) store_h2x4(o + g * 16 + j4 * 4, p.taps_plane_elems, p.taps_planes, v + j4 * 4);
                    }
                }
            }
        }
        for (int cg = 0; cg < cpad; cg += 16) {
            // packed fp32 FMAs (FFMA2, sm_100): two adjacent output channels per instruction, each lane rounded exactly as fmaf
            float2 acc2[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc2[q][j] = make_float2(bs[cg + 2 * j], bs[cg + 2 * j + 1]);
#pragma unroll
            for (int tap = 0; tap < KS * KS; ++tap) {
                const int dy = tap / KS, dx = tap % KS;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float4* w4 = reinterpret_cast<const float4*>(ws + (tap * CIN + ci) * cpad + cg);
                    float2 x2[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (DUP) x2[q] = in2[DUP ? (q >> 1) + dy : 0][DUP ? (q & 1) + dx : 0][DUP ? ci : 0];
                        else { const float xv = in[(q >> 1) + dy][(q & 1) + dx][ci]; x2[q] = make_float2(xv, xv); }
                    }
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w = w4[j4];
                        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc2[q][j4 * 2 + 0] = __ffma2_rn(x2[q], w01, acc2[q][j4 * 2 + 0]);
                            acc2[q][j4 * 2 + 1] = __ffma2_rn(x2[q], w23, acc2[q][j4 * 2 + 1]);
                        }
                    }
                }
            }
            if (cg == 0 && more) lookup();          // the next region's samples have arrived by now
            float acc[4][16];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[q][2 * j] = acc2[q][j].x; acc[q][2 * j + 1] = acc2[q][j].y; }
            // ReLU / leaky-ReLU (slope > 0) are non-decreasing, so max-pool first and activate the survivor only:
            // max(act(a), act(b)) == act(max(a, b)) bit for bit, at a quarter of the activation work
            const int nq = pool ? 1 : 4;
            if (pool) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[0][j] = apply_act(fmaxf(fmaxf(acc[0][j], acc[1][j]), fmaxf(acc[2][j], acc[3][j])), act, p.leaky);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[q][j] = apply_act(acc[q][j], act, p.leaky);
            }
            if (FAST) {
                const int64_t opix = ((int64_t)n * oh + y0 / 2 + py) * oh + x0 / 2 + px;
                __half* o = p.out_h + opix * p.out_cs + cg;
                if (wide_out && cg + 16 <= p.out_cs) store_h16(o, p.out_plane_elems, p.out_planes, acc[0]);
                else {                              // channel strides are multiples of 8: 16-byte stores
                    if (cg < p.out_cs) store_h8(o, p.out_plane_elems, p.out_planes, acc[0]);
                    if (cg + 8 < p.out_cs) store_h8(o + 8, p.out_plane_elems, p.out_planes, acc[0] + 8);
                }
                continue;
            }
            for (int q = 0; q < nq; ++q) {
                const int oy = pool ? (y0 / 2 + py) : (y0 + 2 * py + (q >> 1));
                const int ox = pool ? (x0 / 2 + px) : (x0 + 2 * px + (q & 1));
                const int64_t opix = ((int64_t)n * oh + oy) * oh + ox;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = q == 0 ? acc[0][j] : (q == 1 ? acc[1][j] : (q == 2 ? acc[2][j] : acc[3][j]));
                if (p.out) {
                    float* o = p.out + opix * p.cout + cg;
                    if (cg + 16 <= p.cout && !(p.cout & 3)) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            reinterpret_cast<float4*>(o)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (cg + j < p.cout) o[j] = v[j];
                    }
                }
                if (p.out_h) {
                    if (wide_out && cg + 16 <= p.out_cs)
                        store_h16(p.out_h + opix * p.out_cs + cg, p.out_plane_elems, p.out_planes, v);
                    else {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            if (cg + j4 * 4 < p.out_cs) store_h2x4(p.out_h + opix * p.out_cs + cg + j4 * 4, p.out_plane_elems, p.out_planes, v + j4 * 4);
                    }
                }
            }
        }
        if (more) stash(xin0 + ((rr + 1) & 1) * XN);
        __syncthreads();
    }
}

// Tap expansion of the raw network input: out[n][y][x][tap*cin + c] = src[n][y+dy][x+dx][c] (0 outside the tile),
// as fp16 hi[/lo] planes.  One thread per (pixel, group of 8 output channels): one 16-byte store per plane.
__global__ void __launch_bounds__(256) taps_kernel(const TapsParams p) {
    const int groups = p.cs >> 3;
    const int64_t total = (int64_t)p.n_tiles * p.S * p.S * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int x = (int)(pix % p.S), y = (int)((pix / p.S) % p.S);
    const int64_t n = pix / ((int64_t)p.S * p.S);
    const int r = p.ks / 2, nch = p.ks * p.ks * p.cin;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = g * 8 + j;
        float val = 0.f;
        if (ch < nch) {
            const int tap = ch / p.cin, c = ch - tap * p.cin;
            const int yy = y + tap / p.ks - r, xx = x + tap % p.ks - r;
            if (yy >= 0 && yy < p.S && xx >= 0 && xx < p.S) val = __ldg(p.src + ((n * p.S + yy) * p.S + xx) * p.cin + c);
        }
        v[j] = val;
    }
    __half* o = p.out + pix * p.cs + g * 8;
    store_h2x4(o, p.out_plane_elems, p.out_planes, v);
    store_h2x4(o + 4, p.out_plane_elems, p.out_planes, v + 4);
}

// One thread per pixel: K logits from cin channels, then a numerically stable softmax.
template <int K>
__global__ void __launch_bounds__(256) top_softmax_kernel(const TopParams p) {
    extern __shared__ float wsm[];              // [cin][K] + [K]
    for (int i = threadIdx.x; i < p.cin * K; i += blockDim.x) wsm[i] = p.w[i];
    if (threadIdx.x < K) wsm[p.cin * K + threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    __syncthreads();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.n_pix) return;
    const float* x = p.src + pix * p.cin;
    float z[K];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = 0.f;
    if ((p.cin & 3) == 0) {
        for (int c = 0; c < p.cin; c += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + c));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                z[k] = fmaf(v.x, wsm[(c + 0) * K + k], z[k]);
                z[k] = fmaf(v.y, wsm[(c + 1) * K + k], z[k]);
                z[k] = fmaf(v.z, wsm[(c + 2) * K + k], z[k]);
                z[k] = fmaf(v.w, wsm[(c + 3) * K + k], z[k]);
            }
        }
    } else {
        for (int c = 0; c < p.cin; ++c) {
            const float v = __ldg(x + c);
#pragma unroll
            for (int k = 0; k < K; ++k) z[k] = fmaf(v, wsm[c * K + k], z[k]);
        }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] += wsm[p.cin * K + k]; m = fmaxf(m, z[k]); }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - m); s += z[k]; }
    const float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < K; ++k) p.probs[pix * K + k] = z[k] * inv;
}

__device__ __forceinline__ double load_sample(const void* img, int dtype, int64_t i) {
    switch (dtype) {
        case UMX_U8:  return (double)reinterpret_cast<const uint8_t*>(img)[i];
        case UMX_U16: return (double)reinterpret_cast<const uint16_t*>(img)[i];
        case UMX_F32: return (double)reinterpret_cast<const float*>(img)[i];
        default:      return reinterpret_cast<const double*>(img)[i];
    }
}

// ---- skimage.transform.resize semantics (see umx_kernels.cuh: Resample), float64, scipy's operation order:
// no FMA contraction (__dmul_rn / __dadd_rn), bilinear terms summed (0,0), (0,1), (1,0), (1,1), each (v * wy) * wx.
__device__ __forceinline__ int mirror_idx(int i, int n) {      // ndimage 'mirror': d c b | a b c d | c b a
    if (n <= 1) return 0;
    const int p = 2 * n - 2;
    i = i < 0 ? -i : i;
    if (i >= p) i %= p;
    return i >= n ? p - i : i;
}

struct AxisTaps { int s0, s1; double t; };

// output index k -> the two source samples and the weight of the second one (NI_ZoomShift with grid_mode: the
// coordinate (k + 0.5) * zoom - 0.5 is reflected when negative; a right neighbour past the end mirrors to n - 2)
__device__ __forceinline__ AxisTaps axis_taps(int k, double zoom, int n) {
    double cc = __dsub_rn(__dmul_rn((double)k + 0.5, zoom), 0.5);
    if (cc < 0.0) cc = -cc;
    AxisTaps a;
    a.s0 = (int)floor(cc);
    if (a.s0 > n - 1) a.s0 = n - 1;
    a.t = cc - (double)a.s0;
    a.s1 = a.s0 + 1;
    if (a.s1 >= n) a.s1 = n > 1 ? 2 * n - 2 - a.s1 : 0;
    return a;
}

// Gaussian-filtered source sample at (r, c): axis 0 first, then axis 1 (ndimage.gaussian_filter), each pass in
// correlate1d's symmetric form  x[0]*g[0] + sum_{j<0} (x[j] + x[-j]) * g[j].
template <typename Fetch>
__device__ __forceinline__ double filtered_sample(const Fetch& f, const Resample& rs, int r, int c) {
    if (rs.ry == 0 && rs.rx == 0) return f(r, c);
    auto column = [&](int cc) -> double {
        if (rs.ry == 0) return f(r, cc);
        double acc = __dmul_rn(f(r, cc), rs.gy[rs.ry]);
        for (int j = -rs.ry; j < 0; ++j) {
            const double pair = __dadd_rn(f(mirror_idx(r + j, rs.src_h), cc), f(mirror_idx(r - j, rs.src_h), cc));
            acc = __dadd_rn(acc, __dmul_rn(pair, rs.gy[rs.ry + j]));
        }
        return acc;
    };
    if (rs.rx == 0) return column(c);
    double acc = __dmul_rn(column(c), rs.gx[rs.rx]);
    for (int j = -rs.rx; j < 0; ++j) {
        const double pair = __dadd_rn(column(mirror_idx(c + j, rs.src_w)), column(mirror_idx(c - j, rs.src_w)));
        acc = __dadd_rn(acc, __dmul_rn(pair, rs.gx[rs.rx + j]));
    }
    return acc;
}

template <typename Fetch>
__device__ __forceinline__ double resample_value(const Fetch& f, const Resample& rs, int y, int x) {
    const AxisTaps ay = axis_taps(y, rs.zoom_y, rs.src_h), ax = axis_taps(x, rs.zoom_x, rs.src_w);
    const double wy0 = 1.0 - ay.t, wx0 = 1.0 - ax.t;
    double v = __dmul_rn(__dmul_rn(filtered_sample(f, rs, ay.s0, ax.s0), wy0), wx0);
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(filtered_sample(f, rs, ay.s0, ax.s1), wy0), ax.t));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(filtered_sample(f, rs, ay.s1, ax.s0), ay.t), wx0));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(filtered_sample(f, rs, ay.s1, ax.s1), ay.t), ax.t));
    return v;
}

__device__ __forceinline__ double rescale_sample(double v, const PreMap& pre) {      // rescale_intensity
    v = fmin(fmax(v, pre.imin), pre.imax);
    v = (v - pre.imin) / (pre.imax - pre.imin);
    return v * (pre.omax - pre.omin) + pre.omin;
}

// every 8/16-bit sample code through the float64 map of the gather (entry n: the zero padding outside the image)
__global__ void __launch_bounds__(256) norm_lut_kernel(const NormLutParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > p.n) return;
    double v = 0.0;
    if (i < p.n) {
        v = (double)i;
        if (p.has_pre) {
            v = v * p.pre.in_scale;
            if (p.pre.rescale) v = rescale_sample(v, p.pre);
        }
    }
    p.out[i] = (float)((v - p.mean) / p.std_dev);
}

// out[t][y][x][c] = float((frame - mean)/std); frame = premap(sample) inside the image, 0 outside.
__global__ void __launch_bounds__(256) gather_tiles_kernel(const GatherParams p) {
    const int64_t total = (int64_t)p.n_tiles * p.S * p.S;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int lg = 31 - __clz(p.S);               // imSize is a power of two (checked in umx_create)
    const int x = (int)(i & (p.S - 1));
    const int y = (int)((i >> lg) & (p.S - 1));
    const int tl = (int)(i >> (2 * lg));
    const int t = p.tile0 + tl;
    const int ti = t / p.npc, tj = t % p.npc;
    const int r = ti * p.sub + y - p.margin;     // image row
    const int c = tj * p.sub + x - p.margin;     // image col
    const bool inside = (r >= 0 && r < p.H && c >= 0 && c < p.W);
    if (p.lut) {          // integer samples, no resampling: one table look-up replaces the float64 arithmetic bit for bit
        const int n_codes = p.dtype == UMX_U8 ? 256 : 65536;
        for (int ch = 0; ch < p.C; ++ch) {
            int code = n_codes;
            if (inside) {
                const int64_t idx = (int64_t)((p.n_planes == 1) ? 0 : ch) * p.plane_stride + (int64_t)(r - p.img_row0) * p.W + c;
                code = p.dtype == UMX_U8 ? (int)reinterpret_cast<const uint8_t*>(p.img)[idx] : (int)reinterpret_cast<const uint16_t*>(p.img)[idx];
            }
            p.out[i * p.C + ch] = __ldg(p.lut + ch * 65537 + code);
        }
        return;
    }
    for (int ch = 0; ch < p.C; ++ch) {
        double v = 0.0;
        if (inside) {
            const int plane = (p.n_planes == 1) ? 0 : ch;
            if (p.rs.on) {
                const int64_t base = (int64_t)plane * p.plane_stride - (int64_t)p.img_row0 * p.rs.src_w;
                const double sc = p.has_pre ? p.pre[ch].in_scale : 1.0;
                auto fetch = [&](int rr, int cc) -> double {
                    return __dmul_rn(load_sample(p.img, p.dtype, base + (int64_t)rr * p.rs.src_w + cc), sc);
                };
                v = resample_value(fetch, p.rs, r, c);
                if (p.has_pre && p.pre[ch].rescale) v = rescale_sample(v, p.pre[ch]);
            } else {
                v = load_sample(p.img, p.dtype, (int64_t)plane * p.plane_stride + (int64_t)(r - p.img_row0) * p.W + c);
                if (p.has_pre) {
                    v = v * p.pre[ch].in_scale;
                    if (p.pre[ch].rescale) v = rescale_sample(v, p.pre[ch]);
                }
            }
        }
        p.out[i * p.C + ch] = (float)((v - p.mean) / p.std_dev);
    }
}

// One thread per destination pixel, all K planes: uint8 page * (1/255) -> resize -> np.uint8(255 * x).
__global__ void __launch_bounds__(256) resize_u8_kernel(const ResizeU8Params p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = p.row0 + blockIdx.y;
    if (x >= p.dst_w || y >= p.row1) return;
    for (int k = 0; k < p.K; ++k) {
        const uint8_t* plane = p.src + (int64_t)k * p.src_plane_stride - (int64_t)p.src_row0 * p.rs.src_w;
        auto fetch = [&](int rr, int cc) -> double {
            return __dmul_rn((double)__ldg(plane + (int64_t)rr * p.rs.src_w + cc), 1.0 / 255);
        };
        const double v = resample_value(fetch, p.rs, y, x);
        p.out[(int64_t)k * p.out_plane_stride + (int64_t)(y - p.out_row_base) * p.dst_w + x] = (uint8_t)fmin(255.0, fmax(0.0, floor(__dmul_rn(255.0, v))));
    }
}

__device__ __forceinline__ unsigned long long minmax_encode(double v) {      // order-preserving double -> uint64
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) resample_minmax_kernel(const MinMaxParams p) {
    const int64_t total = (int64_t)p.dst_h * p.dst_w;
    double lo = INFINITY, hi = -INFINITY;
    auto fetch = [&](int rr, int cc) -> double {
        return __dmul_rn(load_sample(p.img, p.dtype, (int64_t)rr * p.rs.src_w + cc), p.in_scale);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.dst_w), x = (int)(i % p.dst_w);
        const double v = p.rs.on ? resample_value(fetch, p.rs, y, x) : fetch(y, x);
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(p.out + 0, minmax_encode(lo));
        atomicMax(p.out + 1, minmax_encode(hi));
    }
}

__device__ __forceinline__ float ramp(int v, int S, int two_m) {
    const int d = min(v, S - 1 - v);
    return fminf(1.f, (float)d / (float)two_m);
}

// One thread per output pixel; every class.  Each pixel is covered by <= 2x2 tiles.
template <int K>
__global__ void __launch_bounds__(256) stitch_kernel(const StitchParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = p.row0 + blockIdx.y;
    if (c >= p.W || r >= p.row1) return;
    const int R = r + p.margin, Cc = c + p.margin;      // padded-frame coordinates
    const int two_m = 2 * p.margin;
    int ti1 = min(R / p.sub, p.npr - 1);
    int tj1 = min(Cc / p.sub, p.npc - 1);
    float num[K];
#pragma unroll
    for (int k = 0; k < K; ++k) num[k] = 0.f;
    float cnt = 0.f;
    if (p.replace) {        // tiles are patched in row-major order: the lower-right tile covering the pixel is the last writer
        const float* q = p.probs + ((((int64_t)(ti1 - p.tr_lo) * p.npc + tj1) * p.S + (R - ti1 * p.sub)) * p.S + (Cc - tj1 * p.sub)) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) num[k] = __ldg(q + k);
        cnt = 1.f;
    } else
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int ti = ti1 - a;
        if (ti < 0) continue;
        const int y = R - ti * p.sub;
        if (y >= p.S) continue;
        const float wy = ramp(y, p.S, two_m);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int tj = tj1 - b;
            if (tj < 0) continue;
            const int x = Cc - tj * p.sub;
            if (x >= p.S) continue;
            const float w = fminf(wy, ramp(x, p.S, two_m));
            cnt += w;
            if (w > 0.f) {
                const float* q = p.probs + ((((int64_t)(ti - p.tr_lo) * p.npc + tj) * p.S + y) * p.S + x) * K;
#pragma unroll
                for (int k = 0; k < K; ++k) num[k] = fmaf(__ldg(q + k), w, num[k]);
            }
        }
    }
    const int64_t o = (int64_t)(r - p.out_row_base) * p.W + c;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float v = num[k] / cnt;
        if (p.out_f32) p.out_f32[k * p.out_plane_stride + o] = v;
        if (p.out_u8) {
            uint8_t q;
            if (p.fp16_quant) {     // p rounded to float16 (PI2D's output dtype), 255 * p rounded to float16 again, then truncated
                const float p16 = __half2float(__float2half_rn(v));
                q = (uint8_t)fminf(255.f, floorf(__half2float(__float2half_rn(255.f * p16))));
            } else q = (uint8_t)fminf(255.f, floorf(255.f * v));
            if (p.requant) q = (uint8_t)(__dmul_rn(255.0, __dmul_rn((double)q, 1.0 / 255)));      // second quantisation at equal size
            p.out_u8[k * p.out_plane_stride + o] = q;
        }
    }
}

}  // namespace

size_t conv_simt_smem_bytes(const ConvParams& p) {
    size_t best = 0;
    for (int t = 0; t < p.nterms; ++t) {
        const ConvTerm& T = p.term[t];
        int mt = 0;
        for (int ph = 0; ph < p.nphase; ++ph) mt = T.ntaps[ph] > mt ? T.ntaps[ph] : mt;
        const size_t npix = (size_t)p.nt * (p.ph + T.hy0 + T.hy1) * (p.pw + T.hx0 + T.hx1);
        const size_t fl = ((npix * CK + 3) & ~(size_t)3) + (size_t)mt * CK * COT;
        best = fl > best ? fl : best;
    }
    return best * sizeof(float);
}

cudaError_t conv_simt_configure() {
    return cudaFuncSetAttribute(conv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t s) {
    dim3 grid;
    if (p.nt > 1) grid.x = (p.n_tiles + p.nt - 1) / p.nt;
    else grid.x = (unsigned)((int64_t)p.n_tiles * (p.in_h / p.ph) * (p.in_w / p.pw));
    grid.y = (p.cout + COT - 1) / COT;
    grid.z = p.nphase;
    conv_simt_kernel<<<grid, NTHR, conv_simt_smem_bytes(p), s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_top_softmax(const TopParams& p, cudaStream_t s) {
    const unsigned blocks = (unsigned)((p.n_pix + 255) / 256);
    const size_t sm = (size_t)(p.cin * p.k + p.k) * sizeof(float);
    if (p.k == 2) top_softmax_kernel<2><<<blocks, 256, sm, s>>>(p);
    else if (p.k == 3) top_softmax_kernel<3><<<blocks, 256, sm, s>>>(p);
    else if (p.k == 4) top_softmax_kernel<4><<<blocks, 256, sm, s>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

size_t first_conv_smem_bytes(int cin, int ks, int cout) {
    const int pw = 32 + 2 * (ks / 2), cpad = (cout + 15) & ~15;
    return 2 * (((size_t)pw * pw * cin + 3) & ~(size_t)3) * 4 + ((size_t)ks * ks * cin * cpad + cpad) * 4;
}

template <int CIN, int KS, bool FAST>
static cudaError_t launch_first_v(const FirstParams& p, unsigned grid, size_t sm, cudaStream_t s) {
    if (sm > 48 * 1024) {       // per device and per call: handles on different GPUs share this process
        cudaError_t e = cudaFuncSetAttribute(first_conv_kernel<CIN, KS, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
    }
    first_conv_kernel<CIN, KS, FAST><<<grid, 256, sm, s>>>(p);
    return cudaGetLastError();
}

template <int CIN, int KS>
static cudaError_t launch_first(const FirstParams& p, unsigned grid, size_t sm, cudaStream_t s) {
    static const bool no_fast = getenv("UMX_FC_FAST") && atoi(getenv("UMX_FC_FAST")) == 0;
    const bool fast = !no_fast && p.pool && p.act == ACT_LEAKY && !p.out && p.out_h && (p.out_cs & 7) == 0 &&
                      (p.out_plane_elems & 7) == 0;
    return fast ? launch_first_v<CIN, KS, true>(p, grid, sm, s) : launch_first_v<CIN, KS, false>(p, grid, sm, s);
}

cudaError_t launch_first_conv(const FirstParams& p_in, cudaStream_t s) {
    if (p_in.n_tiles == 0) return cudaSuccess;
    FirstParams p = p_in;
    const int nb = p.S / 32;
    // four regions per block (weights staged once, patch loads pipelined) when that still leaves a few waves of blocks
    static const int force_rpb = getenv("UMX_FC_RPB") ? atoi(getenv("UMX_FC_RPB")) : 0;
    const int64_t regions = (int64_t)p.n_tiles * nb * nb;
    p.rpb = (regions % 4 == 0 && regions / 4 >= 4 * 148 * 2) ? 4 : 1;
    if (force_rpb > 0 && regions % force_rpb == 0) p.rpb = force_rpb;
    const unsigned grid = (unsigned)(regions / p.rpb);
    const size_t sm = first_conv_smem_bytes(p.cin, p.ks, p.cout);
    if (p.cin == 1 && p.ks == 3) return launch_first<1, 3>(p, grid, sm, s);
    if (p.cin == 2 && p.ks == 3) return launch_first<2, 3>(p, grid, sm, s);
    if (p.cin == 1 && p.ks == 5) return launch_first<1, 5>(p, grid, sm, s);
    if (p.cin == 2 && p.ks == 5) return launch_first<2, 5>(p, grid, sm, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_taps(const TapsParams& p, cudaStream_t s) {
    const int64_t total = (int64_t)p.n_tiles * p.S * p.S * (p.cs >> 3);
    if (total == 0) return cudaSuccess;
    taps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_gather_tiles(const GatherParams& p, cudaStream_t s) {
    const int64_t total = (int64_t)p.n_tiles * p.S * p.S;
    if (total == 0) return cudaSuccess;
    gather_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_norm_lut(const NormLutParams& p, cudaStream_t s) {
    norm_lut_kernel<<<(p.n + 1 + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_resize_u8(const ResizeU8Params& p, cudaStream_t s) {
    if (p.row1 <= p.row0) return cudaSuccess;
    dim3 grid((p.dst_w + 255) / 256, p.row1 - p.row0);
    resize_u8_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_resample_minmax(const MinMaxParams& p, cudaStream_t s) {
    const int64_t total = (int64_t)p.dst_h * p.dst_w;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    resample_minmax_kernel<<<blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

double minmax_decode(unsigned long long code) {
    const unsigned long long b = (code & 0x8000000000000000ull) ? (code & 0x7FFFFFFFFFFFFFFFull) : ~code;
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
}

// ndimage's Gaussian taps: radius int(4 sigma + 0.5), exp(-0.5 x^2 / sigma^2) normalised to sum 1
static int gaussian_taps(double sigma, double* g) {
    if (!(sigma > 1e-15)) return 0;
    const int r = (int)(4.0 * sigma + 0.5);
    if (r > kMaxResampleRadius) return -1;
    double sum = 0.0;
    for (int i = -r; i <= r; ++i) { g[i + r] = exp(-0.5 / (sigma * sigma) * (double)(i * i)); sum += g[i + r]; }
    for (int i = 0; i <= 2 * r; ++i) g[i] /= sum;
    return r;
}

bool make_resample(Resample* rs, int src_h, int src_w, int dst_h, int dst_w) {
    memset(rs, 0, sizeof(*rs));
    rs->src_h = src_h; rs->src_w = src_w;
    if (src_h == dst_h && src_w == dst_w) return true;
    rs->on = 1;
    rs->zoom_y = (double)src_h / (double)dst_h; rs->zoom_x = (double)src_w / (double)dst_w;
    // anti-aliasing whenever some axis shrinks, sigma = max(0, (factor - 1) / 2) per axis (skimage defaults)
    if (dst_h < src_h || dst_w < src_w) {
        const int ry = gaussian_taps(std::max(0.0, (rs->zoom_y - 1.0) / 2.0), rs->gy);
        const int rx = gaussian_taps(std::max(0.0, (rs->zoom_x - 1.0) / 2.0), rs->gx);
        if (ry < 0 || rx < 0) return false;
        rs->ry = ry; rs->rx = rx;
    }
    return true;
}

cudaError_t launch_stitch(const StitchParams& p, cudaStream_t s) {
    if (p.row1 <= p.row0) return cudaSuccess;
    dim3 grid((p.W + 255) / 256, p.row1 - p.row0);
    if (p.K == 2) stitch_kernel<2><<<grid, 256, 0, s>>>(p);
    else if (p.K == 3) stitch_kernel<3><<<grid, 256, 0, s>>>(p);
    else if (p.K == 4) stitch_kernel<4><<<grid, 256, 0, s>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace umx
