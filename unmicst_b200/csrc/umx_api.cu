// C-ABI of the engine (include/unmicst_b200.h): model build (fold / repack / plan),
// tile forward, whole-image tiling driver, profiling.  Host code only; kernels live in
// kernels_simt.cu (fp32 CUDA cores) and kernels_tc.cu (tcgen05 tensor path).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "umx_internal.h"

namespace umx {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

constexpr float kBnEps = 1e-3f;      // FusedBatchNorm epsilon of every shipped graph
constexpr float kLeaky = 0.2f;       // LeakyRelu alpha of the v2 graphs

const HostTensor* find(const umx_handle* h, const std::string& name) {
    auto it = h->tensors.find(name);
    return it == h->tensors.end() ? nullptr : &it->second;
}

int need(const umx_handle* h, const std::string& name, std::vector<int64_t> shape, const HostTensor** out) {
    const HostTensor* t = find(h, name);
    if (!t) { set_error("tensor '%s' required by the graph is missing", name.c_str()); return UMX_ENOTENSOR; }
    if (t->shape != shape) {
        std::string got, want;
        for (auto d : t->shape) got += std::to_string(d) + ",";
        for (auto d : shape) want += std::to_string(d) + ",";
        set_error("tensor '%s' has shape [%s] but the graph needs [%s]", name.c_str(), got.c_str(), want.c_str());
        return UMX_ENOTENSOR;
    }
    *out = t;
    return UMX_OK;
}

int upload(umx_handle* h, const std::vector<float>& v, float** out) {
    float* d = nullptr;
    UMX_CUDA_TRY(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
    h->dev_allocs.push_back(d);
    if (!v.empty()) UMX_CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    *out = d;
    return UMX_OK;
}

int new_buffer(umx_handle* h, const std::string& name, int hh, int ww, int c) {
    Buffer b; b.name = name; b.h = hh; b.w = ww; b.c = c;
    h->bufs.push_back(b);
    return (int)h->bufs.size() - 1;
}

// gamma/sqrt(var+eps) and beta - mu*scale of one batch-norm scope
int bn_affine(const umx_handle* h, const std::string& scope, int c, std::vector<float>* scale, std::vector<float>* shift) {
    const HostTensor *g, *b, *mu, *var;
    UMX_TRY(need(h, scope + "/gamma", {c}, &g));
    UMX_TRY(need(h, scope + "/beta", {c}, &b));
    UMX_TRY(need(h, scope + "/moving_mean", {c}, &mu));
    UMX_TRY(need(h, scope + "/moving_variance", {c}, &var));
    scale->resize(c); shift->resize(c);
    for (int i = 0; i < c; ++i) {
        const float s = g->data[i] / sqrtf(var->data[i] + kBnEps);
        (*scale)[i] = s;
        (*shift)[i] = b->data[i] - mu->data[i] * s;
    }
    return UMX_OK;
}

void conv_taps(ConvTerm* T, int k) {
    const int r = (k - 1) / 2;
    T->ntaps[0] = (int8_t)(k * k);
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            const int t = a * k + b;
            T->dy[0][t] = (int8_t)(a - r); T->dx[0][t] = (int8_t)(b - r); T->wi[0][t] = (int8_t)t;
        }
    T->hy0 = T->hy1 = T->hx0 = T->hx1 = r;
}

// tf.nn.conv2d_transpose stride 2 SAME as four sub-pixel phases:
// out[2q+p] gathers taps a with (p + pb - a) even from input row q + (p + pb - a)/2.
void convt_taps(ConvTerm* T, int k) {
    const int pb = (k - 2) / 2;
    int hy0 = 0, hy1 = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int ph = py * 2 + px;
            int n = 0;
            for (int a = 0; a < k; ++a) {
                if ((py + pb - a) & 1) continue;
                for (int b = 0; b < k; ++b) {
                    if ((px + pb - b) & 1) continue;
                    const int dy = (py + pb - a) / 2, dx = (px + pb - b) / 2;   // exact: even numerators
                    T->dy[ph][n] = (int8_t)dy; T->dx[ph][n] = (int8_t)dx; T->wi[ph][n] = (int8_t)(a * k + b);
                    hy0 = std::max(hy0, -dy); hy1 = std::max(hy1, dy);
                    ++n;
                }
            }
            T->ntaps[ph] = (int8_t)n;
        }
    T->hy0 = T->hx0 = hy0; T->hy1 = T->hx1 = hy1;
}

void pick_patch(ConvParams* cp) {
    cp->pw = std::min(16, cp->in_w);
    cp->ph = std::min(8, cp->in_h);
    while (cp->ph * cp->pw > 128) cp->ph >>= 1;
    cp->nt = 128 / (cp->ph * cp->pw);
}

using TermSpec = TermHost;

// Append a fused conv op (host-side description only; lowered to a kernel by lower_plan):
//   out = pool?(post(act(sum_terms conv + bias)))
int add_conv(umx_handle* h, const std::string& name, std::vector<TermSpec>& terms, int cout, bool transpose,
             const std::vector<float>* bias, int act, const std::vector<float>* post_scale,
             const std::vector<float>* post_shift, bool pool, int* out_buf) {
    Op op; op.kind = OP_CONV; op.name = name;
    ConvSpec& sp = op.spec;
    sp.terms = terms; sp.cout = cout; sp.transpose = transpose; sp.act = act; sp.pool = pool;
    if (bias) { sp.has_bias = true; sp.bias = *bias; }
    if (post_scale) { sp.has_post = true; sp.post_scale = *post_scale; sp.post_shift = *post_shift; }
    const Buffer in0 = h->bufs[terms[0].src0];
    double macs = 0, in_elems = 0, wbytes = 0;
    for (auto& t : terms) {
        const Buffer& a = h->bufs[t.src0];
        if (a.h != in0.h || a.w != in0.w) { set_error("%s: term grid mismatch", name.c_str()); return UMX_EINVAL; }
        const int cin = a.c + (t.src1 >= 0 ? h->bufs[t.src1].c : 0);
        macs += (double)t.k * t.k * cin * cout * in0.h * in0.w;      // conv: per out px; convT: per in px
        in_elems += (double)cin * in0.h * in0.w;
        wbytes += (double)t.w.size() * 4;
    }
    const int os = transpose ? 2 : 1;
    const int oh = pool ? in0.h / 2 : in0.h * os, ow = pool ? in0.w / 2 : in0.w * os;
    op.out_buf = new_buffer(h, name, oh, ow, cout);
    op.flops_per_tile = 2.0 * macs;
    op.bytes_per_tile = 4.0 * (in_elems + (double)oh * ow * cout);
    op.weight_bytes = wbytes;
    h->ops.push_back(std::move(op));
    *out_buf = h->ops.back().out_buf;
    return UMX_OK;
}

std::vector<float> scaled_kernel(const HostTensor& w, const HostTensor* add, const std::vector<float>* scale) {
    // HWIO [k][k][cin][cout] is already [tap][cin][cout]
    std::vector<float> out(w.data);
    const int64_t cout = w.shape[3];
    if (add) for (size_t i = 0; i < out.size(); ++i) out[i] += add->data[i];
    if (scale) for (size_t i = 0; i < out.size(); ++i) out[i] *= (*scale)[i % cout];
    return out;
}

std::vector<float> transposed_kernel(const HostTensor& w) {
    // TF conv2d_transpose filter [a][b][cout][cin] -> [tap][cin][cout]
    const int64_t k = w.shape[0], co = w.shape[2], ci = w.shape[3];
    std::vector<float> out(w.data.size());
    for (int64_t t = 0; t < k * k; ++t)
        for (int64_t o = 0; o < co; ++o)
            for (int64_t i = 0; i < ci; ++i) out[(t * ci + i) * co + o] = w.data[(t * co + o) * ci + i];
    return out;
}

int build_plan(umx_handle* h) {
    const int L = h->L, k = h->desc.ks, E = h->desc.n_extra_convs, K = h->K;
    const std::vector<int>& n = h->chan;
    const bool v2 = h->desc.graph == UMX_GRAPH_V2;
    const int act = v2 ? ACT_LEAKY : ACT_RELU;
    h->in_buf = new_buffer(h, "input", h->S, h->S, h->C);
    std::vector<int> ds{h->in_buf};
    char nm[128];
    for (int i = 0; i < L; ++i) {
        const int src = ds[i];
        const HostTensor *w1, *wsc;
        std::vector<float> scale, shift;
        if (v2) {
            snprintf(nm, sizeof nm, "downsampling/ld%d/kernelD%d", i, i);
            UMX_TRY(need(h, nm, {k, k, n[i], n[i + 1]}, &w1));
            snprintf(nm, sizeof nm, "ld%d/shortcutWeights", i);
            UMX_TRY(need(h, nm, {k, k, n[i], n[i + 1]}, &wsc));
            snprintf(nm, sizeof nm, "ld%d/batch_normalization", i);
            UMX_TRY(bn_affine(h, nm, n[i + 1], &scale, &shift));
        } else {
            snprintf(nm, sizeof nm, "downsampling/ld%d/kernel1", i);
            UMX_TRY(need(h, nm, {k, k, n[i], n[i + 1]}, &w1));
            snprintf(nm, sizeof nm, "downsampling/ld%d/shortcutWeights", i);
            UMX_TRY(need(h, nm, {1, 1, n[i], n[i + 1]}, &wsc));
            UMX_TRY(bn_affine(h, i == 0 ? std::string("batch_normalization") : "batch_normalization_" + std::to_string(i),
                              n[i + 1], &scale, &shift));
        }
        int cur = src, out = -1;
        // chain: kernel1/kernelD, then extras; the shortcut joins the LAST conv of the chain
        for (int e = 0; e <= E; ++e) {
            const HostTensor* we = w1;
            if (e > 0) {
                if (v2) snprintf(nm, sizeof nm, "ld%d/kernelExtra%d", i, e - 1);
                else snprintf(nm, sizeof nm, "downsampling/ld%d/kernelExtra%d", i, e - 1);
                UMX_TRY(need(h, nm, {k, k, n[i + 1], n[i + 1]}, &we));
            }
            const bool last = (e == E);
            std::vector<TermSpec> terms(1);
            terms[0].src0 = cur; terms[0].k = k;
            snprintf(nm, sizeof nm, "ld%d.conv%d", i, e);
            if (!last) {
                terms[0].w = scaled_kernel(*we, nullptr, nullptr);
                UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, nullptr, act, nullptr, nullptr, false, &out));
            } else if (v2) {
                // leaky(BN(c + s)) -> fold BN scale into both linear maps; merge them when they share the input
                if (E == 0) {
                    terms[0].w = scaled_kernel(*we, wsc, &scale);
                } else {
                    terms[0].w = scaled_kernel(*we, nullptr, &scale);
                    TermSpec sc; sc.src0 = src; sc.k = k; sc.w = scaled_kernel(*wsc, nullptr, &scale);
                    terms.push_back(sc);
                }
                UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, &shift, act, nullptr, nullptr, true, &out));
            } else {
                // BN(relu(c + s)): shortcut is a 1x1 conv of the layer input; BN stays a post-activation affine
                terms[0].w = scaled_kernel(*we, nullptr, nullptr);
                TermSpec sc; sc.src0 = src; sc.k = 1; sc.w = scaled_kernel(*wsc, nullptr, nullptr);
                terms.push_back(sc);
                UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, nullptr, act, &scale, &shift, true, &out));
            }
            cur = out;
        }
        ds.push_back(cur);
    }
    // bottom
    int u = -1;
    {
        const HostTensor* wb;
        UMX_TRY(need(h, "lb/kernel1", {k, k, n[L], n[L + 1]}, &wb));
        std::vector<TermSpec> terms(1);
        terms[0].src0 = ds[L]; terms[0].k = k;
        if (v2) {
            std::vector<float> scale, shift;
            UMX_TRY(bn_affine(h, "conv", n[L + 1], &scale, &shift));
            terms[0].w = scaled_kernel(*wb, nullptr, &scale);
            UMX_TRY(add_conv(h, "lb.conv", terms, n[L + 1], false, &shift, act, nullptr, nullptr, false, &u));
        } else {
            terms[0].w = scaled_kernel(*wb, nullptr, nullptr);
            UMX_TRY(add_conv(h, "lb.conv", terms, n[L + 1], false, nullptr, act, nullptr, nullptr, false, &u));
        }
    }
    // up path
    for (int i = L - 1; i >= 0; --i) {
        const HostTensor *wu, *w2;
        if (v2) snprintf(nm, sizeof nm, "lu%d/kernelU%d", i, i); else snprintf(nm, sizeof nm, "upsampling/lu%d/kernel1", i);
        UMX_TRY(need(h, nm, {k, k, n[i + 1], n[i + 2]}, &wu));
        if (v2) snprintf(nm, sizeof nm, "lu%d/kernel2", i); else snprintf(nm, sizeof nm, "upsampling/lu%d/kernel2", i);
        UMX_TRY(need(h, nm, {k, k, n[i] + n[i + 1], n[i + 1]}, &w2));
        int us = -1, cv = -1;
        {
            std::vector<TermSpec> terms(1);
            terms[0].src0 = u; terms[0].k = k; terms[0].w = transposed_kernel(*wu);
            snprintf(nm, sizeof nm, "lu%d.convT", i);
            UMX_TRY(add_conv(h, nm, terms, n[i + 1], true, nullptr, act, nullptr, nullptr, false, &us));
        }
        {
            std::vector<TermSpec> terms(1);
            terms[0].src0 = ds[i]; terms[0].src1 = us; terms[0].k = k;
            snprintf(nm, sizeof nm, "lu%d.conv2", i);
            if (v2) {
                std::vector<float> scale, shift;
                char sc[64]; snprintf(sc, sizeof sc, "lu%d/conv2", i);
                UMX_TRY(bn_affine(h, sc, n[i + 1], &scale, &shift));
                terms[0].w = scaled_kernel(*w2, nullptr, &scale);
                UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, &shift, act, nullptr, nullptr, false, &cv));
            } else {
                terms[0].w = scaled_kernel(*w2, nullptr, nullptr);
                UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, nullptr, act, nullptr, nullptr, false, &cv));
            }
        }
        for (int e = 0; e < E; ++e) {
            const HostTensor* we;
            if (v2) snprintf(nm, sizeof nm, "lu%d/kernel2Extra%d", i, e); else snprintf(nm, sizeof nm, "upsampling/lu%d/kernel2Extra%d", i, e);
            UMX_TRY(need(h, nm, {k, k, n[i + 1], n[i + 1]}, &we));
            std::vector<TermSpec> terms(1);
            terms[0].src0 = cv; terms[0].k = k; terms[0].w = scaled_kernel(*we, nullptr, nullptr);
            snprintf(nm, sizeof nm, "lu%d.extra%d", i, e);
            UMX_TRY(add_conv(h, nm, terms, n[i + 1], false, nullptr, act, nullptr, nullptr, false, &cv));
        }
        u = cv;
    }
    // top: 1x1 conv (+ BN on the logits in v2) + softmax
    {
        const HostTensor* wt;
        UMX_TRY(need(h, "lt/kernel", {1, 1, n[1], K}, &wt));
        Op op; op.kind = OP_TOP; op.name = "lt.softmax";
        std::vector<float> w(wt->data), bias;
        if (v2) {
            std::vector<float> scale, shift;
            UMX_TRY(bn_affine(h, "batch_normalization", K, &scale, &shift));
            for (size_t i = 0; i < w.size(); ++i) w[i] *= scale[i % K];
            bias = shift;
        }
        op.top_w = w; op.top_b = bias;
        op.tp.cin = n[1]; op.tp.k = K;
        op.top_src = u;
        op.flops_per_tile = 2.0 * n[1] * K * h->S * h->S;
        op.bytes_per_tile = 4.0 * (double)(n[1] + K) * h->S * h->S;
        h->ops.push_back(op);
    }
    return UMX_OK;
}

enum { TC_NONE = 0, TC_PLAIN = 1, TC_SHORT_SKIP = 3, TC_SHORT_A1 = 4 };

// hi/lo correction terms of a tensor-path op, per concat source: bit 0 = a_hi * w_lo, bit 1 = a_lo * w_hi.
// 3 = the full split (three MMAs per product), 0 = one MMA per product; packed t0 | t1 << 2.
int op_terms(const umx_handle* h, const Op& op) {
    const size_t idx = (size_t)(&op - h->ops.data());
    if (idx < h->op_terms.size() && h->op_terms[idx] >= 0) return h->op_terms[idx] & 15;
    if (h->precision == UMX_PREC_SINGLE) return 0;
    if (h->precision == UMX_PREC_MIXED) return (idx < 64 && ((h->single_mask >> idx) & 1)) ? 0 : 15;
    return 15;
}
// fp16 operand planes of a tensor-path op: 1 = one MMA per product, 2 = some hi/lo correction term is on
int op_planes(const umx_handle* h, const Op& op) { return op_terms(h, op) ? 2 : 1; }

// Which tensor-path form (if any) can run this op (after rewrite_narrow_sources).
//  - one term: k x k conv / conv-transpose of one or two wide concat sources (TC_PLAIN); k = 1 is the tap-expanded
//    first layer
//  - two terms, the second a 1x1 conv of another buffer (legacy shortcut UnMicst.py:95-97, or the tap-expanded raw
//    input of lu0.conv2): a one-channel source is folded into the epilogue (TC_SHORT_SKIP), a wide one joins the
//    K loop at the centre tap (TC_SHORT_A1)
int tc_mode_of(const umx_handle* h, const Op& op) {
    if (h->precision == UMX_PREC_FP32 || op.kind != OP_CONV) return TC_NONE;
    const ConvSpec& sp = op.spec;
    const TermHost& T0 = sp.terms[0];
    if (T0.k != 1 && T0.k != 3 && T0.k != 5) return TC_NONE;
    if (sp.transpose && (sp.terms.size() != 1 || T0.k == 1)) return TC_NONE;
    const Buffer& a = h->bufs[T0.src0];
    if (a.h < 4 || a.w < 4 || a.c < 8) return TC_NONE;
    if (T0.src1 >= 0 && h->bufs[T0.src1].c < 8) return TC_NONE;
    if (sp.terms.size() == 1) return TC_PLAIN;
    if (sp.terms.size() == 2) {
        const TermHost& T1 = sp.terms[1];
        if (T1.k != 1 || T1.src1 >= 0 || T0.src1 >= 0 || T0.k == 1) return TC_NONE;
        const Buffer& sc = h->bufs[T1.src0];
        if (sc.c >= 8) return TC_SHORT_A1;
        return sc.c == 1 ? TC_SHORT_SKIP : TC_NONE;
    }
    return TC_NONE;
}

// Tensor-path form of the convolutions that read the raw network input (1-2 channels, too narrow for a TMA box):
// the input is expanded once per tile into a k*k*C-channel fp16 "taps" buffer (im2col of the SAME-padded tile,
// channel = tap*C + c) by taps_kernel; then
//   first layer  conv_k(input)            ->  1x1 conv of the taps buffer (K = k*k*C)
//   lu0.conv2    conv_k(concat(input,up)) ->  conv_k(up) + 1x1 conv of the taps buffer at the centre tap
// which are the same linear maps with the same weights (HWIO rows [tap][c] are already in taps-channel order).
bool first_eligible(const umx_handle* h, const Op& op);

int rewrite_narrow_sources(umx_handle* h) {
    if (h->precision == UMX_PREC_FP32) return UMX_OK;
    std::map<std::pair<int, int>, int> taps_of;       // (source buffer, k) -> taps buffer
    std::vector<Op> ops;
    auto taps_buffer = [&](int src, int k) {
        auto key = std::make_pair(src, k);
        auto it = taps_of.find(key);
        if (it != taps_of.end()) return it->second;
        const Buffer sb = h->bufs[src];
        const int tb = new_buffer(h, sb.name + ".taps" + std::to_string(k), sb.h, sb.w, k * k * sb.c);
        Op op; op.kind = OP_TAPS; op.name = h->bufs[tb].name; op.out_buf = tb; op.taps_src = src; op.taps_k = k;
        op.flops_per_tile = 0;
        op.bytes_per_tile = (double)sb.h * sb.w * (4.0 * sb.c + 4.0 * k * k * sb.c);
        ops.push_back(op);
        taps_of[key] = tb;
        return tb;
    };
    for (auto& op : h->ops) {
        if (op.kind == OP_CONV && !op.spec.transpose && op.spec.terms.size() == 1) {
            TermHost& T = op.spec.terms[0];
            const Buffer& a = h->bufs[T.src0];
            const bool narrow = a.c <= 2 && (T.k == 3 || T.k == 5) && a.h >= 4 && a.w >= 4;
            if (narrow && T.src1 < 0) {
                if (!first_eligible(h, op)) {              // (the fp32 first-layer kernel is faster when it applies: K = k*k*C is tiny)
                    T.src0 = taps_buffer(T.src0, T.k);
                    T.k = 1;                               // [tap][c][cout] is already [1][tap*C + c][cout]
                }
            } else if (narrow && h->bufs[T.src1].c >= 8) {
                const int sc = a.c, wc = h->bufs[T.src1].c, taps = T.k * T.k, cout = op.spec.cout;
                TermHost t1; t1.k = 1; t1.src0 = taps_buffer(T.src0, T.k);
                t1.w.resize((size_t)taps * sc * cout);
                std::vector<float> w0((size_t)taps * wc * cout);
                for (int t = 0; t < taps; ++t) {
                    memcpy(&t1.w[(size_t)t * sc * cout], &T.w[(size_t)t * (sc + wc) * cout], (size_t)sc * cout * sizeof(float));
                    memcpy(&w0[(size_t)t * wc * cout], &T.w[((size_t)t * (sc + wc) + sc) * cout], (size_t)wc * cout * sizeof(float));
                }
                T.w = w0; T.src0 = T.src1; T.src1 = -1;
                op.spec.terms.push_back(t1);
            }
        }
        ops.push_back(std::move(op));
    }
    h->ops = std::move(ops);
    return UMX_OK;
}

int pick_n_tile(int cout) {
    const int c16 = (cout + 15) & ~15;
    for (int n = 256; n >= 16; n -= 16)
        if (c16 % n == 0) return n;
    return 16;
}

// fp32 [tap][cin][cout] -> fp16 [plane][tap][cout][cin] (hi, lo = fp16(w - hi))
void split_weights(const std::vector<float>& w, int taps, int cin, int cout, int planes, std::vector<__half>* out) {
    out->assign((size_t)planes * taps * cout * cin, __float2half(0.f));
    const size_t plane = (size_t)taps * cout * cin;
    for (int t = 0; t < taps; ++t)
        for (int i = 0; i < cin; ++i)
            for (int o = 0; o < cout; ++o) {
                const float v = w[((size_t)t * cin + i) * cout + o];
                const __half hi = __float2half_rn(v);
                const size_t idx = ((size_t)t * cout + o) * cin + i;
                (*out)[idx] = hi;
                if (planes == 2) (*out)[plane + idx] = __float2half_rn(v - __half2float(hi));
            }
}

int lower_conv_simt(umx_handle* h, Op& op) {
    const ConvSpec& sp = op.spec;
    ConvParams& cp = op.cp;
    memset(&cp, 0, sizeof(cp));
    const Buffer& in0 = h->bufs[sp.terms[0].src0];
    cp.nterms = (int)sp.terms.size();
    cp.in_h = in0.h; cp.in_w = in0.w; cp.cout = sp.cout;
    cp.os = sp.transpose ? 2 : 1; cp.nphase = sp.transpose ? 4 : 1;
    cp.act = sp.act; cp.leaky = kLeaky; cp.pool = sp.pool ? 1 : 0;
    pick_patch(&cp);
    for (int t = 0; t < cp.nterms; ++t) {
        ConvTerm& T = cp.term[t];
        const Buffer& a = h->bufs[sp.terms[t].src0];
        T.c0 = a.c; T.src0 = a.d;
        if (sp.terms[t].src1 >= 0) { T.c1 = h->bufs[sp.terms[t].src1].c; T.src1 = h->bufs[sp.terms[t].src1].d; }
        if (!T.src0 || (sp.terms[t].src1 >= 0 && !T.src1)) { set_error("%s: fp32 source buffer missing", op.name.c_str()); return UMX_EINVAL; }
        if (sp.transpose) convt_taps(&T, sp.terms[t].k); else conv_taps(&T, sp.terms[t].k);
        float* dw = nullptr;
        UMX_TRY(upload(h, sp.terms[t].w, &dw));
        T.w = dw;
    }
    float* d = nullptr;
    if (sp.has_bias) { UMX_TRY(upload(h, sp.bias, &d)); cp.bias = d; }
    if (sp.has_post) {
        UMX_TRY(upload(h, sp.post_scale, &d)); cp.post_scale = d;
        UMX_TRY(upload(h, sp.post_shift, &d)); cp.post_shift = d;
    }
    const Buffer& ob = h->bufs[op.out_buf];
    cp.out = ob.d; cp.out_h = ob.dh; cp.out_planes = ob.planes; cp.out_plane_elems = ob.plane_elems;
    cp.out_cs = ob.cs();
    if (conv_simt_smem_bytes(cp) > 100 * 1024) { set_error("%s: shared-memory tile too large", op.name.c_str()); return UMX_EINVAL; }
    return UMX_OK;
}

bool first_eligible(const umx_handle* h, const Op& op) {
    if (op.kind != OP_CONV) return false;
    const ConvSpec& sp = op.spec;
    if (sp.terms.size() != 1 || sp.has_post || sp.terms[0].src1 >= 0 || sp.transpose) return false;
    const int k = sp.terms[0].k;
    if (k != 3 && k != 5) return false;
    const Buffer& a = h->bufs[sp.terms[0].src0];
    return a.c <= 2 && a.h % 32 == 0 && a.h == a.w && first_conv_smem_bytes(a.c, k, sp.cout) <= 96 * 1024;
}

int lower_conv_first(umx_handle* h, Op& op) {
    const ConvSpec& sp = op.spec;
    FirstParams& fp = op.fp;
    memset(&fp, 0, sizeof(fp));
    const Buffer& a = h->bufs[sp.terms[0].src0];
    if (!a.d) { set_error("%s: fp32 source buffer missing", op.name.c_str()); return UMX_EINVAL; }
    fp.src = a.d; fp.S = a.h; fp.cin = a.c; fp.cout = sp.cout; fp.act = sp.act; fp.leaky = kLeaky;
    fp.ks = sp.terms[0].k; fp.pool = sp.pool ? 1 : 0;
    float* d = nullptr;
    UMX_TRY(upload(h, sp.terms[0].w, &d)); fp.w = d;
    if (sp.has_bias) { UMX_TRY(upload(h, sp.bias, &d)); fp.bias = d; }
    const Buffer& ob = h->bufs[op.out_buf];
    fp.out = ob.d; fp.out_h = ob.dh; fp.out_planes = ob.planes; fp.out_plane_elems = ob.plane_elems; fp.out_cs = ob.cs();
    return UMX_OK;
}

int lower_conv_tc(umx_handle* h, Op& op) {
    const ConvSpec& sp = op.spec;
    TcConvParams& tp = op.tcp;
    memset(&tp, 0, sizeof(tp));
    const TermHost& T = sp.terms[0];
    const int mode = op.tc_mode, k = T.k, ntaps_w = k * k;
    const Buffer& a0 = h->bufs[T.src0];
    const Buffer* a1 = nullptr;
    if (mode == TC_PLAIN && T.src1 >= 0) a1 = &h->bufs[T.src1];
    if (mode == TC_SHORT_A1) a1 = &h->bufs[sp.terms[1].src0];
    const Buffer* narrow = mode == TC_SHORT_SKIP ? &h->bufs[sp.terms[1].src0] : nullptr;
    const int planes = op_planes(h, op);
    tp.in_h = a0.h; tp.in_w = a0.w;
    tp.bw = std::min(a0.w, 16); tp.bh = std::min(a0.h, 128 / tp.bw); tp.bn = 128 / (tp.bw * tp.bh);
    tp.c0 = a0.cs(); tp.c1 = a1 ? a1->cs() : 0;          // storage channels (zero-padded to a multiple of 8)
    tp.cout = sp.cout; tp.n_t = pick_n_tile(sp.cout); tp.n_ntiles = (sp.cout + tp.n_t - 1) / tp.n_t;
    tp.nphase = sp.transpose ? 4 : 1; tp.os = sp.transpose ? 2 : 1;
    tp.a1_center = mode == TC_SHORT_A1 ? 1 : 0;
    tp.center_tap = ntaps_w / 2;
    ConvTerm tt; memset(&tt, 0, sizeof(tt));
    if (sp.transpose) convt_taps(&tt, k); else conv_taps(&tt, k);
    for (int ph = 0; ph < tp.nphase; ++ph) {
        TcPhaseGrid& g = tp.grid[ph];
        const int nt = tt.ntaps[ph];
        int nx = 1;
        while (nx < nt && tt.dy[ph][nx] == tt.dy[ph][0]) ++nx;
        g.ntaps = nt; g.nx = nx; g.dy0 = tt.dy[ph][0]; g.dx0 = tt.dx[ph][0]; g.wi0 = tt.wi[ph][0];
        g.dstep = nx > 1 ? tt.dx[ph][1] - tt.dx[ph][0] : (nt > nx ? tt.dy[ph][nx] - tt.dy[ph][0] : 1);
        g.wix = nx > 1 ? tt.wi[ph][1] - tt.wi[ph][0] : 0;
        g.wiy = nt > nx ? tt.wi[ph][nx] - tt.wi[ph][0] : 0;
        for (int i = 0; i < nt; ++i) {          // the tap list must be exactly this grid
            const int iy = i / nx, ix = i % nx;
            if (nt % nx || tt.dy[ph][i] != g.dy0 + g.dstep * iy || tt.dx[ph][i] != g.dx0 + g.dstep * ix ||
                tt.wi[ph][i] != g.wi0 + iy * g.wiy + ix * g.wix) {
                set_error("%s: taps of phase %d are not a regular grid", op.name.c_str(), ph); return UMX_EINVAL;
            }
        }
    }
    tp.planes = planes;
    tp.terms0 = op_terms(h, op) & 3; tp.terms1 = (op_terms(h, op) >> 2) & 3;
    if (!a1) tp.terms1 = tp.terms0;             // single source: one setting
    if (const char* e = getenv("UMX_TC_TERMS")) {          // experiment: "name:t0:t1;name:t0:t1" picks the correction terms per source
        std::string spec(e);
        size_t pos = 0;
        while (pos < spec.size()) {
            const size_t end = std::min(spec.find(';', pos), spec.size());
            const std::string item = spec.substr(pos, end - pos);
            const size_t c1 = item.find(':'), c2 = item.find(':', c1 + 1);
            if (c1 != std::string::npos && c2 != std::string::npos && item.substr(0, c1) == op.name.substr(0, c1) &&
                (op.name.size() == c1 || op.name[c1] == '+')) {
                tp.terms0 = atoi(item.substr(c1 + 1, c2 - c1 - 1).c_str()) & 3; tp.terms1 = atoi(item.substr(c2 + 1).c_str()) & 3;
            }
            pos = end + 1;
        }
    }
    tp.planes_a = ((tp.terms0 | tp.terms1) & 2) ? 2 : 1; tp.planes_b = ((tp.terms0 | tp.terms1) & 1) ? 2 : 1;
    {
        const char* e = getenv("UMX_TC_PAIR");
        tp.pair = (e ? atoi(e) : 1) && ((tp.n_t / 2) % 8 == 0) ? 1 : 0;
    }
    // halo mode for the high-resolution layers: A is fetched once per 64-channel slab as a pixel patch
    // (instead of once per tap), cutting the L2->SM traffic of A by ~k*k/1.4
    {
        const char* e = getenv("UMX_TC_HALO");
        if ((e ? atoi(e) : 1) && a0.w >= 16 && a0.h >= 16 && ntaps_w > 1) {
            tp.halo = 1; tp.bw = 8; tp.bh = 16; tp.bn = 1;
            tp.hx0 = tt.hx0; tp.hy0 = tt.hy0;
            tp.pw = tp.bw + tt.hx0 + tt.hx1; tp.ph = tp.bh + tt.hy0 + tt.hy1;
        }
        // 8x8 grids: two tiles per GEMM tile, patch rows interleaved [h][tile][w] (needs a tensor map whose tile
        // dimension comes before the row dimension; probed here, plain mode if the driver refuses it)
        const char* e8 = getenv("UMX_TC_HALO8");
        if ((e ? atoi(e) : 1) && (e8 ? atoi(e8) : 1) && a0.w == 8 && a0.h == 8 && ntaps_w > 1 && !sp.transpose) {       // (conv-transpose measured slower: one patch per phase)
            CUtensorMap probe;
            const int pw = 8 + tt.hx0 + tt.hx1, ph = 8 + tt.hy0 + tt.hy1;
            if (a0.dh && make_act_tensor_map(&probe, a0.dh, a0.planes, a0.plane_elems, h->cap_tiles, a0.h, a0.w, a0.cs(), pw, ph, 2, 1, 1) == 0) {
                tp.halo = 1; tp.halo_nh = 1; tp.bw = 8; tp.bh = 8; tp.bn = 2;
                tp.hx0 = tt.hx0; tp.hy0 = tt.hy0; tp.pw = pw; tp.ph = ph;
            }
        }
    }
    {
        const size_t cpad = (size_t)tp.n_ntiles * tp.n_t;
        const size_t tables = (cpad * ((sp.has_post ? 2 : 0) + (narrow ? 1 : 0) + (op.fuse_top >= 0 ? 4 : 0)) +
                               (op.fuse_top >= 0 ? 2 * 3 * 128 * 4 : 0) + 8) * 4 + 64;
        const size_t budget = 227 * 1024 - 2048 - 512 - tables;
        const size_t bb = tc_conv_b_bytes(tp);
        if (tp.halo) {
            const size_t ab = tc_conv_a_bytes(tp);
            tp.stages = 3;          // patch slots: the next slab's patch is in flight while this one is multiplied
            if (const char* e = getenv("UMX_TC_ASTAGES")) tp.stages = std::max(2, std::min(8, atoi(e)));
            // taps per weight slot: as many as still leave three slots in flight (every slot costs a barrier round trip
            // and a tcgen05.commit in the single issuing thread, so few large slots beat many small ones)
            int max_taps = 1;
            for (int ph = 0; ph < tp.nphase; ++ph) max_taps = std::max(max_taps, (int)tp.grid[ph].ntaps);
            tp.gb = std::min(9, max_taps);
            while (tp.gb > 1 && budget < tp.stages * ab + 3 * (size_t)tp.gb * bb) tp.gb--;
            if (const char* e = getenv("UMX_TC_GB")) tp.gb = std::max(1, std::min(9, atoi(e)));
            while (tp.stages > 2 && budget < tp.stages * ab + 2 * (size_t)tp.gb * bb) tp.stages--;
            while (tp.gb > 1 && budget < tp.stages * ab + 2 * (size_t)tp.gb * bb) tp.gb--;
            tp.b_stages = budget > tp.stages * ab ? (int)std::min<size_t>(6, (budget - tp.stages * ab) / ((size_t)tp.gb * bb)) : 0;
            // weight-stationary: when every weight tile of the layer (all taps of every slab, this CTA's share of the
            // N rows) fits next to >= 3 patch slots, it is loaded once per CTA and the per-item weight traffic, barrier
            // round trips and commits disappear (the 64x64 layers of the v2 graphs in single precision)
            {
                const int nc0 = (tp.c0 + 63) / 64, nc1 = (tp.c1 + 63) / 64, n_chunks = nc0 + nc1;
                // (the slabs of a 1x1 term exist at the centre tap only: one tile each instead of ntaps_w)
                const size_t tile = (size_t)(tp.pair ? tp.n_t / 2 : tp.n_t) * 128;
                // full slabs keep the lo plane too when some full-slab source issues the a_hi*w_lo term (narrow layers: it still fits)
                tp.res_m_planes = ((tp.terms0 & 1) || (!tp.a1_center && nc1 > 0 && (tp.terms1 & 1))) ? 2 : 1;
                tp.res_c_planes = (tp.a1_center && (tp.terms1 & 1)) ? 2 : 1;
                const size_t res = ((size_t)nc0 * ntaps_w * tp.res_m_planes + (size_t)nc1 * (tp.a1_center ? tp.res_c_planes : ntaps_w * tp.res_m_planes)) * tile;
                const char* e = getenv("UMX_TC_RESIDENT");
                // N-concatenated correction (narrow layers, CTA pairs): one MMA of width 2*n_t multiplies a_hi by [w_hi | w_lo]
                // (the even CTA supplies w_hi as its half of the B rows, the odd CTA w_lo) and the epilogue adds the two halves
                // of the accumulator: a product with the a_hi*w_lo term costs one MMA less, and below N ~ 110 an MMA costs the
                // single issuing thread the same ~55 cycles whatever its width.  Per tap a slab keeps [X: n_t rows | Y: n_t/2
                // rows of w_hi for the a_lo*w_hi term] = three tiles.
                const char* en = getenv("UMX_TC_NCAT");
                const char* em = getenv("UMX_TC_MERGE_PX");
                const bool will_merge = sp.transpose && tp.nphase == 4 && 2 * tp.n_t <= 256 && !narrow && op.fuse_top < 0 && (em ? atoi(em) : 1);
                const size_t res3 = ((size_t)nc0 * ntaps_w * 3 + (size_t)nc1 * (tp.a1_center ? tp.res_c_planes : ntaps_w * 3)) * tile;
                // (measured, Cyto2: lu0.conv2+lt, N = 32, 120 -> 96 ms; the conv-transposes - few MMAs per accumulator, epilogue-bound -
                // lose 20 % to the second TMEM read, and from N = 64 on the wide MMA costs what two narrow ones do: not used there.
                // The path is compiled into the kernels with the fused lt epilogue only: in the others the extra epilogue code
                // alone cost the conv-transposes 7 %)
                const bool ncat_pays = op.fuse_top >= 0 && !sp.transpose && tp.n_t <= 48;
                if ((e ? atoi(e) : 1) && (en ? atoi(en) != 0 && ncat_pays : ncat_pays) && tp.pair && tp.planes_b == 2 &&
                    ((tp.terms0 | tp.terms1) & 1) && tp.n_ntiles == 1 &&
                    ntaps_w <= 9 && 2 * tp.n_t * (will_merge ? 2 : 1) <= 256 && budget >= res3 + (tp.planes_a == 2 ? 2 : 3) * ab) {
                    tp.ncat = 1; tp.res_m_planes = 3;
                    tp.b_resident = 1; tp.gb = ntaps_w; tp.b_stages = n_chunks; tp.b_res_bytes = (int32_t)res3;
                    tp.stages = (int)std::min<size_t>(6, (budget - res3) / ab);
                }
                // (with both activation planes in a patch slot two slots are accepted - in this mode nothing but patches
                // moves, so no patch waits behind a weight load)
                else if ((e ? atoi(e) : 1) && tp.n_ntiles == 1 && ntaps_w <= 9 && tp.res_m_planes <= tp.planes_b && budget >= res + (tp.planes_a == 2 ? 2 : 3) * ab) {
                    tp.b_resident = 1; tp.gb = ntaps_w; tp.b_stages = n_chunks; tp.b_res_bytes = (int32_t)res;
                    tp.stages = (int)std::min<size_t>(6, (budget - res) / ab);
                }
            }
            // the producer runs the patch loads stages-1 slabs ahead of the weight loads of the same slab stream; with only
            // two patch slots the patch of slab s+1 would queue behind weights that cannot all land before that very
            // patch is consumed (a weight ring shorter than a slab's taps): a deadlock, so such layers run in plain mode
            if (!tp.b_resident && (tp.b_stages < 2 || tp.stages < 3)) { tp.halo = 0; tp.halo_nh = 0; tp.bw = std::min(a0.w, 16); tp.bh = std::min(a0.h, 128 / tp.bw); tp.bn = 128 / (tp.bw * tp.bh); }
        }
        {
            const char* e = getenv("UMX_TC_MERGE_PX");
            tp.merge_px = (tp.halo && sp.transpose && tp.nphase == 4 && 2 * tp.n_t <= 256 && !narrow && op.fuse_top < 0 && (e ? atoi(e) : 1)) ? 1 : 0;
        }
        tp.kslab = 1;
        if (!tp.halo) {
            // single precision issues one MMA per K step: put two slabs behind each barrier round trip
            const int n_chunks = (tp.c0 + 63) / 64 + (tp.a1_center ? 0 : (tp.c1 + 63) / 64);
            tp.kslab = std::min(planes == 1 ? 2 : 1, std::max(1, n_chunks));
            if (const char* e = getenv("UMX_TC_KSLAB")) tp.kslab = std::max(1, std::min(std::min(4, n_chunks), atoi(e)));
            while (tp.kslab > 1 && budget / (tp.kslab * (tc_conv_a_bytes(tp) + bb)) < 3) tp.kslab--;
            tp.stages = (int)std::min<size_t>(6, budget / (tp.kslab * (tc_conv_a_bytes(tp) + bb)));
        }
    }
    if (const char* e = getenv("UMX_TC_EXP")) tp.exp_flags = atoi(e);
    tp.epi_nap_ns = 0;
    if (const char* e = getenv("UMX_TC_NAP")) tp.epi_nap_ns = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("UMX_TC_STAGES")) {
        const int cap = std::max(2, atoi(e));
        if (tp.halo) { if (!tp.b_resident) tp.b_stages = std::min(tp.b_stages, cap); } else tp.stages = std::min(tp.stages, cap);
    }
    if (tp.stages < 2) { set_error("%s: pipeline does not fit shared memory", op.name.c_str()); return UMX_EINVAL; }
    tp.act = sp.act; tp.leaky = kLeaky; tp.pool = sp.pool ? 1 : 0;
    float* d = nullptr;
    if (tp.n_ntiles * tp.n_t > kTcMaxCols) { set_error("%s: %d output channels exceed the tensor path's tables", op.name.c_str(), sp.cout); return UMX_EINVAL; }
    if (sp.has_bias) for (int i = 0; i < sp.cout; ++i) tp.tab_bias[i] = sp.bias[i];
    if (sp.has_post) {
        UMX_TRY(upload(h, sp.post_scale, &d)); tp.post_scale = d;
        UMX_TRY(upload(h, sp.post_shift, &d)); tp.post_shift = d;
    }
    const Buffer& ob = h->bufs[op.out_buf];
    tp.out_f = ob.d; tp.out_h = ob.dh; tp.out_planes = ob.planes; tp.out_plane_elems = ob.plane_elems; tp.out_cs = ob.cs();
    // weights -> fp16 planes on the device; input-channel rows follow the padded storage layout [c0s | c1s]
    const int cin = tp.c0 + tp.c1;
    std::vector<float> wmain((size_t)ntaps_w * cin * sp.cout, 0.f);
    auto put_rows = [&](const std::vector<float>& w, int w_taps, int w_cin, int row0, int nrows, int dst0, int tap_dst0) {
        for (int t = 0; t < w_taps; ++t)
            for (int i = 0; i < nrows; ++i)
                memcpy(&wmain[((size_t)(tap_dst0 + t) * cin + dst0 + i) * sp.cout], &w[((size_t)t * w_cin + row0 + i) * sp.cout],
                       sp.cout * sizeof(float));
    };
    if (mode == TC_PLAIN) {
        const int r0 = a0.c, r1 = a1 ? a1->c : 0;
        put_rows(T.w, ntaps_w, r0 + r1, 0, r0, 0, 0);
        if (r1) put_rows(T.w, ntaps_w, r0 + r1, r0, r1, tp.c0, 0);
    } else if (mode == TC_SHORT_SKIP) {
        put_rows(T.w, ntaps_w, a0.c, 0, a0.c, 0, 0);
    } else {    // TC_SHORT_A1: the 1x1 shortcut's rows live at the centre tap only
        put_rows(T.w, ntaps_w, a0.c, 0, a0.c, 0, 0);
        put_rows(sp.terms[1].w, 1, a1->c, 0, a1->c, tp.c0, tp.center_tap);
    }
    if (narrow) {       // one-channel 1x1 shortcut: weights [1][1][cout]
        UMX_TRY(upload(h, sp.terms[1].w, &d));
        tp.skip_w = d; tp.skip_c = 1; tp.skip_src = narrow->d;
        if (!narrow->d) { set_error("%s: fp32 skip source missing", op.name.c_str()); return UMX_EINVAL; }
    }
    if (op.fuse_top >= 0) {
        Op& top = h->ops[op.fuse_top];
        UMX_TRY(upload(h, top.top_w, &d)); tp.top_w = d;          // (pointer only marks the fusion; the kernel reads the tables below)
        tp.top_k = h->K;
        if (sp.cout > 256) { set_error("%s: fused lt needs <= 256 channels", op.name.c_str()); return UMX_EINVAL; }
        for (int c = 0; c < sp.cout; ++c)
            for (int k = 0; k < h->K; ++k) tp.tab_topw[c * 4 + k] = top.top_w[(size_t)c * h->K + k];
        for (int k = 0; k < h->K && k < (int)top.top_b.size(); ++k) tp.tab_topb[k] = top.top_b[k];
    }
    // Device tap order = phase by phase (the taps of a phase are contiguous, so halo mode fetches a whole group of
    // taps with one TMA box), followed by `gb` all-zero taps so a box that starts at the last tap never leaves the tensor.
    std::vector<int> order;
    for (int ph = 0; ph < tp.nphase; ++ph) {
        TcPhaseGrid& g = tp.grid[ph];
        const int start = (int)order.size();
        for (int i = 0; i < g.ntaps; ++i) order.push_back(g.wi0 + (i / g.nx) * g.wiy + (i % g.nx) * g.wix);
        g.wi0 = start; g.wix = 1; g.wiy = g.nx;
    }
    if ((int)order.size() != ntaps_w) { set_error("%s: phases do not cover every weight tap once", op.name.c_str()); return UMX_EINVAL; }
    const int box_taps = tp.halo ? tp.gb : 1, taps_dev = ntaps_w + (tp.halo ? tp.gb : 0);
    std::vector<float> wperm((size_t)taps_dev * cin * sp.cout, 0.f);
    for (int i = 0; i < ntaps_w; ++i)
        memcpy(&wperm[(size_t)i * cin * sp.cout], &wmain[(size_t)order[i] * cin * sp.cout], (size_t)cin * sp.cout * sizeof(float));
    for (int i = 0; i < ntaps_w; ++i) if (order[i] == tp.center_tap) { tp.center_tap = i; break; }
    const int planes_b = tp.planes_b, planes_a = tp.planes_a;
    std::vector<__half> wh;
    split_weights(wperm, taps_dev, cin, sp.cout, planes_b, &wh);
    __half* dw = nullptr;
    UMX_CUDA_TRY(cudaMalloc(&dw, wh.size() * sizeof(__half)));
    h->dev_allocs.push_back(reinterpret_cast<float*>(dw));
    UMX_CUDA_TRY(cudaMemcpy(dw, wh.data(), wh.size() * sizeof(__half), cudaMemcpyHostToDevice));
    int rc = make_weight_tensor_map(&op.mapB, dw, planes_b, taps_dev, sp.cout, cin, tp.pair ? tp.n_t / 2 : tp.n_t, tp.ncat ? 1 : (tp.b_resident ? tp.res_m_planes : planes_b), tp.ncat ? 1 : box_taps);     // ncat: one-tile boxes (plane, tap, half of the rows)
    if (rc) { set_error("%s: cuTensorMapEncodeTiled(weights) failed (%d)", op.name.c_str(), rc); return UMX_ECUDA; }
    op.mapB1 = op.mapB;
    if (tp.halo && !tp.b_resident && planes_b == 2 && (!(tp.terms0 & 1) || (a1 && !(tp.terms1 & 1)))) {
        // some source never multiplies the weights' lo plane: it streams one-plane boxes (the slot layout stays [hi taps][lo taps])
        rc = make_weight_tensor_map(&op.mapB1, dw, planes_b, taps_dev, sp.cout, cin, tp.pair ? tp.n_t / 2 : tp.n_t, 1, box_taps);
        if (rc) { set_error("%s: cuTensorMapEncodeTiled(weights, hi plane) failed (%d)", op.name.c_str(), rc); return UMX_ECUDA; }
        tp.b1_hi_only = 1;
    }
    if (tp.b_resident && tp.a1_center && !tp.ncat) {
        rc = make_weight_tensor_map(&op.mapB1, dw, planes_b, taps_dev, sp.cout, cin, tp.pair ? tp.n_t / 2 : tp.n_t, tp.res_c_planes, 1);
        if (rc) { set_error("%s: cuTensorMapEncodeTiled(weights, one tap) failed (%d)", op.name.c_str(), rc); return UMX_ECUDA; }
    }
    if (!a0.dh || (a1 && !a1->dh)) { set_error("%s: fp16 source buffer missing", op.name.c_str()); return UMX_EINVAL; }
    if (!tp.halo && planes_a == 2 && (a0.planes < 2 || (a1 && a1->planes < 2))) { set_error("%s: plain-mode op needs the lo plane of every source", op.name.c_str()); return UMX_EINVAL; }
    if (tp.halo && (((tp.terms0 & 2) && a0.planes < 2) || (a1 && (tp.terms1 & 2) && a1->planes < 2))) { set_error("%s: lo plane of a source is missing", op.name.c_str()); return UMX_EINVAL; }
    const int box_w = tp.halo ? tp.pw : tp.bw, box_h = tp.halo ? tp.ph : tp.bh, box_p = tp.halo ? 1 : planes_a;
    rc = make_act_tensor_map(&op.mapA0, a0.dh, a0.planes, a0.plane_elems, h->cap_tiles, a0.h, a0.w, a0.cs(), box_w, box_h, tp.bn, box_p, tp.halo_nh);
    if (rc) { set_error("%s: cuTensorMapEncodeTiled(A0) failed (%d)", op.name.c_str(), rc); return UMX_ECUDA; }
    if (a1) {
        rc = make_act_tensor_map(&op.mapA1, a1->dh, a1->planes, a1->plane_elems, h->cap_tiles, a1->h, a1->w, a1->cs(), box_w, box_h, tp.bn, box_p, tp.halo_nh);
        if (rc) { set_error("%s: cuTensorMapEncodeTiled(A1) failed (%d)", op.name.c_str(), rc); return UMX_ECUDA; }
    } else {
        op.mapA1 = op.mapA0;
    }
    return UMX_OK;
}

// Decide per op which kernel runs it, which formats every buffer must exist in, allocate the
// workspace and bind device pointers / tensor maps.
int lower_plan(umx_handle* h) {
    for (size_t i = 0; i < h->ops.size(); ++i) {
        Op& op = h->ops[i];
        op.tc_mode = tc_mode_of(h, op);
        op.use_tc = op.tc_mode != TC_NONE;
        op.use_first = !op.use_tc && first_eligible(h, op);
    }
    // lt 1x1 conv + softmax rides in the epilogue of the conv that feeds it when one CTA tile spans all channels
    for (size_t i = 1; i < h->ops.size(); ++i) {
        Op& top = h->ops[i];
        Op& prev = h->ops[i - 1];
        if (top.kind == OP_TOP && prev.kind == OP_CONV && prev.use_tc && prev.out_buf == top.top_src &&
            pick_n_tile(prev.spec.cout) >= prev.spec.cout && !prev.spec.pool && !prev.spec.transpose) {
            prev.fuse_top = (int)i;
            top.fused_away = true;
            const Buffer& ob = h->bufs[prev.out_buf];
            prev.flops_per_tile += top.flops_per_tile;
            prev.bytes_per_tile += 4.0 * ob.h * ob.w * (h->K - ob.c);     // writes K probabilities instead of cout activations
            prev.name += "+lt";
        }
    }
    for (auto& op : h->ops) {
        if (op.kind == OP_CONV) {
            for (size_t ti = 0; ti < op.spec.terms.size(); ++ti) {
                const TermHost& t = op.spec.terms[ti];
                if (op.tc_mode == TC_SHORT_SKIP && ti == 1) { h->bufs[t.src0].need_f = true; continue; }
                // the lo plane of a buffer exists only where a consumer multiplies it (a_lo * w_hi term of that source);
                // source 1 of the kernel = the second concat source, or the 1x1 term's buffer (TC_SHORT_A1)
                const int terms = op_terms(h, op);
                const bool two_src = t.src1 >= 0 || op.tc_mode == TC_SHORT_A1;
                const int si[2] = {t.src0, t.src1};
                for (int j = 0; j < 2; ++j) {
                    const int s = si[j];
                    if (s < 0) continue;
                    if (op.use_tc) {
                        h->bufs[s].need_h = true;
                        const int ts = !two_src ? (terms & 3) : ((ti == 1 || j == 1) ? (terms >> 2) & 3 : terms & 3);
                        // plain (non-halo) mode, i.e. grids below 16 x 16, fetches the planes of both sources with one box shape:
                        // there any a_lo term asks for the lo plane of every source
                        const bool any_a_lo = ((terms | (terms >> 2)) & 2) != 0;
                        if ((ts & 2) || (any_a_lo && h->bufs[t.src0].h < 16)) h->bufs[s].need_lo = true;
                    } else h->bufs[s].need_f = true;
                }
            }
        } else if (op.kind == OP_TAPS) {
            h->bufs[op.taps_src].need_f = true;
            h->bufs[op.out_buf].need_h = true;
        } else if (!op.fused_away) {
            h->bufs[op.top_src].need_f = true;
        }
    }
    h->bufs[h->in_buf].need_f = true;
    for (auto& b : h->bufs) {
        if (b.need_f) UMX_CUDA_TRY(cudaMalloc(&b.d, (size_t)h->cap_tiles * b.per_tile() * sizeof(float)));
        if (b.need_h) {
            const int planes = b.need_lo ? 2 : 1;         // the lo plane exists only where a split consumer reads it
            b.planes = planes; b.plane_elems = (int64_t)h->cap_tiles * b.per_tile_h();
            UMX_CUDA_TRY(cudaMalloc(&b.dh, (size_t)planes * b.plane_elems * sizeof(__half)));
            UMX_CUDA_TRY(cudaMemset(b.dh, 0, (size_t)planes * b.plane_elems * sizeof(__half)));
        }
    }
    UMX_CUDA_TRY(cudaMalloc(&h->probs, (size_t)h->cap_tiles * h->S * h->S * h->K * sizeof(float)));
    for (auto& op : h->ops) {
        if (op.kind == OP_CONV) {
            if (op.use_tc) UMX_TRY(lower_conv_tc(h, op));
            else if (op.use_first) UMX_TRY(lower_conv_first(h, op));
            else UMX_TRY(lower_conv_simt(h, op));
            op.spec = ConvSpec();       // host copies of the weights are no longer needed
        } else if (op.kind == OP_TAPS) {
            const Buffer& sb = h->bufs[op.taps_src];
            const Buffer& ob = h->bufs[op.out_buf];
            op.taps.src = sb.d; op.taps.out = ob.dh; op.taps.out_plane_elems = ob.plane_elems; op.taps.out_planes = ob.planes;
            op.taps.S = sb.h; op.taps.cin = sb.c; op.taps.ks = op.taps_k; op.taps.cs = ob.cs();
            op.bytes_per_tile = (double)sb.h * sb.w * (4.0 * sb.c + 2.0 * ob.planes * ob.cs());
        } else if (!op.fused_away) {
            float* d = nullptr;
            UMX_TRY(upload(h, op.top_w, &d)); op.tp.w = d;
            if (!op.top_b.empty()) { UMX_TRY(upload(h, op.top_b, &d)); op.tp.bias = d; }
            op.tp.src = h->bufs[op.top_src].d;
        }
        ProfSlot ps; ps.name = op.name;
        op.prof_slot = (int)h->prof.size();
        h->prof.push_back(ps);
    }
    // the first-layer kernel already holds every input patch in registers: let it write the tap expansion too
    for (auto& t : h->ops) {
        if (t.kind != OP_TAPS) continue;
        for (auto& f : h->ops) {
            if (f.kind != OP_CONV || !f.use_first || f.fp.src != t.taps.src || f.fp.ks != t.taps.ks || f.fp.cin != t.taps.cin) continue;
            f.fp.taps_out = t.taps.out; f.fp.taps_plane_elems = t.taps.out_plane_elems;
            f.fp.taps_planes = t.taps.out_planes; f.fp.taps_cs = t.taps.cs;
            f.bytes_per_tile += t.bytes_per_tile;
            f.name += "+taps"; h->prof[f.prof_slot].name = f.name;
            t.fused_away = true;
            break;
        }
    }
    // can the tile gather be folded into the first layer?  Only if nothing else reads the gathered input buffer.
    {
        const float* in = h->bufs[h->in_buf].d;
        int readers = 0, first_reader = -1;
        for (size_t i = 0; i < h->ops.size(); ++i) {
            const Op& op = h->ops[i];
            if (op.fused_away) continue;
            bool reads = false;
            if (op.kind == OP_CONV && op.use_first) reads = op.fp.src == in;
            else if (op.kind == OP_CONV && op.use_tc) reads = op.tcp.skip_src == in;
            else if (op.kind == OP_CONV) { for (int t = 0; t < op.cp.nterms; ++t) reads |= op.cp.term[t].src0 == in || op.cp.term[t].src1 == in; }
            else if (op.kind == OP_TAPS) reads = op.taps.src == in;
            else reads = op.tp.src == in;
            if (reads) { ++readers; if (first_reader < 0) first_reader = (int)i; }
        }
        const char* e = getenv("UMX_FUSE_GATHER");
        h->fuse_gather = (e ? atoi(e) : 1) && readers == 1 && first_reader >= 0 && h->ops[first_reader].kind == OP_CONV && h->ops[first_reader].use_first;
    }
    return UMX_OK;
}

cudaEvent_t grab_event(umx_handle* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ScopedTimer {
    umx_handle* h; int slot; double flops, bytes; cudaEvent_t a = nullptr;
    ScopedTimer(umx_handle* hh, int s, double f, double b) : h(hh), slot(s), flops(f), bytes(b) {
        if (h->profiling) { a = grab_event(h); cudaEventRecord(a, h->stream); }
    }
    ~ScopedTimer() {
        if (h->profiling) {
            cudaEvent_t b = grab_event(h);
            cudaEventRecord(b, h->stream);
            h->pending.push_back({slot, a, b, flops, bytes});
        }
    }
};

int aux_slot(umx_handle* h, const char* name) {
    for (size_t i = 0; i < h->prof.size(); ++i) if (h->prof[i].name == name) return (int)i;
    ProfSlot ps; ps.name = name; h->prof.push_back(ps);
    return (int)h->prof.size() - 1;
}

void drain_profile(umx_handle* h) {
    for (auto& pe : h->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            ProfSlot& s = h->prof[pe.slot];
            s.launches += 1; s.ms += ms; s.flops += pe.flops; s.bytes += pe.bytes;
        }
        h->event_pool.push_back(pe.a); h->event_pool.push_back(pe.b);
    }
    h->pending.clear();
}

// Run the network on nb tiles already sitting in the input buffer; probs -> `probs_out` (device).
int run_network(umx_handle* h, int nb, float* probs_out, const FirstImage* fused_input = nullptr) {
    for (auto& op : h->ops) {
        if (op.fused_away) continue;
        ScopedTimer tm(h, op.prof_slot, op.flops_per_tile * nb, op.bytes_per_tile * nb + op.weight_bytes);
        if (op.kind == OP_CONV && op.use_tc) {
            TcConvParams tp = op.tcp;
            tp.n_tiles = nb;
            if (op.fuse_top >= 0) tp.top_probs = probs_out;
            if (tp.exp_flags & 64) {           // UMX_TC_EXP=64: per-role cycle accounting, printed per launch (debug only)
                if (!h->d_dbg) UMX_CUDA_TRY(cudaMalloc(&h->d_dbg, 16 * sizeof(unsigned long long)));
                UMX_CUDA_TRY(cudaMemsetAsync(h->d_dbg, 0, 16 * sizeof(unsigned long long), h->stream));
                tp.dbg = h->d_dbg;
            }
            UMX_CUDA_TRY(launch_tc_conv(op.mapA0, op.mapA1, op.mapB, op.mapB1, tp, h->num_sms, h->stream));
            if (tp.exp_flags & 64) {
                unsigned long long c[16];
                UMX_CUDA_TRY(cudaMemcpyAsync(c, h->d_dbg, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
                UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
                const double n = c[15] ? (double)c[15] : 1.0;      // MMA-issuing CTAs
                fprintf(stderr, "[umx dbg] %-14s tiles %5d halo %d pair %d stages %d/%d gb %d res %d | producer waitA %.0f waitB %.0f work %.0f | mma waitT %.0f waitA %.0f waitB %.0f work %.0f | epi waitT %.0f work %.0f (kcycles per issuing CTA)\n",
                        op.name.c_str(), nb, tp.halo, tp.pair, tp.stages, tp.b_stages, tp.gb, tp.b_resident, c[0] / n / 1e3, c[1] / n / 1e3, c[2] / n / 1e3,
                        c[4] / n / 1e3, c[5] / n / 1e3, c[6] / n / 1e3, c[7] / n / 1e3, c[8] / n / 1e3, c[9] / n / 1e3);
            }
        } else if (op.kind == OP_CONV && op.use_first) {
            FirstParams fp = op.fp;
            fp.n_tiles = nb;
            if (fused_input && fp.src == h->bufs[h->in_buf].d) fp.im = *fused_input;       // the tile gather happens inside the kernel
            UMX_CUDA_TRY(launch_first_conv(fp, h->stream));
        } else if (op.kind == OP_CONV) {
            ConvParams cp = op.cp;
            cp.n_tiles = nb;
            UMX_CUDA_TRY(launch_conv_simt(cp, h->stream));
        } else if (op.kind == OP_TAPS) {
            TapsParams tp = op.taps;
            tp.n_tiles = nb;
            UMX_CUDA_TRY(launch_taps(tp, h->stream));
        } else {
            TopParams tp = op.tp;
            tp.n_pix = (int64_t)nb * h->S * h->S;
            tp.probs = probs_out;
            UMX_CUDA_TRY(launch_top_softmax(tp, h->stream));
        }
        h->launches += 1;
    }
    return UMX_OK;
}

int ensure(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return UMX_OK;
    if (*p) UMX_CUDA_TRY(cudaFree(*p));
    *p = nullptr; *cap = 0;
    UMX_CUDA_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    return UMX_OK;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

size_t dtype_size(int dtype) {
    switch (dtype) { case UMX_U8: return 1; case UMX_U16: return 2; case UMX_F32: return 4; case UMX_F64: return 8; }
    return 0;
}

}  // namespace
}  // namespace umx

using namespace umx;

extern "C" {

const char* umx_last_error(void) { return g_err; }
const char* umx_version(void) { return "unmicst_b200 0.1 (sm_100a)"; }

int umx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int umx_device_free_mem(int device, int64_t* free_bytes, int64_t* total_bytes) {
    int cur = 0;
    UMX_CUDA_TRY(cudaGetDevice(&cur));
    UMX_CUDA_TRY(cudaSetDevice(device));
    size_t f = 0, t = 0;
    cudaError_t e = cudaMemGetInfo(&f, &t);
    cudaSetDevice(cur);
    UMX_CUDA_TRY(e);
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return UMX_OK;
}

void* umx_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        set_error("cudaHostAlloc(%lld) failed", (long long)bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void umx_host_free(void* p) { if (p) cudaFreeHost(p); }

int umx_create(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights, int32_t device, umx_handle** out) {
    return umx_create_ex(desc, weights, n_weights, device, nullptr, 0, out);
}

int umx_set_op_terms(umx_handle* h, int32_t op_index, int32_t terms) {
    if (!h || op_index < 0 || op_index >= (int)h->ops.size() || terms < 0 || terms > 15) { set_error("umx_set_op_terms: bad argument"); return UMX_EINVAL; }
    Op& op = h->ops[op_index];
    if (op.kind != OP_CONV || !op.use_tc) { set_error("op %d is not on the tensor path", op_index); return UMX_EINVAL; }
    TcConvParams& tp = op.tcp;
    if (tp.planes_a != 2 || tp.planes_b != 2 || !h->full_split) { set_error("op %d was not built with the full hi/lo split: its lo planes do not exist", op_index); return UMX_EINVAL; }
    // every plane the full split needs exists (buffers, weights, shared-memory slots), so any subset of the terms can
    // be switched on this live handle; the slot layout stays that of the full split (a calibration aid: the arithmetic
    // equals that of a handle built with these terms, the loads are those of the full split)
    tp.terms0 = terms & 3; tp.terms1 = tp.c1 > 0 ? (terms >> 2) & 3 : tp.terms0;
    return UMX_OK;
}

int umx_create_ex(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights, int32_t device,
                  const int32_t* op_terms_in, int32_t n_op_terms, umx_handle** out) {
    if (!desc || !out || (!weights && n_weights > 0)) { set_error("umx_create: null argument"); return UMX_EINVAL; }
    *out = nullptr;
    if (desc->abi_version != UMX_ABI_VERSION) { set_error("umx_create: ABI version %d != %d", desc->abi_version, UMX_ABI_VERSION); return UMX_EINVAL; }
    if (desc->graph != UMX_GRAPH_LEGACY && desc->graph != UMX_GRAPH_V2) { set_error("unknown graph %d", desc->graph); return UMX_EINVAL; }
    if (desc->down_samp_fact != 2) { set_error("downSampFact %d unsupported (only 2)", desc->down_samp_fact); return UMX_EINVAL; }
    if (desc->ks != 1 && desc->ks != 3 && desc->ks != 5) { set_error("ks %d unsupported (1, 3 or 5)", desc->ks); return UMX_EINVAL; }
    if (desc->n_classes < 2 || desc->n_classes > 4) { set_error("nClasses %d unsupported (2..4)", desc->n_classes); return UMX_EINVAL; }
    if (desc->n_channels < 1 || desc->n_out0 < 1 || desc->n_layers < 1 || desc->feat_maps_fact < 1 || desc->n_extra_convs < 0) {
        set_error("invalid hyper-parameters"); return UMX_EINVAL;
    }
    const int S = desc->im_size;
    if (S < 8 || (S & (S - 1)) || (S >> desc->n_layers) < 4) {
        set_error("imSize %d with %d layers unsupported (power of two, deepest grid >= 4)", S, desc->n_layers);
        return UMX_EINVAL;
    }
    int ndev = umx_device_count();
    if (ndev <= 0) { set_error("no CUDA device visible (this engine has no CPU path)"); return UMX_ENODEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return UMX_EINVAL; }
    cudaDeviceProp prop;
    UMX_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major, prop.minor); return UMX_ENODEVICE; }
    UMX_CUDA_TRY(cudaSetDevice(device));

    umx_handle* h = new (std::nothrow) umx_handle();
    if (!h) { set_error("out of host memory"); return UMX_ENOMEM; }
    h->device = device; h->desc = *desc;
    h->S = S; h->C = desc->n_channels; h->K = desc->n_classes; h->L = desc->n_layers;
    h->margin = S / 8; h->sub = S - 2 * h->margin;
    h->chan = {desc->n_channels, desc->n_out0};
    for (int i = 0; i < desc->n_layers; ++i) h->chan.push_back(h->chan.back() * desc->feat_maps_fact);
    for (int i = 0; i < n_weights; ++i) {
        const umx_tensor& t = weights[i];
        if (!t.name || !t.data || t.ndim < 1 || t.ndim > 4) { set_error("weights[%d] malformed", i); umx_destroy(h); return UMX_EINVAL; }
        HostTensor ht;
        for (int d = 0; d < t.ndim; ++d) ht.shape.push_back(t.shape[d]);
        ht.data.assign(t.data, t.data + ht.numel());
        h->tensors[t.name] = std::move(ht);
    }
    int rc = build_plan(h);
    if (rc != UMX_OK) { umx_destroy(h); return rc; }
    // batch: enough tiles to fill 148 SMs several times over, bounded by workspace memory
    h->num_sms = prop.multiProcessorCount;
    h->precision = desc->precision == UMX_PREC_DEFAULT ? UMX_PREC_SPLIT3 : desc->precision;
    if (h->precision == UMX_PREC_MIXED) h->single_mask = (uint64_t)(uint32_t)desc->reserved[0] | ((uint64_t)(uint32_t)desc->reserved[1] << 32);
    if (h->precision < UMX_PREC_FP32 || h->precision > UMX_PREC_MIXED) { set_error("unknown precision %d", desc->precision); umx_destroy(h); return UMX_EINVAL; }
    rc = rewrite_narrow_sources(h);
    if (rc != UMX_OK) { umx_destroy(h); return rc; }
    h->full_split = h->precision == UMX_PREC_SPLIT3 && !op_terms_in;
    if (op_terms_in && h->precision != UMX_PREC_FP32) {
        h->op_terms.assign(h->ops.size(), -1);
        for (int i = 0; i < n_op_terms && i < (int)h->ops.size(); ++i) h->op_terms[i] = op_terms_in[i] < 0 ? -1 : (int8_t)(op_terms_in[i] & 15);
    }
    int64_t per_tile = 0;
    for (auto& b : h->bufs) per_tile += b.per_tile() * 4;
    per_tile += (int64_t)S * S * h->K * 4;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int64_t mb = desc->max_batch_tiles > 0 ? desc->max_batch_tiles : std::max<int64_t>(64, (int64_t)(16.0 * 1024 * 1024 * 1024 / per_tile));
    mb = std::min<int64_t>(mb, std::max<int64_t>(1, (int64_t)(free_b * 0.5 / per_tile)));
    mb = std::min<int64_t>(mb, 4096);
    h->max_batch = (int)mb;
    h->cap_tiles = (h->max_batch + 7) & ~7;
    rc = lower_plan(h);
    if (rc != UMX_OK) { umx_destroy(h); return rc; }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("stream creation failed"); umx_destroy(h); return UMX_ECUDA;
    }
    h->stream = h->own_stream;
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&h->ev_stitch[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming);
    }
    if (conv_simt_configure() != cudaSuccess || tc_conv_configure() != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); umx_destroy(h); return UMX_ECUDA; }
    h->tensors.clear();   // host copies no longer needed
    *out = h;
    return UMX_OK;
}

void umx_destroy(umx_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    drain_profile(h);
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (auto p : h->dev_allocs) cudaFree(p);
    for (auto& b : h->bufs) { if (b.d) cudaFree(b.d); if (b.dh) cudaFree(b.dh); }
    if (h->probs) cudaFree(h->probs);
    if (h->d_dbg) cudaFree(h->d_dbg);
    if (h->d_band_u8) cudaFree(h->d_band_u8);
    if (h->d_out_u8) cudaFree(h->d_out_u8);
    if (h->d_lut) cudaFree(h->d_lut);
    if (h->d_minmax) cudaFree(h->d_minmax);
    if (h->d_img) cudaFree(h->d_img);
    if (h->d_probs_rows) cudaFree(h->d_probs_rows);
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage_u8[i]) cudaFree(h->d_stage_u8[i]);
        if (h->d_stage_f32[i]) cudaFree(h->d_stage_f32[i]);
        if (h->ev_stitch[i]) cudaEventDestroy(h->ev_stitch[i]);
        if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
}

int umx_set_stream(umx_handle* h, uint64_t cuda_stream) {
    if (!h) { set_error("null handle"); return UMX_EINVAL; }
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return UMX_OK;
}

int umx_profile_enable(umx_handle* h, int32_t on) {
    if (!h) { set_error("null handle"); return UMX_EINVAL; }
    h->profiling = on != 0;
    return UMX_OK;
}

int umx_profile_read(umx_handle* h, umx_prof_entry* out, int32_t capacity, int32_t reset) {
    if (!h) { set_error("null handle"); return UMX_EINVAL; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain_profile(h);
    int n = 0;
    for (auto& s : h->prof) {
        if (out && n < capacity) {
            memset(&out[n], 0, sizeof(out[n]));
            strncpy(out[n].name, s.name.c_str(), sizeof(out[n].name) - 1);
            out[n].launches = s.launches; out[n].ms_total = s.ms; out[n].flops = s.flops; out[n].bytes = s.bytes;
        }
        ++n;
        if (reset) { s.launches = 0; s.ms = 0; s.flops = 0; s.bytes = 0; }
    }
    return n;
}

int64_t umx_launch_count(umx_handle* h) { return h ? h->launches : 0; }

int umx_op_info(umx_handle* h, int32_t op_index, int32_t* out, int32_t capacity) {
    if (!h || !out || capacity < 12 || op_index < 0 || op_index >= (int)h->ops.size()) { set_error("umx_op_info: bad argument"); return UMX_EINVAL; }
    const Op& op = h->ops[op_index];
    const TcConvParams& tp = op.tcp;
    const bool tc = op.kind == OP_CONV && op.use_tc;
    const int v[12] = {tc ? 1 : 0, tc ? tp.halo : 0, tc ? tp.pair : 0, tc ? tp.stages : 0, tc ? tp.b_stages : 0, tc ? tp.gb : 0,
                       tc ? (tp.ncat ? 2 : tp.b_resident) : 0, tc ? tp.merge_px : 0, tc ? tp.planes_a : 0, tc ? tp.planes_b : 0,
                       tc ? tp.terms0 : 0, tc ? tp.terms1 : 0};
    for (int i = 0; i < 12; ++i) out[i] = v[i];
    return UMX_OK;
}

int64_t umx_debug_buffer(umx_handle* h, const char* name, int32_t n_tiles, float* out, int64_t capacity) {
    if (!h || !name || !out || n_tiles < 0) { set_error("umx_debug_buffer: bad argument"); return UMX_EINVAL; }
    UMX_CUDA_TRY(cudaSetDevice(h->device));
    for (auto& b : h->bufs) {
        if (b.name != name) continue;
        const int64_t n = (int64_t)std::min(n_tiles, h->cap_tiles) * b.per_tile();
        if (n > capacity) { set_error("umx_debug_buffer: capacity %lld < %lld", (long long)capacity, (long long)n); return UMX_EINVAL; }
        UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (b.d) {
            UMX_CUDA_TRY(cudaMemcpy(out, b.d, n * sizeof(float), cudaMemcpyDeviceToHost));
        } else if (b.dh) {
            const int64_t npix = n / b.c, cs = b.cs();
            std::vector<__half> tmp(npix * cs);
            for (int pl = 0; pl < b.planes; ++pl) {
                UMX_CUDA_TRY(cudaMemcpy(tmp.data(), b.dh + pl * b.plane_elems, tmp.size() * sizeof(__half), cudaMemcpyDeviceToHost));
                for (int64_t px = 0; px < npix; ++px)
                    for (int c = 0; c < b.c; ++c) out[px * b.c + c] = (pl ? out[px * b.c + c] : 0.f) + __half2float(tmp[px * cs + c]);
            }
        } else { set_error("buffer '%s' is not materialised", name); return UMX_EINVAL; }
        return b.per_tile();
    }
    set_error("no buffer named '%s'", name);
    return UMX_EINVAL;
}

int umx_forward_tiles(umx_handle* h, const float* tiles, int32_t n_tiles, float* probs, int32_t precision) {
    if (!h || n_tiles < 0 || (n_tiles > 0 && (!tiles || !probs))) { set_error("umx_forward_tiles: bad argument"); return UMX_EINVAL; }
    if (precision != UMX_PREC_DEFAULT && precision != h->precision) {
        set_error("umx_forward_tiles: precision %d requested but the handle was built with %d (the arithmetic is fixed at umx_create)", precision, h->precision);
        return UMX_EINVAL;
    }
    UMX_CUDA_TRY(cudaSetDevice(h->device));
    const size_t in_tile = (size_t)h->S * h->S * h->C, out_tile = (size_t)h->S * h->S * h->K;
    for (int t0 = 0; t0 < n_tiles; t0 += h->max_batch) {
        const int nb = std::min(h->max_batch, n_tiles - t0);
        UMX_CUDA_TRY(cudaMemcpyAsync(h->bufs[h->in_buf].d, tiles + (size_t)t0 * in_tile, nb * in_tile * sizeof(float),
                                     cudaMemcpyDefault, h->stream));
        UMX_TRY(run_network(h, nb, h->probs));
        UMX_CUDA_TRY(cudaMemcpyAsync(probs + (size_t)t0 * out_tile, h->probs, nb * out_tile * sizeof(float),
                                     cudaMemcpyDefault, h->stream));
        // the host buffers may be pageable: finish this group before the input buffer is reused
        UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    drain_profile(h);
    return UMX_OK;
}

int umx_band_rows(umx_handle* h, int32_t H, int32_t tr0, int32_t tr1, int32_t* row0, int32_t* row1) {
    if (!h || H <= 0) { set_error("umx_band_rows: bad argument"); return UMX_EINVAL; }
    const int npr = (H + h->sub - 1) / h->sub;
    if (tr1 <= 0 || tr1 > npr) tr1 = npr;
    if (tr0 < 0 || tr0 >= tr1) { set_error("umx_band_rows: empty band"); return UMX_EINVAL; }
    const int frame_rows = npr * h->sub + 2 * h->margin;
    const int p0 = tr0 * h->sub, p1 = (tr1 == npr) ? frame_rows : tr1 * h->sub;
    if (row0) *row0 = std::min(H, std::max(0, p0 - h->margin));
    if (row1) *row1 = std::min(H, std::max(0, p1 - h->margin));
    return UMX_OK;
}

}  // extern "C"

namespace umx {
namespace {

// first / last source row the resize of destination row y touches (before mirroring, which only acts at the borders)
void resize_window(const Resample& rs, int y, int* lo, int* hi) {
    if (!rs.on) { *lo = *hi = y; return; }
    double cc = ((double)y + 0.5) * rs.zoom_y - 0.5;
    if (cc < 0) cc = -cc;
    const int s0 = (int)floor(cc);
    *lo = s0 - rs.ry; *hi = s0 + 1 + rs.ry;
}

// Raw-grid rows a tile-row band owns when its pages are resized back (see umx_band_out_rows): the cut between two
// bands is the first raw row whose window reaches the inference rows the upper band does not emit.
int raw_cut(const umx_handle* h, const Resample& rs_out, int infer_h, int raw_h, int t) {
    const int npr = (infer_h + h->sub - 1) / h->sub;
    if (t <= 0) return 0;
    if (t >= npr) return raw_h;
    const int bound = t * h->sub - h->margin;          // first inference row the bands above do not emit
    int y = std::max(0, (int)floor((bound - 2 - rs_out.ry) / std::max(rs_out.zoom_y, 1e-9)) - 2);
    for (; y < raw_h; ++y) {
        int lo, hi; resize_window(rs_out, y, &lo, &hi);
        if (hi >= bound) break;
    }
    return std::min(y, raw_h);
}

}  // namespace
}  // namespace umx

extern "C" {

int umx_band_out_rows(umx_handle* h, int32_t infer_h, int32_t tr0, int32_t tr1, int32_t raw_h, int32_t* row0, int32_t* row1) {
    if (!h || infer_h <= 0 || raw_h <= 0) { set_error("umx_band_out_rows: bad argument"); return UMX_EINVAL; }
    if (raw_h == infer_h) return umx_band_rows(h, infer_h, tr0, tr1, row0, row1);
    const int npr = (infer_h + h->sub - 1) / h->sub;
    if (tr1 <= 0 || tr1 > npr) tr1 = npr;
    if (tr0 < 0 || tr0 >= tr1) { set_error("umx_band_out_rows: empty band"); return UMX_EINVAL; }
    Resample rs;
    if (!make_resample(&rs, infer_h, 1, raw_h, 1)) { set_error("scaling factor too large for the resize kernel"); return UMX_EINVAL; }
    if (row0) *row0 = raw_cut(h, rs, infer_h, raw_h, tr0);
    if (row1) *row1 = raw_cut(h, rs, infer_h, raw_h, tr1);
    return UMX_OK;
}

int umx_resample_minmax(umx_handle* h, const void* plane, int32_t dtype, int32_t H, int32_t W, int32_t out_h, int32_t out_w,
                        double in_scale, double* min_out, double* max_out) {
    if (!h || !plane || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || dtype_size(dtype) == 0) { set_error("umx_resample_minmax: bad argument"); return UMX_EINVAL; }
    UMX_CUDA_TRY(cudaSetDevice(h->device));
    MinMaxParams mp{};
    if (!make_resample(&mp.rs, H, W, out_h, out_w)) { set_error("scaling factor too small for the resize kernel (Gaussian radius > %d)", kMaxResampleRadius); return UMX_EINVAL; }
    const size_t bytes = (size_t)H * W * dtype_size(dtype);
    const void* src = plane;
    if (!is_device_ptr(plane)) {
        void* p = h->d_img;
        UMX_TRY(ensure(&p, &h->d_img_bytes, bytes));
        h->d_img = p;
        UMX_CUDA_TRY(cudaMemcpyAsync(h->d_img, plane, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_img;
    }
    if (!h->d_minmax) UMX_CUDA_TRY(cudaMalloc(&h->d_minmax, 2 * sizeof(unsigned long long)));
    const unsigned long long init[2] = {~0ull, 0ull};
    UMX_CUDA_TRY(cudaMemcpyAsync(h->d_minmax, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    mp.img = src; mp.dtype = dtype; mp.dst_h = out_h; mp.dst_w = out_w; mp.in_scale = in_scale; mp.out = h->d_minmax;
    UMX_CUDA_TRY(launch_resample_minmax(mp, h->stream));
    h->launches += 1;
    unsigned long long res[2];
    UMX_CUDA_TRY(cudaMemcpyAsync(res, h->d_minmax, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (min_out) *min_out = minmax_decode(res[0]);
    if (max_out) *max_out = minmax_decode(res[1]);
    return UMX_OK;
}

int umx_infer_image(umx_handle* h, const void* img, int32_t dtype, int32_t n_planes, int32_t H, int32_t W,
                    int64_t plane_stride, double mean, double std_dev, uint8_t* out_u8, float* out_f32,
                    const umx_opts* opts) {
    if (!h || !img || H <= 0 || W <= 0 || dtype_size(dtype) == 0) { set_error("umx_infer_image: bad argument"); return UMX_EINVAL; }
    if (n_planes != 1 && n_planes != h->C) { set_error("image has %d planes, network takes %d channels", n_planes, h->C); return UMX_EINVAL; }
    if (!out_u8 && !out_f32) { set_error("umx_infer_image: no output requested"); return UMX_EINVAL; }
    if (std_dev == 0.0) { set_error("std is zero"); return UMX_EINVAL; }
    if (opts && opts->precision != UMX_PREC_DEFAULT && opts->precision != h->precision) {
        set_error("umx_infer_image: precision %d requested but the handle was built with %d (the arithmetic is fixed at umx_create)", opts->precision, h->precision);
        return UMX_EINVAL;
    }
    UMX_CUDA_TRY(cudaSetDevice(h->device));
    // RH x RW: the samples in img; IH x IW: the grid the network runs on (a resized view when --scalingFactor != 1)
    const int RH = H, RW = W;
    const int IH = (opts && opts->infer_h > 0) ? opts->infer_h : RH, IW = (opts && opts->infer_w > 0) ? opts->infer_w : RW;
    const bool cli_quant = opts && (opts->flags & UMX_F_CLI_QUANT);
    Resample rs_in, rs_out;
    if (!make_resample(&rs_in, RH, RW, IH, IW) || !make_resample(&rs_out, IH, IW, RH, RW)) {
        set_error("scaling %dx%d -> %dx%d needs a Gaussian radius > %d", RH, RW, IH, IW, kMaxResampleRadius); return UMX_EINVAL;
    }
    const bool resize_back = cli_quant && rs_in.on;
    if (cli_quant && (!out_u8 || out_f32)) { set_error("UMX_F_CLI_QUANT produces out_u8 only"); return UMX_EINVAL; }
    const int S = h->S, m = h->margin, sub = h->sub, K = h->K;
    const int npr = (IH + sub - 1) / sub, npc = (IW + sub - 1) / sub;
    const int frame_rows = npr * sub + 2 * m;
    int ta = opts ? opts->tile_row0 : 0, tb = opts ? opts->tile_row1 : 0;
    if (tb <= 0 || tb > npr) tb = npr;
    if (ta < 0 || ta >= tb) { set_error("empty tile-row band [%d,%d) of %d", ta, tb, npr); return UMX_EINVAL; }
    if (plane_stride == 0) plane_stride = (int64_t)RH * RW;
    const int64_t out_grid = resize_back ? (int64_t)RH * RW : (int64_t)IH * IW;
    int64_t out_ps = (opts && opts->out_plane_stride) ? opts->out_plane_stride : out_grid;
    const int out_row_base = opts ? opts->out_row_base : 0;
    // UMX_F_CONTINUE: the tile row above the seam was left in slot 0 of d_probs_rows by the previous call
    const bool cont = opts && (opts->flags & UMX_F_CONTINUE) && ta > 0 && h->carry_row == ta - 1 && h->carry_h == IH && h->carry_w == IW &&
                      h->d_probs_rows != nullptr;
    const int t_first = cont ? ta : std::max(ta - 1, 0);
    const size_t esz = dtype_size(dtype);

    // ---- rows of the network grid this band reads, and the source rows behind them
    const int ir0 = std::max(0, t_first * sub - m), ir1 = std::min(IH, (tb - 1) * sub + S - m);
    int sr0 = ir0, sr1 = ir1;
    if (rs_in.on) {
        int lo, hi, lo2, hi2;
        resize_window(rs_in, ir0, &lo, &hi); resize_window(rs_in, std::max(ir0, ir1 - 1), &lo2, &hi2);
        sr0 = std::max(0, std::min(lo, lo2) - 1); sr1 = std::min(RH, std::max(hi, hi2) + 2);
        if (ir0 == 0) sr0 = 0;
        if (ir1 == IH) sr1 = RH;
        // mirrored taps near the image borders stay within the rows just selected (they reflect into [0, 2R])
        sr1 = std::max(sr1, std::min(RH, 2 * (rs_in.ry + 2))); if (sr0 > 0 && sr0 < 2 * (rs_in.ry + 2)) sr0 = 0;
        if (RH - sr1 < 2 * (rs_in.ry + 2)) sr1 = RH;
    }
    GatherParams gp{};
    gp.dtype = dtype; gp.n_planes = n_planes; gp.H = IH; gp.W = IW; gp.S = S; gp.margin = m; gp.sub = sub; gp.npc = npc;
    gp.C = h->C; gp.mean = mean; gp.std_dev = std_dev; gp.rs = rs_in;
    if (h->C > kMaxImagePlanes) { set_error("networks with more than %d input channels are not supported by the tiling driver", kMaxImagePlanes); return UMX_EINVAL; }
    if (opts && opts->premap) {
        gp.has_pre = 1;
        const bool per_plane = (opts->flags & UMX_F_PREMAP_PER_PLANE) != 0;
        for (int c = 0; c < h->C; ++c) {
            const umx_premap& pm = opts->premap[per_plane ? c : 0];
            gp.pre[c].in_scale = pm.in_scale; gp.pre[c].rescale = pm.rescale;
            gp.pre[c].imin = pm.imin; gp.pre[c].imax = pm.imax; gp.pre[c].omin = pm.omin; gp.pre[c].omax = pm.omax;
        }
    }
    if (is_device_ptr(img)) {
        gp.img = img; gp.plane_stride = plane_stride; gp.img_row0 = 0; gp.img_rows = RH;
    } else {
        const int nrows = sr1 - sr0;
        void* p = h->d_img;
        UMX_TRY(ensure(&p, &h->d_img_bytes, (size_t)n_planes * nrows * RW * esz));
        h->d_img = p;
        for (int pl = 0; pl < n_planes; ++pl) {
            const char* src = (const char*)img + ((size_t)pl * plane_stride + (size_t)sr0 * RW) * esz;
            char* dst = (char*)h->d_img + (size_t)pl * nrows * RW * esz;
            UMX_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)nrows * RW * esz, cudaMemcpyHostToDevice, h->stream));
        }
        gp.img = h->d_img; gp.plane_stride = (int64_t)nrows * RW; gp.img_row0 = sr0; gp.img_rows = nrows;
    }
    // integer samples at native size: the float64 normalisation of every possible code once, then one look-up per pixel
    if (!rs_in.on && (dtype == UMX_U8 || dtype == UMX_U16)) {
        if (!h->d_lut) UMX_CUDA_TRY(cudaMalloc(&h->d_lut, (size_t)kMaxImagePlanes * 65537 * sizeof(float)));
        for (int c = 0; c < h->C; ++c) {
            NormLutParams lp{};
            lp.n = dtype == UMX_U8 ? 256 : 65536; lp.mean = mean; lp.std_dev = std_dev; lp.pre = gp.pre[c]; lp.has_pre = gp.has_pre;
            lp.out = h->d_lut + (size_t)c * 65537;
            UMX_CUDA_TRY(launch_norm_lut(lp, h->stream));
            h->launches += 1;
        }
        gp.lut = h->d_lut;
    }

    // ---- tile-row groups; probs of a group (+ the carried previous tile row) stay on the device
    const int rpg = std::max(1, h->max_batch / npc);
    const size_t row_elems = (size_t)npc * S * S * K;
    {
        void* p = h->d_probs_rows;
        UMX_TRY(ensure(&p, &h->d_probs_rows_bytes, (size_t)(rpg + 1) * row_elems * sizeof(float)));
        if (p != h->d_probs_rows && cont) { set_error("UMX_F_CONTINUE: the workspace changed size between the calls"); h->d_probs_rows = (float*)p; h->carry_row = -1; return UMX_EINVAL; }
        h->d_probs_rows = (float*)p;
    }
    h->carry_row = -1;
    if (out_u8 && out_f32 && is_device_ptr(out_u8) != is_device_ptr(out_f32)) {
        set_error("out_u8 and out_f32 must both be host or both be device pointers"); return UMX_EINVAL;
    }
    const bool u8_direct = out_u8 && is_device_ptr(out_u8);
    const bool f32_direct = out_f32 && is_device_ptr(out_f32);
    const int max_rows = rpg * sub + 2 * m;
    // resize_back: the band's maps stay on the device at the network's size, rows [E0, E1): everything the tile rows
    // t_first .. tb-1 complete (the seam tile row is recomputed anyway, so the rows above the band's own come for free
    // and give the resize window its upper halo); the pages are resized in one pass after the last group.
    const int E0 = (resize_back && ta > 0) ? std::min(IH, (ta - 1) * sub + m) : 0;
    const int E1 = std::min(IH, std::max(0, ((tb == npr) ? frame_rows : tb * sub) - m));
    int Y0 = 0, Y1 = 0;
    if (resize_back) {
        Y0 = raw_cut(h, rs_out, IH, RH, ta); Y1 = raw_cut(h, rs_out, IH, RH, tb);
        int lo, hi; resize_window(rs_out, Y0, &lo, &hi);
        if (ta > 0 && lo < E0) { set_error("scaling factor too extreme for banded output (resize window exceeds the seam tile row)"); return UMX_EINVAL; }
        void* p = h->d_band_u8;
        UMX_TRY(ensure(&p, &h->d_band_u8_bytes, (size_t)K * (E1 - E0) * IW));
        h->d_band_u8 = (uint8_t*)p;
    } else {
        if (out_u8 && !u8_direct && h->d_stage_u8_bytes < (size_t)K * max_rows * IW) {
            for (int i = 0; i < 2; ++i) { if (h->d_stage_u8[i]) cudaFree(h->d_stage_u8[i]); h->d_stage_u8[i] = nullptr; }
            for (int i = 0; i < 2; ++i) UMX_CUDA_TRY(cudaMalloc(&h->d_stage_u8[i], (size_t)K * max_rows * IW));
            h->d_stage_u8_bytes = (size_t)K * max_rows * IW;
        }
        if (out_f32 && !f32_direct && h->d_stage_f32_bytes < (size_t)K * max_rows * IW * 4) {
            for (int i = 0; i < 2; ++i) { if (h->d_stage_f32[i]) cudaFree(h->d_stage_f32[i]); h->d_stage_f32[i] = nullptr; }
            for (int i = 0; i < 2; ++i) UMX_CUDA_TRY(cudaMalloc(&h->d_stage_f32[i], (size_t)K * max_rows * IW * 4));
            h->d_stage_f32_bytes = (size_t)K * max_rows * IW * 4;
        }
    }
    const int s_gather = aux_slot(h, "gather_tiles"), s_stitch = aux_slot(h, "stitch_quantize");
    int gi = 0;
    bool copies_pending[2] = {false, false};
    for (int g0 = t_first; g0 < tb; g0 += rpg, ++gi) {
        const int g1 = std::min(g0 + rpg, tb);
        const int lo = (g0 > t_first || cont) ? g0 - 1 : g0;    // first tile row held in d_probs_rows
        float* group_base = h->d_probs_rows + (size_t)(g0 - lo) * row_elems;
        const int tile_begin = g0 * npc, tile_end = g1 * npc;
        for (int t0 = tile_begin; t0 < tile_end; t0 += h->max_batch) {
            const int nb = std::min(h->max_batch, tile_end - t0);
            gp.tile0 = t0; gp.n_tiles = nb; gp.out = h->bufs[h->in_buf].d;
            if (h->fuse_gather && gp.lut) {
                // integer samples at native size and nothing but the first layer reads the tiles: no gathered tile buffer
                // at all, the first-layer kernel looks the samples up itself
                FirstImage fi{};
                fi.img = gp.img; fi.lut = gp.lut; fi.plane_stride = gp.plane_stride; fi.dtype = gp.dtype; fi.n_planes = gp.n_planes;
                fi.img_row0 = gp.img_row0; fi.H = gp.H; fi.W = gp.W; fi.margin = gp.margin; fi.sub = gp.sub; fi.npc = gp.npc; fi.tile0 = t0;
                UMX_TRY(run_network(h, nb, group_base + (size_t)(t0 - tile_begin) * S * S * K, &fi));
                continue;
            }
            {
                ScopedTimer tm(h, s_gather, 0, (double)nb * S * S * (h->C * 4.0 + esz));
                UMX_CUDA_TRY(launch_gather_tiles(gp, h->stream));
                h->launches += 1;
            }
            UMX_TRY(run_network(h, nb, group_base + (size_t)(t0 - tile_begin) * S * S * K));
        }
        // ---- emit the padded-frame rows this group completes
        int p0 = std::max(g0, ta) * sub;
        if (resize_back && g0 == t_first && ta > 0) p0 = (ta - 1) * sub + 2 * m;       // the halo rows above the band (tile row ta-1 is in slot 0 or recomputed)
        const int p1 = (g1 == npr) ? frame_rows : g1 * sub;
        const int r0 = std::min(IH, std::max(0, p0 - m)), r1 = std::min(IH, std::max(0, p1 - m));
        if (r1 > r0) {
            const int sb = gi & 1;
            if (copies_pending[sb]) { UMX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy[sb], 0)); copies_pending[sb] = false; }
            StitchParams sp{};
            sp.probs = h->d_probs_rows; sp.tr_lo = lo; sp.tr_hi = g1;
            sp.S = S; sp.margin = m; sp.sub = sub; sp.npc = npc; sp.npr = npr; sp.K = K; sp.H = IH; sp.W = IW;
            sp.row0 = r0; sp.row1 = r1;
            sp.requant = (cli_quant && !resize_back) ? 1 : 0;
            sp.replace = (opts && (opts->flags & UMX_F_STITCH_REPLACE)) ? 1 : 0;
            sp.fp16_quant = (opts && (opts->flags & UMX_F_FP16_QUANT)) ? 1 : 0;
            const bool staged = !resize_back && (out_u8 ? !u8_direct : !f32_direct);
            if (resize_back) {
                sp.out_u8 = h->d_band_u8; sp.out_plane_stride = (int64_t)(E1 - E0) * IW; sp.out_row_base = E0;
            } else if (staged) {
                sp.out_u8 = out_u8 ? h->d_stage_u8[sb] : nullptr;
                sp.out_f32 = out_f32 ? h->d_stage_f32[sb] : nullptr;
                sp.out_plane_stride = (int64_t)max_rows * IW; sp.out_row_base = r0;
            } else {
                sp.out_u8 = out_u8; sp.out_f32 = out_f32; sp.out_plane_stride = out_ps; sp.out_row_base = out_row_base;
            }
            {
                ScopedTimer tm(h, s_stitch, 0, (double)(r1 - r0) * IW * K * (4.0 * 1.78 + (out_u8 ? 1 : 0) + (out_f32 ? 4 : 0)));
                UMX_CUDA_TRY(launch_stitch(sp, h->stream));
                h->launches += 1;
            }
            if (staged) {
                UMX_CUDA_TRY(cudaEventRecord(h->ev_stitch[sb], h->stream));
                UMX_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_stitch[sb], 0));
                const size_t n = (size_t)(r1 - r0) * IW;
                for (int k = 0; k < K; ++k) {
                    if (out_u8)
                        UMX_CUDA_TRY(cudaMemcpyAsync(out_u8 + (size_t)k * out_ps + (size_t)(r0 - out_row_base) * IW,
                                                     h->d_stage_u8[sb] + (size_t)k * max_rows * IW, n, cudaMemcpyDeviceToHost, h->copy_stream));
                    if (out_f32)
                        UMX_CUDA_TRY(cudaMemcpyAsync(out_f32 + (size_t)k * out_ps + (size_t)(r0 - out_row_base) * IW,
                                                     h->d_stage_f32[sb] + (size_t)k * max_rows * IW, n * 4, cudaMemcpyDeviceToHost, h->copy_stream));
                }
                UMX_CUDA_TRY(cudaEventRecord(h->ev_copy[sb], h->copy_stream));
                copies_pending[sb] = true;
            }
        }
        // ---- carry the last tile row of this group into slot 0 for the next group (or the next UMX_F_CONTINUE call)
        {
            const float* last = h->d_probs_rows + (size_t)(g1 - 1 - lo) * row_elems;
            if (last != h->d_probs_rows)
                UMX_CUDA_TRY(cudaMemcpyAsync(h->d_probs_rows, last, row_elems * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    h->carry_row = tb - 1; h->carry_h = IH; h->carry_w = IW;
    if (resize_back && Y1 > Y0) {
        // uint8 pages at the network's size -> resize to the raw grid -> second quantisation (UnMicst1-5.py:850-853)
        ResizeU8Params rp{};
        rp.src = h->d_band_u8; rp.src_plane_stride = (int64_t)(E1 - E0) * IW; rp.src_row0 = E0; rp.src_rows = E1 - E0;
        rp.K = K; rp.dst_h = RH; rp.dst_w = RW; rp.row0 = Y0; rp.row1 = Y1; rp.rs = rs_out;
        if (u8_direct) { rp.out = out_u8; rp.out_plane_stride = out_ps; rp.out_row_base = out_row_base; }
        else {
            void* p = h->d_out_u8;
            UMX_TRY(ensure(&p, &h->d_out_u8_bytes, (size_t)K * (Y1 - Y0) * RW));
            h->d_out_u8 = (uint8_t*)p;
            rp.out = h->d_out_u8; rp.out_plane_stride = (int64_t)(Y1 - Y0) * RW; rp.out_row_base = Y0;
        }
        {
            ScopedTimer tm(h, aux_slot(h, "resize_pages"), 0, (double)(Y1 - Y0) * RW * K * (1.0 + rs_out.zoom_y * rs_out.zoom_x));
            UMX_CUDA_TRY(launch_resize_u8(rp, h->stream));
            h->launches += 1;
        }
        if (!u8_direct)
            for (int k = 0; k < K; ++k)
                UMX_CUDA_TRY(cudaMemcpyAsync(out_u8 + (size_t)k * out_ps + (size_t)(Y0 - out_row_base) * RW,
                                             h->d_out_u8 + (size_t)k * (Y1 - Y0) * RW, (size_t)(Y1 - Y0) * RW, cudaMemcpyDeviceToHost, h->stream));
    }
    // make the caller's stream wait for outstanding D2H copies, then (by default) the host too
    for (int sb = 0; sb < 2; ++sb)
        if (copies_pending[sb]) UMX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy[sb], 0));
    if (!(opts && (opts->flags & UMX_F_NO_SYNC))) {
        UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        drain_profile(h);
    }
    return UMX_OK;
}

int umx_infer_images(umx_handle* h, const umx_image* images, int32_t n_images, double mean, double std_dev, int32_t flags) {
    if (!h || n_images < 0 || (n_images > 0 && !images)) { set_error("umx_infer_images: bad argument"); return UMX_EINVAL; }
    if (std_dev == 0.0) { set_error("std is zero"); return UMX_EINVAL; }
    if (h->C > kMaxImagePlanes) { set_error("networks with more than %d input channels are not supported by the tiling driver", kMaxImagePlanes); return UMX_EINVAL; }
    UMX_CUDA_TRY(cudaSetDevice(h->device));
    const int S = h->S, m = h->margin, sub = h->sub, K = h->K;
    std::vector<int> tiles(n_images);
    for (int i = 0; i < n_images; ++i) {
        const umx_image& im = images[i];
        if (!im.img || im.H <= 0 || im.W <= 0 || dtype_size(im.dtype) == 0 || (!im.out_u8 && !im.out_f32)) { set_error("umx_infer_images: image %d malformed", i); return UMX_EINVAL; }
        if (im.n_planes != 1 && im.n_planes != h->C) { set_error("image %d has %d planes, network takes %d channels", i, im.n_planes, h->C); return UMX_EINVAL; }
        if ((flags & UMX_F_CLI_QUANT) && (!im.out_u8 || im.out_f32)) { set_error("UMX_F_CLI_QUANT produces out_u8 only"); return UMX_EINVAL; }
        tiles[i] = ((im.H + sub - 1) / sub) * ((im.W + sub - 1) / sub);
    }
    const int s_gather = aux_slot(h, "gather_tiles"), s_stitch = aux_slot(h, "stitch_quantize");
    h->carry_row = -1;
    int i = 0;
    while (i < n_images) {
        if (tiles[i] > h->max_batch) {        // a large image: the tile-row-group driver
            const umx_image& im = images[i];
            umx_opts o{};
            o.flags = flags & UMX_F_CLI_QUANT; o.premap = im.premap;
            UMX_TRY(umx_infer_image(h, im.img, im.dtype, im.n_planes, im.H, im.W, im.plane_stride, mean, std_dev, im.out_u8, im.out_f32, &o));
            ++i;
            continue;
        }
        int j = i, total = 0;
        size_t in_bytes = 0, u8_bytes = 0, f32_bytes = 0;
        while (j < n_images && tiles[j] <= h->max_batch && total + tiles[j] <= h->max_batch) {
            const umx_image& im = images[j];
            total += tiles[j];
            if (!is_device_ptr(im.img)) in_bytes += (((size_t)im.n_planes * im.H * im.W * dtype_size(im.dtype)) + 255) & ~(size_t)255;
            if (im.out_u8 && !is_device_ptr(im.out_u8)) u8_bytes += (size_t)K * im.H * im.W;
            if (im.out_f32 && !is_device_ptr(im.out_f32)) f32_bytes += (size_t)K * im.H * im.W * 4;
            ++j;
        }
        {
            void* p = h->d_img; UMX_TRY(ensure(&p, &h->d_img_bytes, std::max<size_t>(in_bytes, 256))); h->d_img = p;
            p = h->d_out_u8; UMX_TRY(ensure(&p, &h->d_out_u8_bytes, std::max<size_t>(u8_bytes, 256))); h->d_out_u8 = (uint8_t*)p;
            if (f32_bytes) {
                if (h->d_stage_f32_bytes < f32_bytes) {
                    for (int b = 0; b < 2; ++b) { if (h->d_stage_f32[b]) cudaFree(h->d_stage_f32[b]); h->d_stage_f32[b] = nullptr; }
                    for (int b = 0; b < 2; ++b) UMX_CUDA_TRY(cudaMalloc(&h->d_stage_f32[b], f32_bytes));
                    h->d_stage_f32_bytes = f32_bytes;
                }
            }
        }
        // ---- tiles of every image of the group into one launch batch
        size_t in_off = 0;
        int tile_off = 0;
        for (int k = i; k < j; ++k) {
            const umx_image& im = images[k];
            const size_t esz = dtype_size(im.dtype);
            const int64_t ps = im.plane_stride ? im.plane_stride : (int64_t)im.H * im.W;
            GatherParams gp{};
            gp.dtype = im.dtype; gp.n_planes = im.n_planes; gp.H = im.H; gp.W = im.W; gp.S = S; gp.margin = m; gp.sub = sub;
            gp.npc = (im.W + sub - 1) / sub; gp.C = h->C; gp.mean = mean; gp.std_dev = std_dev;
            make_resample(&gp.rs, im.H, im.W, im.H, im.W);
            if (im.premap) {
                gp.has_pre = 1;
                for (int c = 0; c < h->C; ++c) {
                    gp.pre[c].in_scale = im.premap->in_scale; gp.pre[c].rescale = im.premap->rescale; gp.pre[c].imin = im.premap->imin;
                    gp.pre[c].imax = im.premap->imax; gp.pre[c].omin = im.premap->omin; gp.pre[c].omax = im.premap->omax;
                }
            }
            if (is_device_ptr(im.img)) { gp.img = im.img; gp.plane_stride = ps; }
            else {
                char* dst = (char*)h->d_img + in_off;
                for (int pl = 0; pl < im.n_planes; ++pl)
                    UMX_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)pl * im.H * im.W * esz, (const char*)im.img + (size_t)pl * ps * esz,
                                                 (size_t)im.H * im.W * esz, cudaMemcpyHostToDevice, h->stream));
                gp.img = dst; gp.plane_stride = (int64_t)im.H * im.W;
                in_off += (((size_t)im.n_planes * im.H * im.W * esz) + 255) & ~(size_t)255;
            }
            gp.img_row0 = 0; gp.img_rows = im.H; gp.tile0 = 0; gp.n_tiles = tiles[k];
            gp.out = h->bufs[h->in_buf].d + (size_t)tile_off * S * S * h->C;
            {
                ScopedTimer tm(h, s_gather, 0, (double)tiles[k] * S * S * (h->C * 4.0 + esz));
                UMX_CUDA_TRY(launch_gather_tiles(gp, h->stream));
                h->launches += 1;
            }
            tile_off += tiles[k];
        }
        UMX_TRY(run_network(h, total, h->probs));
        // ---- stitch each image from its slice of the batch
        tile_off = 0;
        size_t u8_off = 0, f32_off = 0;
        for (int k = i; k < j; ++k) {
            const umx_image& im = images[k];
            const int npr = (im.H + sub - 1) / sub, npc = (im.W + sub - 1) / sub;
            const size_t n = (size_t)im.H * im.W;
            StitchParams sp{};
            sp.probs = h->probs + (size_t)tile_off * S * S * K; sp.tr_lo = 0; sp.tr_hi = npr;
            sp.S = S; sp.margin = m; sp.sub = sub; sp.npc = npc; sp.npr = npr; sp.K = K; sp.H = im.H; sp.W = im.W;
            sp.row0 = 0; sp.row1 = im.H; sp.out_plane_stride = (int64_t)n; sp.out_row_base = 0;
            sp.requant = (flags & UMX_F_CLI_QUANT) ? 1 : 0;
            const bool u8_dev = im.out_u8 && is_device_ptr(im.out_u8), f32_dev = im.out_f32 && is_device_ptr(im.out_f32);
            sp.out_u8 = im.out_u8 ? (u8_dev ? im.out_u8 : h->d_out_u8 + u8_off) : nullptr;
            sp.out_f32 = im.out_f32 ? (f32_dev ? im.out_f32 : (float*)((char*)h->d_stage_f32[0] + f32_off)) : nullptr;
            {
                ScopedTimer tm(h, s_stitch, 0, (double)n * K * (4.0 * 1.78 + (im.out_u8 ? 1 : 0) + (im.out_f32 ? 4 : 0)));
                UMX_CUDA_TRY(launch_stitch(sp, h->stream));
                h->launches += 1;
            }
            if (im.out_u8 && !u8_dev) { UMX_CUDA_TRY(cudaMemcpyAsync(im.out_u8, sp.out_u8, (size_t)K * n, cudaMemcpyDeviceToHost, h->stream)); u8_off += (size_t)K * n; }
            if (im.out_f32 && !f32_dev) { UMX_CUDA_TRY(cudaMemcpyAsync(im.out_f32, sp.out_f32, (size_t)K * n * 4, cudaMemcpyDeviceToHost, h->stream)); f32_off += (size_t)K * n * 4; }
            tile_off += tiles[k];
        }
        // the staging buffers are reused by the next group
        UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
        i = j;
    }
    UMX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    drain_profile(h);
    return UMX_OK;
}

// Host-only: the op list umx_create would build for this model (graph construction, BN folding, raw-input
// rewrite, which kernel family runs each op), one line per op.  Needs no GPU: for tests and for inspecting a
// checkpoint.  Returns the number of characters written (excluding the terminator) or a negative UMX_E* code.
int64_t umx_describe_plan(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights, char* out, int64_t capacity) {
    if (!desc || !out || capacity <= 0 || (!weights && n_weights > 0)) { set_error("umx_describe_plan: bad argument"); return UMX_EINVAL; }
    if (desc->abi_version != UMX_ABI_VERSION) { set_error("umx_describe_plan: ABI version %d != %d", desc->abi_version, UMX_ABI_VERSION); return UMX_EINVAL; }
    if (desc->graph != UMX_GRAPH_LEGACY && desc->graph != UMX_GRAPH_V2) { set_error("unknown graph %d", desc->graph); return UMX_EINVAL; }
    const int S = desc->im_size;
    if (S < 8 || (S & (S - 1)) || desc->n_layers < 1 || (S >> desc->n_layers) < 4 || desc->n_channels < 1 || desc->n_out0 < 1 ||
        desc->feat_maps_fact < 1 || desc->n_extra_convs < 0 || desc->n_classes < 2 || desc->n_classes > 4) {
        set_error("invalid hyper-parameters"); return UMX_EINVAL;
    }
    std::unique_ptr<umx_handle> h(new (std::nothrow) umx_handle());
    if (!h) { set_error("out of host memory"); return UMX_ENOMEM; }
    h->desc = *desc;
    h->S = S; h->C = desc->n_channels; h->K = desc->n_classes; h->L = desc->n_layers;
    h->margin = S / 8; h->sub = S - 2 * h->margin;
    h->chan = {desc->n_channels, desc->n_out0};
    for (int i = 0; i < desc->n_layers; ++i) h->chan.push_back(h->chan.back() * desc->feat_maps_fact);
    for (int i = 0; i < n_weights; ++i) {
        const umx_tensor& t = weights[i];
        if (!t.name || !t.data || t.ndim < 1 || t.ndim > 4) { set_error("weights[%d] malformed", i); return UMX_EINVAL; }
        HostTensor ht;
        for (int d = 0; d < t.ndim; ++d) ht.shape.push_back(t.shape[d]);
        ht.data.assign(t.data, t.data + ht.numel());
        h->tensors[t.name] = std::move(ht);
    }
    int rc = build_plan(h.get());
    if (rc != UMX_OK) return rc;
    h->precision = desc->precision == UMX_PREC_DEFAULT ? UMX_PREC_SPLIT3 : desc->precision;
    if (h->precision == UMX_PREC_MIXED) h->single_mask = (uint64_t)(uint32_t)desc->reserved[0] | ((uint64_t)(uint32_t)desc->reserved[1] << 32);
    if (h->precision < UMX_PREC_FP32 || h->precision > UMX_PREC_MIXED) { set_error("unknown precision %d", desc->precision); return UMX_EINVAL; }
    rc = rewrite_narrow_sources(h.get());
    if (rc != UMX_OK) return rc;
    std::string text;
    char line[512];
    for (size_t i = 0; i < h->ops.size(); ++i) {
        const Op& op = h->ops[i];
        if (op.kind == OP_TAPS) {
            const Buffer& ob = h->bufs[op.out_buf];
            snprintf(line, sizeof line, "%zu taps %s src=%s k=%d out=%dx%dx%d\n", i, op.name.c_str(), h->bufs[op.taps_src].name.c_str(), op.taps_k, ob.h, ob.w, ob.c);
        } else if (op.kind == OP_TOP) {
            snprintf(line, sizeof line, "%zu top %s src=%s cin=%d k=%d\n", i, op.name.c_str(), h->bufs[op.top_src].name.c_str(), op.tp.cin, op.tp.k);
        } else {
            const int mode = tc_mode_of(h.get(), op);
            const char* family = mode != TC_NONE ? "tensor" : (first_eligible(h.get(), op) ? "first" : "simt");
            const Buffer& ob = h->bufs[op.out_buf];
            std::string terms;
            for (auto& t : op.spec.terms) {
                terms += " [k=" + std::to_string(t.k) + " " + h->bufs[t.src0].name + ":" + std::to_string(h->bufs[t.src0].c);
                if (t.src1 >= 0) terms += "|" + h->bufs[t.src1].name + ":" + std::to_string(h->bufs[t.src1].c);
                terms += "]";
            }
            snprintf(line, sizeof line, "%zu conv %s %s mode=%d planes=%d%s%s%s%s out=%dx%dx%d flops=%.0f terms=%s\n", i, op.name.c_str(), family, mode,
                     mode != TC_NONE ? op_planes(h.get(), op) : 0, op.spec.transpose ? " transpose" : "", op.spec.pool ? " pool" : "",
                     op.spec.has_bias ? " bias" : "", op.spec.has_post ? " post" : "", ob.h, ob.w, ob.c, op.flops_per_tile, terms.c_str());
        }
        text += line;
    }
    const int64_t n = std::min<int64_t>((int64_t)text.size(), capacity - 1);
    memcpy(out, text.data(), (size_t)n);
    out[n] = 0;
    return n;
}

}  // extern "C"
