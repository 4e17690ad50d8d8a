// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Computes tf.nn.conv2d (3x3, SAME, stride 1) and tf.nn.conv2d_transpose (3x3, stride 2, SAME, as
// four sub-pixel phases) over NHWC activations of many independent PI2D tiles at once:
//     GEMM M = 128 output pixels (a TMA box of bn tiles x bh rows x bw cols),
//          N = n_t output channels (<= 256), K = taps x input channels in 64-channel slabs.
// A (activations) is fetched per filter tap by a 5-D TMA tiled load whose box origin is shifted by
// the tap offset; rows/cols outside an image tile are zero-filled by the TMA unit, which is exactly
// TensorFlow's per-tile SAME padding.  The channel concat [skip, up] of the up path
// (UnMicst1-5.py:196) is never materialised: the K loop walks two tensor maps.  B (weights, BN
// folded) is K-major [tap][cout][cin].  Both land in 128B-swizzled shared memory and feed
// tcgen05.mma (kind::f16, fp32 accumulate in TMEM).  Operands are fp16; in split mode every
// activation/weight is carried as hi + lo fp16 planes and each product issues three MMAs
// (hi*hi + hi*lo + lo*hi), which restores ~fp32 accuracy.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..5 = epilogue
// (TMEM -> registers -> bias / leaky-ReLU / 2x2 max-pool -> fp16 hi/lo planes or fp32).  Persistent
// CTAs, two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdio.h>

#include "umx_kernels.cuh"
#include "umx_tc.cuh"

namespace umx {

namespace {

constexpr int kThreads = 320;               // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kAPlaneBytes = 128 * 128;       // 128 rows x 64 fp16
constexpr int kAccStride = 256;               // TMEM columns between the two accumulator stages
constexpr uint32_t kSpinLimit = 1u << 28;     // trap instead of hanging the GPU on a pipeline bug

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp; ptxas recognises the elect.sync idiom and keeps the single-thread region's
// tcgen05 / TMA operands in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) { printf("umx tc_conv: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
// Polite wait for threads that are not on the critical path (epilogue, producer): back off between polls so the
// MMA-issuing thread's barrier traffic is not contended.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
        if (++spins > (kSpinLimit >> 4)) { printf("umx tc_conv: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the MMA spans two SMs (M = 256); each CTA stages its own 128
// A rows and half of the B rows, all signalling the leader CTA's barriers.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sbo(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TileCoord { int phase, n0, y0, x0, n_idx; };

// work item -> (phase, M unit, N tile); an M unit is one M tile, or a pair of M tiles in CTA-pair mode
__device__ __forceinline__ TileCoord decode_tile(const TcConvParams& p, int item, int m_units, int per_unit, int rank) {
    TileCoord t;
    // N tile fastest, then the conv-transpose phase, then the M unit: the (<= 4) phases and the N tiles that read
    // the same input pixels run back to back, so the input is fetched from HBM once and re-read from L2
    t.n_idx = item % p.n_ntiles;
    const int rest = item / p.n_ntiles;
    t.phase = rest % p.nphase;
    const int mt = (rest / p.nphase) * per_unit + rank;
    if (p.bn > 1) { t.n0 = mt * p.bn; t.y0 = 0; t.x0 = 0; }
    else {
        const int bx = p.in_w / p.bw, by = p.in_h / p.bh;
        t.n0 = mt / (bx * by);
        const int r = mt % (bx * by);
        t.y0 = (r / bx) * p.bh; t.x0 = (r % bx) * p.bw;
    }
    return t;
}

__device__ __forceinline__ float act_fn(float v, int act, float leaky) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LEAKY) return v > 0.f ? v : v * leaky;
    return v;
}

// SKIPC = channels of the narrow fp32 source folded into the epilogue (0 = none), SKT = its tap count.
template <int SKIPC, int SKT, bool PAIR, bool HALO>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ TcConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // Plain mode: one ring of `stages` slots, each [A tile planes][B tile planes] for one (tap, 64-channel slab).
    // Halo mode (high-resolution layers): ring A holds one (bh+halo) x (bw+halo) pixel patch per slab that serves
    // every tap (A is fetched from L2 once instead of once per tap); ring B holds the weights of `gb` taps per slot.
    const int planes = p.planes;
    const int a_box_bytes = HALO ? p.ph * p.pw * 128 : kAPlaneBytes;           // bytes one TMA box delivers per plane
    const int a_plane_bytes = HALO ? ((a_box_bytes + 1023) & ~1023) : kAPlaneBytes;
    const int a_bytes = planes * a_plane_bytes;
    const int b_plane_bytes = (PAIR ? p.n_t / 2 : p.n_t) * 128;     // pair mode: each CTA stages half of the N rows
    const int b_bytes = planes * b_plane_bytes;
    const int ks = HALO ? 1 : p.kslab;                               // plain mode: 64-channel slabs per ring slot
    const int slab_bytes = a_bytes + b_bytes;
    const int stage_bytes = HALO ? a_bytes : ks * slab_bytes;
    const int n_stages = p.stages;                                   // plain: ring slots; halo: patch slots
    const int nb_stages = HALO ? p.b_stages : 0;
    const int gb = HALO ? p.gb : 1;
    uint8_t* smem_b = smem + (size_t)n_stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)nb_stages * gb * b_bytes);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * n_stages;
    const uint32_t fullB0 = empty0 + 8 * n_stages;
    const uint32_t emptyB0 = fullB0 + 8 * nb_stages;
    const uint32_t tfull0 = emptyB0 + 8 * nb_stages;
    const uint32_t tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * n_stages + 2 * nb_stages + 4);
    // small fp32 tables for the epilogue: skip-term weights [9][SKIPC][cout], lt weights [cout][K], lt bias [K]
    // (rows padded with zeros to cpad = n_ntiles * n_t columns so the epilogue needs no channel guards)
    const int cpad = p.n_ntiles * p.n_t;
    float* s_skipw = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_topw = s_skipw + SKT * SKIPC * cpad;
    float* s_topb = s_topw + (p.top_w ? cpad * p.top_k : 0);
    float* s_z = s_topb + 4;                     // [2 acc stages][128 pixels][4]: partial lt logits of the upper column half
    float* s_bias = s_z + (p.top_w ? 2 * 128 * 4 : 0);       // [cpad] bias, then [cpad] post scale, [cpad] post shift
    float* s_ps = s_bias + cpad;
    float* s_pt = s_ps + (p.post_scale ? cpad : 0);
    for (int i = threadIdx.x; i < cpad; i += kThreads) {
        s_bias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
        if (p.post_scale) { s_ps[i] = i < p.cout ? p.post_scale[i] : 1.f; s_pt[i] = i < p.cout ? p.post_shift[i] : 0.f; }
    }
    if (SKIPC > 0)
        for (int i = threadIdx.x; i < SKT * SKIPC * cpad; i += kThreads) {
            const int c = i % cpad;
            s_skipw[i] = c < p.cout ? p.skip_w[(i / cpad) * p.cout + c] : 0.f;
        }
    if (p.top_w) {
        for (int i = threadIdx.x; i < cpad * p.top_k; i += kThreads) s_topw[i] = (i / p.top_k) < p.cout ? p.top_w[i] : 0.f;
        if (threadIdx.x < p.top_k) s_topb[threadIdx.x] = p.top_b ? p.top_b[threadIdx.x] : 0.f;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        if (p.c1 > 0) tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < n_stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
            for (int s = 0; s < nb_stages; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, PAIR ? 16 : 8); }
            fence_barrier_init();
        }
        __syncwarp();
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // peer barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = (p.bn > 1) ? (p.n_tiles + p.bn - 1) / p.bn : p.n_tiles * (p.in_w / p.bw) * (p.in_h / p.bh);
    const int per_unit = PAIR ? 2 : 1;
    const int m_units = (m_tiles + per_unit - 1) / per_unit;
    const int total = p.nphase * m_units * p.n_ntiles;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nch0 = (p.c0 + 63) >> 6, nch1 = (p.c1 + 63) >> 6;

    if (warp == 0) {
        if (elect_one()) {
            // ================= TMA producer =================
            if constexpr (HALO) {
                const int hx0 = p.hx0, hy0 = p.hy0, n_chunks = nch0 + nch1, c0s = p.c0, n_t = p.n_t;
                const bool a1c = p.a1_center != 0; const int ctap = p.center_tap;
                const uint32_t txA = (PAIR ? 2u : 1u) * (uint32_t)(planes * a_box_bytes);
                int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
                for (int tile = item0; tile < total; tile += item_step) {
                    const TileCoord t = decode_tile(p, tile, m_units, per_unit, rank);
                    const int ntap = p.ntaps[t.phase];
                    const int ncol = t.n_idx * n_t + (PAIR ? rank * (n_t / 2) : 0);
                    for (int cb = 0; cb < n_chunks; ++cb) {
                        const bool second = cb >= nch0;
                        const int cc = (second ? cb - nch0 : cb) * 64;
                        const CUtensorMap* mapA = second ? &mapA1 : &mapA0;
                        mbar_wait(empty0 + 8 * sa, pa ^ 1);
                        const uint32_t fa = full0 + 8 * sa;
                        const uint32_t da = smem_u32(smem + (size_t)sa * stage_bytes);
                        if (leader) mbar_expect_tx(fa, txA);
                        for (int pl = 0; pl < planes; ++pl) {
                            if (PAIR) tma_load_5d_pair(da + pl * a_plane_bytes, mapA, fa, cc, t.x0 - hx0, t.y0 - hy0, t.n0, pl);
                            else tma_load_5d(da + pl * a_plane_bytes, mapA, fa, cc, t.x0 - hx0, t.y0 - hy0, t.n0, pl);
                        }
                        if (++sa == n_stages) { sa = 0; pa ^= 1; }
                        const int tb = (second && a1c) ? ctap : 0, te = (second && a1c) ? ctap + 1 : ntap;
                        for (int t0 = tb; t0 < te; t0 += gb) {
                            const int ng = min(gb, te - t0);
                            mbar_wait(emptyB0 + 8 * sb, pb ^ 1);
                            const uint32_t fb = fullB0 + 8 * sb;
                            const uint32_t db = smem_u32(smem_b + (size_t)sb * gb * b_bytes);
                            if (leader) mbar_expect_tx(fb, (PAIR ? 2u : 1u) * (uint32_t)(ng * b_bytes));
                            for (int j = 0; j < ng; ++j) {
                                const int wi = p.taps[t.phase][t0 + j].wi;
                                if (PAIR) tma_load_4d_pair(db + j * b_bytes, &mapB, fb, (second ? c0s : 0) + cc, ncol, wi, 0);
                                else tma_load_4d(db + j * b_bytes, &mapB, fb, (second ? c0s : 0) + cc, ncol, wi, 0);
                            }
                            if (++sb == nb_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            } else {
            int stage = 0; uint32_t phase = 0;
            for (int tile = item0; tile < total; tile += item_step) {
                const TileCoord t = decode_tile(p, tile, m_units, per_unit, rank);
                const int ntap = p.ntaps[t.phase];
                for (int tp = 0; tp < ntap; ++tp) {
                    const TcTap tap = p.taps[t.phase][tp];
                    const int nchunks_tp = nch0 + ((p.a1_center && tp != p.center_tap) ? 0 : nch1);
                    for (int c0 = 0; c0 < nchunks_tp; c0 += ks) {
                        const int ns = min(ks, nchunks_tp - c0);
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t fb = full0 + 8 * stage;
                        const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
                        if (leader) mbar_expect_tx(fb, (PAIR ? 2u : 1u) * (uint32_t)(ns * slab_bytes));
                        for (int j = 0; j < ns; ++j) {
                            const int cb = c0 + j;
                            const bool second = cb >= nch0;
                            const int cc = (second ? cb - nch0 : cb) * 64;
                            const uint32_t sa = sbase + j * slab_bytes;
                            if (PAIR) {
                                tma_load_5d_pair(sa, second ? &mapA1 : &mapA0, fb, cc, t.x0 + tap.dx, t.y0 + tap.dy, t.n0, 0);
                                tma_load_4d_pair(sa + a_bytes, &mapB, fb, (second ? p.c0 : 0) + cc,
                                                 t.n_idx * p.n_t + rank * (p.n_t / 2), tap.wi, 0);
                            } else {
                                tma_load_5d(sa, second ? &mapA1 : &mapA0, fb, cc, t.x0 + tap.dx, t.y0 + tap.dy, t.n0, 0);
                                tma_load_4d(sa + a_bytes, &mapB, fb, (second ? p.c0 : 0) + cc, t.n_idx * p.n_t, tap.wi, 0);
                            }
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            // ================= MMA issuer (leader CTA only in pair mode) =================
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_t >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
            int acc = 0; uint32_t acc_phase = 0;
            if constexpr (HALO) {
                const int pw = p.pw, hx0 = p.hx0, hy0 = p.hy0, n_chunks = nch0 + nch1, c0s = p.c0, c1s = p.c1;
                const bool a1c = p.a1_center != 0; const int ctap = p.center_tap;
                const uint32_t a_sbo = (uint32_t)pw * 128u;         // consecutive 8-pixel rows are one patch row apart
                int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
                for (int tile = item0; tile < total; tile += item_step) {
                    const TileCoord t = decode_tile(p, tile, m_units, per_unit, rank);
                    const int ntap = p.ntaps[t.phase];
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
                    uint32_t accumulate = 0;
                    for (int cb = 0; cb < n_chunks; ++cb) {
                        const bool second = cb >= nch0;
                        const int kvalid = min(64, (second ? c1s : c0s) - (second ? cb - nch0 : cb) * 64);
                        const int nk = (kvalid + 15) >> 4;
                        mbar_wait(full0 + 8 * sa, pa);
                        tc_fence_after();
                        const uint32_t abase = smem_u32(smem + (size_t)sa * stage_bytes);
                        const int tb = (second && a1c) ? ctap : 0, te = (second && a1c) ? ctap + 1 : ntap;
                        for (int t0 = tb; t0 < te; t0 += gb) {
                            const int ng = min(gb, te - t0);
                            mbar_wait(fullB0 + 8 * sb, pb);
                            tc_fence_after();
                            const uint32_t bbase = smem_u32(smem_b + (size_t)sb * gb * b_bytes);
                            for (int j = 0; j < ng; ++j) {
                                const TcTap tap = p.taps[t.phase][t0 + j];
                                const uint32_t sa_t = abase + (uint32_t)((tap.dy + hy0) * pw + tap.dx + hx0) * 128u;
                                const uint32_t sb_t = bbase + j * b_bytes;
                                for (int k = 0; k < nk; ++k) {
                                    const uint64_t ah = umma_desc_sbo(sa_t + k * 32, a_sbo), bh = umma_desc(sb_t + k * 32);
                                    if (PAIR) umma_f16_pair(tmem_d, ah, bh, idesc, accumulate); else umma_f16(tmem_d, ah, bh, idesc, accumulate);
                                    accumulate = 1;
                                    if (planes == 2) {
                                        const uint64_t al = umma_desc_sbo(sa_t + a_plane_bytes + k * 32, a_sbo);
                                        const uint64_t bl = umma_desc(sb_t + b_plane_bytes + k * 32);
                                        if (PAIR) { umma_f16_pair(tmem_d, ah, bl, idesc, 1); umma_f16_pair(tmem_d, al, bh, idesc, 1); }
                                        else { umma_f16(tmem_d, ah, bl, idesc, 1); umma_f16(tmem_d, al, bh, idesc, 1); }
                                    }
                                }
                            }
                            if (PAIR) umma_commit_pair(emptyB0 + 8 * sb); else umma_commit(emptyB0 + 8 * sb);
                            if (++sb == nb_stages) { sb = 0; pb ^= 1; }
                        }
                        if (PAIR) umma_commit_pair(empty0 + 8 * sa); else umma_commit(empty0 + 8 * sa);
                        if (++sa == n_stages) { sa = 0; pa ^= 1; }
                    }
                    if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else {
            int stage = 0; uint32_t phase = 0;
            for (int tile = item0; tile < total; tile += item_step) {
                const TileCoord t = decode_tile(p, tile, m_units, per_unit, rank);
                const int ntap = p.ntaps[t.phase];
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
                uint32_t accumulate = 0;
                for (int tp = 0; tp < ntap; ++tp) {
                    const int nchunks_tp = nch0 + ((p.a1_center && tp != p.center_tap) ? 0 : nch1);
                    for (int c0 = 0; c0 < nchunks_tp; c0 += ks) {
                        const int ns = min(ks, nchunks_tp - c0);
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
                        for (int j = 0; j < ns; ++j) {
                        const int cb = c0 + j;
                        const bool second = cb >= nch0;
                        const int cc = (second ? cb - nch0 : cb) * 64;
                        const int kvalid = min(64, (second ? p.c1 : p.c0) - cc);
                        const int nk = (p.exp_flags & 4) ? 1 : (kvalid + 15) >> 4;
                        const uint32_t sa = sbase + j * slab_bytes;
                        const uint32_t sb = sa + a_bytes;
                        for (int k = 0; k < nk; ++k) {
                            const uint64_t ah = umma_desc(sa + k * 32), bh = umma_desc(sb + k * 32);
                            if (PAIR) umma_f16_pair(tmem_d, ah, bh, idesc, accumulate); else umma_f16(tmem_d, ah, bh, idesc, accumulate);
                            accumulate = 1;
                            if (p.planes == 2) {
                                const uint64_t al = umma_desc(sa + kAPlaneBytes + k * 32);
                                const uint64_t bl = umma_desc(sb + b_plane_bytes + k * 32);
                                if (PAIR) { umma_f16_pair(tmem_d, ah, bl, idesc, 1); umma_f16_pair(tmem_d, al, bh, idesc, 1); }
                                else { umma_f16(tmem_d, ah, bl, idesc, 1); umma_f16(tmem_d, al, bh, idesc, 1); }
                            }
                        }
                        }
                        // frees the smem slot (in both CTAs) when these MMAs retire
                        if (PAIR) umma_commit_pair(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
                // accumulator ready for the epilogue (of both CTAs)
                if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            }
        }
    } else {
        // ================= epilogue (4 warps, one TMEM lane quarter each) =================
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                           // which half of the N columns it handles
        const int n16 = p.n_t >> 4;
        const int c_lo = half == 0 ? 0 : (n16 + 1) >> 1;
        const int c_hi = (p.exp_flags & 8) ? min(c_lo + 1, n16) : (half == 0 ? (n16 + 1) >> 1 : n16);
        const int m = q * 32 + lane;
        const int xl = m % p.bw, yl = (m / p.bw) % p.bh, nl = m / (p.bw * p.bh);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = item0; tile < total; tile += item_step) {
            const TileCoord t = decode_tile(p, tile, m_units, per_unit, rank);
            const int n = t.n0 + nl, y = t.y0 + yl, x = t.x0 + xl;
            const bool valid = n < p.n_tiles;
            int oh, ow, oy, ox; bool writer = valid;
            if (p.pool) { oh = p.in_h >> 1; ow = p.in_w >> 1; oy = y >> 1; ox = x >> 1; writer = valid && !(y & 1) && !(x & 1); }
            else if (p.os == 2) { oh = p.in_h * 2; ow = p.in_w * 2; oy = 2 * y + (t.phase >> 1); ox = 2 * x + (t.phase & 1); }
            else { oh = p.in_h; ow = p.in_w; oy = y; ox = x; }
            const int64_t opix = ((int64_t)n * oh + oy) * ow + ox;
            // narrow fp32 source (raw input channels of lu0.conv2 / the legacy 1x1 shortcut): its taps in registers
            float xs[(SKT > 0 ? SKT : 1) * (SKIPC > 0 ? SKIPC : 1)];
            if (SKIPC > 0) {
#pragma unroll
                for (int tp = 0; tp < SKT; ++tp) {
                    const TcTap tap = p.skip_taps[tp];
                    const int yy = y + tap.dy, xx = x + tap.dx;
                    const bool inb = valid && yy >= 0 && yy < p.in_h && xx >= 0 && xx < p.in_w;
#pragma unroll
                    for (int cs = 0; cs < SKIPC; ++cs)
                        xs[tp * SKIPC + cs] = inb ? __ldg(p.skip_src + (((int64_t)n * p.in_h + yy) * p.in_w + xx) * SKIPC + cs) : 0.f;
                }
            }
            float z[4] = {0.f, 0.f, 0.f, 0.f};           // fused lt logits
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
            uint32_t rn[16];                                   // TMEM load of the next chunk is in flight while this one is processed
            if (c_lo < c_hi) tmem_ld16_issue(tbase + c_lo * 16, rn);
            for (int c16 = c_lo; c16 < c_hi; ++c16) {
                uint32_t r[16];
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = rn[j];
                if (c16 + 1 < c_hi) tmem_ld16_issue(tbase + (c16 + 1) * 16, rn);
                const int co = t.n_idx * p.n_t + c16 * 16;
                float v[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 b = *reinterpret_cast<const float4*>(s_bias + co + j4 * 4);
                    v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) + b.x; v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) + b.y;
                    v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) + b.z; v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) + b.w;
                }
                if (SKIPC > 0) {
#pragma unroll
                    for (int tp = 0; tp < SKT; ++tp) {
#pragma unroll
                        for (int cs = 0; cs < SKIPC; ++cs) {
                            const float xv = xs[tp * SKIPC + cs];
                            const float4* w4 = reinterpret_cast<const float4*>(s_skipw + (tp * SKIPC + cs) * cpad + co);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 w = w4[j4];
                                v[j4 * 4 + 0] = fmaf(xv, w.x, v[j4 * 4 + 0]); v[j4 * 4 + 1] = fmaf(xv, w.y, v[j4 * 4 + 1]);
                                v[j4 * 4 + 2] = fmaf(xv, w.z, v[j4 * 4 + 2]); v[j4 * 4 + 3] = fmaf(xv, w.w, v[j4 * 4 + 3]);
                            }
                        }
                    }
                }
                if (p.act == ACT_LEAKY) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * p.leaky);       // leaky slope < 1
                } else if (p.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (p.post_scale) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], s_ps[co + j], s_pt[co + j]);
                }
                if (p.top_w) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float* wk = s_topw + (co + j) * p.top_k;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < p.top_k) z[k] = fmaf(v[j], wk[k], z[k]);
                    }
                }
                if (p.pool) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], p.bw));
                    }
                }
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    const int c = co + h8 * 8;
                    if (!writer) continue;
                    if (p.out_f && c < p.cout) {
                        float* of = p.out_f + opix * p.cout + c;
                        if (c + 8 <= p.cout && !(p.cout & 3)) {
                            reinterpret_cast<float4*>(of)[0] = make_float4(v[h8 * 8 + 0], v[h8 * 8 + 1], v[h8 * 8 + 2], v[h8 * 8 + 3]);
                            reinterpret_cast<float4*>(of)[1] = make_float4(v[h8 * 8 + 4], v[h8 * 8 + 5], v[h8 * 8 + 6], v[h8 * 8 + 7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (c + j < p.cout) of[j] = v[h8 * 8 + j];
                        }
                    }
                    if (p.out_h && c < p.out_cs) {       // storage channels are padded to a multiple of 8; pad lanes hold 0
                        __half2 hi[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) hi[j] = __floats2half2_rn(v[h8 * 8 + 2 * j], v[h8 * 8 + 2 * j + 1]);
                        __half* o = p.out_h + opix * p.out_cs + c;
                        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(hi);
                        if (p.out_planes == 2) {
                            __half2 lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __half22float2(hi[j]);
                                lo[j] = __floats2half2_rn(v[h8 * 8 + 2 * j] - f.x, v[h8 * 8 + 2 * j + 1] - f.y);
                            }
                            *reinterpret_cast<uint4*>(o + p.out_plane_elems) = *reinterpret_cast<uint4*>(lo);
                        }
                    }
                }
                __syncwarp();        // reconverge before the next .sync.aligned TMEM load
            }
            if (p.top_w) {
                // the two warps of a lane quarter hold the two column halves of each pixel: combine the partial logits
                float* zs = s_z + ((size_t)acc * 128 + m) * 4;
                if (half == 1) { zs[0] = z[0]; zs[1] = z[1]; zs[2] = z[2]; zs[3] = z[3]; }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                if (half == 0) { z[0] += zs[0]; z[1] += zs[1]; z[2] += zs[2]; z[3] += zs[3]; }
            }
            if (p.top_w && writer && half == 0) {
                float mx = -INFINITY, sum = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < p.top_k) { z[k] += s_topb[k]; mx = fmaxf(mx, z[k]); }
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < p.top_k) { z[k] = expf(z[k] - mx); sum += z[k]; }
                const float inv = 1.f / sum;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < p.top_k) p.top_probs[opix * p.top_k + k] = z[k] * inv;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * acc); else mbar_arrive(tempty0 + 8 * acc); }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // the peer may still be reading this CTA's smem / arriving on its barriers
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

size_t tc_conv_fixed_bytes(const TcConvParams& p) {
    const size_t cpad = (size_t)p.n_ntiles * p.n_t;
    const size_t tables = ((size_t)(p.skip_src ? p.skip_ntaps : 0) * p.skip_c * cpad + (p.top_w ? cpad * p.top_k + 4 + 2 * 128 * 4 : 0) + cpad * (p.post_scale ? 3 : 1) + 8) * sizeof(float);
    return (2 * (size_t)(p.stages + (p.halo ? p.b_stages : 0)) + 4) * 8 + 16 + tables + 1024;
}
size_t tc_conv_a_bytes(const TcConvParams& p) {
    const size_t box = p.halo ? (((size_t)p.ph * p.pw * 128 + 1023) & ~(size_t)1023) : (size_t)kAPlaneBytes;
    return (size_t)p.planes * box;
}
size_t tc_conv_b_bytes(const TcConvParams& p) { return (size_t)p.planes * (size_t)(p.pair ? p.n_t / 2 : p.n_t) * 128; }
size_t tc_conv_smem_bytes(const TcConvParams& p) {
    if (p.halo) return p.stages * tc_conv_a_bytes(p) + (size_t)p.b_stages * p.gb * tc_conv_b_bytes(p) + tc_conv_fixed_bytes(p);
    return p.stages * (size_t)(p.kslab > 0 ? p.kslab : 1) * (tc_conv_a_bytes(p) + tc_conv_b_bytes(p)) + tc_conv_fixed_bytes(p);
}

template <int SKIPC, int SKT, bool PAIR, bool HALO>
static cudaError_t configure_one() {
    return cudaFuncSetAttribute(tc_conv_kernel<SKIPC, SKT, PAIR, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

template <int SKIPC, int SKT>
static cudaError_t configure_skip() {
    cudaError_t e = configure_one<SKIPC, SKT, false, false>();
    if (e == cudaSuccess) e = configure_one<SKIPC, SKT, true, false>();
    if (e == cudaSuccess) e = configure_one<SKIPC, SKT, false, true>();
    if (e == cudaSuccess) e = configure_one<SKIPC, SKT, true, true>();
    return e;
}

cudaError_t tc_conv_configure() {
    cudaError_t e = configure_skip<0, 0>();
    if (e == cudaSuccess) e = configure_skip<1, 1>();
    if (e == cudaSuccess) e = configure_skip<1, 9>();
    if (e == cudaSuccess) e = configure_skip<2, 9>();
    if (e == cudaSuccess) e = configure_skip<1, 25>();
    return e;
}

template <int SKIPC, int SKT, bool PAIR, bool HALO>
static cudaError_t launch_one(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcConvParams& p,
                              int grid, size_t smem, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_conv_kernel<SKIPC, SKT, PAIR, HALO>, a0, a1, b, p);
}

template <int SKIPC, int SKT>
static cudaError_t launch_skip(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcConvParams& p,
                               int grid, size_t smem, cudaStream_t s) {
    if (p.pair) return p.halo ? launch_one<SKIPC, SKT, true, true>(a0, a1, b, p, grid, smem, s) : launch_one<SKIPC, SKT, true, false>(a0, a1, b, p, grid, smem, s);
    return p.halo ? launch_one<SKIPC, SKT, false, true>(a0, a1, b, p, grid, smem, s) : launch_one<SKIPC, SKT, false, false>(a0, a1, b, p, grid, smem, s);
}

cudaError_t launch_tc_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcConvParams& p,
                           int num_sms, cudaStream_t s) {
    const int m_tiles = (p.bn > 1) ? (p.n_tiles + p.bn - 1) / p.bn : p.n_tiles * (p.in_w / p.bw) * (p.in_h / p.bh);
    const int per_unit = p.pair ? 2 : 1;
    const int total = p.nphase * ((m_tiles + per_unit - 1) / per_unit) * p.n_ntiles;
    if (total == 0) return cudaSuccess;
    int grid = total * per_unit < num_sms ? total * per_unit : num_sms;
    if (p.pair) grid &= ~1;
    // at least ~120 KB so that exactly one CTA (and its 512 TMEM columns) lives on an SM
    size_t smem = tc_conv_smem_bytes(p);
    if (smem < 120 * 1024) smem = 120 * 1024;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const int skipc = p.skip_src ? p.skip_c : 0, skt = p.skip_src ? p.skip_ntaps : 0;
    if (skipc == 0) return launch_skip<0, 0>(a0, a1, b, p, grid, smem, s);
    if (skipc == 1 && skt == 1) return launch_skip<1, 1>(a0, a1, b, p, grid, smem, s);
    if (skipc == 1 && skt == 9) return launch_skip<1, 9>(a0, a1, b, p, grid, smem, s);
    if (skipc == 2 && skt == 9) return launch_skip<2, 9>(a0, a1, b, p, grid, smem, s);
    if (skipc == 1 && skt == 25) return launch_skip<1, 25>(a0, a1, b, p, grid, smem, s);
    return cudaErrorInvalidValue;
}

int make_act_tensor_map(CUtensorMap* out, const __half* base, int planes, int64_t plane_elems, int n, int h, int w, int c,
                        int bw, int bh, int bn, int box_planes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)planes};
    cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)plane_elems * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)box_planes};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int make_weight_tensor_map(CUtensorMap* out, const __half* base, int planes, int taps, int cout, int cin, int n_t,
                           int box_planes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps, (cuuint64_t)planes};
    cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2, (cuuint64_t)taps * cout * cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)n_t, 1, (cuuint32_t)box_planes};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace umx
