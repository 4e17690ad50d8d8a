// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Computes tf.nn.conv2d (1x1 / 3x3 / 5x5, SAME, stride 1) and tf.nn.conv2d_transpose (stride 2, SAME, as four
// sub-pixel phases) over NHWC activations of many independent PI2D tiles at once:
//     GEMM M = 128 output pixels (a TMA box of bn tiles x bh rows x bw cols),
//          N = n_t output channels (<= 256), K = taps x input channels in 64-channel slabs.
// A (activations): plain mode fetches one box per (tap, slab) with the box origin shifted by the tap offset; halo
// mode fetches one (bh+halo) x (bw+halo) pixel patch per slab and moves the UMMA descriptor inside it per tap.
// Rows/cols outside an image tile are zero-filled by the TMA unit, which is exactly TensorFlow's per-tile SAME
// padding.  The channel concat [skip, up] of the up path (UnMicst1-5.py:196) is never materialised: the K loop
// walks two tensor maps.  B (weights, BN folded) is K-major [tap][cout][cin], streamed through a ring of slots of
// `gb` taps or, when the whole layer fits, resident in shared memory.  Both operands land in 128B-swizzled shared
// memory and feed tcgen05.mma (kind::f16, fp32 accumulate in TMEM).  Operands are fp16; in split mode every
// activation/weight is carried as hi + lo fp16 planes and each product issues three MMAs (hi*hi + hi*lo + lo*hi),
// which restores ~fp32 accuracy.
// Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (+ TMEM allocator), warps 2..9 =
// epilogue, two per TMEM lane quarter (TMEM -> registers -> bias / activation / 2x2 max-pool -> fp16 hi/lo planes
// or fp32, or the fused lt 1x1 conv + softmax).  Persistent CTAs (optionally CTA pairs, cta_group::2), two TMEM
// accumulator stages so the epilogue of item i overlaps the main loop of item i+1.
#include <stdio.h>

#include <utility>

#include "umx_kernels.cuh"
#include "umx_tc.cuh"

namespace umx {

namespace {

#ifdef UMX_NO_NCAT                           // A/B builds: the N-concatenated correction MMA compiled out
constexpr bool kNcat = false;
#else
constexpr bool kNcat = true;
#endif
constexpr int kEpiWarps = 8;                 // epilogue warps: kEpiSub per TMEM lane quarter, each a share of the N columns
constexpr int kEpiSub = kEpiWarps / 4;
constexpr int kThreads = 64 + 32 * kEpiWarps; // warp 0 TMA, warp 1 MMA, then the epilogue warps
constexpr int kAPlaneBytes = 128 * 128;       // 128 rows x 64 fp16
constexpr int kAccStride = 256;               // TMEM columns between the two accumulator stages

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
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~the hint elapses)
// instead of spinning through the issue slots; for waits that are not on the critical path (epilogue, producer).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// A pipeline bug must end in a trap, not in a hung GPU: a wait gives up after kSpinLimit failed polls.  A poll sleeps in
// hardware until the phase completes or the hint elapses, so the limit is reached after 84 s (20 us hints) to 7 min
// (100 us hints) of genuine waiting, and no sooner than ~0.2 s even if every poll returned at once (slow tools,
// time slicing); no legitimate wait of these kernels lasts a millisecond.  The check is a counter on purpose:
// anything heavier in this loop (a %globaltimer read, an out-of-line slow path: both tried) costs the single-thread
// roles 5-15 % on the issue-bound layers and, as an ABI call, 5 MB of code.
constexpr uint32_t kSpinLimit = 1u << 22;
__device__ __forceinline__ void mbar_timed_out(uint32_t bar, uint32_t parity) {
#ifdef UMX_TC_DEBUG     // which barrier each role is stuck on (warp 0 = TMA producer, 1 = MMA issuer, 2.. = epilogue); UMX_TC_EXP=2048 prints the layout
    if ((threadIdx.x & 31) == 0 || threadIdx.x < 64)
        printf("umx tc_conv: mbarrier wait timed out (block %d warp %d thread %d, barrier at shared 0x%x, parity %u)\n",
               blockIdx.x, threadIdx.x >> 5, threadIdx.x, bar, parity);
#else                   // (kept to two arguments: every wait site inlines this call's argument buffer)
    printf("umx tc_conv: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
#endif
    __trap();
}
template <uint32_t HINT_NS>
__device__ __forceinline__ void mbar_wait_impl(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, HINT_NS)) {
        if (++spins > kSpinLimit) mbar_timed_out(bar, parity);
    }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) { mbar_wait_impl<20000u>(bar, parity); }
// epilogue wait with back-off: the hinted try_wait returns every ~70 ns in practice, so eight waiting warps would spend
// their time (and a third of the SM's issued instructions) polling; after a few failed polls they nap between polls
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t nap_ns) {
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++spins > kSpinLimit) mbar_timed_out(bar, parity);
        if (spins > 4) __nanosleep(nap_ns);
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_impl<100000u>(bar, parity); }
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
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
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
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Cycle accounting for UMX_TC_DBG (exp_flags & 64): one thread per role adds its wait / work cycles to p.dbg[].
struct DbgClock {
    long long t; bool on;
    __device__ __forceinline__ void start(bool enable) { on = enable; if (on) t = clock64(); }
    __device__ __forceinline__ void lap(unsigned long long& acc) { if (on) { const long long n = clock64(); acc += (unsigned long long)(n - t); t = n; } }
};

struct TileCoord { int phase, n0, y0, x0, n_idx; };

// work item -> (phase, M unit, N tile); an M unit is one M tile, or a pair of M tiles in CTA-pair mode.
// N tile fastest, then the conv-transpose phase, then the M unit: the (<= 4) phases and the N tiles that read the
// same input pixels run back to back, so the input is fetched from HBM once and re-read from L2.
// Division-free: grids and boxes are powers of two, nphase is 1 or 4, the N-tile count uses a magic multiplier.
struct TileDecoder {
    uint32_t nt_mul; int n_ntiles, ph_shift, per_unit, rank, bn, bx_shift, bxy_shift, bw, bh;
    __device__ __forceinline__ void init(const TcConvParams& p, int per_unit_, int rank_) {
        n_ntiles = p.n_ntiles; nt_mul = (uint32_t)((0x100000000ull + n_ntiles - 1) / (uint32_t)n_ntiles);
        ph_shift = p.nphase == 4 ? (p.merge_px ? 1 : 2) : 0; per_unit = per_unit_; rank = rank_; bn = p.bn; bw = p.bw; bh = p.bh;
        bx_shift = 31 - __clz(p.in_w / p.bw); bxy_shift = bx_shift + 31 - __clz(p.in_h / p.bh);
    }
    __device__ __forceinline__ TileCoord operator()(int item) const {
        TileCoord t;
        int rest = item; t.n_idx = 0;
        if (n_ntiles > 1) { rest = (int)__umulhi((uint32_t)item, nt_mul); t.n_idx = item - rest * n_ntiles; }
        t.phase = rest & ((1 << ph_shift) - 1);
        const int mt = (rest >> ph_shift) * per_unit + rank;
        if (bn > 1) { t.n0 = mt * bn; t.y0 = 0; t.x0 = 0; }
        else {
            t.n0 = mt >> bxy_shift;
            const int r = mt & ((1 << bxy_shift) - 1);
            t.y0 = (r >> bx_shift) * bh; t.x0 = (r & ((1 << bx_shift) - 1)) * bw;
        }
        return t;
    }
};

// tcgen05.mma with the 64-bit shared-memory descriptors given as (lo, hi) register pairs: the hi words are loop
// invariant and the lo words advance by plain 32-bit adds, so one MMA costs two adds and the instruction itself.
template <bool PAIR>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    if (PAIR)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// lo / hi words of the shared-memory descriptor of a K-major, 128B-swizzled operand tile (rows of 128 B, 8-row swizzle
// atoms `sbo` bytes apart): lo = start address >> 4 | LBO (unused, 1) << 16; hi = SBO >> 4 | version 1 << 14 | SWIZZLE_128B << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14) | (2u << 29); }

// tcgen05.mma that always accumulates (every MMA of a work item but the first)
template <bool PAIR>
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (PAIR)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.b32 p, 0, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.b32 p, 0, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}

// The nk <= 4 MMAs (one per 16 channels) of a 64-channel slab; in split mode each product is hi*hi + hi*lo + lo*hi.
// Only the very first MMA of a work item carries a run-time accumulate flag.
// `terms` (split layers): bit 0 = the a_hi * w_lo correction, bit 1 = the a_lo * w_hi correction (3 = full hi/lo split)
template <bool PAIR>
__device__ __forceinline__ void umma_slab(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          int nk, bool split, int terms, uint32_t a_plane16, uint32_t b_plane16, uint32_t& accumulate, int exp = 0) {
    if (exp & 128) return;                       // timing experiment: the issue loop without the MMAs
    umma_issue<PAIR>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
    accumulate = 1;
    if (!split) {
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (k < nk) umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc);
    } else {
        if (terms & 1) umma_acc<PAIR>(tmem_d, a_lo, a_hi, b_lo + b_plane16, b_hi, idesc);
        if (terms & 2) umma_acc<PAIR>(tmem_d, a_lo + a_plane16, a_hi, b_lo, b_hi, idesc);
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (k < nk) {
                umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc);
                if (terms & 1) umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + b_plane16 + 2 * k, b_hi, idesc);
                if (terms & 2) umma_acc<PAIR>(tmem_d, a_lo + a_plane16 + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc);
            }
    }
}

// Halo mode: the `ng` taps of one weight slot against one patch, as straight-line code for a compile-time number of
// 16-channel steps (the gap between two MMAs of the single issuing thread must stay below the ~40-50 cycles one
// N = 80 MMA takes, or the tensor pipe idles at every tap boundary).
// Split layers add up to two correction MMAs per product (uniform in the single issuing thread): terms bit 0 = a_hi * w_lo,
// bit 1 = a_lo * w_hi.  (One compile-time variant per term set made the kernel 16 straight-line copies large and cost the
// issue-bound layers ~10 %: only "none" and "a_lo only" are compiled in, the rest branches on two uniform flags.)
// TERMS: 0 = no correction, 2 = a_lo * w_hi only (the weight-stationary partial split of the 64 x 64 layers), -1 = chosen at run time.
template <bool PAIR, int NK, int TERMS>
__device__ __forceinline__ void halo_taps(int terms, int ng, uint32_t tmem_d, uint32_t& a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t a_plane16, uint32_t bp16, uint32_t b16, int sx, int row_back,
                                          int nx, int& ix, uint32_t& accumulate) {
    const bool w_lo = TERMS == -1 && (terms & 1), a_lo_t = TERMS == 2 || (TERMS == -1 && (terms & 2));
    for (int j = 0; j < ng; ++j, b_lo += b16) {
        umma_issue<PAIR>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
        accumulate = 1;
        if (TERMS != 0) {
            if (w_lo) umma_acc<PAIR>(tmem_d, a_lo, a_hi, b_lo + bp16, b_hi, idesc);
            if (a_lo_t) umma_acc<PAIR>(tmem_d, a_lo + a_plane16, a_hi, b_lo, b_hi, idesc);
        }
#pragma unroll
        for (int k = 1; k < NK; ++k) {
            umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc);
            if (TERMS != 0) {
                if (w_lo) umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + bp16 + 2 * k, b_hi, idesc);
                if (a_lo_t) umma_acc<PAIR>(tmem_d, a_lo + a_plane16 + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc);
            }
        }
        a_lo += sx;
        if (++ix == nx) { ix = 0; a_lo += row_back; }
    }
}

template <bool PAIR, int TERMS>
__device__ __forceinline__ void halo_taps_nk(int terms, int nk, int ng, uint32_t tmem_d, uint32_t& a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t a_plane16, uint32_t bp16, uint32_t b16, int sx, int row_back,
                                             int nx, int& ix, uint32_t& accumulate) {
    if (nk == 4) halo_taps<PAIR, 4, TERMS>(terms, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
    else if (nk == 1) halo_taps<PAIR, 1, TERMS>(terms, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
    else if (nk == 2) halo_taps<PAIR, 2, TERMS>(terms, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
    else halo_taps<PAIR, 3, TERMS>(terms, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
}

// N-concatenated variant (p.ncat): per product one MMA of width 2*n_t, a_hi x X = [w_hi | w_lo] (or [w_hi | 0] for a source without
// the a_hi*w_lo term), plus a_lo x Y (w_hi, width n_t) when the source has the a_lo*w_hi term.  y16 = offset of Y behind X.
template <bool PAIR, int NK>
__device__ __forceinline__ void halo_taps_ncat(bool a_lo_t, int ng, uint32_t tmem_d, uint32_t& a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t idesc2, uint32_t a_plane16, uint32_t y16, uint32_t btap16, int sx, int row_back,
                                               int nx, int& ix, uint32_t& accumulate) {
    for (int j = 0; j < ng; ++j, b_lo += btap16) {
        umma_issue<PAIR>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc2, accumulate);
        accumulate = 1;
        if (a_lo_t) umma_acc<PAIR>(tmem_d, a_lo + a_plane16, a_hi, b_lo + y16, b_hi, idesc);
#pragma unroll
        for (int k = 1; k < NK; ++k) {
            umma_acc<PAIR>(tmem_d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc2);
            if (a_lo_t) umma_acc<PAIR>(tmem_d, a_lo + a_plane16 + 2 * k, a_hi, b_lo + y16 + 2 * k, b_hi, idesc);
        }
        a_lo += sx;
        if (++ix == nx) { ix = 0; a_lo += row_back; }
    }
}
template <bool PAIR>
__device__ __forceinline__ void halo_taps_ncat_nk(int terms, int nk, int ng, uint32_t tmem_d, uint32_t& a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t idesc2, uint32_t a_plane16, uint32_t y16, uint32_t btap16, int sx, int row_back,
                                                  int nx, int& ix, uint32_t& accumulate) {
    const bool al = (terms & 2) != 0;
    if (nk == 4) halo_taps_ncat<PAIR, 4>(al, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, idesc2, a_plane16, y16, btap16, sx, row_back, nx, ix, accumulate);
    else if (nk == 1) halo_taps_ncat<PAIR, 1>(al, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, idesc2, a_plane16, y16, btap16, sx, row_back, nx, ix, accumulate);
    else if (nk == 2) halo_taps_ncat<PAIR, 2>(al, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, idesc2, a_plane16, y16, btap16, sx, row_back, nx, ix, accumulate);
    else halo_taps_ncat<PAIR, 3>(al, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, idesc2, a_plane16, y16, btap16, sx, row_back, nx, ix, accumulate);
}

// per-slab dispatch on the correction terms of the slab's source
template <bool PAIR>
__device__ __forceinline__ void halo_taps_terms(int terms, int nk, int ng, uint32_t tmem_d, uint32_t& a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t a_plane16, uint32_t bp16, uint32_t b16, int sx, int row_back,
                                                int nx, int& ix, uint32_t& accumulate) {
    if (terms == 0) halo_taps_nk<PAIR, 0>(0, nk, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
    else if (terms == 2) halo_taps_nk<PAIR, 2>(2, nk, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
    else halo_taps_nk<PAIR, -1>(terms, nk, ng, tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, row_back, nx, ix, accumulate);
}

// Everything the epilogue of one accumulator tile needs that does not change from item to item.
struct EpiCtx {
    const float* s_ps; const float* s_pt; const float* s_skipw; const float4* s_topw4;
    float* out_f; __half* out_h; int64_t out_plane_elems;
    int cout, out_cs, out_planes, act, pool, bw;
    float leaky;
    bool has_post;
    bool wide_store;            // fp16-only output whose 16-channel chunks are 32-byte aligned and inside the row
};

// Bias / shortcut / activation / post-affine of NC accumulator columns of one pixel, then either the fused lt logits
// (TOPK > 0) or the 2x2 max-pool and the fp32 / fp16 hi[/lo] stores.  co = absolute output channel of column 0.
template <int NC, int SKIPC, int TOPK>
__device__ __forceinline__ void epi_chunk(const TcConvParams& p, const EpiCtx& e, const uint32_t (&r)[NC], int co, float xs, bool writer,
                                          float* of, __half* oh, float (&z)[4]) {
    float v[NC];
#pragma unroll
    for (int j4 = 0; j4 < NC / 4; ++j4) {      // co is warp-uniform: constant-bank loads, no shared-memory traffic
        const float4 b = *reinterpret_cast<const float4*>(p.tab_bias + co + j4 * 4);
        v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) + b.x; v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) + b.y;
        v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) + b.z; v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) + b.w;
    }
    if (SKIPC > 0) {            // 1x1 shortcut of a one-channel source (legacy down layers, UnMicst.py:95-97)
#pragma unroll
        for (int j4 = 0; j4 < NC / 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(e.s_skipw + co + j4 * 4);
            v[j4 * 4 + 0] = fmaf(xs, w.x, v[j4 * 4 + 0]); v[j4 * 4 + 1] = fmaf(xs, w.y, v[j4 * 4 + 1]);
            v[j4 * 4 + 2] = fmaf(xs, w.z, v[j4 * 4 + 2]); v[j4 * 4 + 3] = fmaf(xs, w.w, v[j4 * 4 + 3]);
        }
    }
    if (e.act == ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], v[j] * e.leaky);       // leaky slope < 1
    } else if (e.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (e.has_post) {
#pragma unroll
        for (int j4 = 0; j4 < NC / 4; ++j4) {
            const float4 a = *reinterpret_cast<const float4*>(e.s_ps + co + j4 * 4);
            const float4 b = *reinterpret_cast<const float4*>(e.s_pt + co + j4 * 4);
            v[j4 * 4 + 0] = fmaf(v[j4 * 4 + 0], a.x, b.x); v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], a.y, b.y);
            v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], a.z, b.z); v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], a.w, b.w);
        }
    }
    if (TOPK > 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float4 w = e.s_topw4[co + j];       // one broadcast LDS.128 (the constant bank only offers 64-bit loads)
            z[0] = fmaf(v[j], w.x, z[0]); z[1] = fmaf(v[j], w.y, z[1]);
            if (TOPK > 2) z[2] = fmaf(v[j], w.z, z[2]);
            if (TOPK > 3) z[3] = fmaf(v[j], w.w, z[3]);
        }
        return;
    }
    if (e.pool) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], e.bw));
        }
    }
    if (!writer) return;
    if (NC == 16 && e.wide_store) {      // fp16 only, 32-byte aligned rows: one 256-bit store per plane
        __half2 hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hi[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const uint32_t* h = reinterpret_cast<const uint32_t*>(hi);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(oh), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]),
                     "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
        if (e.out_planes == 2) {
            __half2 lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 f = __half22float2(hi[j]);
                lo[j] = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
            }
            const uint32_t* l = reinterpret_cast<const uint32_t*>(lo);
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(oh + e.out_plane_elems), "r"(l[0]), "r"(l[1]), "r"(l[2]),
                         "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
        }
        return;
    }
#pragma unroll
    for (int h8 = 0; h8 < NC / 8; ++h8) {
        const int c = co + h8 * 8;
        if (e.out_f != nullptr && c < e.cout) {
            float* o = of + h8 * 8;
            if (c + 8 <= e.cout && !(e.cout & 3)) {
                reinterpret_cast<float4*>(o)[0] = make_float4(v[h8 * 8 + 0], v[h8 * 8 + 1], v[h8 * 8 + 2], v[h8 * 8 + 3]);
                reinterpret_cast<float4*>(o)[1] = make_float4(v[h8 * 8 + 4], v[h8 * 8 + 5], v[h8 * 8 + 6], v[h8 * 8 + 7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (c + j < e.cout) o[j] = v[h8 * 8 + j];
            }
        }
        if (e.out_h != nullptr && c < e.out_cs) {       // storage channels are padded to a multiple of 8; pad lanes hold 0
            __half2 hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hi[j] = __floats2half2_rn(v[h8 * 8 + 2 * j], v[h8 * 8 + 2 * j + 1]);
            __half* o = oh + h8 * 8;
            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(hi);
            if (e.out_planes == 2) {
                __half2 lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(hi[j]);
                    lo[j] = __floats2half2_rn(v[h8 * 8 + 2 * j] - f.x, v[h8 * 8 + 2 * j + 1] - f.y);
                }
                *reinterpret_cast<uint4*>(o + e.out_plane_elems) = *reinterpret_cast<uint4*>(lo);
            }
        }
    }
}

// SKIPC = 1: a one-channel fp32 source enters as a 1x1 shortcut in the epilogue.  TOPK = classes of the fused
// lt 1x1 conv + softmax (0: the activation is stored instead).
template <int SKIPC, int TOPK, bool PAIR, bool HALO>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ TcConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle; computed as an offset so the pointers stay provably shared-space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // Plain mode: one ring of `stages` slots, each [A tile planes][B tile planes] for one (tap, 64-channel slab).
    // Halo mode (high-resolution layers): ring A holds one (bh+halo) x (bw+halo) pixel patch per slab that serves
    // every tap (A is fetched from L2 once instead of once per tap); ring B holds the weights of `gb` taps per slot.
    // operand planes held per slot: A (activations) hi [+ lo when some source uses the a_lo * w_hi term], B (weights)
    // hi [+ lo when some source uses the a_hi * w_lo term]
    const int planes_a = p.planes_a, planes_b = p.planes_b;
    const int a_box_bytes = HALO ? p.bn * p.ph * p.pw * 128 : kAPlaneBytes;    // bytes one TMA box delivers per plane
    const int a_plane_bytes = HALO ? ((a_box_bytes + 1023) & ~1023) : kAPlaneBytes;
    const int a_bytes = planes_a * a_plane_bytes;
    const int b_plane_bytes = (PAIR ? p.n_t / 2 : p.n_t) * 128;     // pair mode: each CTA stages half of the N rows
    const int b_bytes = planes_b * b_plane_bytes;
    const int ks = HALO ? 1 : p.kslab;                               // plain mode: 64-channel slabs per ring slot
    const int slab_bytes = a_bytes + b_bytes;
    const int stage_bytes = HALO ? a_bytes : ks * slab_bytes;
    const int n_stages = p.stages;                                   // plain: ring slots; halo: patch slots
    const int nb_stages = HALO ? p.b_stages : 0;
    const int gb = HALO ? p.gb : 1;
    uint8_t* smem_b = smem + (size_t)n_stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (HALO && p.b_resident ? (size_t)p.b_res_bytes : (size_t)nb_stages * gb * b_bytes));
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = full0 + 8 * n_stages;
    const uint32_t fullB0 = empty0 + 8 * n_stages;
    const uint32_t emptyB0 = fullB0 + 8 * nb_stages;
    const uint32_t tfull0 = emptyB0 + 8 * nb_stages;
    const uint32_t tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * n_stages + 2 * nb_stages + 4);
    // small fp32 tables for the epilogue (rows padded with zeros to cpad = n_ntiles * n_t columns so the epilogue
    // needs no channel guards): bias [cpad], post scale / shift [cpad] each, shortcut weights [cpad],
    // lt weights [cpad] x float4, lt bias [4], partial lt logits [2 acc stages][128 pixels][4]
    const int cpad = p.n_ntiles * p.n_t;
    float* s_ps = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_pt = s_ps + (p.post_scale ? cpad : 0);
    float* s_skipw = s_pt + (p.post_scale ? cpad : 0);
    float* s_topw = s_skipw + (SKIPC > 0 ? cpad : 0);
    float* s_z = s_topw + (TOPK > 0 ? 4 * cpad : 0);
    for (int i = threadIdx.x; i < cpad; i += kThreads) {
        if (p.post_scale) { s_ps[i] = i < p.cout ? p.post_scale[i] : 1.f; s_pt[i] = i < p.cout ? p.post_shift[i] : 0.f; }
        if (SKIPC > 0) s_skipw[i] = i < p.cout ? p.skip_w[i] : 0.f;
    }
    if (TOPK > 0)
        for (int i = threadIdx.x; i < cpad * 4; i += kThreads) s_topw[i] = p.tab_topw[i];

    // warp index through a shuffle: the compiler then knows it (and every table index derived from it) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        if (p.c1 > 0) tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        if ((p.b_resident && p.a1_center) || p.b1_hi_only) tma_prefetch_desc(&mapB1);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < n_stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
            for (int s = 0; s < nb_stages; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, (PAIR ? 2 : 1) * kEpiWarps); }
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
    if ((p.exp_flags & 2048) && blockIdx.x == 0 && threadIdx.x == 0)
        printf("umx tc_conv layout: full0 0x%x empty0 0x%x fullB0 0x%x emptyB0 0x%x tfull0 0x%x tempty0 0x%x stages %d/%d gb %d\n",
               full0, empty0, fullB0, emptyB0, tfull0, tempty0, n_stages, nb_stages, gb);

    const int m_tiles = (p.bn > 1) ? (p.n_tiles + p.bn - 1) / p.bn : p.n_tiles * (p.in_w / p.bw) * (p.in_h / p.bh);
    const int per_unit = PAIR ? 2 : 1;
    const int m_units = (m_tiles + per_unit - 1) / per_unit;
    // conv-transpose with merge_px: a work item covers the two sub-pixel phases (px = 0, 1) of one output row parity;
    // they share every input patch and fill two n_t-column halves of one accumulator stage
    const int nsub = (TOPK == 0 && SKIPC == 0 && p.merge_px) ? 2 : 1;      // (never set for the fused-lt / shortcut variants: keeps them lean)
    const int total = (p.nphase / nsub) * m_units * p.n_ntiles;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nch0 = (p.c0 + 63) >> 6, nch1 = (p.c1 + 63) >> 6;
    TileDecoder decode; decode.init(p, per_unit, rank);
    const bool dbg_on = (p.exp_flags & 64) && p.dbg != nullptr;

    if (warp == 0) {
        if (elect_one()) {
            // ================= TMA producer =================
            const int n_t = p.n_t, c0s = p.c0;
            const bool a1c = p.a1_center != 0; const int ctap = p.center_tap;
            const bool noload = (p.exp_flags & 1) != 0;          // timing experiment: barriers complete without any TMA traffic
            if constexpr (HALO) {
                const int hx0 = p.hx0, hy0 = p.hy0, n_chunks = nch0 + nch1;
                const bool nh = p.halo_nh != 0;
                const uint32_t txA1 = (PAIR ? 2u : 1u) * (uint32_t)a_box_bytes;      // per plane; a source loads its lo plane only when its a_lo term is on
                int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
                unsigned long long c_wa = 0, c_wb = 0, c_work = 0; DbgClock clk; clk.start(dbg_on);
                const uint32_t txB = (PAIR ? 2u : 1u) * (uint32_t)(gb * b_bytes);      // a weight box always carries gb taps
                // Patch loads run n_stages - 1 slabs ahead of the weight loads of the same (item, slab) stream: the
                // slot they wait for was released by a slab whose weights have already been requested.
                int a_tile = item0, a_cb = 0;
                TileCoord ta = decode(a_tile);
                auto issue_a = [&]() {
                    const bool second = a_cb >= nch0;
                    const int cc = (second ? a_cb - nch0 : a_cb) * 64;
                    const CUtensorMap* mapA = second ? &mapA1 : &mapA0;
                    clk.lap(c_work);
                    mbar_wait(empty0 + 8 * sa, pa ^ 1);
                    clk.lap(c_wa);
                    const uint32_t fa = full0 + 8 * sa;
                    const uint32_t da = smem_u32(smem) + (uint32_t)(sa * stage_bytes);
                    const int npl = (((second ? p.terms1 : p.terms0) & 2) && planes_a == 2) ? 2 : 1;
                    if (noload) { if (leader) mbar_arrive(fa); }
                    else if (leader) mbar_expect_tx(fa, txA1 * (uint32_t)npl);
                    const int c2 = nh ? ta.n0 : ta.y0 - hy0, c3 = nh ? ta.y0 - hy0 : ta.n0;     // {C, W, tiles, H} order on 8x8 grids
                    for (int pl = 0; pl < npl && !noload; ++pl) {
                        if (PAIR) tma_load_5d_pair(da + pl * a_plane_bytes, mapA, fa, cc, ta.x0 - hx0, c2, c3, pl);
                        else tma_load_5d(da + pl * a_plane_bytes, mapA, fa, cc, ta.x0 - hx0, c2, c3, pl);
                    }
                    if (++sa == n_stages) { sa = 0; pa ^= 1; }
                    if (++a_cb == n_chunks) { a_cb = 0; a_tile += item_step; if (a_tile < total) ta = decode(a_tile); }
                };
                // weight loads of one slab (one box per slot of gb taps); they run one slab ahead of the MMA thread
                int b_tile = item0, b_cb = 0;
                TileCoord tb = decode(b_tile);
                auto issue_b = [&]() {
                  for (int sub = 0; sub < nsub; ++sub) {
                    const TcPhaseGrid g = p.grid[tb.phase * nsub + sub];
                    const int ncol = tb.n_idx * n_t + (PAIR ? rank * (n_t / 2) : 0);
                    const bool second = b_cb >= nch0;
                    const int cc = (second ? b_cb - nch0 : b_cb) * 64;
                    const bool centre_only = second && a1c;        // 1x1 term: its slabs exist at the centre tap only
                    const int te = centre_only ? 1 : g.ntaps;
                    const int kc = (second ? c0s : 0) + cc;
                    const int wi0 = centre_only ? ctap : g.wi0;
                    // a source without the a_hi*w_lo term fetches the hi plane of the weights only (mapB1: one-plane box)
                    const bool hi_only = planes_b == 2 && p.b1_hi_only && !((second ? p.terms1 : p.terms0) & 1);
                    for (int t0 = 0; t0 < te; t0 += gb) {
                        clk.lap(c_work);
                        mbar_wait(emptyB0 + 8 * sb, pb ^ 1);
                        clk.lap(c_wb);
                        const uint32_t fb = fullB0 + 8 * sb;
                        const uint32_t db = smem_u32(smem_b) + (uint32_t)(sb * gb * b_bytes);
                        if (noload) { if (leader) mbar_arrive(fb); }
                        else {
                            if (leader) mbar_expect_tx(fb, hi_only ? txB / 2 : txB);
                            if (PAIR) tma_load_4d_pair(db, hi_only ? &mapB1 : &mapB, fb, kc, ncol, wi0 + t0, 0);
                            else tma_load_4d(db, hi_only ? &mapB1 : &mapB, fb, kc, ncol, wi0 + t0, 0);
                        }
                        if (++sb == nb_stages) { sb = 0; pb ^= 1; }
                    }
                  }
                    if (++b_cb == n_chunks) { b_cb = 0; b_tile += item_step; if (b_tile < total) tb = decode(b_tile); }
                };
                // order per slab s: weights(s+1), then patch(s+n_stages-1): the wait for a free patch slot (= the MMAs of
                // slab s-1 have retired) never holds back weights the MMA thread needs next
                if (p.b_resident) {
                    // every weight tile once: one box of all taps per slab, all on the first weight barrier
                    if (item0 < total) {
                        if (noload) { if (leader) mbar_arrive(fullB0); }
                        else {
                            if (leader) mbar_expect_tx(fullB0, (PAIR ? 2u : 1u) * (uint32_t)p.b_res_bytes);
                            const int ncol = PAIR ? rank * (n_t / 2) : 0;
                            uint32_t db = smem_u32(smem_b);
                            for (int cb = 0; cb < n_chunks; ++cb) {
                                const bool second = cb >= nch0;
                                const int kc = (second ? c0s : 0) + (second ? cb - nch0 : cb) * 64;
                                if (kNcat && TOPK > 0 && PAIR && p.ncat) {
                                    // one-tile boxes.  Full slab, per tap: X = n_t rows of B for the wide MMA (even CTA: w_hi; odd CTA:
                                    // w_lo, or zeros - rows beyond cout - for a source without the a_hi*w_lo term), then Y = this
                                    // CTA's half of w_hi.  Centre-only slab: hi [+ lo] tile as in the plain resident layout.
                                    if (second && a1c) {
                                        tma_load_4d_pair(db, &mapB, fullB0, kc, ncol, ctap, 0);
                                        if (p.res_c_planes == 2) tma_load_4d_pair(db + b_plane_bytes, &mapB, fullB0, kc, ncol, ctap, 1);
                                        db += p.res_c_planes * b_plane_bytes;
                                    } else {
                                        const bool w_lo = ((second ? p.terms1 : p.terms0) & 1) != 0;
                                        const int xrow = (rank == 0 || w_lo) ? 0 : 2 * n_t, xplane = (rank == 1 && w_lo) ? 1 : 0;
                                        for (int t = 0; t < gb; ++t) {
                                            tma_load_4d_pair(db, &mapB, fullB0, kc, xrow, t, xplane);
                                            tma_load_4d_pair(db + b_plane_bytes, &mapB, fullB0, kc, xrow + n_t / 2, t, xplane);
                                            tma_load_4d_pair(db + 2 * b_plane_bytes, &mapB, fullB0, kc, ncol, t, 0);
                                            db += 3 * b_plane_bytes;
                                        }
                                    }
                                } else
                                if (second && a1c) {        // 1x1 term: only its centre tap exists (one-tap box)
                                    if (PAIR) tma_load_4d_pair(db, &mapB1, fullB0, kc, ncol, ctap, 0);
                                    else tma_load_4d(db, &mapB1, fullB0, kc, ncol, ctap, 0);
                                    db += p.res_c_planes * b_plane_bytes;          // hi [+ lo when that source uses the a_hi*w_lo term]
                                } else {
                                    if (PAIR) tma_load_4d_pair(db, &mapB, fullB0, kc, ncol, 0, 0);
                                    else tma_load_4d(db, &mapB, fullB0, kc, ncol, 0, 0);
                                    db += p.res_m_planes * gb * b_plane_bytes;     // [plane][tap][rows]: hi, + lo when a full-slab source uses a_hi*w_lo
                                }
                            }
                        }
                    }
                    while (a_tile < total) issue_a();
                } else {
                    for (int i = 0; i < n_stages - 1 && a_tile < total; ++i) issue_a();
                    if (b_tile < total) issue_b();
                    while (b_tile < total) {
                        issue_b();
                        if (a_tile < total) issue_a();
                    }
                }
                clk.lap(c_work);
                if (dbg_on && leader) { atomicAdd(p.dbg + 0, c_wa); atomicAdd(p.dbg + 1, c_wb); atomicAdd(p.dbg + 2, c_work); }
            } else {
                int stage = 0; uint32_t phase = 0;
                unsigned long long c_wa = 0, c_work = 0; DbgClock clk; clk.start(dbg_on);
                for (int tile = item0; tile < total; tile += item_step) {
                    const TileCoord t = decode(tile);
                    const TcPhaseGrid g = p.grid[t.phase];
                    const int ncol = t.n_idx * n_t + (PAIR ? rank * (n_t / 2) : 0);
                    int wi = g.wi0, ix = 0, ax = t.x0 + g.dx0, ay = t.y0 + g.dy0;
                    for (int tp = 0; tp < g.ntaps; ++tp) {
                        const int nchunks_tp = nch0 + ((a1c && tp != ctap) ? 0 : nch1);
                        for (int c0 = 0; c0 < nchunks_tp; c0 += ks) {
                            const int ns = min(ks, nchunks_tp - c0);
                            clk.lap(c_work);
                            mbar_wait(empty0 + 8 * stage, phase ^ 1);
                            clk.lap(c_wa);
                            const uint32_t fb = full0 + 8 * stage;
                            uint32_t sa = smem_u32(smem) + (uint32_t)(stage * stage_bytes);
                            if (noload) { if (leader) mbar_arrive(fb); }
                            else if (leader) mbar_expect_tx(fb, (PAIR ? 2u : 1u) * (uint32_t)(ns * slab_bytes));
                            for (int j = 0; j < ns && !noload; ++j, sa += slab_bytes) {
                                const int cb = c0 + j;
                                const bool second = cb >= nch0;
                                const int cc = (second ? cb - nch0 : cb) * 64;
                                if (PAIR) {
                                    tma_load_5d_pair(sa, second ? &mapA1 : &mapA0, fb, cc, ax, ay, t.n0, 0);
                                    tma_load_4d_pair(sa + a_bytes, &mapB, fb, (second ? c0s : 0) + cc, ncol, wi, 0);
                                } else {
                                    tma_load_5d(sa, second ? &mapA1 : &mapA0, fb, cc, ax, ay, t.n0, 0);
                                    tma_load_4d(sa + a_bytes, &mapB, fb, (second ? c0s : 0) + cc, ncol, wi, 0);
                                }
                            }
                            if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        }
                        wi += g.wix; ax += g.dstep;
                        if (++ix == g.nx) { ix = 0; wi += g.wiy - g.nx * g.wix; ax = t.x0 + g.dx0; ay += g.dstep; }
                    }
                }
                clk.lap(c_work);
                if (dbg_on && leader) { atomicAdd(p.dbg + 0, c_wa); atomicAdd(p.dbg + 2, c_work); }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            // ================= MMA issuer (leader CTA only in pair mode) =================
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_t >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
            const int terms0 = p.terms0, terms1 = p.terms1;
            const bool split = (terms0 | terms1) != 0;
            const uint32_t a_plane16 = (uint32_t)a_plane_bytes >> 4, b_plane16 = (uint32_t)b_plane_bytes >> 4;
            const uint32_t b_hi = desc_hi(1024);
            const bool a1c = p.a1_center != 0; const int ctap = p.center_tap;
            const int nk_last0 = (((p.c0 - 1) & 63) + 16) >> 4, nk_last1 = nch1 ? (((p.c1 - 1) & 63) + 16) >> 4 : 0;   // MMAs of the last slab
            const bool one_mma = (p.exp_flags & 4) != 0;
            const int exp_mma = p.exp_flags & (128 | 256);
            int acc = 0; uint32_t acc_phase = 0;
            unsigned long long c_wt = 0, c_wa = 0, c_wb = 0, c_work = 0; DbgClock clk; clk.start(dbg_on);
            if constexpr (HALO) {
                const int pw = p.pw, hx0 = p.hx0, hy0 = p.hy0, n_chunks = nch0 + nch1;
                const int prow = p.halo_nh ? pw * p.bn : pw;           // pixels between two image rows inside the patch
                const uint32_t a_hi = desc_hi((uint32_t)pw * 128u);     // consecutive 8-pixel rows are one patch row apart
                const bool ncat = kNcat && TOPK > 0 && PAIR && p.ncat != 0;      // (only the kernels with the fused lt epilogue carry this path: see lower_conv_tc)
                const uint32_t idesc2 = (1u << 4) | ((uint32_t)(p.n_t >> 2) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);   // width 2*n_t
                const uint32_t acc_cols = ncat ? 2u * (uint32_t)p.n_t : (uint32_t)p.n_t;      // accumulator columns of one (sub-)phase
                const uint32_t b16 = ncat ? 3 * b_plane16 : b_plane16;  // weight slot = [plane][tap][rows]: taps one plane tile apart (ncat: [tap][X, X, Y])
                const uint32_t bp16 = (uint32_t)gb * b_plane16;
                int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
                const bool resident = p.b_resident != 0;
                const uint32_t c_off = (uint32_t)((hy0 * prow + hx0) * 8);      // centre tap of the patch (1x1 term)
                if (resident && item0 < total) { mbar_wait(fullB0, 0); tc_fence_after(); }
                for (int tile = item0; tile < total; tile += item_step) {
                    const int ph0 = decode(tile).phase * nsub;
                    // per sub-phase tap walk, once per item (the slab loop below must stay a handful of instructions per
                    // slab: its single thread paces the N = 80 layers).  A offset (16-byte units) of tap (iy, ix) inside the
                    // patch: off0 + iy*sy + ix*sx
                    // (scalars, not arrays: anything indexed by the sub-phase would live in local memory)
                    const TcPhaseGrid ga = p.grid[ph0], gb2 = p.grid[ph0 + nsub - 1];
                    const int sx0 = ga.dstep * 8, rb0 = ga.dstep * prow * 8 - ga.nx * ga.dstep * 8, nx0 = ga.nx, nt0 = ga.ntaps;
                    const uint32_t off0 = (uint32_t)(((ga.dy0 + hy0) * prow + ga.dx0 + hx0) * 8), wi0 = (uint32_t)ga.wi0 * b16;
                    const int sx1 = gb2.dstep * 8, rb1 = gb2.dstep * prow * 8 - gb2.nx * gb2.dstep * 8, nx1 = gb2.nx, nt1 = gb2.ntaps;
                    const uint32_t off1 = (uint32_t)(((gb2.dy0 + hy0) * prow + gb2.dx0 + hx0) * 8), wi1 = (uint32_t)gb2.wi0 * b16;
                    clk.lap(c_work);
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                    clk.lap(c_wt);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
                    uint32_t accum0 = 0, accum1 = 0;      // per sub-phase accumulator (merge_px)
                    for (int cb = 0; cb < n_chunks; ++cb) {
                        const bool second = cb >= nch0;
                        const int nk = one_mma ? 1 : (cb == nch0 - 1 ? nk_last0 : (cb == n_chunks - 1 ? nk_last1 : 4));
                        const int tsel = split ? (second ? terms1 : terms0) : 0;
                        clk.lap(c_work);
                        mbar_wait(full0 + 8 * sa, pa);
                        clk.lap(c_wa);
                        tc_fence_after();
                        const uint32_t a_lo0 = desc_lo(smem_u32(smem) + (uint32_t)(sa * stage_bytes));
                        const bool centre_only = second && a1c;
                        for (int sub = 0; sub < nsub; ++sub) {        // the taps of one sub-phase against this slab's patch
                            const int sx = sub ? sx1 : sx0, rb = sub ? rb1 : rb0, nx = sub ? nx1 : nx0;
                            const uint32_t tmem_ds = sub ? tmem_d + acc_cols : tmem_d;
                            uint32_t accumulate = sub ? accum1 : accum0;
                            const int te = centre_only ? 1 : (sub ? nt1 : nt0);
                            uint32_t a_lo = a_lo0 + (centre_only ? c_off : (sub ? off1 : off0));
                            int ix = 0;
                            if (resident) {
                                // weights stay put: slab cb's taps start at cb * gb tiles, addressed by their device tap index
                                // (a centre-only slab keeps just that one tap, so the slabs after it start gb - 1 tiles earlier)
                                const uint32_t r_off = (uint32_t)((centre_only ? nch0 * gb * p.res_m_planes + (cb - nch0) * p.res_c_planes : cb * gb * p.res_m_planes) * b_plane_bytes);
                                const uint32_t b_lo = desc_lo(smem_u32(smem_b) + r_off) + (centre_only ? 0u : (sub ? wi1 : wi0));
                                if (exp_mma & 128) {}
                                else if (ncat && !centre_only) halo_taps_ncat_nk<PAIR>(tsel, nk, te, tmem_ds, a_lo, a_hi, b_lo, b_hi, idesc, idesc2, a_plane16, 2 * b_plane16, b16, sx, rb, nx, ix, accumulate);
                                else halo_taps_terms<PAIR>(tsel, nk, te, tmem_ds, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, centre_only ? b_plane16 : bp16, b16, sx, rb, nx, ix, accumulate);
                            } else
                            for (int t0 = 0; t0 < te; t0 += gb) {
                                const int ng = min(gb, te - t0);
                                clk.lap(c_work);
                                mbar_wait(fullB0 + 8 * sb, pb);
                                clk.lap(c_wb);
                                tc_fence_after();
                                const uint32_t b_lo = desc_lo(smem_u32(smem_b) + (uint32_t)(sb * gb * b_bytes));
                                if (exp_mma & 128) {}                 // timing experiment: the issue loop without the MMAs
                                else halo_taps_terms<PAIR>(tsel, nk, ng, tmem_ds, a_lo, a_hi, b_lo, b_hi, idesc, a_plane16, bp16, b16, sx, rb, nx, ix, accumulate);
                                if (PAIR) umma_commit_pair(emptyB0 + 8 * sb); else umma_commit(emptyB0 + 8 * sb);
                                if (++sb == nb_stages) { sb = 0; pb ^= 1; }
                            }
                            if (sub) accum1 = accumulate; else accum0 = accumulate;
                        }
                        if (PAIR) umma_commit_pair(empty0 + 8 * sa); else umma_commit(empty0 + 8 * sa);
                        if (++sa == n_stages) { sa = 0; pa ^= 1; }
                    }
                    if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else {
                const uint32_t a_hi = b_hi;
                const uint32_t slab16 = (uint32_t)slab_bytes >> 4, ab16 = (uint32_t)a_bytes >> 4;
                int stage = 0; uint32_t phase = 0;
                for (int tile = item0; tile < total; tile += item_step) {
                    const int ntap = p.grid[decode(tile).phase].ntaps;
                    clk.lap(c_work);
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                    clk.lap(c_wt);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
                    uint32_t accumulate = 0;
                    for (int tp = 0; tp < ntap; ++tp) {
                        const int nchunks_tp = nch0 + ((a1c && tp != ctap) ? 0 : nch1);
                        for (int c0 = 0; c0 < nchunks_tp; c0 += ks) {
                            const int ns = min(ks, nchunks_tp - c0);
                            clk.lap(c_work);
                            mbar_wait(full0 + 8 * stage, phase);
                            clk.lap(c_wa);
                            tc_fence_after();
                            uint32_t a_lo = desc_lo(smem_u32(smem) + (uint32_t)(stage * stage_bytes));
                            for (int j = 0; j < ns; ++j, a_lo += slab16) {
                                const int cb = c0 + j;
                                const int nk = one_mma ? 1 : (cb == nch0 - 1 ? nk_last0 : (cb == nch0 + nch1 - 1 ? nk_last1 : 4));
                                umma_slab<PAIR>(tmem_d, a_lo, a_hi, a_lo + ab16, b_hi, idesc, nk, split, cb >= nch0 ? terms1 : terms0, a_plane16, b_plane16, accumulate, exp_mma);
                            }
                            // frees the smem slot (in both CTAs) when these MMAs retire
                            if (PAIR) umma_commit_pair(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
                            if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        }
                    }
                    // accumulator ready for the epilogue (of both CTAs)
                    if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
            clk.lap(c_work);
            if (dbg_on) { atomicAdd(p.dbg + 4, c_wt); atomicAdd(p.dbg + 5, c_wa); atomicAdd(p.dbg + 6, c_wb); atomicAdd(p.dbg + 7, c_work); atomicAdd(p.dbg + 15, 1ull); }
        }
    } else {
        // ================= epilogue (kEpiWarps warps: kEpiSub per TMEM lane quarter, each a share of the N columns) =================
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        const int sub = (warp - 2) >> 2;                            // which share of the N columns it handles
        EpiCtx e;
        e.s_ps = s_ps; e.s_pt = s_pt; e.s_skipw = s_skipw; e.s_topw4 = reinterpret_cast<const float4*>(s_topw);
        e.out_f = p.out_f; e.out_h = p.out_h; e.out_plane_elems = p.out_plane_elems;
        e.cout = p.cout; e.out_cs = p.out_cs; e.out_planes = p.out_planes; e.act = p.act; e.pool = p.pool; e.bw = p.halo_nh ? p.bw * p.bn : p.bw;      // lane distance of the pixel one image row down
        e.leaky = p.leaky; e.has_post = p.post_scale != nullptr;
        e.wide_store = p.out_f == nullptr && p.out_h != nullptr && (p.out_cs & 15) == 0 && p.n_ntiles * p.n_t <= p.out_cs &&
                       (p.out_plane_elems & 15) == 0 && !(p.exp_flags & 512);
        // the N tile is dealt out in groups of 8 columns (16-byte fp16 stores): n_t = 80 -> 3, 3, 2, 2 groups
        const int n_t = p.n_t, n8 = n_t >> 3;
        // merge_px (two px phases side by side in one stage): the two warps of a lane quarter take one phase each, i.e. all
        // n_t columns of it and the output pixel `sub` to the right; otherwise they share the columns of one phase
        const bool px_split = nsub == 2;
        const int g_lo = px_split ? 0 : sub * (n8 / kEpiSub) + min(sub, n8 % kEpiSub);          // first column group (channels)
        const int g_n = px_split ? n8 : n8 / kEpiSub + (sub < n8 % kEpiSub ? 1 : 0);
        const bool fold = kNcat && TOPK > 0 && PAIR && HALO && p.ncat != 0;      // the accumulator is [0, n_t) + [n_t, 2 n_t): see halo_taps_ncat
        const uint32_t col0 = (uint32_t)((px_split ? sub * (fold ? 2 * n_t : n_t) : 0) + g_lo * 8);                 // first TMEM column
        const int px_off = px_split ? sub : 0;
        // x16 TMEM loads per warp, plus one x8 load for an odd group count; timing experiments: 8 = first chunk
        // only, 32 = barrier handshake only, 16 = no global stores
        const bool noepi = (p.exp_flags & 32) != 0, one_chunk = (p.exp_flags & 8) != 0;
        const int lead8 = (!noepi && !one_chunk && (g_lo & 1) && g_n > 0) ? 1 : 0;
        const int n_full = noepi ? 0 : (one_chunk ? min(1, g_n >> 1) : ((g_n - lead8) >> 1));
        const bool tail8 = ((g_n - lead8) & 1) && !noepi && !one_chunk;
        const bool nostore = (p.exp_flags & 16) != 0;
        const uint32_t nap_ns = p.epi_nap_ns;
        const int in_h = p.in_h, in_w = p.in_w, os = p.os, n_tiles = p.n_tiles;
        const int m = q * 32 + lane;
        const int xl = m % p.bw;
        const int yl = p.halo_nh ? m / (p.bw * p.bn) : (m / p.bw) % p.bh, nl = p.halo_nh ? (m / p.bw) % p.bn : m / (p.bw * p.bh);
        const int oh_ = e.pool ? in_h >> 1 : in_h * os, ow_ = e.pool ? in_w >> 1 : in_w * os;
        int acc = 0; uint32_t acc_phase = 0;
        unsigned long long c_wt = 0, c_work = 0; DbgClock clk; clk.start(dbg_on && leader && lane == 0 && warp == 2);
        for (int tile = item0; tile < total; tile += item_step) {
            const TileCoord t = decode(tile);
            const int n = t.n0 + nl, y = t.y0 + yl, x = t.x0 + xl;
            const bool valid = n < n_tiles && !nostore;
            int oy, ox; bool writer = valid;
            if (e.pool) { oy = y >> 1; ox = x >> 1; writer = valid && !(y & 1) && !(x & 1); }
            else if (os == 2) { oy = 2 * y + (px_split ? t.phase : (t.phase >> 1)); ox = 2 * x + (px_split ? px_off : (t.phase & 1)); }
            else { oy = y; ox = x; }
            const int64_t opix = ((int64_t)n * oh_ + oy) * ow_ + ox;
            const int co0 = t.n_idx * n_t + g_lo * 8;              // first output channel of this warp's columns
            float* of = e.out_f + opix * e.cout + co0;              // only dereferenced when the base is non-null
            __half* oh = e.out_h + opix * e.out_cs + co0;
            float xs = 0.f;
            if (SKIPC > 0 && valid) xs = __ldg(p.skip_src + ((int64_t)n * in_h + y) * in_w + x);
            float z[4] = {0.f, 0.f, 0.f, 0.f};           // fused lt logits
            clk.lap(c_work);
            if (nap_ns) mbar_wait_backoff(tfull0 + 8 * acc, acc_phase, nap_ns); else mbar_wait_relaxed(tfull0 + 8 * acc, acc_phase);
            clk.lap(c_wt);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride) + col0;
            // the TMEM load of the next chunk is in flight while the current one is processed; an odd first group is
            // taken alone so that every 16-column chunk starts on a 32-byte boundary of the fp16 row (256-bit stores)
            uint32_t ra[16], rb[16], rt[8];
            uint32_t ra2[16], rb2[16], rt2[8];            // ncat: the same columns of the second accumulator half
            auto ld16 = [&](uint32_t addr, uint32_t (&r)[16], uint32_t (&r2)[16]) { tmem_ld16_issue(addr, r); if (fold) tmem_ld16_issue(addr + (uint32_t)n_t, r2); };
            auto ld8 = [&](uint32_t addr, uint32_t (&r)[8], uint32_t (&r2)[8]) { tmem_ld8_issue(addr, r); if (fold) tmem_ld8_issue(addr + (uint32_t)n_t, r2); };
            auto add16 = [&](uint32_t (&r)[16], const uint32_t (&r2)[16]) {
                if (fold) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
            };
            auto add8 = [&](uint32_t (&r)[8], const uint32_t (&r2)[8]) {
                if (fold) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
            };
            const uint32_t tb16 = tbase + lead8 * 8;
            const int cb16 = co0 + lead8 * 8;
            float* of16 = of + lead8 * 8; __half* oh16 = oh + lead8 * 8;
            if (lead8) ld8(tbase, rt, rt2); else if (n_full > 0) ld16(tb16, ra, ra2); else if (tail8) ld8(tb16, rt, rt2);
            if (lead8) {
                tmem_ld_wait();
                add8(rt, rt2);
                if (n_full > 0) ld16(tb16, ra, ra2);
                epi_chunk<8, SKIPC, TOPK>(p, e, rt, co0, xs, writer, of, oh, z);
                __syncwarp();        // reconverge before the next .sync.aligned TMEM instruction
                if (n_full == 0 && tail8) ld8(tb16, rt, rt2);
            }
            for (int i = 0; i < n_full; i += 2) {
                tmem_ld_wait();
                add16(ra, ra2);
                if (i + 1 < n_full) ld16(tb16 + (i + 1) * 16, rb, rb2); else if (tail8) ld8(tb16 + n_full * 16, rt, rt2);
                epi_chunk<16, SKIPC, TOPK>(p, e, ra, cb16 + i * 16, xs, writer, of16 + i * 16, oh16 + i * 16, z);
                __syncwarp();
                if (i + 1 < n_full) {
                    tmem_ld_wait();
                    add16(rb, rb2);
                    if (i + 2 < n_full) ld16(tb16 + (i + 2) * 16, ra, ra2); else if (tail8) ld8(tb16 + n_full * 16, rt, rt2);
                    epi_chunk<16, SKIPC, TOPK>(p, e, rb, cb16 + (i + 1) * 16, xs, writer, of16 + (i + 1) * 16, oh16 + (i + 1) * 16, z);
                    __syncwarp();
                }
            }
            if (tail8) {
                tmem_ld_wait();
                add8(rt, rt2);
                epi_chunk<8, SKIPC, TOPK>(p, e, rt, cb16 + n_full * 16, xs, writer, of16 + n_full * 16, oh16 + n_full * 16, z);
                __syncwarp();
            }
            if (TOPK > 0) {
                // the warps of a lane quarter hold the column shares of each pixel: combine the partial logits
                float4* zs = reinterpret_cast<float4*>(s_z) + (acc * (kEpiSub - 1) * 128 + m);
                if (sub > 0) zs[(sub - 1) * 128] = make_float4(z[0], z[1], z[2], z[3]);
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kEpiSub) : "memory");
                if (sub == 0 && writer) {
#pragma unroll
                    for (int o_ = 0; o_ < kEpiSub - 1; ++o_) {
                        const float4 o = zs[o_ * 128];
                        z[0] += o.x; z[1] += o.y; z[2] += o.z; z[3] += o.w;
                    }
                    float mx = -INFINITY, sum = 0.f;
#pragma unroll
                    for (int k = 0; k < TOPK; ++k) { z[k] += p.tab_topb[k]; mx = fmaxf(mx, z[k]); }
#pragma unroll
                    for (int k = 0; k < TOPK; ++k) { z[k] = expf(z[k] - mx); sum += z[k]; }
                    const float inv = 1.f / sum;
                    float* pr = p.top_probs + opix * TOPK;
#pragma unroll
                    for (int k = 0; k < TOPK; ++k) pr[k] = z[k] * inv;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * acc); else mbar_arrive(tempty0 + 8 * acc); }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        clk.lap(c_work);
        if (clk.on) { atomicAdd(p.dbg + 8, c_wt); atomicAdd(p.dbg + 9, c_work); }
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
    const size_t tables = (cpad * ((p.post_scale ? 2 : 0) + (p.skip_src ? 1 : 0) + (p.top_w ? 4 : 0)) + (p.top_w ? 2 * (kEpiSub - 1) * 128 * 4 : 0) + 8) * sizeof(float);
    return (2 * (size_t)(p.stages + (p.halo ? p.b_stages : 0)) + 4) * 8 + 16 + tables + 1024;
}
size_t tc_conv_a_bytes(const TcConvParams& p) {
    const size_t box = p.halo ? (((size_t)p.bn * p.ph * p.pw * 128 + 1023) & ~(size_t)1023) : (size_t)kAPlaneBytes;
    return (size_t)p.planes_a * box;
}
size_t tc_conv_b_bytes(const TcConvParams& p) { return (size_t)p.planes_b * (size_t)(p.pair ? p.n_t / 2 : p.n_t) * 128; }
size_t tc_conv_smem_bytes(const TcConvParams& p) {
    if (p.halo) return p.stages * tc_conv_a_bytes(p) + (p.b_resident ? (size_t)p.b_res_bytes : (size_t)p.b_stages * p.gb * tc_conv_b_bytes(p)) + tc_conv_fixed_bytes(p);
    return p.stages * (size_t)(p.kslab > 0 ? p.kslab : 1) * (tc_conv_a_bytes(p) + tc_conv_b_bytes(p)) + tc_conv_fixed_bytes(p);
}

template <int SKIPC, int TOPK, bool PAIR, bool HALO>
static cudaError_t configure_one() {
    return cudaFuncSetAttribute(tc_conv_kernel<SKIPC, TOPK, PAIR, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

template <int SKIPC, int TOPK>
static cudaError_t configure_epi() {
    cudaError_t e = configure_one<SKIPC, TOPK, false, false>();
    if (e == cudaSuccess) e = configure_one<SKIPC, TOPK, true, false>();
    if (e == cudaSuccess) e = configure_one<SKIPC, TOPK, false, true>();
    if (e == cudaSuccess) e = configure_one<SKIPC, TOPK, true, true>();
    return e;
}

cudaError_t tc_conv_configure() {
    cudaError_t e = configure_epi<0, 0>();
    if (e == cudaSuccess) e = configure_epi<1, 0>();
    if (e == cudaSuccess) e = configure_epi<0, 2>();
    if (e == cudaSuccess) e = configure_epi<0, 3>();
    if (e == cudaSuccess) e = configure_epi<0, 4>();
    return e;
}

template <int SKIPC, int TOPK, bool PAIR, bool HALO>
static cudaError_t launch_one(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b1, const TcConvParams& p,
                              int grid, size_t smem, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_conv_kernel<SKIPC, TOPK, PAIR, HALO>, a0, a1, b, b1, p);
}

template <int SKIPC, int TOPK>
static cudaError_t launch_epi(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b1, const TcConvParams& p,
                              int grid, size_t smem, cudaStream_t s) {
    if (p.pair) return p.halo ? launch_one<SKIPC, TOPK, true, true>(a0, a1, b, b1, p, grid, smem, s) : launch_one<SKIPC, TOPK, true, false>(a0, a1, b, b1, p, grid, smem, s);
    return p.halo ? launch_one<SKIPC, TOPK, false, true>(a0, a1, b, b1, p, grid, smem, s) : launch_one<SKIPC, TOPK, false, false>(a0, a1, b, b1, p, grid, smem, s);
}

cudaError_t launch_tc_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b1,
                           const TcConvParams& p, int num_sms, cudaStream_t s) {
    const int m_tiles = (p.bn > 1) ? (p.n_tiles + p.bn - 1) / p.bn : p.n_tiles * (p.in_w / p.bw) * (p.in_h / p.bh);
    const int per_unit = p.pair ? 2 : 1;
    const int total = (p.nphase / (p.merge_px ? 2 : 1)) * ((m_tiles + per_unit - 1) / per_unit) * p.n_ntiles;
    if (total == 0) return cudaSuccess;
    int grid = total * per_unit < num_sms ? total * per_unit : num_sms;
    if (p.pair) grid &= ~1;
    // at least ~120 KB so that exactly one CTA (and its 512 TMEM columns) lives on an SM
    size_t smem = tc_conv_smem_bytes(p);
    if (smem < 120 * 1024) smem = 120 * 1024;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const bool skip = p.skip_src != nullptr;
    if (skip && (p.skip_c != 1 || p.top_w)) return cudaErrorInvalidValue;
    if (skip) return launch_epi<1, 0>(a0, a1, b, b1, p, grid, smem, s);
    if (!p.top_w) return launch_epi<0, 0>(a0, a1, b, b1, p, grid, smem, s);
    if (p.top_k == 2) return launch_epi<0, 2>(a0, a1, b, b1, p, grid, smem, s);
    if (p.top_k == 3) return launch_epi<0, 3>(a0, a1, b, b1, p, grid, smem, s);
    if (p.top_k == 4) return launch_epi<0, 4>(a0, a1, b, b1, p, grid, smem, s);
    return cudaErrorInvalidValue;
}

int make_act_tensor_map(CUtensorMap* out, const __half* base, int planes, int64_t plane_elems, int n, int h, int w, int c,
                        int bw, int bh, int bn, int box_planes, int nh_order) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)planes};
    cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)plane_elems * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)box_planes};
    if (nh_order) {          // tiles before rows: box memory order [h][tile][w][c]
        std::swap(dims[2], dims[3]); std::swap(strides[1], strides[2]); std::swap(box[2], box[3]);
    }
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int make_weight_tensor_map(CUtensorMap* out, const __half* base, int planes, int taps, int cout, int cin, int n_t,
                           int box_planes, int box_taps) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps, (cuuint64_t)planes};
    cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2, (cuuint64_t)taps * cout * cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)n_t, (cuuint32_t)box_taps, (cuuint32_t)box_planes};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace umx
