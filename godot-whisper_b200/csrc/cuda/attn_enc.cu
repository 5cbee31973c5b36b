// Fused encoder self-attention for sm_100a:  O = softmax(Q K^T / 8) V  for one (chunk, head, 128-query tile) per CTA.
//
// Replaces the KQ mul_mat -> scale -> soft_max -> V·P mul_mat -> permute/cpy chain of whisper_build_graph_encoder
// (/root/reference/thirdparty/whisper.cpp/whisper.cpp:1880-1917) with ggml's arithmetic
// (/root/reference/thirdparty/whisper.cpp/ggml.c:9737-9948 mul_mat, :11116-11201 soft_max): f16 operands, f32 accumulation,
// row maximum first, exp through the f16 table on f16-rounded (s - max), an exact sum of the table values, probabilities
// p = f16(e * (float)(1/sum)) and only then the product with V.  Scores never leave the chip: because the reference rounds the
// normalised probabilities, the row statistics must be final before the first P·V product, so the score tiles are recomputed
// three times on the tensor cores (max pass, sum pass, product pass) instead of being stored — 2 x 4·T²·64 extra flops per
// head against 12 bytes of HBM traffic per score.
//
//   TMA (128B swizzle): Q tile once, K tiles (3 passes) and V^T tiles (last pass) through shared-memory rings
//   tcgen05.mma kind::f16: S = Q K^T (128 x 128 x 64) into a double-buffered TMEM accumulator; O += P V (128 x 64 x 128)
//   softmax warps: tcgen05.ld of their 32 lanes (one query row per thread), exp table held in shared memory (the 19 584 entries
//   that are not zero), P written to shared memory in the UMMA K-major swizzled layout, O drained through tcgen05.ld at the end.
//
// Warp roles: 0..4 WPQ - 1 softmax (warp w: TMEM lane quadrant w & 3, key columns (128 / WPQ) (w >> 2) .. of every tile), then the
// K/Q producer (also the TMEM allocator), the MMA issuer, the V^T producer.
//
// The kernel is a template over its resources — softmax warps per lane quadrant (WPQ), ring depths (KS, VS), and the form of the
// on-chip exp table (ITAB) — because what bounds it is the softmax warps' instruction issue and the shared-memory look-ups, not the
// tensor pipe:
//   ITAB = false: the table holds the f16 values; the sum pass converts each to 2^-24 units (HADD2.F32, FMUL, F2I per score)
//   ITAB = true:  the table holds e * 2^24 as 32-bit integers (built from the f16 table when the CTA starts): the sum pass is
//                 LDS + IADD, the product pass multiplies (float) F by inv * 2^-24 — the same product, scaled by a power of two,
//                 so it rounds to the same f16 probability.
#include "dev.cuh"
#include "kernels.cuh"
#include "tc.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>

namespace wb200 {

bool gemm_tc_make_map(const Operand & op, int K, int nb1, int nb2, int box_rows, void * out_map);

namespace {

using namespace tc;

constexpr int kQRows   = 128;                 // query rows per CTA (UMMA M)
constexpr int kKeys    = 128;                 // keys per score tile (UMMA N)
constexpr int kTile    = 128 * 64 * 2;        // bytes of a 128-row x 64-column f16 tile (one swizzle atom wide)
constexpr int kExpTab  = 19584;               // entries of the exp table kept on chip: exp(x) rounds to zero in f16 beyond
constexpr int kAttnVariants = 24, kAttnDefaultVariant = 5;

// dynamic shared memory: [exp table | row-statistics scratch | mbarriers | TMEM slot] at fixed offsets from its start (the table
// look-ups compile to LDS with an immediate offset), then the TMA / UMMA tiles from the next 1024-byte boundary
template <int WPQ, int KS, int VS, bool ITAB, int NSB = 2, int PIPE = 1>
struct AttnCfg {
    static constexpr int kWPQ     = WPQ;                 // softmax warps per TMEM lane quadrant
    static constexpr int kCols    = kKeys / WPQ;         // key columns of a tile per softmax thread
    static constexpr int kSmWarps = 4 * WPQ;
    static constexpr int kKS = KS, kVS = VS;             // ring depths
    static constexpr int kPipe = PIPE;                   // 0: load, compute, release   1: release as soon as the copy has landed + next tile prefetched   2: early release only   3: 0 + reuse
    static constexpr int kNSB = NSB;                     // score buffers in TMEM (128 columns each; O sits behind them)
    static constexpr int kThreads = (kSmWarps + 3) * 32;
    static constexpr int kNumBar = 1 + 2 * KS + 2 * VS + 2 * NSB + 2 + 2 + 1 + 1 + 2;
    static_assert(NSB * kKeys + 64 <= 512, "TMEM columns");
    static constexpr int kOffTab = 0;
    static constexpr int kOffRed = kOffTab + kExpTab * (ITAB ? 4 : 2);
    static constexpr int kOffBar = kOffRed + WPQ * kQRows * 8;
    static constexpr int kHeadBytes = kOffBar + kNumBar * 8 + 16;
    static constexpr int kOffQ   = 0;                                   // tile offsets, relative to the aligned tile base
    static constexpr int kOffK   = kOffQ + kTile;
    static constexpr int kOffV   = kOffK + KS * kTile;
    static constexpr int kOffP   = kOffV + VS * kTile;                 // V^T tile: 2 atoms of 64 rows x 64 keys = 16 KB
    static constexpr int kTileBytes = kOffP + 2 * 2 * kTile;            // P tile: 2 atoms of 128 rows x 64 keys = 32 KB, double-buffered
    static constexpr int kSmemBytes = kHeadBytes + 1024 + kTileBytes;
    static_assert(kSmemBytes <= 227 * 1024, "shared-memory plan does not fit");
    static_assert(kOffRed % 8 == 0 && kOffBar % 8 == 0, "alignment");
    static_assert(kCols == 32 || kCols == 64, "a softmax thread reads one or two 32-column slices per tile");
    static_assert(PIPE != 3 || !ITAB, "kept table values are f16 bits");
    static_assert(kThreads * 104 <= 65536 || WPQ <= 2, "register file");
};

// f16 bit patterns of two softmax arguments y = max - s >= 0 (the table index of x = -y), clamped into the table, packed in halves
__device__ __forceinline__ uint32_t exp_index2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return __vminu2(*(const uint32_t *) &h, (uint32_t) (kExpTab - 1) * 0x10001u);
}

template <bool ITAB> struct ExpTab;
template <> struct ExpTab<false> {                  // f16 entries
    static __device__ __forceinline__ void fetch(const uint8_t * tab, uint32_t idx2, uint32_t & a, uint32_t & b) {
        const uint32_t off = idx2 << 1;
        a = *(const uint16_t *) (tab + (off & 0xffffu));
        b = *(const uint16_t *) (tab + (off >> 16));
    }
    static __device__ __forceinline__ unsigned int units(uint32_t v) { return (unsigned int) (__half2float(__ushort_as_half((uint16_t) v)) * 16777216.0f); }
    static __device__ __forceinline__ float value(uint32_t v) { return __half2float(__ushort_as_half((uint16_t) v)); }
    static __device__ __forceinline__ float inv_for(float inv) { return inv; }
};
template <> struct ExpTab<true> {                   // e * 2^24 as integers
    static __device__ __forceinline__ void fetch(const uint8_t * tab, uint32_t idx2, uint32_t & a, uint32_t & b) {
        a = *(const uint32_t *) (tab + ((idx2 & 0xffffu) << 2));
        b = *(const uint32_t *) (tab + ((idx2 >> 16) << 2));
    }
    static __device__ __forceinline__ unsigned int units(uint32_t v) { return v; }
    static __device__ __forceinline__ float value(uint32_t v) { return (float) v; }              // exact: v <= 2^24
    static __device__ __forceinline__ float inv_for(float inv) { return inv * (1.0f / 16777216.0f); }   // exact scaling, no underflow (inv >= 2^-11)
};

// 32 scores of one row -> sum of their table exponentials in units of 2^-24 (exact: every f16 value is a multiple of 2^-24).
// y = max - s / 8: the product by 1/8 is exact, so one fused operation rounds like the reference's scale followed by its subtract
// (negated: rounding to nearest is symmetric, and the largest score gives +0 = the index of exp(-0)).
// FULL = every column is a real key; otherwise only the first n_valid are.
template <bool FULL, bool ITAB, int DBG>
__device__ __forceinline__ unsigned int sum_slice(const uint32_t (&r)[32], const uint8_t * tab, float mxs, int n_valid) {
    unsigned int isum = 0;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const uint32_t idx2 = exp_index2(fmaf(__uint_as_float(r[i]), -0.125f, mxs), fmaf(__uint_as_float(r[i + 1]), -0.125f, mxs));
        uint32_t v0, v1;
        if (DBG & 1) { v0 = idx2 & 0xffffu; v1 = idx2 >> 16; } else ExpTab<ITAB>::fetch(tab, idx2, v0, v1);     // (DBG: timing experiments only)
        unsigned int u0 = ExpTab<ITAB>::units(v0), u1 = ExpTab<ITAB>::units(v1);
        if (!FULL) { if (i >= n_valid) u0 = 0; if (i + 1 >= n_valid) u1 = 0; }
        isum += u0 + u1;
    }
    return isum;
}

// ... and the table values themselves (f16 bits, two per word; zero for padding columns) for the product pass (f16 table only)
template <bool FULL, int DBG>
__device__ __forceinline__ unsigned int sum_slice_keep(const uint32_t (&r)[32], const uint8_t * tab, float mxs, int n_valid, uint32_t (&keep)[16]) {
    unsigned int isum = 0;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const uint32_t idx2 = exp_index2(fmaf(__uint_as_float(r[i]), -0.125f, mxs), fmaf(__uint_as_float(r[i + 1]), -0.125f, mxs));
        uint32_t v0, v1;
        if (DBG & 1) { v0 = idx2 & 0xffffu; v1 = idx2 >> 16; } else ExpTab<false>::fetch(tab, idx2, v0, v1);
        if (!FULL) { if (i >= n_valid) v0 = 0; if (i + 1 >= n_valid) v1 = 0; }
        isum += ExpTab<false>::units(v0) + ExpTab<false>::units(v1);
        keep[i >> 1] = v0 | (v1 << 16);
    }
    return isum;
}

// 32 scores of one row -> 32 probabilities p = f16(e * inv), packed in pairs (`inv` as ExpTab::inv_for made it)
template <bool FULL, bool ITAB, int DBG>
__device__ __forceinline__ void prob_slice(const uint32_t (&r)[32], const uint8_t * tab, float mxs, float inv, int n_valid, uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const uint32_t idx2 = exp_index2(fmaf(__uint_as_float(r[i]), -0.125f, mxs), fmaf(__uint_as_float(r[i + 1]), -0.125f, mxs));
        uint32_t v0, v1;
        if (DBG & 1) { v0 = idx2 & 0xffffu; v1 = idx2 >> 16; } else ExpTab<ITAB>::fetch(tab, idx2, v0, v1);     // (DBG: timing experiments only)
        float e0 = ExpTab<ITAB>::value(v0), e1 = ExpTab<ITAB>::value(v1);
        if (!FULL) { if (i >= n_valid) e0 = 0.0f; if (i + 1 >= n_valid) e1 = 0.0f; }
        const __half2 p2 = __floats2half2_rn(__fmul_rn(e0, inv), __fmul_rn(e1, inv));
        pk[i >> 1] = *(const uint32_t *) &p2;
    }
}

template <class C, bool ITAB, int DBG>
__global__ void __launch_bounds__(C::kThreads, 1)
k_attn_enc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
           __half * __restrict__ out, int T, int d, const uint16_t * __restrict__ exp_lut, int n_qt, int n_head, int n_items) {
    constexpr int kWPQ = C::kWPQ, kCols = C::kCols, kSmWarps = C::kSmWarps, kKS = C::kKS, kVS = C::kVS, kThreads = C::kThreads, kNSB = C::kNSB;
    constexpr int kNumBar = C::kNumBar, kOffTab = C::kOffTab, kOffRed = C::kOffRed, kOffBar = C::kOffBar, kHeadBytes = C::kHeadBytes;
    constexpr int kOffQ = C::kOffQ, kOffK = C::kOffK, kOffV = C::kOffV, kOffP = C::kOffP;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + (uint32_t) kHeadBytes + 1023u) & ~1023u;     // tile base
    const uint8_t * tab = smem_raw + kOffTab;
    unsigned long long * red = (unsigned long long *) (smem_raw + kOffRed);  // [kWPQ][128]: row maxima (as floats), then row sums
    const uint32_t bar0 = smem_u32(smem_raw) + kOffBar;
    uint32_t * tmem_slot = (uint32_t *) (smem_raw + kOffBar + kNumBar * 8);

    const uint32_t q_full = bar0;
    const uint32_t k_full = q_full + 8, k_empty = k_full + 8 * kKS;
    const uint32_t v_full = k_empty + 8 * kKS, v_empty = v_full + 8 * kVS;
    const uint32_t s_full = v_empty + 8 * kVS, s_empty = s_full + 8 * kNSB;
    const uint32_t p_full = s_empty + 8 * kNSB, p_empty = p_full + 16;
    const uint32_t o_full = p_empty + 16, tab_full = o_full + 8, q_empty = tab_full + 8, o_empty = q_empty + 8;

    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // (provably warp-uniform: the role branches stay converged)
    // work item w = (query tile, head, chunk), query tile fastest: CTA c takes w = c, c + gridDim.x, ... (one item per CTA, or a persistent CTA per SM:
    // the barriers, the TMEM allocation and the exp table are set up once, and the loads of the next item run under the tail of this one)
    auto item_coords = [&](int w, int & q0, int & head, int & chunk) { q0 = (w % n_qt) * kQRows; const int r = w / n_qt; head = r % n_head; chunk = r / n_head; };
    const int nt = (T + kKeys - 1) / kKeys;            // key tiles per pass
    // Score tiles of one work item.  Plain order: the nt key tiles three times (max pass, sum pass, product pass).  With reuse
    // (C::kPipe == 3) three of them are never computed: the sum pass starts on the scores of key tile nt - 1, which the softmax threads
    // still hold from the max pass, and then runs DOWN to key tile 0; the table values of its last two tiles (1 and 0) stay in
    // registers, packed, and become the first two P tiles of the product pass.  P tiles are still produced in the order 0 .. nt - 1,
    // so the output bits do not change.
    const bool reuse = C::kPipe == 3 && nt >= 3;
    const int NT = reuse ? 3 * nt - 3 : 3 * nt;
    auto key_tile_of = [&](int t) { return !reuse ? t % nt : t < nt ? t : t < 2 * nt - 1 ? 2 * nt - 2 - t : t - (2 * nt - 1) + 2; };

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < kKS; ++s) { mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1); }
        for (int s = 0; s < kVS; ++s) { mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1); }
        for (int s = 0; s < kNSB; ++s) { mbar_init(s_full + 8 * s, 1); mbar_init(s_empty + 8 * s, kSmWarps); }
        for (int s = 0; s < 2; ++s) { mbar_init(p_full + 8 * s, kSmWarps); mbar_init(p_empty + 8 * s, 1); }
        mbar_init(o_full, 1);
        mbar_init(tab_full, 1);
        mbar_init(q_empty, 1);
        mbar_init(o_empty, kSmWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kSmWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the non-zero part of the exp table (arguments -0 .. -17.3), indexed by the f16 bit pattern without its sign.  The f16 form is one
    // bulk copy that lands while the max pass runs (tab_full); the integer form is converted by all threads here.
    if (ITAB) {
        for (int i = threadIdx.x; i < kExpTab / 2; i += kThreads) {
            const uint32_t w = __ldg((const uint32_t *) (exp_lut + 0x8000) + i);
            const uint32_t lo = (uint32_t) (__half2float(__ushort_as_half((uint16_t) (w & 0xffffu))) * 16777216.0f);
            const uint32_t hi = (uint32_t) (__half2float(__ushort_as_half((uint16_t) (w >> 16))) * 16777216.0f);
            *(uint2 *) (smem_raw + kOffTab + 8 * i) = make_uint2(lo, hi);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + kNSB * kKeys;

    // The three single-purpose warps walk their loops and wait on the barriers as whole warps; one elected lane (elect.sync) issues the
    // TMA / tcgen05 instructions, so their uniform-register operands need no one-lane-at-a-time loop around every instruction.
    if (warp == kSmWarps) {
        // ---- Q / K producer: the K tiles of the three passes ----
        int kt = 0, n_done = 0;                                  // K tiles / items so far: ring positions and barrier parities run on across items
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n_done) {
            int q0, head, chunk; item_coords(w, q0, head, chunk);
            if (n_done > 0) mbar_wait(q_empty, (n_done - 1) & 1);  // the last score tile of the previous item has been computed: its Q tile may go
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, kTile);
                tma_load_4d(sb + kOffQ, &tmQ, q_full, 0, q0, head, chunk);
            }
            for (int t = 0; t < NT; ++t, ++kt) {
                const int s = kt % kKS;
                mbar_wait(k_empty + 8 * s, ((kt / kKS) & 1) ^ 1);
                if (elect_one()) {
                    if ((DBG & 8) && kt >= kKS) mbar_arrive(k_full + 8 * s);                     // (timing experiment: no K traffic)
                    else {
                        mbar_arrive_expect_tx(k_full + 8 * s, kTile);
                        tma_load_4d(sb + kOffK + s * kTile, &tmK, k_full + 8 * s, 0, key_tile_of(t) * kKeys, head, chunk);
                    }
                }
            }
        }
    } else if (warp == kSmWarps + 2) {
        if (elect_one()) {
            if (ITAB) {
                mbar_arrive(tab_full);                               // (written before the block-wide barrier above)
            } else {
                mbar_arrive_expect_tx(tab_full, kExpTab * 2);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem_raw) + (uint32_t) kOffTab), "l"((uint64_t) (exp_lut + 0x8000)), "r"((uint32_t) (kExpTab * 2)), "r"(tab_full) : "memory");
            }
        }
        // ---- V^T producer: two 64-key atoms per tile ----
        int vj = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            int q0, head, chunk; item_coords(w, q0, head, chunk);
            for (int j = 0; j < nt; ++j, ++vj) {
                const int s = vj % kVS;
                mbar_wait(v_empty + 8 * s, ((vj / kVS) & 1) ^ 1);
                if (elect_one()) {
                    if ((DBG & 8) && vj >= kVS) mbar_arrive(v_full + 8 * s);
                    else {
                        mbar_arrive_expect_tx(v_full + 8 * s, kTile);
                        tma_load_4d(sb + kOffV + s * kTile,             &tmV, v_full + 8 * s, j * kKeys,      0, head, chunk);
                        tma_load_4d(sb + kOffV + s * kTile + kTile / 2, &tmV, v_full + 8 * s, j * kKeys + 64, 0, head, chunk);
                    }
                }
            }
        }
    } else if (warp == kSmWarps + 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc_s = (1u << 4) | ((uint32_t) (kKeys >> 3) << 17) | ((uint32_t) (kQRows >> 4) << 24);   // f16 x f16 -> f32, K-major, 128 x 128
        constexpr uint32_t idesc_o = (1u << 4) | ((uint32_t) (64 >> 3) << 17)    | ((uint32_t) (kQRows >> 4) << 24);   // 128 x 64
        const uint64_t qdesc = umma_desc_sw128(sb + kOffQ);
        int n_done = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n_done) {
            const int sbase = n_done * NT, pbase = n_done * nt;    // score / product tiles so far
            mbar_wait(q_full, n_done & 1);
            auto issue_s = [&](int t) {
                const int g = sbase + t, s = g % kKS, b = g % kNSB;
                mbar_wait(k_full + 8 * s, (g / kKS) & 1);
                mbar_wait(s_empty + 8 * b, ((g / kNSB) & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t kdesc = umma_desc_sw128(sb + kOffK + s * kTile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (!(DBG & 16)) umma_f16(tmem_base + b * kKeys, qdesc + (uint64_t) (2 * k), kdesc + (uint64_t) (2 * k), idesc_s, k != 0);
                    umma_commit(k_empty + 8 * s);
                    umma_commit(s_full + 8 * b);
                    if (t == NT - 1) umma_commit(q_empty);           // every score tile of this item has been issued: the Q tile is free when they retire
                }
            };
            for (int t = 0; t < kNSB - 1 && t < NT; ++t) issue_s(t);         // the score tiles run kNSB - 1 ahead of the product tiles
            for (int t = 0; t < NT; ++t) {
                if (t + kNSB - 1 < NT) issue_s(t + kNSB - 1);
                // product tiles whose probabilities follow score tile t
                const int j_lo = !reuse ? t - 2 * nt : t == 2 * nt - 2 ? 0 : t - (2 * nt - 1) + 2;
                const int j_hi = !reuse ? j_lo : t == 2 * nt - 2 ? 1 : j_lo;
                for (int j = j_lo; j <= j_hi && (reuse ? t >= 2 * nt - 2 : t >= 2 * nt); ++j) {
                    const int gj = pbase + j, b = gj & 1, s = gj % kVS;
                    mbar_wait(v_full + 8 * s, (gj / kVS) & 1);
                    mbar_wait(p_full + 8 * b, (gj >> 1) & 1);
                    if (j == 0 && n_done > 0) mbar_wait(o_empty, (n_done - 1) & 1);      // the previous item's output has left TMEM
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint64_t pdesc = umma_desc_sw128(sb + kOffP + b * 2 * kTile + (kk >> 2) * kTile) + (uint64_t) (2 * (kk & 3));
                            const uint64_t vdesc = umma_desc_sw128(sb + kOffV + s * kTile + (kk >> 2) * (kTile / 2)) + (uint64_t) (2 * (kk & 3));
                            if (!(DBG & 16)) umma_f16(tmem_o, pdesc, vdesc, idesc_o, (j | kk) != 0);
                        }
                        umma_commit(v_empty + 8 * s);
                        umma_commit(p_empty + 8 * b);
                        if (j == nt - 1) umma_commit(o_full);
                    }
                }
            }
        }
    } else if (warp < kSmWarps) {
        // ---- softmax warps: thread <-> query row, kCols key columns of every tile ----
        const int quad = warp & 3, part = warp >> 2;
        const int row = quad * 32 + lane;
        const int col0 = part * kCols;
        const uint32_t t_row = tmem_base + ((uint32_t) (quad * 32) << 16) + (uint32_t) col0;
        float * red_f = (float *) red;
        int n_done = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n_done) {
        int q0, head, chunk; item_coords(w, q0, head, chunk);
        const int sbase = n_done * NT, pbase = n_done * nt;          // score / product tiles so far
        float mx = -INFINITY, mxs = 0.0f, inv = 0.0f;
        unsigned long long tot = 0;

        // Score tile t of this item lives in score buffer (sbase + t) mod kNSB.  kPipe 0: copy, arithmetic, release.  kPipe 2: release as
        // soon as the copy has landed.  kPipe 1: that, and the copy of tile t + 1 is in flight during the arithmetic of tile t.  kPipe 3:
        // kPipe 0 with three score tiles per item replaced by values the threads still hold (see `reuse` above).
        auto load_scores = [&](int t, uint32_t (&r)[kCols]) {
            const int g = sbase + t;
            mbar_wait(s_full + 8 * (g % kNSB), (g / kNSB) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < kCols; c += 32) if (!(DBG & 2)) tmem_ld32(t_row + (uint32_t) ((g % kNSB) * kKeys + c), *(uint32_t (*)[32]) &r[c]);
        };
        auto release_scores = [&](int t) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + 8 * ((sbase + t) % kNSB));
        };
        auto n_valid_of = [&](int j) { return T - (j * kKeys + col0); };   // real keys in this thread's slice of key tile j (may be <= 0 or >= kCols)
        // ---- row maxima (ggml.c:11170-11172) ----
        auto max_tile = [&](int j, const uint32_t (&r)[kCols]) {
            const int n_valid = n_valid_of(j);
#pragma unroll
            for (int c = 0; c < kCols; c += 32) {
                if (DBG & 4) continue;
                if (n_valid - c >= 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[c + i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (c + i < n_valid) mx = fmaxf(mx, __uint_as_float(r[c + i]));
                }
            }
        };
        auto max_finish = [&]() {
            red_f[part * kQRows + row] = mx;
            asm volatile("bar.sync 1, %0;" :: "n"(kSmWarps * 32) : "memory");
            float m = red_f[row];
#pragma unroll
            for (int p2 = 1; p2 < kWPQ; ++p2) m = fmaxf(m, red_f[p2 * kQRows + row]);
            mxs = __fmul_rn(m, 0.125f);                 // KQ / sqrt(64): an exact scaling, applied after the product like whisper.cpp:1897
            asm volatile("bar.sync 1, %0;" :: "n"(kSmWarps * 32) : "memory");
            mbar_wait(tab_full, 0);                     // the exp table has landed (copied while the first pass ran)
        };
        // ---- exact sum of the table exponentials (ggml.c:11174-11192): every f16 value is a multiple of 2^-24 ----
        auto sum_tile = [&](int j, const uint32_t (&r)[kCols]) {
            const int n_valid = n_valid_of(j);
            unsigned int isum = 0;
#pragma unroll
            for (int c = 0; c < kCols; c += 32) {
                const uint32_t (&rs)[32] = *(const uint32_t (*)[32]) &r[c];
                if (DBG & 4) continue;
                if (n_valid - c >= 32) isum += sum_slice<true, ITAB, DBG>(rs, tab, mxs, 32);
                else                   isum += sum_slice<false, ITAB, DBG>(rs, tab, mxs, n_valid - c);
            }
            tot += isum;
        };
        // ... leaving the table values (f16 bits, zero for padding columns) in `keep`, two per word
        auto sum_tile_keep = [&](int j, const uint32_t (&r)[kCols], uint32_t (&keep)[kCols / 2]) {
            const int n_valid = n_valid_of(j);
            unsigned int isum = 0;
#pragma unroll
            for (int c = 0; c < kCols; c += 32) {
                const uint32_t (&rs)[32] = *(const uint32_t (*)[32]) &r[c];
                uint32_t (&ks)[16] = *(uint32_t (*)[16]) &keep[c / 2];
                if (n_valid - c >= 32) isum += sum_slice_keep<true, DBG>(rs, tab, mxs, 32, ks);
                else                   isum += sum_slice_keep<false, DBG>(rs, tab, mxs, n_valid - c, ks);
            }
            tot += isum;
        };
        auto sum_finish = [&]() {
            red[part * kQRows + row] = tot;
            asm volatile("bar.sync 1, %0;" :: "n"(kSmWarps * 32) : "memory");
            unsigned long long sm = 0;
#pragma unroll
            for (int p2 = 0; p2 < kWPQ; ++p2) sm += red[p2 * kQRows + row];
            inv = ExpTab<ITAB>::inv_for((float) (1.0 / ((double) sm * (1.0 / 16777216.0))));      // ggml.c:11196-11197
        };
        // ---- p = f16(e * inv) into the UMMA operand layout; the MMA warp adds P V on the tensor cores ----
        // KEPT: the table values come from registers (sum_tile_keep left them in r, two per word) instead of being looked up for the scores in r
        auto prob_tile = [&](int j, const auto & r, auto kept_tag) {
            constexpr bool KEPT = decltype(kept_tag)::value;
            const int n_valid = n_valid_of(j);
            const int gj = pbase + j, b = gj & 1;
            mbar_wait(p_empty + 8 * b, ((gj >> 1) & 1) ^ 1);
            const uint32_t pbase_addr = sb + kOffP + b * 2 * kTile + row * 128;
#pragma unroll
            for (int c = 0; c < kCols; c += 32) {
                uint32_t pk[16];
                if (DBG & 4) continue;
                if constexpr (KEPT) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t e2 = r[c / 2 + i];
                        const __half2 p2 = __floats2half2_rn(__fmul_rn(__half2float(__ushort_as_half((uint16_t) (e2 & 0xffffu))), inv),
                                                             __fmul_rn(__half2float(__ushort_as_half((uint16_t) (e2 >> 16))), inv));
                        pk[i] = *(const uint32_t *) &p2;
                    }
                } else {
                    const uint32_t (&rs)[32] = *(const uint32_t (*)[32]) &r[c];
                    if (n_valid - c >= 32) prob_slice<true, ITAB, DBG>(rs, tab, mxs, inv, 32, pk);
                    else                   prob_slice<false, ITAB, DBG>(rs, tab, mxs, inv, n_valid - c, pk);
                }
                // 16-byte chunk q of the row inside its 64-key atom sits at ((q ^ (row & 7)) << 4): the 128-byte swizzle of the UMMA descriptor
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk16 = (((col0 + c) & 63) >> 3) + q;
                    const uint32_t addr = pbase_addr + (uint32_t) (((col0 + c) >> 6) * kTile + ((chunk16 ^ (row & 7)) << 4));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + 8 * b);
        };
        // the arithmetic of score tile t (plain order: pass = t / nt, key tile = t % nt)
        int pass = 0, j = 0;
        auto compute = [&](const uint32_t (&r)[kCols]) {
            if (pass == 0)      { max_tile(j, r); if (j == nt - 1) max_finish(); }
            else if (pass == 1) { sum_tile(j, r); if (j == nt - 1) sum_finish(); }
            else                prob_tile(j, r, std::false_type{});
            if (++j == nt) { j = 0; ++pass; }
        };
        uint32_t ra[kCols], rb[kCols];
        if (DBG & 2) {
#pragma unroll
            for (int c = 0; c < kCols; ++c) { ra[c] = 0x3f800000u + c; rb[c] = 0x3f000000u + c; }
        }
        if (C::kPipe == 1) {
            auto step = [&](int t, uint32_t (&cur)[kCols], uint32_t (&nxt)[kCols]) {
                tmem_ld_wait();                                      // tile t is in `cur`
                release_scores(t);
                if (t + 1 < NT) load_scores(t + 1, nxt);
                compute(cur);
            };
            load_scores(0, ra);
            for (int t = 0; t < NT; t += 2) {
                step(t, ra, rb);
                if (t + 1 < NT) step(t + 1, rb, ra);
            }
        } else if (!reuse) {
            for (int t = 0; t < NT; ++t) {
                load_scores(t, ra);
                tmem_ld_wait();
                if (C::kPipe == 2) release_scores(t);
                compute(ra);
                if (C::kPipe != 2) release_scores(t);
            }
        } else {
            uint32_t ke0[kCols / 2], ke1[kCols / 2];                 // table values of key tiles 0 and 1, kept from the sum pass for the product pass
            for (int t = 0; t < NT; ++t) {
                load_scores(t, ra);
                tmem_ld_wait();
                const int kt = key_tile_of(t);
                if (t < nt) {
                    max_tile(kt, ra);
                    if (t == nt - 1) { max_finish(); sum_tile(kt, ra); }                // the sum pass starts on the scores still in registers
                } else if (t < 2 * nt - 1) {
                    if (kt == 0) sum_tile_keep(kt, ra, ke0); else if (kt == 1) sum_tile_keep(kt, ra, ke1); else sum_tile(kt, ra);
                } else {
                    prob_tile(kt, ra, std::false_type{});
                }
                release_scores(t);
                if (t == 2 * nt - 2) {                               // the sum pass is complete: the first two P tiles come from the kept values
                    sum_finish();
                    prob_tile(0, ke0, std::true_type{});
                    prob_tile(1, ke1, std::true_type{});
                }
            }
        }

        // ---- O: 64 / kWPQ output features per thread, merged heads layout [T][d] (whisper.cpp:1913-1917) ----
        mbar_wait(o_full, n_done & 1);
        tc_fence_after();
        constexpr int kOC = 64 / kWPQ;
        static_assert(kOC % 16 == 0, "the drain reads 16 columns at a time");
#pragma unroll
        for (int c = 0; c < kOC; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_o + ((uint32_t) (quad * 32) << 16) + (uint32_t) (part * kOC + c), r);
            tmem_ld_wait();
            if (q0 + row < T) {
                __half * op = out + ((int64_t) chunk * T + q0 + row) * d + head * 64 + part * kOC + c;
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __half2 h2 = __floats2half2_rn(__uint_as_float(r[i + 2 * u]), __uint_as_float(r[i + 2 * u + 1]));
                        w[u] = *(const uint32_t *) &h2;
                    }
                    *(uint4 *) (op + i) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);                     // the accumulator may take the next item's products
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == kSmWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

int attention_enc_table_entries() { return kExpTab; }

namespace {

template <int WPQ, int KS, int VS, bool ITAB, int DBG = 0, int NSB = 2, int PIPE = 1>
bool launch_variant(int slot, const CUtensorMap & tmQ, const CUtensorMap & tmK, const CUtensorMap & tmV, __half * out16, dim3 grid3, int T, int d,
                    const uint16_t * exp_lut, cudaStream_t st, bool persistent = false) {
    using C = AttnCfg<WPQ, KS, VS, ITAB, NSB, PIPE>;
    static bool attr_set[16][kAttnVariants] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_set[dev][slot]) {
        if (cudaFuncSetAttribute(k_attn_enc<C, ITAB, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) {
            fprintf(stderr, "whisper_b200: cannot reserve %d bytes of shared memory for the fused attention kernel\n", C::kSmemBytes);
            return false;
        }
        attr_set[dev][slot] = true;
    }
    const int n_items = (int) (grid3.x * grid3.y * grid3.z);
    static int sms[16] = {};
    if (dev < 16 && sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    const int grid = persistent && dev < 16 && sms[dev] > 0 ? std::min(n_items, sms[dev]) : n_items;
    k_attn_enc<C, ITAB, DBG><<<grid, C::kThreads, C::kSmemBytes, st>>>(tmQ, tmK, tmV, out16, T, d, exp_lut, (int) grid3.x, (int) grid3.y, n_items);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace

// variant: 0 = 8 softmax warps, f16 table (the round-1 configuration)   1 = 16 softmax warps, f16 table (4 % faster)
//          2 = 8 softmax warps, integer table, two-deep K / V^T rings    3 = 16 softmax warps, integer table, two-deep rings
//          4 = 16 softmax warps, three score buffers, score buffers released before the arithmetic, next tile prefetched into registers
//          5 = variant 1 as a persistent CTA per SM walking over the work items (the default: another 6 %)   6 = variant 0, persistent
//          < 0 = the default (WHISPER_B200_ATTN_VARIANT overrides it).  All variants produce the same bits; measured on a B200 they are
//          within 6 % of each other because the kernel is bound by the shared-memory / TMEM load pipe (3.5 bank-conflict wavefronts per
//          table look-up, two look-ups per score), not by issue slots, the tensor pipe or K / V traffic (profiles/r02_attn_enc_experiments.md).
bool launch_attention_enc(const __half * q16, const __half * k16, const __half * vt16, __half * out16, int B, int T, int Tp, int d,
                          int n_head, const uint16_t * exp_lut, cudaStream_t st, int variant) {
    if (variant < 0) {
        static const int env_variant = [] { const char * e = getenv("WHISPER_B200_ATTN_VARIANT"); return e ? atoi(e) : kAttnDefaultVariant; }();
        variant = env_variant;
    }
    if (d != n_head * 64 || T <= 0) return false;
    Operand Q; Q.p = q16;  Q.ld = d;  Q.bs1 = 64;               Q.bs2 = (int64_t) T * d;  Q.rows = T;
    Operand K; K.p = k16;  K.ld = d;  K.bs1 = 64;               K.bs2 = (int64_t) T * d;  K.rows = T;
    Operand V; V.p = vt16; V.ld = Tp; V.bs1 = (int64_t) 64 * Tp; V.bs2 = (int64_t) d * Tp; V.rows = 64;
    alignas(64) CUtensorMap tmQ, tmK, tmV;
    if (!gemm_tc_make_map(Q, 64, n_head, B, kQRows, &tmQ) || !gemm_tc_make_map(K, 64, n_head, B, kKeys, &tmK) ||
        !gemm_tc_make_map(V, T, n_head, B, 64, &tmV)) return false;
    const dim3 grid((T + kQRows - 1) / kQRows, n_head, B);
    switch (variant) {
        case 0:  return launch_variant<2, 3, 3, false, 0, 2, 0>(0, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 1:  return launch_variant<4, 3, 3, false, 0, 2, 0>(1, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 2:  return launch_variant<2, 2, 2, true, 0, 2, 0>(2, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 3:  return launch_variant<4, 2, 2, true, 0, 2, 0>(3, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 4:  return launch_variant<4, 4, 2, false, 0, 3, 1>(4, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 5:  return launch_variant<4, 3, 3, false, 0, 2, 0>(1, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st, true);    // variant 1, a persistent CTA per SM
        case 6:  return launch_variant<2, 3, 3, false, 0, 2, 0>(0, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st, true);    // variant 0, persistent
        case 7:  return launch_variant<4, 3, 3, false, 0, 2, 3>(7, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st, true);    // variant 5 + three score tiles per item reused from registers
        case 8:  return launch_variant<2, 3, 3, false, 0, 2, 3>(8, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st, true);    // variant 6 + reuse
#ifdef WB200_ATTN_EXPERIMENTS
        // timing experiments (wrong results by construction; profiles/r02_attn_enc_experiments.md): variant 4 without table look-ups (DBG 1),
        // without TMEM reads (2), without any softmax arithmetic (4), without K / V traffic (8), without MMAs (16)
        case 15: return launch_variant<4, 4, 2, false, 1, 3, 1>(5, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 16: return launch_variant<4, 4, 2, false, 2, 3, 1>(6, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 17: return launch_variant<4, 4, 2, false, 3, 3, 1>(7, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 18: return launch_variant<4, 4, 2, false, 7, 3, 1>(8, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 19: return launch_variant<4, 4, 2, false, 15, 3, 1>(9, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 20: return launch_variant<4, 4, 2, false, 23, 3, 1>(10, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
        case 21: return launch_variant<4, 4, 2, false, 31, 3, 1>(11, tmQ, tmK, tmV, out16, grid, T, d, exp_lut, st);
#endif
        default: fprintf(stderr, "whisper_b200: no attention kernel variant %d\n", variant); return false;
    }
}

}  // namespace wb200
