// The encoder's weight contractions for sm_100a, second generation:  D[n][m] = sum_k A[n][k] * W[m][k]  (f16 x f16 -> f32) with an
// epilogue that never touches HBM from a per-thread row walk.
//
// Replaces ggml_compute_forward_mul_mat (/root/reference/thirdparty/whisper.cpp/ggml.c:9737-9948) plus the add / scale / gelu / cpy
// nodes behind it for the Q K V projections, the attention output projection, both FFN maps and the cross-attention K / V projections
// of whisper_build_graph_encoder / _cross (whisper.cpp:1831-1970, 2038-2066).  Same arithmetic as gemm_tc.cu's epilogue (dev.cuh).
//
//   persistent CTA per SM, ten warps:  0 = TMA producer (operand ring, residual tiles)   1 = MMA issuer (tcgen05.mma 128 x 128 x 16,
//   two TMEM accumulators)   2..9 = epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4)
//   epilogue:  tcgen05.ld 32 columns at a time (the next load in flight while this one is worked on)  ->  bias / scale / GELU / residual
//              in registers  ->  the output tile assembled in SHARED MEMORY in the 128-byte-swizzled box layout  ->  one elected thread
//              writes it with TMA stores (cp.async.bulk.tensor ... global.shared::cta; tile tails are clipped by the tensor map).
//   HALF  mode: f16 outputs [n][m] or transposed [m][n] (V^T), up to three feature segments with their own bias / scale / GELU / layout;
//               the GELU table look-ups (ggml.c:1416-1423: one per FC1 output) hit a copy of the table's live range in shared memory.
//   RES32 mode: f32 output = acc + bias + residual.  Long contractions (K > 512: FC2): every epilogue thread reads its row segment of the
//               residual straight from global memory (32 contiguous floats, requested before it waits for the accumulator), so that shared
//               memory holds a five-stage operand ring — the main loop is what bounds those.  Short ones (out-proj, RES32T): the residual tile
//               arrives by TMA, one tile ahead, into the buffer the result is stored from (two such tiles, three ring stages) — the epilogue
//               is what bounds those, and a residual that is already on chip when the accumulator completes keeps it short.
#include "dev.cuh"
#include "tc.cuh"
#include "gemm_enc.cuh"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace wb200 {

namespace {

using namespace tc;

constexpr int kTile   = 128;                  // rows and features per tile
constexpr int kBK     = 64;                   // one 128-byte swizzle atom of f16
constexpr int kMaxStages = 8;
constexpr int kOpBytes    = kTile * kBK * 2;  // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kOpBytes;
constexpr int kHalfTile   = kTile * kTile * 2;                // 32 KB: f16 output tile
constexpr int kF32Tile    = kTile * kTile * 4;                // 64 KB: f32 residual / output tile
constexpr int kGeluLo     = 0x0800, kGeluMag = 0x4800;        // table copy covers 2^-13 <= |x| < 8 (f16 patterns 0x0800 .. 0x47ff of either sign)
constexpr int kGeluSpan   = kGeluMag - kGeluLo;               // 16 384 entries per sign
constexpr int kGeluBytes  = 2 * kGeluSpan * 2;                // 64 KB
constexpr int kNumBars    = 2 * kMaxStages + 4 + 4 + 1;

// Three shared-memory plans.  The operand ring is as deep as the rest allows: the kernel's main loop is bound by the round trip of a ring
// stage (MMAs retire -> commit arrives -> producer wakes -> TMA fetch from L2 lands -> MMA warp wakes: ~3 300 cycles, measured by switching
// the traffic / the MMAs / the epilogue off, profiles/r02_gemm_enc_skeleton_experiment.md) against 256 cycles of tensor-pipe work per
// stage, so the time per k-block is that round trip divided by the number of stages in flight.
enum { MODE_HALF = 0, MODE_GELU = 1, MODE_RES32 = 2, MODE_RES32T = 3 };
template <int MODE> struct EncCfg {
    static constexpr int kStages = MODE == MODE_HALF || MODE == MODE_RES32 ? 5 : 3;
    static constexpr int kRingBytes = kStages * kStageBytes;
    // two output tiles wherever two epilogue groups alternate (all modes but the direct-residual one)
    static constexpr int kTailBytes = MODE == MODE_HALF ? 2 * kHalfTile : MODE == MODE_GELU ? 2 * kHalfTile + kGeluBytes : MODE == MODE_RES32 ? kF32Tile : 2 * kF32Tile;
    static constexpr int kSmem = kRingBytes + kTailBytes + kNumBars * 8 + 16 + 1024;
    static_assert(kSmem <= 227 * 1024, "shared-memory plan does not fit");
    static_assert(kStages <= kMaxStages, "ring barriers");
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap * map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"((uint64_t) map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap * map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"((uint64_t) map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int THREADS> __device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(THREADS) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_b16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.b16 [%0], %1;" :: "r"(addr), "h"(v) : "memory");
}

// GELU through the f16 table (ggml.c:1416-1423) for 32 values in place: the live range of the table (2^-13 <= |x| < 8, both signs) sits in shared
// memory and every look-up goes there with a clamped index, without a branch; the rare slice that holds a larger or tinier |x| is patched from the
// full table in HBM afterwards (one warp vote per slice).
__device__ __forceinline__ void gelu_slice(float (&v)[32], uint32_t tab_smem, const uint16_t * __restrict__ lut) {      // tab_smem: shared-space address
    uint32_t hb[16];
    uint32_t oob = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const __half2 h2 = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        hb[q] = *(const uint32_t *) &h2;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const uint32_t u = hb[q];
        const uint32_t mag = u & 0x7fff7fffu;
        const uint32_t cl = __vminu2(__vmaxu2(mag, (uint32_t) kGeluLo * 0x10001u), (uint32_t) (kGeluMag - 1) * 0x10001u);
        oob |= cl ^ mag;
        const uint32_t i0 = (cl & 0xffffu) - (uint32_t) kGeluLo + ((u >> 15) & 1u) * (uint32_t) kGeluSpan;
        const uint32_t i1 = (cl >> 16) - (uint32_t) kGeluLo + (u >> 31) * (uint32_t) kGeluSpan;
        uint16_t g0, g1;                                  // (explicit shared-space loads: through a generic pointer these were LD.E with 64-bit address arithmetic)
        asm("ld.shared.u16 %0, [%1];" : "=h"(g0) : "r"(tab_smem + 2 * i0));
        asm("ld.shared.u16 %0, [%1];" : "=h"(g1) : "r"(tab_smem + 2 * i1));
        v[2 * q]     = __half2float(__ushort_as_half(g0));
        v[2 * q + 1] = __half2float(__ushort_as_half(g1));
    }
    if (__any_sync(0xffffffffu, oob != 0)) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint32_t u = hb[q];
            if ((u & 0x7fffu) - (uint32_t) kGeluLo >= (uint32_t) kGeluSpan) v[2 * q] = __half2float(__ushort_as_half(__ldg(lut + (u & 0xffffu))));
            if (((u >> 16) & 0x7fffu) - (uint32_t) kGeluLo >= (uint32_t) kGeluSpan) v[2 * q + 1] = __half2float(__ushort_as_half(__ldg(lut + (u >> 16))));
        }
    }
}

template <int MODE, int EW, int GROUPS>
__global__ void __launch_bounds__((2 + EW) * 32, 1)
k_gemm_enc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,
           const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmRes,
           const EncGemmArgs a) {
    constexpr bool RES32T = MODE == MODE_RES32T;             // residual by TMA into one of two output tiles
    constexpr bool RES32 = MODE == MODE_RES32 || RES32T;
    // Epilogue warps work in kGroups independent groups on alternating tiles (group g <-> TMEM accumulator g <-> output tile g): while one
    // group sits in its barrier / fence / store phase the other is in its arithmetic, so the per-tile latency chain of the epilogue
    // is paid once per two tiles.  (The direct-residual mode keeps one group: it needs the shared memory for a deep ring.)
    constexpr int kGroups = (EW == 16 && MODE != MODE_RES32) ? GROUPS : 1;
    constexpr int kEpiWarps = EW, kGroupWarps = EW / kGroups, kColsW = kTile * 4 / kGroupWarps;     // features of a tile per epilogue warp (64 or 32)
    static_assert(kColsW == 32 || kColsW == 64, "four or two epilogue warps per TMEM lane quadrant");
    constexpr int kStages = EncCfg<MODE>::kStages, kRingBytes = EncCfg<MODE>::kRingBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    const uint32_t ring = smem_u32(smem);
    const uint32_t stage0 = ring + kRingBytes;                               // HALF: output tile; RES32: two residual / output tiles
    const uint32_t gelu_s = ring + kRingBytes + 2 * kHalfTile;              // shared-space address of the GELU table copy (behind the two output tiles)
    uint64_t * bars = (uint64_t *) (smem + kRingBytes + EncCfg<MODE>::kTailBytes);
    uint32_t * tmem_slot = (uint32_t *) (bars + kNumBars);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
    const uint32_t acc_full = empty0 + 8 * kMaxStages, acc_empty = acc_full + 16;
    const uint32_t res_full = acc_empty + 16, res_free = res_full + 16, gelu_full = res_free + 16;

    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // (provably warp-uniform: the role branches stay converged)
    const int num_k = (a.K + kBK - 1) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, kGroups == 2 ? kGroupWarps : kEpiWarps);
            mbar_init(res_full + 8 * i, 1); mbar_init(res_free + 8 * i, 1);
        }
        mbar_init(gelu_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t: n tile fastest (CTAs that run side by side share the weight tile), then m tile, then batch entry
    auto tile_coords = [&](int t, int & n0, int & m0, int & bz) {
        const int tn = t % a.tiles_n, r = t / a.tiles_n;
        const int tm = r % a.tiles_m;
        bz = r / a.tiles_m;
        n0 = tn * kTile; m0 = tm * kTile;
    };

    if (warp == 0) {
        // ---- TMA producer (the whole warp walks the loop and waits; one elected lane issues): the operand ring runs on across tiles ----
        if (MODE == MODE_GELU && elect_one()) {
            // the live range of the GELU table: f16 patterns kGeluLo .. kGeluMag - 1 of both signs, two bulk copies that land under the first tile
            const uint32_t dst = ring + kRingBytes + 2 * kHalfTile;
            mbar_arrive_expect_tx(gelu_full, kGeluBytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"((uint64_t) (a.gelu_lut + kGeluLo)), "r"((uint32_t) (kGeluSpan * 2)), "r"(gelu_full) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst + kGeluSpan * 2), "l"((uint64_t) (a.gelu_lut + 0x8000 + kGeluLo)), "r"((uint32_t) (kGeluSpan * 2)), "r"(gelu_full) : "memory");
        }
        int it = 0, i = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++i) {
            int n0, m0, bz; tile_coords(t, n0, m0, bz);
            if (RES32T) {
                const int rb = i & 1;
                mbar_wait(res_free + 8 * rb, ((i >> 1) & 1) ^ 1);          // the store that last read this buffer is done with it
                if (elect_one()) {
                    mbar_arrive_expect_tx(res_full + 8 * rb, kF32Tile);
                    const int rz = a.res_batched ? bz : 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) tma_load_3d(stage0 + rb * kF32Tile + b * (kTile * 128), &tmRes, res_full + 8 * rb, m0 + 32 * b, n0, rz);
                }
            }
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(empty0 + 8 * s, ((it / kStages) & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t a_dst = ring + s * kStageBytes;
                    const bool skip_w = (a.dbg & 1) && i > 0, skip_a = (a.dbg & 2) && i > 0;       // (timing experiments: stale operands)
                    const uint32_t bytes = (skip_a ? 0 : kOpBytes) + (skip_w ? 0 : kOpBytes);
                    if (bytes) mbar_arrive_expect_tx(full0 + 8 * s, bytes); else mbar_arrive(full0 + 8 * s);
                    if (!skip_a) tma_load_3d(a_dst, &tmA, full0 + 8 * s, kb * kBK, n0, bz);
                    if (!skip_w) tma_load_3d(a_dst + kOpBytes, &tmW, full0 + 8 * s, kb * kBK, m0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (same idiom) ----
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t) (kTile >> 3) << 17) | ((uint32_t) (kTile >> 4) << 24);   // f16 x f16 -> f32, K-major, 128 x 128
        int it = 0, i = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++i) {
            const int ab = i & 1;
            mbar_wait(acc_empty + 8 * ab, ((i >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t) (ab * kTile);
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(full0 + 8 * s, (it / kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = ring + s * kStageBytes;
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + kOpBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) if (!(a.dbg & 8)) umma_f16(acc, adesc + (uint64_t) (2 * k), bdesc + (uint64_t) (2 * k), idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * s);
                    if (kb == num_k - 1) umma_commit(acc_full + 8 * ab);
                }
            }
        }
    } else if (warp >= 2) {
        // ---- epilogue: thread <-> token row (TMEM lane), kColsW of the tile's 128 features per warp ----
        const int grp = kGroups == 2 ? (warp - 2) / kGroupWarps : 0;         // which group: tiles grp, grp + kGroups, ... of this CTA
        const int wg = (warp - 2) - grp * kGroupWarps;                       // warp within its group
        const int quad = warp & 3, part = wg >> 2;
        const int row = quad * 32 + lane;
        const bool lead_warp = wg == 0;                                      // its elected lane issues the group's TMA stores
        const int bar_id = 1 + grp;
        if (MODE == MODE_GELU) mbar_wait(gelu_full, 0);
        const int my_tiles = (a.n_tiles - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;
        for (int i = grp; i < my_tiles; i += kGroups) {
            const int t = (int) blockIdx.x + i * (int) gridDim.x;
            const int ab = i & 1;
            int n0, m0, bz; tile_coords(t, n0, m0, bz);
            const int seg_i = a.nseg > 1 ? m0 / a.seg_m : 0;
            const EncSeg sg = a.seg[seg_i];
            const int m_seg0 = m0 - seg_i * a.seg_m;
            const uint32_t t_row = tmem_base + ((uint32_t) (quad * 32) << 16) + (uint32_t) (ab * kTile + part * kColsW);
            float4 resv[RES32 && !RES32T ? kColsW / 4 : 1];
            if (RES32 && !RES32T) {
                // this thread's residual values: requested now, needed after the accumulator has arrived
                const bool live = n0 + row < a.N && n0 + row < a.res_rows;
                const float4 * rp = (const float4 *) (a.res + (a.res_batched ? (int64_t) bz * a.res_bs : 0) + (int64_t) (n0 + row) * a.res_ld + m0 + part * kColsW);
#pragma unroll
                for (int q = 0; q < kColsW / 4; ++q) resv[q] = live ? rp[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            mbar_wait(acc_full + 8 * ab, (i >> 1) & 1);
            tc_fence_after();
            if (a.dbg & 4) {                                                     // (timing experiment: no epilogue work at all)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * ab);
                if (RES32T && lead_warp && elect_one()) { mbar_wait(res_full + 8 * (i & 1), (i >> 1) & 1); mbar_arrive(res_free + 8 * (i & 1)); }
                continue;
            }
            uint32_t ra[32], rb[32];
            tmem_ld32(t_row, ra);
            if (kColsW == 64) tmem_ld32(t_row + 32, rb);
            if (!RES32T) {
                // the TMA store of the previous tile must be done reading the staging tile before anybody overwrites it
                if (lead_warp && elect_one()) bulk_wait_read0();               // (this group's previous store: bulk groups are per thread)
                epi_bar<kGroupWarps * 32>(bar_id);
            } else {
                mbar_wait(res_full + 8 * (i & 1), (i >> 1) & 1);
            }
            const uint32_t stage = stage0 + (RES32T ? (uint32_t) ((i & 1) * kF32Tile) : (uint32_t) (grp * kHalfTile));
#pragma unroll
            for (int j = 0; j < kColsW / 32; ++j) {
                if (j == 0) tmem_ld_wait();                                    // (the loads were issued back to back: one wait covers them)
                const uint32_t (&r)[32] = j == 0 ? ra : rb;
                const int c0 = part * kColsW + 32 * j;                         // first feature of this slice inside the tile
                if (RES32) {
                    // v = (acc + bias) + residual into the output tile: four boxes of 32 f32 columns, 16-byte pieces swizzled by the row
                    const uint32_t rbase = stage + (uint32_t) ((c0 >> 5) * (kTile * 128) + row * 128);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t addr = rbase + (uint32_t) ((q ^ (row & 7)) << 4);
                        float4 rs;
                        if (RES32T) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(rs.x), "=f"(rs.y), "=f"(rs.z), "=f"(rs.w) : "r"(addr) : "memory");
                        else rs = resv[RES32 && !RES32T ? 8 * j + q : 0];
                        float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (sg.bias) b = __ldg((const float4 *) (sg.bias + m_seg0 + c0 + 4 * q));
                        float4 v;
                        v.x = __fadd_rn(__fadd_rn(__uint_as_float(r[4 * q + 0]), b.x), rs.x);
                        v.y = __fadd_rn(__fadd_rn(__uint_as_float(r[4 * q + 1]), b.y), rs.y);
                        v.z = __fadd_rn(__fadd_rn(__uint_as_float(r[4 * q + 2]), b.z), rs.z);
                        v.w = __fadd_rn(__fadd_rn(__uint_as_float(r[4 * q + 3]), b.w), rs.w);
                        st_shared_f4(addr, v);
                    }
                } else {
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
                    if (sg.bias) {
#pragma unroll
                        for (int q = 0; q < 32; q += 4) {
                            const float4 b = __ldg((const float4 *) (sg.bias + m_seg0 + c0 + q));
                            v[q] = __fadd_rn(v[q], b.x); v[q + 1] = __fadd_rn(v[q + 1], b.y); v[q + 2] = __fadd_rn(v[q + 2], b.z); v[q + 3] = __fadd_rn(v[q + 3], b.w);
                        }
                    }
                    if (sg.scale != 1.0f) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = __fmul_rn(v[q], sg.scale);
                    }
                    if (MODE == MODE_GELU && sg.gelu) gelu_slice(v, gelu_s, a.gelu_lut);
                    if (!sg.transposed) {
                        // two boxes of 64 f16 columns: 16-byte piece p of row r sits at (p ^ (r & 7)) << 4
                        const uint32_t rbase = stage + (uint32_t) ((c0 >> 6) * (kTile * 128) + row * 128);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t pk[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const __half2 h = __floats2half2_rn(v[8 * q + 2 * u], v[8 * q + 2 * u + 1]);
                                pk[u] = *(const uint32_t *) &h;
                            }
                            const int piece = ((c0 & 63) >> 3) + q;
                            st_shared_v4(rbase + (uint32_t) ((piece ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                    } else {
                        // transposed tile [feature][token]: two boxes of 64 token columns; this thread owns token column `row`
                        const uint32_t cbase = stage + (uint32_t) ((row >> 6) * (kTile * 128) + (row & 7) * 2);
                        const int piece = (row & 63) >> 3;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const int m = c0 + q;
                            st_shared_b16(cbase + (uint32_t) (m * 128 + ((piece ^ (m & 7)) << 4)), __half_as_ushort(__float2half_rn(v[q])));
                        }
                    }
                }
            }
            // the accumulator is in registers / shared memory now: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * ab);
            fence_async_smem();
            epi_bar<kGroupWarps * 32>(bar_id);
            if (lead_warp && elect_one()) {
                const int bo = sg.bmap ? __ldg(sg.bmap + bz) : bz;
                if (RES32) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) tma_store_3d(&tmO0, stage + b * (kTile * 128), m0 + 32 * b, n0, bo);
                    bulk_commit();
                    if (RES32T) {
                        bulk_wait_read0();                                     // short: the engine only has to READ 64 KB of shared memory
                        mbar_arrive(res_free + 8 * (i & 1));
                    }
                } else {
                    const CUtensorMap * om = seg_i == 0 ? &tmO0 : seg_i == 1 ? &tmO1 : &tmO2;
                    if (!sg.transposed) {
                        tma_store_3d(om, stage, m_seg0, n0, bo);
                        tma_store_3d(om, stage + kTile * 128, m_seg0 + 64, n0, bo);
                    } else {
                        tma_store_3d(om, stage, n0, m_seg0, bo);
                        tma_store_3d(om, stage + kTile * 128, n0 + 64, m_seg0, bo);
                    }
                    bulk_commit();
                }
            }
        }
        if (lead_warp && elect_one()) bulk_wait_all();                                           // global writes complete before the kernel ends
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- tensor maps -------------------------------------------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled) p;
    });
    return fn;
}

struct Map3Key {
    const void * p; int f32; int64_t d0, d1, d2, s1, s2; int b0, b1;
    bool operator<(const Map3Key & o) const {
        return std::tie(p, f32, d0, d1, d2, s1, s2, b0, b1) < std::tie(o.p, o.f32, o.d0, o.d1, o.d2, o.s1, o.s2, o.b0, o.b1);
    }
};
std::mutex g_mu;
std::map<Map3Key, CUtensorMap> g_maps3;

// 3-D tensor map, 128-byte swizzle: element (c0, c1, c2) at base + c0 * esize + c1 * s1 + c2 * s2 (strides in BYTES); box = b0 x b1 x 1
bool make_map3(const void * base, bool f32, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2, int b0, int b1, CUtensorMap & out) {
    const Map3Key key{base, f32 ? 1 : 0, d0, d1, d2, s1, s2, b0, b1};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_maps3.find(key);
        if (it != g_maps3.end()) { out = it->second; return true; }
    }
    PFN_encodeTiled enc = encode_fn();
    if (!enc) { fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled unavailable\n"); return false; }
    if (d2 <= 1) { d2 = 1; s2 = s1 * d1; }
    if (((uintptr_t) base & 15) || (s1 & 15) || (s2 & 15) || s1 <= 0 || s2 <= 0) {
        fprintf(stderr, "whisper_b200: tensor not 16-byte aligned for TMA (p=%p s1=%lld s2=%lld)\n", base, (long long) s1, (long long) s2);
        return false;
    }
    cuuint64_t dims[3] = {(cuuint64_t) d0, (cuuint64_t) d1, (cuuint64_t) d2};
    cuuint64_t strides[2] = {(cuuint64_t) s1, (cuuint64_t) s2};
    cuuint32_t box[3] = {(cuuint32_t) b0, (cuuint32_t) b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = enc(&out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void *) base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled (3-D) failed (%d) dims %lld x %lld x %lld strides %lld %lld\n", (int) rc, (long long) d0,
                (long long) d1, (long long) d2, (long long) s1, (long long) s2);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps3[key] = out;
    return true;
}

int g_sms = 0;

}  // namespace

void gemm_enc_forget_maps() {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps3.clear();
}

bool gemm_enc_usable(const EncGemm & g) {
    const int seg = g.nseg > 1 ? g.seg_m : g.M;
    if (g.nseg < 1 || g.nseg > 3 || seg % kTile != 0 || g.M != seg * g.nseg || g.K % 8 != 0 || g.N <= 0 || g.nb < 1) return false;
    if (const char * e = getenv("WHISPER_B200_GEMM_V2")) { if (atoi(e) == 0) return false; }
    return true;
}

bool launch_gemm_enc(const EncGemm & g, cudaStream_t st) {
    if (!gemm_enc_usable(g)) return false;
    int dev = 0;
    cudaGetDevice(&dev);
    const bool res32 = g.res32;
    bool any_gelu = false;
    for (int i = 0; i < g.nseg; ++i) any_gelu |= g.out[i].gelu != 0;
    const int mode = res32 ? (g.K > 512 ? MODE_RES32 : MODE_RES32T) : any_gelu ? MODE_GELU : MODE_HALF;
    // epilogue warps: sixteen (four per TMEM lane quadrant) hide the look-up / TMEM latencies of the epilogue better than eight
    static const int ew_env = [] { const char * e = getenv("WHISPER_B200_GEMM_EPI_WARPS"); return e ? atoi(e) : 16; }();
    const int ew = ew_env == 8 ? 8 : 16;
    static const int groups_env = [] { const char * e = getenv("WHISPER_B200_GEMM_GROUPS"); return e ? atoi(e) : 2; }();
    if (g_sms == 0) cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    alignas(64) CUtensorMap tmA, tmW, tmO[3], tmRes;
    // A: [nb][N rows][K] (rows may overlap: ld < K is fine for TMA), W: [M][K]
    if (!make_map3(g.A, false, g.K, g.a_rows, g.nb, g.a_ld * 2, g.a_bs * 2, kBK, kTile, tmA)) return false;
    if (!make_map3(g.W, false, g.K, g.M, 1, g.w_ld * 2, g.w_ld * 2 * (int64_t) g.M, kBK, kTile, tmW)) return false;
    tmRes = tmA;
    for (int i = 0; i < 3; ++i) tmO[i] = tmA;
    EncGemmArgs a;
    a.N = g.N; a.M = g.M; a.K = g.K; a.nseg = g.nseg; a.seg_m = g.nseg > 1 ? g.seg_m : g.M;
    a.tiles_n = (g.N + kTile - 1) / kTile; a.tiles_m = g.M / kTile; a.n_tiles = a.tiles_n * a.tiles_m * g.nb;
    a.gelu_lut = g.gelu_lut; a.any_gelu = 0; a.res_batched = g.res_bs != 0;
    a.res = g.res; a.res_ld = g.res_ld; a.res_bs = g.res_bs; a.res_rows = g.res_rows;
    if (const char * e = getenv("WHISPER_B200_GEMM_DBG")) a.dbg = atoi(e);
    const int seg = a.seg_m;
    for (int i = 0; i < g.nseg; ++i) {
        const EncOut & o = g.out[i];
        a.seg[i].bias = o.bias; a.seg[i].scale = o.scale; a.seg[i].gelu = o.gelu; a.seg[i].transposed = o.transposed; a.seg[i].bmap = o.bmap;
        a.any_gelu |= o.gelu;
        if (o.gelu && !g.gelu_lut) return false;
        const int64_t nbo = o.n_batch_out > 0 ? o.n_batch_out : g.nb;
        if (res32) {
            if (!make_map3(o.p, true, seg, g.N, nbo, o.ld * 4, o.bs * 4, 32, kTile, tmO[i])) return false;
        } else if (!o.transposed) {
            if (!make_map3(o.p, false, seg, g.N, nbo, o.ld * 2, o.bs * 2, 64, kTile, tmO[i])) return false;
        } else {
            if (!make_map3(o.p, false, g.N, seg, nbo, o.ld * 2, o.bs * 2, 64, kTile, tmO[i])) return false;
        }
    }
    if (res32) {
        if (g.nseg != 1 || !g.res) return false;
        if (((uintptr_t) g.res & 15) || (g.res_ld & 3) || (g.res_bs & 3)) return false;       // the epilogue reads the residual with 16-byte loads
        if (mode == MODE_RES32T &&
            !make_map3(g.res, true, seg, g.res_rows, g.res_bs != 0 ? g.nb : 1, g.res_ld * 4, (g.res_bs != 0 ? g.res_bs : g.res_ld * (int64_t) g.res_rows) * 4, 32, kTile, tmRes)) return false;
    }
    const int grid = std::min(a.n_tiles, g_sms);
    auto go = [&](auto kernel, int smem, int slot, int threads) -> bool {
        static bool attr_set[16][8] = {};
        if (dev < 16 && !attr_set[dev][slot]) {
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) { fprintf(stderr, "whisper_b200: cannot reserve shared memory for k_gemm_enc: %s\n", cudaGetErrorString(e)); return false; }
            attr_set[dev][slot] = true;
        }
        kernel<<<grid, threads, smem, st>>>(tmA, tmW, tmO[0], tmO[1], tmO[2], tmRes, a);
        return cudaGetLastError() == cudaSuccess;
    };
    if (ew == 16 && groups_env != 1) {
        if (mode == MODE_RES32) return go(k_gemm_enc<MODE_RES32, 16, 2>, EncCfg<MODE_RES32>::kSmem, 0, 18 * 32);
        if (mode == MODE_RES32T) return go(k_gemm_enc<MODE_RES32T, 16, 2>, EncCfg<MODE_RES32T>::kSmem, 6, 18 * 32);
        if (mode == MODE_GELU)  return go(k_gemm_enc<MODE_GELU, 16, 2>, EncCfg<MODE_GELU>::kSmem, 1, 18 * 32);
        return go(k_gemm_enc<MODE_HALF, 16, 2>, EncCfg<MODE_HALF>::kSmem, 2, 18 * 32);
    }
    if (ew == 16) {
        if (mode == MODE_RES32) return go(k_gemm_enc<MODE_RES32, 16, 1>, EncCfg<MODE_RES32>::kSmem, 0, 18 * 32);
        if (mode == MODE_RES32T) return go(k_gemm_enc<MODE_RES32T, 16, 1>, EncCfg<MODE_RES32T>::kSmem, 6, 18 * 32);
        if (mode == MODE_GELU)  return go(k_gemm_enc<MODE_GELU, 16, 1>, EncCfg<MODE_GELU>::kSmem, 1, 18 * 32);
        return go(k_gemm_enc<MODE_HALF, 16, 1>, EncCfg<MODE_HALF>::kSmem, 2, 18 * 32);
    }
    if (mode == MODE_RES32) return go(k_gemm_enc<MODE_RES32, 8, 1>, EncCfg<MODE_RES32>::kSmem, 3, 10 * 32);
    if (mode == MODE_RES32T) return go(k_gemm_enc<MODE_RES32T, 8, 1>, EncCfg<MODE_RES32T>::kSmem, 7, 10 * 32);
    if (mode == MODE_GELU)  return go(k_gemm_enc<MODE_GELU, 8, 1>, EncCfg<MODE_GELU>::kSmem, 4, 10 * 32);
    return go(k_gemm_enc<MODE_HALF, 8, 1>, EncCfg<MODE_HALF>::kSmem, 5, 10 * 32);
}

}  // namespace wb200
