// Tensor-core contraction for sm_100a:  D[n][m] = sum_k A[n][k] * W[m][k],  f16 x f16 -> f32.
//
//   TMA (cp.async.bulk.tensor.4d, 128B swizzle)  ->  shared-memory ring (kStages x {A 128x64, W BNx64})
//   tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16, issued by one thread)  ->  f32 accumulators in TMEM
//   tcgen05.ld (32 lanes x 32 bit x 16 columns per warp)  ->  fused epilogue (dev.cuh: bias / scale / GELU table /
//   residual / f16 and transposed-f16 stores)  ->  HBM.
//
// Replaces, for every weight and attention mat-mul of the path, ggml_compute_forward_mul_mat
// (/root/reference/thirdparty/whisper.cpp/ggml.c:9737-9948) plus the element-wise graph nodes that follow it in
// whisper_build_graph_{conv,encoder,cross,decoder} (whisper.cpp:1660-2505).
//
// Roles inside one 128-thread CTA: warp 0 / lane 0 = TMA producer, warp 1 / lane 0 = MMA issuer, warp 2 = TMEM
// allocator; afterwards all four warps drain their 32 TMEM lanes (one token row per thread).
#include "dev.cuh"
#include "tc.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace wb200 {

namespace {

using namespace tc;

constexpr int kBlockN  = 128;   // token rows per CTA  (UMMA M)
constexpr int kBlockK  = 64;    // 64 f16 = 128 bytes = one swizzle atom row
constexpr int kUmmaK   = 16;

template <int BM, int ST>   // BM = feature columns per CTA (UMMA N): 64, 128 or 256; ST = shared-memory stages of the operand ring
struct Cfg {
    static constexpr int kStages   = ST;
    static constexpr int kABytes   = kBlockN * kBlockK * 2;       // 16 KB
    static constexpr int kWBytes   = BM * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kWBytes;
    static constexpr int kSmemBytes  = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int kTmemCols   = BM < 32 ? 32 : BM;
};

// Epilogue of one 128 x BM accumulator tile: thread <-> token row (TMEM lane quad * 32 + lane), 16 feature columns at a time.
template <int BM>
__device__ __forceinline__ void gemm_epilogue(const GemmEpi & epi, int N, int M, int n0, int m0, int b1, int b2, uint32_t tmem_acc, int quad, int lane) {
    const int seg_i = epi.nseg > 1 ? (m0 / epi.seg_m) : 0;
    const EpiSeg & sg = epi.seg[seg_i];
    const int m_seg0 = m0 - seg_i * epi.seg_m;                    // first feature of this tile inside its segment
    const int m_lim  = (epi.nseg > 1 ? epi.seg_m : M) - m_seg0;   // valid features in this tile (may exceed BM)
    const int n = n0 + quad * 32 + lane;
    const bool n_ok = n < N;
    const uint32_t t_lane = tmem_acc + ((uint32_t) (quad * 32) << 16);

    const bool vec_ok = (m_lim >= BM)
        && (!sg.out32 || ((sg.out32_ld & 3) == 0 && (sg.out32_bs1 & 3) == 0 && (sg.out32_bs2 & 3) == 0 && ((uintptr_t) sg.out32 & 15) == 0))
        && (!sg.out16 || ((sg.out16_ld & 7) == 0 && (sg.out16_bs1 & 7) == 0 && (sg.out16_bs2 & 7) == 0 && ((uintptr_t) sg.out16 & 15) == 0))
        && (!sg.res   || ((sg.res_ld & 3) == 0 && ((uintptr_t) sg.res & 15) == 0));

    const int64_t r16  = (sg.out16 && n_ok)  ? (sg.rowmap16  ? (int64_t) sg.rowmap16[n]  : (int64_t) n) : 0;
    const int64_t c16t = (sg.out16t && n_ok) ? (sg.rowmap16t ? (int64_t) sg.rowmap16t[n] : (int64_t) n) : 0;
    const int rn = sg.res_mod ? (n % sg.res_mod) : n;
    const int b2o = sg.bmap2 ? __ldg(sg.bmap2 + b2) : b2;       // outer batch index as the output strides see it

    // the residual of the next 16 columns is fetched before this chunk's stores are issued (one HBM round trip per tile row instead
    // of one per chunk; in-place residual updates are safe: a thread only ever reads what it has not written yet)
    const bool res_pre = vec_ok && sg.res && n_ok;
    float4 rpre[4];
    if (res_pre && m_lim > 0) {
        const float * rp = sg.res + (int64_t) rn * sg.res_ld + m_seg0;
#pragma unroll
        for (int i = 0; i < 4; ++i) rpre[i] = __ldg((const float4 *) (rp + 4 * i));
    }
#pragma unroll 1
    for (int c = 0; c < BM / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(t_lane + (uint32_t) (c * 16), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = m_seg0 + c * 16;                            // feature index inside the segment
        float v[16];
        if (!n_ok || c * 16 >= m_lim) {
            // nothing to store for this row / these columns (tile tail)
        } else if (vec_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            if (sg.bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b = __ldg((const float4 *) (sg.bias + m + i));
                    v[i] = __fadd_rn(v[i], b.x); v[i + 1] = __fadd_rn(v[i + 1], b.y);
                    v[i + 2] = __fadd_rn(v[i + 2], b.z); v[i + 3] = __fadd_rn(v[i + 3], b.w);
                }
            }
            if (sg.scale != 1.0f) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], sg.scale);
            }
            if (sg.gelu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = gelu_table(epi.gelu_lut, v[i]);
            }
            uint32_t pk[8];
            if (sg.out16 && sg.out16_pre) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    pk[i] = *(const uint32_t *) &h;
                }
            }
            if (sg.res) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b = rpre[i >> 2];
                    v[i] = __fadd_rn(v[i], b.x); v[i + 1] = __fadd_rn(v[i + 1], b.y);
                    v[i + 2] = __fadd_rn(v[i + 2], b.z); v[i + 3] = __fadd_rn(v[i + 3], b.w);
                }
                if ((c + 1) * 16 < BM && (c + 1) * 16 < m_lim) {
                    const float * rp = sg.res + (int64_t) rn * sg.res_ld + m + 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) rpre[i] = __ldg((const float4 *) (rp + 4 * i));
                }
            }
            if (sg.out32) {
                float * op = sg.out32 + (int64_t) b2o * sg.out32_bs2 + (int64_t) b1 * sg.out32_bs1 + (int64_t) n * sg.out32_ld + m;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *(float4 *) (op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            if (sg.out16) {
                __half * op = sg.out16 + (int64_t) b2o * sg.out16_bs2 + (int64_t) b1 * sg.out16_bs1 + r16 * sg.out16_ld + m;
                if (!sg.out16_pre) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        pk[i] = *(const uint32_t *) &h;
                    }
                }
                *(uint4 *) op       = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *(uint4 *) (op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            if (sg.out16t) {
                __half * op = sg.out16t + (int64_t) b2o * sg.out16t_bs2 + (int64_t) b1 * sg.out16t_bs1 + (int64_t) m * sg.out16t_ld + c16t;
#pragma unroll
                for (int i = 0; i < 16; ++i) op[(int64_t) i * sg.out16t_ld] = __float2half_rn(v[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (c * 16 + i < m_lim) {
                    float pre;
                    const float x = epi_value(sg, epi.gelu_lut, __uint_as_float(r[i]), n, m + i, &pre);
                    epi_store(sg, x, pre, n, m + i, b1, b2);
                }
            }
        }
        __syncwarp();
    }

}

template <int BM, int ST>
__global__ void __launch_bounds__(128)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
          int N, int M, int K, int nb1, int w_batched, const GemmEpi epi) {
    using C = Cfg<BM, ST>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    uint64_t * bars = (uint64_t *) (smem + C::kStages * C::kStageBytes);      // full[kStages], empty[kStages], done
    uint32_t * tmem_slot = (uint32_t *) (bars + 2 * C::kStages + 1);

    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // (provably warp-uniform: role branches stay converged)
    const int n0 = blockIdx.x * kBlockN;
    const int m0 = blockIdx.y * BM;
    const int b1 = blockIdx.z % nb1, b2 = blockIdx.z / nb1;
    const int num_k = (K + kBlockK - 1) / kBlockK;

    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + C::kStages), done = smem_u32(bars + 2 * C::kStages);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t) C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // producer and issuer walk their loops as whole warps; one elected lane (elect.sync) issues the TMA / tcgen05 instructions
    if (warp == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < num_k; ++kb) {
            const int s = kb % C::kStages;
            const uint32_t ph = (kb / C::kStages) & 1;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (elect_one()) {
                const uint32_t a_dst = smem_u32(smem + s * C::kStageBytes);
                const uint32_t w_dst = a_dst + C::kABytes;
                mbar_arrive_expect_tx(full0 + 8 * s, C::kStageBytes);
                tma_load_4d(a_dst, &tmA, full0 + 8 * s, kb * kBlockK, n0, b1, b2);
                tma_load_4d(w_dst, &tmW, full0 + 8 * s, kb * kBlockK, m0, w_batched ? b1 : 0, w_batched ? b2 : 0);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = (1u << 4)                      // D = f32
                                 | (0u << 7) | (0u << 10)         // A, B = f16
                                 | (0u << 15) | (0u << 16)        // A, B K-major
                                 | ((uint32_t) (BM >> 3) << 17)   // UMMA N
                                 | ((uint32_t) (kBlockN >> 4) << 24);  // UMMA M = 128
        for (int kb = 0; kb < num_k; ++kb) {
            const int s = kb % C::kStages;
            const uint32_t ph = (kb / C::kStages) & 1;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t a_addr = smem_u32(smem + s * C::kStageBytes);
                const uint64_t adesc = umma_desc_sw128(a_addr);
                const uint64_t bdesc = umma_desc_sw128(a_addr + C::kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // advance 16 elements = 32 bytes inside the swizzle atom: +2 in 16-byte units
                    umma_f16(tmem_base, adesc + (uint64_t) (2 * k), bdesc + (uint64_t) (2 * k), idesc, (kb | k) != 0);
                }
                umma_commit(empty0 + 8 * s);                      // frees the smem stage when the MMAs retire
                if (kb == num_k - 1) umma_commit(done);           // accumulator complete
            }
        }
    }
    __syncwarp();

    // ---- epilogue: thread <-> token row, 16 feature columns at a time ----
    mbar_wait(done, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    gemm_epilogue<BM>(epi, N, M, n0, m0, b1, b2, tmem_base, warp, lane);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t) C::kTmemCols) : "memory");
    }
}

// ---- persistent variant ------------------------------------------------------------------------------------------------
// One CTA per SM slot walks over the tiles of the launch (tile t -> CTA t mod grid).  Six warps: TMA producer, MMA issuer, four
// epilogue warps.  The operand ring runs on across tile boundaries and the accumulator is double-buffered in TMEM, so the loads and
// MMAs of tile i + 1 run under the epilogue of tile i, and barrier setup / TMEM allocation / descriptor fetch are paid once per CTA
// instead of once per tile.  Same operands, same epilogue (gemm_epilogue) as k_gemm_tc.
template <int BM, int ST>
__global__ void __launch_bounds__(192)
k_gemm_tc_persistent(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     int N, int M, int K, int nb1, int w_batched, int tiles_n, int tiles_m, int n_tiles, const GemmEpi epi) {
    using C = Cfg<BM, ST>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    uint64_t * bars = (uint64_t *) (smem + C::kStages * C::kStageBytes);      // full[kStages], empty[kStages], acc_full[2], acc_empty[2]
    uint32_t * tmem_slot = (uint32_t *) (bars + 2 * C::kStages + 4);
    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // (provably warp-uniform: role branches stay converged)
    const int num_k = (K + kBlockK - 1) / kBlockK;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + C::kStages);
    const uint32_t acc_full = smem_u32(bars + 2 * C::kStages), acc_empty = acc_full + 16;
    constexpr uint32_t kAccCols = 2 * (BM < 32 ? 32 : BM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t: n tile fastest (CTAs that run side by side share the weight tile), then m tile, then batch
    auto tile_coords = [&](int t, int & n0, int & m0, int & b1, int & b2) {
        const int tn = t % tiles_n, r = t / tiles_n;
        const int tm = r % tiles_m, bz = r / tiles_m;
        n0 = tn * kBlockN; m0 = tm * BM; b1 = bz % nb1; b2 = bz / nb1;
    };

    if (warp == 0) {
        // ---- TMA producer (whole warp in the loop, one elected lane issues) ----
        int it = 0;                                                   // k blocks issued so far (ring position)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int n0, m0, b1, b2; tile_coords(t, n0, m0, b1, b2);
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % C::kStages;
                mbar_wait(empty0 + 8 * s, ((it / C::kStages) & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t a_dst = smem_u32(smem + s * C::kStageBytes);
                    mbar_arrive_expect_tx(full0 + 8 * s, C::kStageBytes);
                    tma_load_4d(a_dst, &tmA, full0 + 8 * s, kb * kBlockK, n0, b1, b2);
                    tma_load_4d(a_dst + C::kABytes, &tmW, full0 + 8 * s, kb * kBlockK, m0, w_batched ? b1 : 0, w_batched ? b2 : 0);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t) (BM >> 3) << 17) | ((uint32_t) (kBlockN >> 4) << 24);   // f16 x f16 -> f32, K-major, 128 x BM
        int it = 0, i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int ab = i & 1;
            mbar_wait(acc_empty + 8 * ab, ((i >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t) (ab * (kAccCols / 2));
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % C::kStages;
                mbar_wait(full0 + 8 * s, (it / C::kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + s * C::kStageBytes);
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + C::kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_f16(acc, adesc + (uint64_t) (2 * k), bdesc + (uint64_t) (2 * k), idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * s);
                    if (kb == num_k - 1) umma_commit(acc_full + 8 * ab);
                }
            }
        }
    } else if (warp >= 2) {
        // ---- epilogue warps: TMEM lane quadrant = warp % 4 ----
        const int quad = warp & 3;
        int i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int ab = i & 1;
            int n0, m0, b1, b2; tile_coords(t, n0, m0, b1, b2);
            mbar_wait(acc_full + 8 * ab, (i >> 1) & 1);
            tc_fence_after();
            gemm_epilogue<BM>(epi, N, M, n0, m0, b1, b2, tmem_base + (uint32_t) (ab * (kAccCols / 2)), quad, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * ab);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kAccCols) : "memory");
    }
}

// ---- tensor maps -------------------------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = (PFN_encodeTiled) p;
        }
    });
    return fn;
}

struct MapKey {
    const void * p; int64_t ld, bs1, bs2; int rows, K, nb1, nb2, box_rows;
    bool operator<(const MapKey & o) const {
        return std::tie(p, ld, bs1, bs2, rows, K, nb1, nb2, box_rows) < std::tie(o.p, o.ld, o.bs1, o.bs2, o.rows, o.K, o.nb1, o.nb2, o.box_rows);
    }
};

std::mutex g_map_mutex;
std::map<MapKey, CUtensorMap> g_maps;

bool make_map(const Operand & op, int K, int nb1, int nb2, int box_rows, CUtensorMap & out) {
    const MapKey key{op.p, op.ld, op.bs1, op.bs2, op.rows, K, nb1, nb2, box_rows};
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { out = it->second; return true; }
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled unavailable\n"); return false; }
    // strides of size-1 dims are irrelevant but must still be multiples of 16 bytes
    const int64_t bs1 = nb1 > 1 ? op.bs1 : op.ld * (int64_t) op.rows;
    const int64_t bs2 = nb2 > 1 ? op.bs2 : bs1 * nb1;
    cuuint64_t dims[4]    = {(cuuint64_t) K, (cuuint64_t) op.rows, (cuuint64_t) nb1, (cuuint64_t) nb2};
    cuuint64_t strides[3] = {(cuuint64_t) op.ld * 2, (cuuint64_t) (bs1 > 0 ? bs1 : 8) * 2, (cuuint64_t) (bs2 > 0 ? bs2 : 8) * 2};
    cuuint32_t box[4]     = {(cuuint32_t) kBlockK, (cuuint32_t) box_rows, 1, 1};
    cuuint32_t estr[4]    = {1, 1, 1, 1};
    if (((uintptr_t) op.p & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        fprintf(stderr, "whisper_b200: operand not 16-byte aligned for TMA (p=%p ld=%lld bs1=%lld bs2=%lld)\n",
                (const void *) op.p, (long long) op.ld, (long long) op.bs1, (long long) op.bs2);
        return false;
    }
    const CUresult rc = enc(&out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *) op.p, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled failed (%d) K=%d rows=%d ld=%lld nb=(%d,%d)\n", (int) rc, K, op.rows,
                (long long) op.ld, nb1, nb2);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_map_mutex);
    g_maps[key] = out;
    return true;
}

int g_persistent_ctas = -1;    // CTAs of a persistent launch (SM count x 2); 0 = never use the persistent kernel (WHISPER_B200_GEMM_PERSISTENT=0)

template <int BM, int ST>
bool launch_bm_persistent(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, int max_ctas, cudaStream_t st) {
    using C = Cfg<BM, ST>;
    static bool attr_set[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(k_gemm_tc_persistent<BM, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) return false;
        attr_set[dev] = true;
    }
    CUtensorMap tmA, tmW;
    if (!make_map(A, sh.K, sh.nb1, sh.nb2, kBlockN, tmA)) return false;
    const int w_batched = (sh.nb1 * sh.nb2 > 1) && (W.bs1 != 0 || W.bs2 != 0);
    if (!make_map(W, sh.K, w_batched ? sh.nb1 : 1, w_batched ? sh.nb2 : 1, BM, tmW)) return false;
    const int tiles_n = (sh.N + kBlockN - 1) / kBlockN, tiles_m = (sh.M + BM - 1) / BM;
    const int n_tiles = tiles_n * tiles_m * sh.nb1 * sh.nb2;
    k_gemm_tc_persistent<BM, ST><<<std::min(n_tiles, max_ctas), 192, C::kSmemBytes, st>>>(tmA, tmW, sh.N, sh.M, sh.K, sh.nb1, w_batched, tiles_n, tiles_m, n_tiles, epi);
    return cudaGetLastError() == cudaSuccess;
}

template <int BM, int ST>
bool launch_bm(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, cudaStream_t st) {
    using C = Cfg<BM, ST>;
    static bool attr_set[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(k_gemm_tc<BM, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) {
            fprintf(stderr, "whisper_b200: cannot reserve %d bytes of shared memory for the tcgen05 GEMM\n", C::kSmemBytes);
            return false;
        }
        attr_set[dev] = true;
    }
    CUtensorMap tmA, tmW;
    if (!make_map(A, sh.K, sh.nb1, sh.nb2, kBlockN, tmA)) return false;
    const int w_batched = (sh.nb1 * sh.nb2 > 1) && (W.bs1 != 0 || W.bs2 != 0);
    if (!make_map(W, sh.K, w_batched ? sh.nb1 : 1, w_batched ? sh.nb2 : 1, BM, tmW)) return false;
    dim3 grid((sh.N + kBlockN - 1) / kBlockN, (sh.M + BM - 1) / BM, sh.nb1 * sh.nb2);
    k_gemm_tc<BM, ST><<<grid, 128, C::kSmemBytes, st>>>(tmA, tmW, sh.N, sh.M, sh.K, sh.nb1, w_batched, epi);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace

// 2-D f16 tensor map with 128-byte swizzle (rows of `cols` elements, `row_stride_bytes` apart; box = box_cols x box_rows).
// Used by the decode-step kernel to stream cross-attention K / V^T chunks with one TMA instruction per box.
bool make_tensor_map_2d_f16(void * out_map, const void * base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                            uint32_t box_cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled unavailable\n"); return false; }
    cuuint64_t dims[2]    = {(cuuint64_t) cols, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) row_stride_bytes};
    cuuint32_t box[2]     = {box_cols, box_rows};
    cuuint32_t estr[2]    = {1, 1};
    const CUresult rc = enc((CUtensorMap *) out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *) base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        fprintf(stderr, "whisper_b200: cuTensorMapEncodeTiled (2-D) failed (%d) cols=%llu rows=%llu stride=%llu\n", (int) rc,
                (unsigned long long) cols, (unsigned long long) rows, (unsigned long long) row_stride_bytes);
        return false;
    }
    return true;
}

bool gemm_tc_make_map(const Operand & op, int K, int nb1, int nb2, int box_rows, void * out_map) {
    return make_map(op, K, nb1, nb2, box_rows, *(CUtensorMap *) out_map);
}

void gemm_tc_forget_maps() {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    g_maps.clear();
}

bool launch_gemm_tc(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, cudaStream_t st) {
    // tile width: segments must be tile-aligned; 64-wide tiles for the per-head P·V contraction (M = 64)
    const int seg = epi.nseg > 1 ? epi.seg_m : sh.M;
    // few rows against a small weight matrix (the linear maps of a wide decoder pass: 33..512 rows, M <= 8192): the tile count, not
    // the tensor pipe, bounds these — 32-wide tiles put 4x as many CTAs to work, each with a two-iteration epilogue, and an
    // eight-stage ring has the whole K extent of the narrow maps in flight at once
    if (sh.nb1 * sh.nb2 == 1 && sh.N <= 512 && sh.M <= 8192 && seg % 32 == 0) return launch_bm<32, 8>(A, W, sh, epi, st);
    if (seg % 128 == 0 || seg > 128) {
        if (epi.nseg > 1 && seg % 128 != 0) {
            fprintf(stderr, "whisper_b200: segment width %d is not a multiple of the 128-wide tile\n", seg);
            return false;
        }
        // a contraction of one 64-wide k block (attention scores, K = d_head) needs no operand ring: one stage leaves room
        // for four CTAs per SM (the TMEM limit at 128 columns each), which is what hides the store-bound epilogue
        if (sh.K <= kBlockK) return launch_bm<128, 1>(A, W, sh, epi, st);
        if (g_persistent_ctas < 0) {
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            g_persistent_ctas = 2 * sms;
            if (const char * e = getenv("WHISPER_B200_GEMM_PERSISTENT")) g_persistent_ctas = atoi(e) > 0 ? atoi(e) * sms : 0;
        }
        // many tiles per SM: the persistent kernel pipelines them (loads and MMAs of the next tile under the epilogue of this one)
        const int64_t n_tiles = (int64_t) ((sh.N + kBlockN - 1) / kBlockN) * ((sh.M + 127) / 128) * sh.nb1 * sh.nb2;
        if (g_persistent_ctas > 0 && n_tiles > (int64_t) g_persistent_ctas) return launch_bm_persistent<128, 3>(A, W, sh, epi, g_persistent_ctas, st);
        return launch_bm<128, 3>(A, W, sh, epi, st);
    }
    return launch_bm<64, 4>(A, W, sh, epi, st);
}

}  // namespace wb200
