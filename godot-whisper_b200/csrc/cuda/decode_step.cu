// k_decode_step — one decoder token step (n <= 16 sequences, one new token each) as a single persistent cooperative kernel.
// See decode_step.cuh for what it replaces.  Structure:
//
//   phases   : embed | per layer { QKV, self-attention, out-proj, cross-Q, cross-attention, cross-out-proj, FC1, FC2 } |
//              logits (+ logits rules and per-CTA sampler partials) | sampler finalize.   The phase table is built on the host.
//   barrier  : grid-wide, split into arrive (one release-add on a 64-bit counter) and wait (relaxed spin, bounded => trap)
//   jobs     : a phase is cut into jobs (16*tj weight rows of a linear map, possibly in column slabs; or one key chunk of one
//              (row, head) of attention); job j belongs to CTA j mod grid.  Everything a job needs that does NOT depend on this
//              step (weights, cross K/V, older self-attention cells) is fetched with cp.async into a 3-slot shared-memory ring,
//              two jobs ahead of the one being computed — the look-ahead crosses phase boundaries, so a CTA's weights are
//              already on chip when the barrier in front of them opens.
//   math     : linear maps = mma.sync.m16n8k16 (f16 x f16 -> f32), activations (<= 16 rows, rounded to f16 exactly where the
//              reference rounds them, ggml.c:9841-9857) as the B operand; K is split over the warps of a CTA when a phase has
//              few weight rows.  LayerNorm (f64 sums, ggml.c:9301-9352), table GELU / exp (ggml.c:1416-1423, 11170-11192),
//              f64 softmax sums, f16 P: the same arithmetic as the multi-kernel path in kernels.cu.
#include "decode_step.cuh"

#include <algorithm>
#include <cstdio>

namespace wb200 {

namespace {

constexpr int kThreads  = 256;
constexpr int kWarps    = 8;
constexpr int kSlots    = 3;
constexpr int kKeysCap  = 1536;                 // >= max(n_audio_ctx, self-attention cells)
constexpr int kKRow     = 72;                   // halves between consecutive key rows of a K chunk in shared memory (64 + 8: conflict-free ldmatrix)
constexpr int kLnMax4   = 6;                    // 4-feature groups per lane in the LayerNorm prologue (d <= 768)

struct Misc {
    StepPhase ph[kStepMaxPhases];
    float  red[kWarps * 16 * 17];               // K-split partial accumulators: [ks * tj + tile][row][17]
    float  sc[kKeysCap];                        // attention scores, then exp values
    __half p16[kKeysCap];                       // normalised probabilities (f16, as the reference's P operand)
    float  redf[kWarps];
    float  redf2[64];
    double redd[kWarps];
    // per-row step metadata, copied from the staging block once per launch
    int64_t koff_self[kStepMaxRows], voff_self[kStepMaxRows], koff_cross[kStepMaxRows], voff_cross[kStepMaxRows];
    int     wslot[kStepMaxRows], rule[kStepMaxRows][4], rowmap_k[kStepMaxRows], rowmap_v[kStepMaxRows], own[kStepMaxRows];
    int     ticket;
    alignas(8) uint64_t tma_bar[kSlots];       // one mbarrier per ring slot for TMA-fetched jobs
    uint8_t cls_job[128];                       // token class bits of the vocabulary rows of the current logits job
};

// ---- small device helpers ----------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void * dst_smem, const void * src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) __trap();       // a protocol bug must fail the launch, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap * map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"((uint64_t) map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier over all CTAs of the (co-resident) grid, split in two so that work which does not depend on other CTAs
// (issuing prefetches) sits between arrival and wait.  `target` is the running arrival count thread 0 waits for.
__device__ __forceinline__ void barrier_arrive(unsigned long long * bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // release: the CTA's writes (ordered before by bar.sync) are visible GPU-wide
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" :: "l"(bar) : "memory");
    }
}
// Every cross-CTA read after the barrier goes through L2 (ld.global.cg / cp.async.cg), the coherence point, and is issued only
// after thread 0 has observed the full arrival count, so no cache invalidation is needed on this side.
__device__ __forceinline__ void barrier_wait(unsigned long long * bar, unsigned long long target) {
    if (threadIdx.x == 0) {
        if (ld_relaxed_u64(bar) < target) {
            const long long t0 = clock64();
            int spins = 0;
            while (ld_relaxed_u64(bar) < target) {
                if ((++spins & 1023) == 0 && clock64() - t0 > 6000000000LL) __trap();   // ~3 s: fail the launch, never hang the GPU
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void * smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(smem_ptr)));
}

__device__ __forceinline__ float exp_table(const uint16_t * __restrict__ lut, float x) {
    const uint16_t h = __half_as_ushort(__float2half_rn(x));
    return __half2float(__ushort_as_half(__ldg(lut + h)));
}

// LayerNorm of two rows by one warp with the reference's arithmetic (ggml.c:9301-9352, whisper.cpp:2236-2246: f64 sums, mean
// and variance rounded to f32, one rounding per operation).  Rows are read once through L2 and kept in registers; gamma and
// beta are fetched alongside; the two rows' dependency chains interleave.  With `emb` set the rows are the token + positional
// embedding (whisper.cpp:2229-2233) and, if x_out is given, are also stored as the residual stream.
struct LnSrc { const float * x; const __half * te; const float * pe; };
// s / d for an integer-valued d, correctly rounded (Markstein: q = s*r, one FMA residual, one FMA correction; r = RN(1/d))
__device__ __forceinline__ double div_by_d(double s, double dd, double inv_d) {
    const double q = s * inv_d;
    const double rem = fma(-q, dd, s);
    return fma(rem, inv_d, q);
}
// lane l holds features 128 j + 4 l .. + 3 (j < per_lane4): 16-byte loads, a quarter of the memory requests of a scalar layout
__device__ __forceinline__ void ln_load(const LnSrc & src, float4 (&v)[kLnMax4], int per_lane4, int lane) {
#pragma unroll
    for (int j = 0; j < kLnMax4; ++j) {
        float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (j < per_lane4) {
            const int e = 128 * j + 4 * lane;
            if (src.te) {
                const uint2 h = __ldg((const uint2 *) (src.te + e));
                const float4 p = __ldg((const float4 *) (src.pe + e));
                const float2 a = __half22float2(*(const __half2 *) &h.x), b = __half22float2(*(const __half2 *) &h.y);
                t = make_float4(__fadd_rn(a.x, p.x), __fadd_rn(a.y, p.y), __fadd_rn(b.x, p.z), __fadd_rn(b.y, p.w));
            } else {
                t = __ldcg((const float4 *) (src.x + e));
            }
        }
        v[j] = t;
    }
}
__device__ __forceinline__ double sum4(const float4 & t) { return ((double) t.x + (double) t.y) + ((double) t.z + (double) t.w); }
__device__ __forceinline__ double sq4(const float4 & t, float mean) {
    const float a = __fsub_rn(t.x, mean), b = __fsub_rn(t.y, mean), c = __fsub_rn(t.z, mean), d = __fsub_rn(t.w, mean);
    return ((double) __fmul_rn(a, a) + (double) __fmul_rn(b, b)) + ((double) __fmul_rn(c, c) + (double) __fmul_rn(d, d));
}
__device__ __forceinline__ void ln_store4(__half * out, const float4 & t, float mean, float sc, const float4 & g, const float4 & b) {
    const float y0 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t.x, mean), sc), g.x), b.x);
    const float y1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t.y, mean), sc), g.y), b.y);
    const float y2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t.z, mean), sc), g.z), b.z);
    const float y3 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t.w, mean), sc), g.w), b.w);
    const __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
    uint2 pk;
    pk.x = *(const uint32_t *) &h0; pk.y = *(const uint32_t *) &h1;
    *(uint2 *) out = pk;
}
// f64 sums as independent chains (f64 adds have a long latency on this part); summing <= 768 floats in f64 is exact to ~1e-16,
// so the order does not reach the f32 result
__device__ __forceinline__ void ln_rows2(const LnSrc & src0, const LnSrc & src1, bool on0, bool on1, float * xo0, float * xo1,
                                         const float * __restrict__ gamma, const float * __restrict__ beta,
                                         __half * out0, __half * out1, int d, double inv_d, int per_lane4, float eps, int lane) {
    const double dd = (double) d;
    float4 v0[kLnMax4], v1[kLnMax4], gm[kLnMax4], bt[kLnMax4];
    if (on0) ln_load(src0, v0, per_lane4, lane);
    else {
#pragma unroll
        for (int j = 0; j < kLnMax4; ++j) v0[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (on1) ln_load(src1, v1, per_lane4, lane);
    else {
#pragma unroll
        for (int j = 0; j < kLnMax4; ++j) v1[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
#pragma unroll
    for (int j = 0; j < kLnMax4; ++j) {
        gm[j] = j < per_lane4 ? __ldg((const float4 *) (gamma + 128 * j + 4 * lane)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        bt[j] = j < per_lane4 ? __ldg((const float4 *) (beta + 128 * j + 4 * lane)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (xo0 && on0) {
#pragma unroll
        for (int j = 0; j < kLnMax4; ++j) if (j < per_lane4) *(float4 *) (xo0 + 128 * j + 4 * lane) = v0[j];
    }
    if (xo1 && on1) {
#pragma unroll
        for (int j = 0; j < kLnMax4; ++j) if (j < per_lane4) *(float4 *) (xo1 + 128 * j + 4 * lane) = v1[j];
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < kLnMax4; ++j) if (j < per_lane4) { s0 += sum4(v0[j]); s1 += sum4(v1[j]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    const float mean0 = (float) div_by_d(s0, dd, inv_d), mean1 = (float) div_by_d(s1, dd, inv_d);
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int j = 0; j < kLnMax4; ++j) if (j < per_lane4) { q0 += sq4(v0[j], mean0); q1 += sq4(v1[j], mean1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
    const float var0 = (float) div_by_d(q0, dd, inv_d), var1 = (float) div_by_d(q1, dd, inv_d);
    const float sc0 = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var0, eps))), sc1 = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var1, eps)));
#pragma unroll
    for (int j = 0; j < kLnMax4; ++j) {
        if (j < per_lane4) {
            if (on0) ln_store4(out0 + 128 * j + 4 * lane, v0[j], mean0, sc0, gm[j], bt[j]);
            if (on1) ln_store4(out1 + 128 * j + 4 * lane, v1[j], mean1, sc1, gm[j], bt[j]);
        }
    }
}

__device__ __forceinline__ bool token_masked(int i, int flags, int cls, int beg, int eot, int tid0_initial, int tid0_seek) {
    if (cls & 1) return true;                                                     // <|notimestamps|>, sot, nosp, translate, transcribe, prev, languages
    if ((cls & 2) && (flags & 8)) return true;                                    // non-speech symbols
    if ((cls & 4) && (flags & 1)) return true;                                    // blank / eot at the start of a sequence
    if ((cls & 8) && (flags & 4)) return true;                                    // solm unless tinydiarize
    if ((flags & 2) && i >= beg) return true;                                     // no_timestamps
    if (flags & 16) {                                                             // last token was a timestamp
        if (flags & 32) { if (i >= beg) return true; } else { if (i < eot) return true; }
    }
    if ((flags & 64) && i >= beg + tid0_initial + 1) return true;                 // max_initial_ts
    if ((flags & 128) && i >= beg && i < beg + tid0_seek) return true;            // timestamps do not decrease
    return false;
}

// running (max, first index of max, sum of exp(x - max)) and its merge
struct Stat { float m; int i; double s; };
__device__ __forceinline__ Stat stat_merge(const Stat & a, const Stat & b) {
    Stat r;
    if (b.m > a.m || (b.m == a.m && b.i < a.i)) { r.m = b.m; r.i = b.i; } else { r.m = a.m; r.i = a.i; }
    double s = 0.0;
    if (a.s > 0.0) s += a.s * (double) expf(a.m - r.m);
    if (b.s > 0.0) s += b.s * (double) expf(b.m - r.m);
    r.s = s;
    return r;
}

__device__ __forceinline__ Stat stat_shfl_xor(const Stat & a, int o) {
    Stat r;
    r.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    r.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    r.s = __shfl_xor_sync(0xffffffffu, a.s, o);
    return r;
}
// Per-thread running statistics of the logits phase: a thread sees at most a few hundred terms, so its partial sum lives in f32
// and uses the fast exponential (the reference sums all 51 864 terms sequentially in f32, whisper.cpp:4644-4649); the per-CTA
// and final merges run in f64.
struct StatF { float m; int i; float s; };
__device__ __forceinline__ void stat_add(StatF & a, float x, int idx) {
    if (x > a.m) { a.s = (a.s > 0.0f ? a.s * __expf(a.m - x) : 0.0f) + 1.0f; a.m = x; a.i = idx; }
    else         { a.s += __expf(x - a.m); }
}

// ---- job cursor: the ordered list of this CTA's jobs over the whole step ----------------------------------------------------------

struct Cursor { int ph, j, sub; };

struct Geo {            // CTA-uniform values every job needs
    int cta, n_cta, n_phases, n_kv, kc_keys, kc_cross, nc_self, nc_cross;     // cta / n_cta: index inside / size of this CTA's row group
};

__device__ __forceinline__ int n_sub_of(const Misc & mi, const StepArgs & a, const Geo & g, int ph) {
    const int type = mi.ph[ph].type;
    if (type == STEP_GEMM) return mi.ph[ph].ksplit;
    return 2 * (type == STEP_SELF ? g.nc_self : g.nc_cross);
}
__device__ __forceinline__ void cursor_seek(const Misc & mi, const Geo & g, Cursor & c, int ph_from) {
    int ph = ph_from;
    while (ph < g.n_phases && g.cta >= mi.ph[ph].n_jobs) ++ph;
    c.ph = ph; c.j = g.cta; c.sub = 0;
}
__device__ __forceinline__ void cursor_advance(const Misc & mi, const StepArgs & a, const Geo & g, Cursor & c) {
    if (c.ph >= g.n_phases) return;
    if (++c.sub < n_sub_of(mi, a, g, c.ph)) return;
    c.sub = 0; c.j += g.n_cta;
    if (c.j < mi.ph[c.ph].n_jobs) return;
    cursor_seek(mi, g, c, c.ph + 1);
}

// issues the cp.async copies of job c into `slot` (the caller commits the group)
// Cross-attention chunks come through the TMA unit (one instruction per box, issued by thread 0, completion on the slot's
// mbarrier); returns true for such a job.  Everything else is cp.async by all threads.
__device__ bool fetch_job(Misc & mi, const StepArgs & a, const Geo & g, const Cursor & c, uint8_t * slot, int slot_idx) {
    if (c.ph >= g.n_phases) return false;
    const StepPhase & p = mi.ph[c.ph];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __half * dst0 = (__half *) slot;
    if (p.type == STEP_GEMM && p.tma) {
        // one box of 64 columns x 16 tj rows per 64-column block of K; rows past n_vocab read as zeros
        if (threadIdx.x == 0) {
            const int rows = p.tj * 16, nkb = p.K >> 6;
            const uint32_t bar = smem_addr(&mi.tma_bar[slot_idx]), dst = smem_addr(slot);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(bar, (uint32_t) (rows * p.K * 2));
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(dst + kb * rows * 128, &a.tm_te, bar, kb * 64, c.j * rows);
        }
        return true;
    }
    if (p.type == STEP_GEMM) {
        const int kc = p.kc, Ks = kc + 32;
        const int k_lo = c.sub * kc, k_n = min(kc, p.K - k_lo), kc8 = k_n >> 3;
        const int row0 = c.j * p.tj * 16;
        const int rows = min(p.tj * 16, p.M - row0);
        for (int r = warp; r < rows; r += kWarps) {
            const __half * src = p.W + (int64_t) (row0 + r) * p.K + k_lo;
            __half * dst = dst0 + (int64_t) r * Ks;
            for (int q = lane; q < kc8; q += 32) cp_async16(dst + q * 8, src + q * 8);
        }
        return false;
    }
    if (p.type == STEP_CROSS) {
        if (threadIdx.x == 0) {
            const int r = c.j / a.n_head, hh = c.j - r * a.n_head;
            const int kck = g.kc_cross, nc = g.nc_cross;
            const bool is_v = c.sub >= nc;
            const int k0 = (is_v ? c.sub - nc : c.sub) * kck;
            const uint32_t bar = smem_addr(&mi.tma_bar[slot_idx]), dst = smem_addr(slot);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the slot's earlier (generic-proxy) readers are done
            mbar_arrive_expect_tx(bar, (uint32_t) kck * 128u);
            if (!is_v) {
                // K rows of this (slot, layer): row index in the [slots*Lt*Tmax][d] view
                const int row0 = (int) (mi.koff_cross[r] / a.d) + p.layer * a.Tmax + k0;
                for (int b = 0; b < kck / 128; ++b) tma_load_2d(dst + b * 16384, &a.tm_cross_k, bar, hh * 64, row0 + b * 128);
            } else {
                // V^T rows (features) of this (slot, layer, head): row index in the [slots*Lt*d][Tpmax] view; one box per 64 keys
                const int row0 = (int) (mi.voff_cross[r] / a.Tpmax) + p.layer * a.d + hh * 64;
                for (int b = 0; b < kck / 64; ++b) tma_load_2d(dst + b * 8192, &a.tm_cross_v, bar, k0 + b * 64, row0);
            }
        }
        return true;
    }
    // attention: item j = (row r, head hh); sub-job = key chunk of K, then of V^T
    const int r = c.j / a.n_head, hh = c.j - r * a.n_head;
    const int il = p.layer, d = a.d, kck = g.kc_keys;
    const int n_keys = g.n_kv;
    const int nc = g.nc_self;
    const bool is_v = c.sub >= nc;
    const int k0 = (is_v ? c.sub - nc : c.sub) * kck;
    if (!is_v) {
        const int k1 = min(n_keys, k0 + kck);
        const __half * Kb = a.self_k + (int64_t) il * a.kv_cells * d + mi.koff_self[r];
        Kb += hh * 64 + 8 * (lane & 7);
        for (int jk = k0 + warp * 4 + (lane >> 3); jk < k1; jk += kWarps * 4)
            cp_async16(dst0 + (int64_t) (jk - k0) * kKRow + 8 * (lane & 7), Kb + (int64_t) jk * d);
    } else {
        const int n_pad = (n_keys + 7) & ~7;
        const int k1 = min(n_pad, k0 + kck);
        const int64_t ld_v = a.kv_cells;
        const __half * Vb = a.self_v + (int64_t) il * d * a.kv_cells + mi.voff_self[r];
        Vb += (int64_t) (hh * 64) * ld_v;
        const int pieces = (k1 - k0) >> 3;
        const int vrow = kck + 8;
        for (int f = warp; f < 64; f += kWarps) {
            const __half * src = Vb + (int64_t) f * ld_v + k0;
            __half * dst = dst0 + (int64_t) f * vrow;
            for (int q = lane; q < pieces; q += 32) cp_async16(dst + q * 8, src + q * 8);
        }
        // the P·V contraction runs in steps of 16 keys: keys between the last fetched one and the next multiple of 16 must be
        // finite (their probabilities are zero); the slot is free, so plain stores are fine
        if (((k1 - k0) & 15) && threadIdx.x < 64)
            *(uint4 *) (dst0 + (int64_t) threadIdx.x * vrow + (k1 - k0)) = make_uint4(0, 0, 0, 0);
    }
    return false;
}

// Sampler finalize for row r by one warp: merges the per-CTA partials and applies whisper_process_logits' timestamp-vs-text rule
// and whisper_sample_token's greedy pick (whisper.cpp:4637-4720, 4777-4834).
__device__ void finalize_row(const StepArgs & a, const double * records, int r, int r_global, int n_cta, int lane) {
    const int ws = a.wslot[r_global] - a.n_full;
    if (ws < 0) return;
    Stat tx{-INFINITY, 0x7fffffff, 0.0}, ts{-INFINITY, 0x7fffffff, 0.0};
    for (int c = lane; c < n_cta; c += 32) {
        const double * rec = records + ((int64_t) c * kStepMaxRows + r) * 6;
        Stat b0{(float) __ldcg(rec + 0), (int) __ldcg(rec + 1), __ldcg(rec + 2)};
        Stat b1{(float) __ldcg(rec + 3), (int) __ldcg(rec + 4), __ldcg(rec + 5)};
        tx = stat_merge(tx, b0); ts = stat_merge(ts, b1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { tx = stat_merge(tx, stat_shfl_xor(tx, o)); ts = stat_merge(ts, stat_shfl_xor(ts, o)); }
    if (lane != 0) return;
    const int beg = a.token_beg;
    const float M = fmaxf(tx.m, ts.m);
    double S = 0.0;
    if (tx.s > 0.0) S += tx.s * (double) expf(tx.m - M);
    if (ts.s > 0.0) S += ts.s * (double) expf(ts.m - M);
    const float lse = logf((float) S) + M;
    // timestamp mass vs the best text token (whisper.cpp:4659-4684)
    float ts_logprob = -INFINITY;
    if ((float) ts.s > 0.0f) ts_logprob = logf((float) ts.s) + (ts.m - lse);
    const float text_max = tx.s > 0.0 ? tx.m - lse : -INFINITY;
    const bool text_off = ts_logprob > text_max;
    const float p_text = tx.s > 0.0 ? expf(tx.m - lse) : 0.0f;
    const float p_tsb  = ts.s > 0.0 ? expf(ts.m - lse) : 0.0f;
    int id = 0, tid = 0;
    float pbest = 0.0f, plog = 0.0f;
    if (!text_off && p_text > 0.0f && p_text >= p_tsb) { id = tx.i; pbest = p_text; plog = tx.m - lse; }   // text ids precede timestamp ids: ties go to text
    else if (p_tsb > 0.0f)                             { id = ts.i; pbest = p_tsb;  plog = ts.m - lse; }
    if (p_tsb > 0.0f) tid = ts.i;
    const double p_ts_sum = ts.s > 0.0 ? ts.s * (double) expf(ts.m - lse) : 0.0;
    float pt = (float) ((double) p_tsb / (p_ts_sum + 1e-10));
    const float ptsum = (float) p_ts_sum;
    if (id >= beg) { tid = id; pt = pbest; }
    float * o = a.sampled + 6 * (int64_t) ws;
    o[0] = __int_as_float(id); o[1] = __int_as_float(tid); o[2] = pbest; o[3] = plog; o[4] = pt; o[5] = ptsum;
}

#define TRACE(slot_) do { if (a.trace && threadIdx.x == 0) a.trace[((int64_t) blockIdx.x * kStepMaxPhases + ph) * 8 + (slot_)] = globaltimer_ns(); } while (0)

// ---- the kernel ----------------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads, 1)
k_decode_step(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __half * xs = (__half *) smem;
    uint8_t * ring = smem + a.xs_bytes;
    Misc & mi = *(Misc *) (smem + a.xs_bytes + kSlots * a.slot_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int d = a.d, V = a.n_vocab;
    // row group of this CTA
    const int cta_per = gridDim.x / a.n_groups;
    const int grp = blockIdx.x / cta_per;
    if (grp >= a.n_groups) return;                         // left-over CTAs of an uneven split take no part (not even in barriers)
    int row0 = 0;
    for (int i = 0; i < grp; ++i) row0 += a.n_grp[i];
    const int n = a.n_groups > 1 ? a.n_grp[grp] : a.n;
    float * const x32 = a.x32 + (int64_t) row0 * d;
    __half * const q16 = a.q16 + (int64_t) row0 * d, * const attn16 = a.attn16 + (int64_t) row0 * d, * const h16 = a.h16 + (int64_t) row0 * 4 * d;
    double * const records = a.records + (int64_t) grp * cta_per * kStepMaxRows * 6;
    unsigned long long * const bar = a.bar + 2 * grp;
    const int nt_count = (n + 7) >> 3;                     // activation row tiles of 8
    const double inv_d = 1.0 / (double) d;
    Geo geo;
    geo.cta = blockIdx.x - grp * cta_per; geo.n_cta = cta_per; geo.n_phases = a.n_phases; geo.kc_keys = a.chunk_keys;
    geo.n_kv = min(__ldg(a.n_kv_dev), a.kv_cells);
    geo.nc_self = (geo.n_kv + geo.kc_keys - 1) / geo.kc_keys;
    geo.kc_cross = a.chunk_keys_cross;
    geo.nc_cross = (a.n_audio_ctx + geo.kc_cross - 1) / geo.kc_cross;

    // phase table -> shared memory; sampler partials of this CTA
    {
        const int words = a.n_phases * (int) (sizeof(StepPhase) / 4);
        const uint32_t * src = (const uint32_t *) (a.n_groups > 1 ? a.phases_grp[grp] : a.phases);
        uint32_t * dst = (uint32_t *) mi.ph;
        for (int i = threadIdx.x; i < words; i += kThreads) dst[i] = __ldg(src + i);
        if (threadIdx.x < n) {
            const int r = threadIdx.x, rg = row0 + r;
            mi.koff_self[r] = a.koff_self[rg]; mi.voff_self[r] = a.voff_self[rg];
            mi.koff_cross[r] = a.koff_cross[rg]; mi.voff_cross[r] = a.voff_cross[rg];
            const int ws = a.wslot[rg];
            mi.wslot[r] = ws;
#pragma unroll
            for (int i = 0; i < 4; ++i) mi.rule[r][i] = ws >= a.n_full ? a.rule[4 * (ws - a.n_full) + i] : 0;
            mi.rowmap_k[r] = a.rowmap_k[rg]; mi.rowmap_v[r] = a.rowmap_v[rg];
            mi.own[r] = a.rowmap_k[rg] - (int) (a.koff_self[rg] / a.d);
        }
    }
    unsigned long long bar_target = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) mbar_init(smem_addr(&mi.tma_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // compute cursor, issue cursor (kSlots jobs ahead once the pipeline is primed)
    Cursor cur, iss;
    cursor_seek(mi, geo, cur, 0);
    iss = cur;
    int q_cur = 0, q_iss = 0;                              // running job numbers: job q lives in ring slot q % kSlots
    unsigned tma_par = 0;                                  // per slot: parity of the next TMA completion to wait for
    for (int i = 0; i < kSlots; ++i) {
        fetch_job(mi, a, geo, iss, ring + (q_iss % kSlots) * a.slot_bytes, q_iss % kSlots);
        cp_async_commit();
        cursor_advance(mi, a, geo, iss); ++q_iss;
    }

    // state that lives across the sub-jobs of one job
    float acc[2][4];
    uint32_t qb[8];
    float pv_acc[4];
    StatF st_tx{-INFINITY, 0x7fffffff, 0.0f}, st_ts{-INFINITY, 0x7fffffff, 0.0f};  // sampler partials of this thread's (row, column class)
    StatF st_tx1 = st_tx, st_ts1 = st_ts;                                          // second set: even / odd elements of a job
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) qb[i] = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) pv_acc[i] = 0.0f;

    for (int ph = 0; ph < a.n_phases; ++ph) {
        const int type = mi.ph[ph].type;
        bool deferred_issue = false;
        if (cur.ph == ph) {
            const StepPhase & P = mi.ph[ph];
            const bool is_gemm = type == STEP_GEMM;
            if (is_gemm) {
                // ---- stage the activation operand: rows 0..n-1 as f16 [8 nt_count][K + 32], zero rows above n ----
                const int K = P.K, Ks = K + (P.tma ? 8 : 32);       // ldmatrix wants rows 16 bytes apart mod 128, the 16-byte fragment loads 64
                if (P.src_ln) {
                    const int per_lane4 = d >> 7;
                    const int r0 = warp, r1 = warp + kWarps;
                    const bool emb = P.src_ln == 2;
                    LnSrc s0{x32 + (int64_t) r0 * d, nullptr, nullptr}, s1{x32 + (int64_t) r1 * d, nullptr, nullptr};
                    if (emb) {
                        if (r0 < n) { s0.te = a.te + (int64_t) a.token[row0 + r0] * d; s0.pe = a.pe + (int64_t) a.pos[row0 + r0] * d; }
                        if (r1 < n) { s1.te = a.te + (int64_t) a.token[row0 + r1] * d; s1.pe = a.pe + (int64_t) a.pos[row0 + r1] * d; }
                    }
                    const bool wr = emb && geo.cta == 0;           // one CTA stores the embedding as the residual stream
                    if (r0 < n) ln_rows2(s0, s1, true, r1 < n, wr ? x32 + (int64_t) r0 * d : nullptr, wr ? x32 + (int64_t) r1 * d : nullptr,
                                         P.g, P.b, xs + (int64_t) r0 * Ks, xs + (int64_t) r1 * Ks, d, inv_d, per_lane4, a.eps, lane);
                    if (r0 >= n && r0 < nt_count * 8) for (int i = lane; i < K; i += 32) xs[(int64_t) r0 * Ks + i] = __float2half_rn(0.0f);
                    if (r1 >= n && r1 < nt_count * 8) for (int i = lane; i < K; i += 32) xs[(int64_t) r1 * Ks + i] = __float2half_rn(0.0f);
                } else {
                    const int kc8 = K >> 3, total = nt_count * 8 * kc8;
                    for (int i = threadIdx.x; i < total; i += kThreads) {
                        const int r = i / kc8, c = i - r * kc8;
                        const uint4 v = r < n ? __ldcg((const uint4 *) (P.x16 + (int64_t) (row0 + r) * P.x16_ld + c * 8)) : make_uint4(0, 0, 0, 0);
                        *(uint4 *) (xs + (int64_t) r * Ks + c * 8) = v;
                    }
                }
            }
            TRACE(2);
            bool first_job = true;
            long long tw = 0, tc = 0, ti = 0, t_a = clock64();
            while (cur.ph == ph) {
                const int slot_idx = q_cur % kSlots;
                cp_async_wait<kSlots - 1>();
                if (type == STEP_CROSS || (is_gemm && P.tma)) {
                    mbar_wait(smem_addr(&mi.tma_bar[slot_idx]), (tma_par >> slot_idx) & 1u);
                    tma_par ^= 1u << slot_idx;
                }
                __syncthreads();
                { const long long t_b = clock64(); tw += t_b - t_a; t_a = t_b; }
                if (first_job) TRACE(3);
                uint8_t * slot = ring + slot_idx * a.slot_bytes;

                if (is_gemm) {
                    // ---- 16*tj weight rows, columns [sub*kc, +kc), against the staged rows ----
                    const int tj = P.tj, kz = kWarps / tj, kc = P.kc, KsW = kc + 32, KsX = P.K + 32;
                    const int tile = warp / kz, ks = warp - tile * kz;
                    const int m_tile = cur.j * tj + tile;
                    const bool live = m_tile * 16 < P.M;
                    if (P.epi == EPI_LOGITS && threadIdx.x < tj * 16) {
                        const int m = cur.j * tj * 16 + threadIdx.x;          // consumed after the __syncthreads of the reduction below
                        mi.cls_job[threadIdx.x] = m < V ? __ldg(a.cls + m) : (uint8_t) 1;
                    }
                    if (cur.sub == 0) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
                    }
                    if (live && P.tma) {
                        // swizzled boxes [K / 64][16 tj rows][128 bytes]; A and B fragments through ldmatrix
                        const int rows = tj * 16, KsL = P.K + 8;
                        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = lane >> 4;      // A: row / 16-byte piece of the 8x8 matrices
                        const int arow = tile * 16 + lrow;
                        const uint8_t * abase = slot + arow * 128;
                        const __half * bbase = xs + (int64_t) ((lane & 7) + ((lane >> 4) << 3)) * KsL + ((lane >> 3) & 1) * 8;   // B: rows 0..15 of the activations
                        const int n_steps = P.K >> 4;
                        for (int st = ks; st < n_steps; st += kz) {
                            uint32_t af[4], bf[4];
                            ldmatrix_x4(af, abase + (st >> 2) * rows * 128 + ((((st & 3) * 2 + lcol) ^ (arow & 7)) << 4));
                            ldmatrix_x4(bf, bbase + 16 * st);
                            mma_16816(acc[0], af[0], af[1], af[2], af[3], bf[0], bf[1]);
                            if (nt_count > 1) mma_16816(acc[1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
                        }
                    } else if (live) {
                        const int k_lo = cur.sub * kc;
                        const __half * w0 = (const __half *) slot + (int64_t) (tile * 16 + g8) * KsW + 8 * t4;
                        const __half * w1 = w0 + 8 * KsW;
                        const __half * xb = xs + (int64_t) g8 * KsX + k_lo + 8 * t4;
                        const int n_blk = min(kc, P.K - k_lo) >> 5;
                        for (int b = ks; b < n_blk; b += kz) {
                            const uint4 alo = *(const uint4 *) (w0 + 32 * b);
                            const uint4 ahi = *(const uint4 *) (w1 + 32 * b);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
                                if (nt < nt_count) {
                                    const uint4 bb = *(const uint4 *) (xb + (int64_t) nt * 8 * KsX + 32 * b);
                                    // a lane's 8 consecutive k are split 4 + 4 over two MMAs; A and B use the same permutation of k
                                    mma_16816(acc[nt], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
                                    mma_16816(acc[nt], alo.z, ahi.z, alo.w, ahi.w, bb.z, bb.w);
                                }
                            }
                        }
                    }
                    if (cur.sub == P.ksplit - 1) {
                        // partial accumulators -> shared memory as [ks * tj + tile][row][17], then one output per thread
                        {
                            float * mine = mi.red + (ks * tj + tile) * (16 * 17);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    mine[(nt * 8 + 2 * t4 + (c & 1)) * 17 + g8 + ((c & 2) ? 8 : 0)] = acc[nt][c];
                        }
                        __syncthreads();
                        const int row0 = cur.j * tj * 16;
                        const bool logits_phase = P.epi == EPI_LOGITS;
                        const int row = threadIdx.x >> 4;                           // 16 threads per activation row, for every job
                        if (logits_phase) {
                            // ---- logits: host rows are stored; sampled rows get the rules applied and feed the running statistics.
                            // The (up to 8) elements of a thread are handled in an unrolled loop with two sets of accumulators, so
                            // that their dependency chains overlap (8 warps per SM hide little latency by themselves). ----
                            const int ws_row = row < n ? mi.wslot[row] : 0x7fffffff;
                            const int rule0 = mi.rule[row][0], rule1 = mi.rule[row][1], rule2 = mi.rule[row][2];
                            const bool to_host = ws_row < a.n_full;
                            const int M = P.M, beg = a.token_beg;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (i < tj) {
                                    const int ml = (threadIdx.x & 15) + 16 * i;
                                    const int m = row0 + ml;
                                    const float * src = mi.red + i * (16 * 17) + row * 17 + (ml & 15);
                                    float v = src[0];
                                    for (int k2 = 1; k2 < kz; ++k2) v += src[k2 * tj * (16 * 17)];
                                    if (m < M && row < n) {
                                        if (to_host) a.logits[(int64_t) ws_row * V + m] = v;
                                        else if (!token_masked(m, rule0, (int) mi.cls_job[ml], beg, a.token_eot, rule1, rule2)) {
                                            if (m >= beg) stat_add((i & 1) ? st_ts1 : st_ts, v, m); else stat_add((i & 1) ? st_tx1 : st_tx, v, m);
                                        }
                                    }
                                }
                            }
                        } else {
#pragma unroll 2
                        for (int i = 0; i < tj; ++i) {
                            const int ml = (threadIdx.x & 15) + 16 * i;
                            const int m = row0 + ml;
                            float v = 0.0f;
                            {
                                const float * src = mi.red + i * (16 * 17) + row * 17 + (ml & 15);
                                for (int k2 = 0; k2 < kz; ++k2) v += src[k2 * tj * (16 * 17)];
                            }
                            if (m >= P.M || row >= n) continue;
                            switch (P.epi) {
                                case EPI_QKV: {
                                    const int seg = m / d, mseg = m - seg * d;
                                    if (seg == 0) {
                                        v = __fmul_rn(__fadd_rn(v, __ldg(P.bias + m)), a.qscale);
                                        q16[(int64_t) row * d + mseg] = __float2half_rn(v);
                                    } else if (seg == 1) {
                                        v = __fmul_rn(v, a.qscale);
                                        a.self_k[(int64_t) P.layer * a.kv_cells * d + (int64_t) mi.rowmap_k[row] * d + mseg] = __float2half_rn(v);
                                    } else {
                                        v = __fadd_rn(v, __ldg(P.bias + m));
                                        a.self_v[(int64_t) P.layer * d * a.kv_cells + (int64_t) mseg * a.kv_cells + mi.rowmap_v[row]] = __float2half_rn(v);
                                    }
                                } break;
                                case EPI_RESID: {
                                    float * px = x32 + (int64_t) row * d + m;
                                    *px = __fadd_rn(__fadd_rn(v, __ldg(P.bias + m)), __ldcg(px));
                                } break;
                                case EPI_Q:
                                    v = __fmul_rn(__fadd_rn(v, __ldg(P.bias + m)), a.qscale);
                                    q16[(int64_t) row * d + m] = __float2half_rn(v);
                                    break;
                                default:     // EPI_FC1
                                    v = gelu_table(a.gelu_lut, __fadd_rn(v, __ldg(P.bias + m)));
                                    h16[(int64_t) row * (4 * d) + m] = __float2half_rn(v);
                                    break;
                            }
                        }
                        }
                        (void) logits_phase;
                    }
                } else {
                    // ---- attention item (row r, head hh), sub-job = one key chunk of K (scores) or of V^T (P V) ----
                    const int r = cur.j / a.n_head, hh = cur.j - r * a.n_head;
                    const bool self = type == STEP_SELF;
                    const int il = P.layer, kck = self ? geo.kc_keys : geo.kc_cross;
                    const int n_keys = self ? geo.n_kv : a.n_audio_ctx;
                    const int nc = self ? geo.nc_self : geo.nc_cross;
                    const bool is_v = cur.sub >= nc;
                    const int k0 = (is_v ? cur.sub - nc : cur.sub) * kck;
                    __half * chunk = (__half *) slot;
                    // the cell this row wrote in this step was fetched before it existed: patch it from global
                    const int own = self ? mi.own[r] : -1;
                    if (self && own >= k0 && own < k0 + kck) {
                        if (!is_v) {
                            if (threadIdx.x < 8) {
                                const __half * src = a.self_k + (int64_t) il * a.kv_cells * d + mi.koff_self[r] + (int64_t) own * d + hh * 64 + 8 * threadIdx.x;
                                *(uint4 *) (chunk + (int64_t) (own - k0) * kKRow + 8 * threadIdx.x) = __ldcg((const uint4 *) src);
                            }
                        } else if (threadIdx.x < 64) {
                            const __half * src = a.self_v + (int64_t) il * d * a.kv_cells + mi.voff_self[r] + (int64_t) (hh * 64 + threadIdx.x) * a.kv_cells + own;
                            chunk[(int64_t) threadIdx.x * (kck + 8) + (own - k0)] = __ushort_as_half(__ldcg((const unsigned short *) src));
                        }
                        __syncthreads();
                    }
                    // lane -> row / column of the 8x8 matrices ldmatrix.x4 delivers as the A fragment of m16n8k16
                    const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_col = (lane >> 4) * 8;
                    if (!is_v) {
                        // scores = K q on the tensor cores: A = 16 keys x 64, B = q in column 0 (other columns zero)
                        if (cur.sub == 0) {
                            const __half * qp = q16 + (int64_t) r * d + hh * 64 + 2 * t4;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                qb[2 * ks]     = g8 == 0 ? __ldcg((const unsigned int *) (qp + 16 * ks))     : 0u;
                                qb[2 * ks + 1] = g8 == 0 ? __ldcg((const unsigned int *) (qp + 16 * ks + 8)) : 0u;
                            }
                        }
                        const int k1 = min(n_keys, k0 + kck);
                        const int n_tiles = (k1 - k0 + 15) >> 4;
                        for (int tk = warp; tk < n_tiles; tk += kWarps) {
                            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                            uint32_t af[4][4];
                            const int row = tk * 16 + lm_row;
                            if (self) {
                                // padded rows (kKRow halves apart)
                                const __half * arow = chunk + (int64_t) row * kKRow + lm_col;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(af[ks], arow + 16 * ks);
                            } else {
                                // TMA box layout: 128-byte rows, 16-byte piece j of row r stored at piece j ^ (r & 7)
                                const uint8_t * rbase = (const uint8_t *) chunk + (int64_t) row * 128;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(af[ks], rbase + (((2 * ks + (lane >> 4)) ^ (row & 7)) << 4));
                            }
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_16816(c, af[ks][0], af[ks][1], af[ks][2], af[ks][3], qb[2 * ks], qb[2 * ks + 1]);
                            if (t4 == 0) {
                                const int j0 = k0 + tk * 16 + g8;
                                if (j0 < k1)     mi.sc[j0] = c[0];
                                if (j0 + 8 < k1) mi.sc[j0 + 8] = c[2];
                            }
                        }
                        if (cur.sub == nc - 1) {
                            // softmax over all keys (ggml.c:11116-11201): global max, table exp, f64 sum, p rounded to f16.
                            // kKeysCap / kThreads = 6 scores per thread; mask and table look-ups are issued together before their first use
                            constexpr int kPer = kKeysCap / kThreads;
                            const float * mrow = self ? a.mask + (int64_t) (row0 + r) * a.ld_mask : nullptr;
                            __syncthreads();                                   // scores of every warp are in shared memory
                            float sv[kPer];
                            float mx = -INFINITY;
#pragma unroll
                            for (int u = 0; u < kPer; ++u) {
                                const int jk = threadIdx.x + kThreads * u;
                                float v = -INFINITY;
                                if (jk < n_keys) { v = mi.sc[jk]; if (mrow) v = __fadd_rn(v, __ldg(mrow + jk)); }
                                sv[u] = v;
                                mx = fmaxf(mx, v);
                            }
                            mx = warp_max(mx);
                            if (lane == 0) mi.redf[warp] = mx;
                            __syncthreads();
                            mx = mi.redf[0];
#pragma unroll
                            for (int i = 1; i < kWarps; ++i) mx = fmaxf(mx, mi.redf[i]);
                            double sum = 0.0;
                            float ev[kPer];
#pragma unroll
                            for (int u = 0; u < kPer; ++u) ev[u] = sv[u] != -INFINITY ? exp_table(a.exp_lut, __fsub_rn(sv[u], mx)) : 0.0f;
#pragma unroll
                            for (int u = 0; u < kPer; ++u) sum += (double) ev[u];
                            sum = warp_sum(sum);
                            if (lane == 0) mi.redd[warp] = sum;
                            __syncthreads();
                            sum = 0.0;
#pragma unroll
                            for (int i = 0; i < kWarps; ++i) sum += mi.redd[i];
                            const float inv = (float) (1.0 / sum);
                            const int n_pad = (n_keys + 15) & ~15;
#pragma unroll
                            for (int u = 0; u < kPer; ++u) {
                                const int jk = threadIdx.x + kThreads * u;
                                if (jk < n_pad) mi.p16[jk] = __float2half_rn(jk < n_keys ? __fmul_rn(ev[u], inv) : 0.0f);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) pv_acc[i] = 0.0f;
                        }
                    } else {
                        // P V on the tensor cores: A = V^T (64 features x keys), B = p in column 0; warp w owns features
                        // 16 (w & 3) .. +15 and every second step of 16 keys
                        const int n_pad = (n_keys + 7) & ~7;
                        const int k1 = min(n_pad, k0 + kck);
                        const int n_steps = (k1 - k0 + 15) >> 4;
                        const int vrow = kck + 8;
                        const int frow = (warp & 3) * 16 + lm_row;                  // feature row of this lane's ldmatrix address
                        const __half * pp = mi.p16 + k0 + 2 * t4;
                        float acc2[4] = {0.0f, 0.0f, 0.0f, 0.0f};                   // second accumulator: two independent mma chains
                        for (int st = warp >> 2; st < n_steps; st += 4) {
                            uint32_t af0[4], af1[4];
                            const int st1 = st + 2;
                            if (self) {
                                const __half * arow = chunk + (int64_t) frow * vrow + lm_col;
                                ldmatrix_x4(af0, arow + 16 * st);
                                if (st1 < n_steps) ldmatrix_x4(af1, arow + 16 * st1);
                            } else {
                                // boxes of 64 keys x 64 features (8 KB), rows of 128 bytes, 16-byte pieces swizzled by the row
                                const uint8_t * b0p = (const uint8_t *) chunk + (st >> 2) * 8192 + frow * 128;
                                ldmatrix_x4(af0, b0p + (((2 * (st & 3) + (lane >> 4)) ^ (frow & 7)) << 4));
                                if (st1 < n_steps) {
                                    const uint8_t * b1p = (const uint8_t *) chunk + (st1 >> 2) * 8192 + frow * 128;
                                    ldmatrix_x4(af1, b1p + (((2 * (st1 & 3) + (lane >> 4)) ^ (frow & 7)) << 4));
                                }
                            }
                            const uint32_t b0 = g8 == 0 ? *(const uint32_t *) (pp + 16 * st)     : 0u;
                            const uint32_t b1 = g8 == 0 ? *(const uint32_t *) (pp + 16 * st + 8) : 0u;
                            mma_16816(pv_acc, af0[0], af0[1], af0[2], af0[3], b0, b1);
                            if (st1 < n_steps) {
                                const uint32_t c0 = g8 == 0 ? *(const uint32_t *) (pp + 16 * st1)     : 0u;
                                const uint32_t c1 = g8 == 0 ? *(const uint32_t *) (pp + 16 * st1 + 8) : 0u;
                                mma_16816(acc2, af1[0], af1[1], af1[2], af1[3], c0, c1);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) pv_acc[i] += acc2[i];
                        if (cur.sub == 2 * nc - 1) {
                            // the two warps of a feature tile add their halves of the key range
                            if (t4 == 0 && warp >= 4) { mi.redf2[(warp & 3) * 16 + g8] = pv_acc[0]; mi.redf2[(warp & 3) * 16 + g8 + 8] = pv_acc[2]; }
                            __syncthreads();
                            if (t4 == 0 && warp < 4) {
                                const int f = warp * 16 + g8;
                                attn16[(int64_t) r * d + hh * 64 + f]     = __float2half_rn(pv_acc[0] + mi.redf2[f]);
                                attn16[(int64_t) r * d + hh * 64 + f + 8] = __float2half_rn(pv_acc[2] + mi.redf2[f + 8]);
                            }
                        }
                    }
                }
                if (first_job) TRACE(4);
                first_job = false;
                cursor_advance(mi, a, geo, cur); ++q_cur;
                if (cur.ph == ph) {
                    // more work in this phase: refill the slot just consumed right away
                    __syncthreads();
                    { const long long t_b = clock64(); tc += t_b - t_a; t_a = t_b; }
                    fetch_job(mi, a, geo, iss, ring + (q_iss % kSlots) * a.slot_bytes, q_iss % kSlots);
                    cp_async_commit();
                    cursor_advance(mi, a, geo, iss); ++q_iss;
                    { const long long t_b = clock64(); ti += t_b - t_a; t_a = t_b; }
                } else {
                    deferred_issue = true;              // last job of the phase: arrive at the barrier first, prefetch afterwards
                }
            }
            if (a.trace && threadIdx.x == 0) {
                unsigned long long * tr = a.trace + ((int64_t) blockIdx.x * kStepMaxPhases + ph) * 8;
                tr[5] = (unsigned long long) tw; tr[6] = (unsigned long long) tc; tr[7] = (unsigned long long) ti;
            }
        }
        if (ph == a.n_phases - 1) {
            // ---- every CTA publishes its sampler partials; the CTA that publishes last finalizes all rows ----
            {
                Stat tx{st_tx.m, st_tx.i, (double) st_tx.s}, ts{st_ts.m, st_ts.i, (double) st_ts.s};
                tx = stat_merge(tx, Stat{st_tx1.m, st_tx1.i, (double) st_tx1.s}); ts = stat_merge(ts, Stat{st_ts1.m, st_ts1.i, (double) st_ts1.s});
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { tx = stat_merge(tx, stat_shfl_xor(tx, o)); ts = stat_merge(ts, stat_shfl_xor(ts, o)); }
                if ((threadIdx.x & 15) == 0) {
                    double * rec = records + ((int64_t) geo.cta * kStepMaxRows + (threadIdx.x >> 4)) * 6;
                    rec[0] = (double) tx.m; rec[1] = (double) tx.i; rec[2] = tx.s;
                    rec[3] = (double) ts.m; rec[4] = (double) ts.i; rec[5] = ts.s;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                mi.ticket = (int) atomicAdd(bar + 1, 1ull);
            }
            __syncthreads();
            if (mi.ticket == geo.n_cta - 1) {
                for (int r = warp; r < n; r += kWarps) finalize_row(a, records, r, row0 + r, geo.n_cta, lane);
                // every CTA of the group is past its last barrier: the arrival counter and the ticket start the next launch at zero
                if (threadIdx.x == 0) { bar[0] = 0; bar[1] = 0; }
            }
            TRACE(0);
        }
        if (ph + 1 < a.n_phases) {
            TRACE(0);
            barrier_arrive(bar);                // (__syncthreads inside: the slot just consumed is free from here on)
            if (deferred_issue) {
                fetch_job(mi, a, geo, iss, ring + (q_iss % kSlots) * a.slot_bytes, q_iss % kSlots);
                cp_async_commit();
                cursor_advance(mi, a, geo, iss); ++q_iss;
            }
            bar_target += (unsigned long long) geo.n_cta;
            barrier_wait(bar, bar_target);
            TRACE(1);
        }
    }
    cp_async_wait<0>();
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------------------------------

size_t decode_step_smem_bytes(int d, int * xs_bytes, int * slot_bytes, int * chunk_keys) {
    if (d <= 0 || (d & 127) || d > 128 * kLnMax4) return 0;
    const int k_max = 4 * d;
    const int xs = kStepMaxRows * (k_max + 32) * 2;
    const int total = 227 * 1024;
    const int misc = (int) ((sizeof(Misc) + 127) & ~(size_t) 127);
    int slot = (total - xs - misc - 128) / kSlots;
    slot &= ~1023;
    // a 16-row block of the narrowest map (K = d) and a 64-key attention chunk must fit one slot
    if (slot < 16 * (d + 32) * 2 || slot < 64 * 128) return 0;
    *xs_bytes = xs; *slot_bytes = slot; *chunk_keys = std::min(slot / (kKRow * 2), slot / 128 - 8) & ~15;
    return (size_t) xs + (size_t) kSlots * slot + misc;
}

int decode_step_grid(size_t smem_bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || !prop.cooperativeLaunch) return 0;
    if (cudaFuncSetAttribute(k_decode_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_step, kThreads, smem_bytes) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    return prop.multiProcessorCount;
}

int decode_step_plan(const StepLayerW * L, int n_layer, int d, int n_head, int n_vocab, const __half * te, const float * ln_g,
                     const float * ln_b, const __half * attn16, const __half * h16, int n, int grid, int slot_bytes, StepPhase * out) {
    if (1 + 8 * n_layer > kStepMaxPhases) return 0;
    int np = 0;
    auto gemm = [&](int layer, int epi, const __half * W, int M, int K, const float * g, const float * b, const __half * x16, int x16_ld,
                    const float * bias) {
        StepPhase p;
        p.type = STEP_GEMM; p.epi = epi; p.layer = layer; p.W = W; p.M = M; p.K = K; p.bias = bias;
        p.src_ln = g != nullptr; p.g = g; p.b = b; p.x16 = x16; p.x16_ld = x16_ld;
        // column slabs: a 16-row block must fit one slot
        p.ksplit = 1;
        while (K % p.ksplit != 0 || (K / p.ksplit) % 32 != 0 || 16 * (K / p.ksplit + 32) * 2 > slot_bytes) ++p.ksplit;
        p.kc = K / p.ksplit;
        // rows per job: the smallest tj that gives every CTA at most one job, bounded by the slot size
        const int n_tiles = (M + 15) / 16;
        p.tj = 1;
        for (int tj = 1; tj <= 8; tj *= 2) {
            if (tj * 16 * (p.kc + 32) * 2 > slot_bytes) break;
            p.tj = tj;
            if ((n_tiles + tj - 1) / tj <= grid) break;
        }
        p.n_jobs = (n_tiles + p.tj - 1) / p.tj;
        out[np++] = p;
    };
    auto attn = [&](int layer, int type) {
        StepPhase p;
        p.type = type; p.layer = layer; p.n_jobs = n * n_head;
        out[np++] = p;
    };
    for (int il = 0; il < n_layer; ++il) {
        gemm(il, EPI_QKV,   L[il].wqkv, 3 * d, d, L[il].ln1_g, L[il].ln1_b, nullptr, 0, L[il].bqkv);
        attn(il, STEP_SELF);
        gemm(il, EPI_RESID, L[il].wo,   d, d, nullptr, nullptr, attn16, d, L[il].bo);
        gemm(il, EPI_Q,     L[il].wcq,  d, d, L[il].lnc_g, L[il].lnc_b, nullptr, 0, L[il].bcq);
        attn(il, STEP_CROSS);
        gemm(il, EPI_RESID, L[il].wco,  d, d, nullptr, nullptr, attn16, d, L[il].bco);
        gemm(il, EPI_FC1,   L[il].w1,   4 * d, d, L[il].ln2_g, L[il].ln2_b, nullptr, 0, L[il].b1);
        gemm(il, EPI_RESID, L[il].w2,   d, 4 * d, nullptr, nullptr, h16, 4 * d, L[il].b2);
    }
    gemm(0, EPI_LOGITS, te, n_vocab, d, ln_g, ln_b, nullptr, 0, nullptr);
    {
        // the logits weights stream through TMA boxes: no padding, the largest block of 16 tj rows that fits a slot
        StepPhase & p = out[np - 1];
        int tj = 4;
        while (tj > 1 && tj * 16 * d * 2 > slot_bytes) tj >>= 1;
        if (tj * 16 * d * 2 <= slot_bytes && (d & 63) == 0) {
            p.tma = 1; p.tj = tj; p.ksplit = 1; p.kc = d;
            p.n_jobs = ((n_vocab + 15) / 16 + tj - 1) / tj;
        }
    }
    out[0].src_ln = 2;      // the first phase normalises the token + positional embedding itself
    return np;
}

bool launch_decode_step(const StepArgs & a, int grid, size_t smem_bytes, cudaStream_t st) {
    if (a.chunk_keys_cross < 128 || (a.chunk_keys_cross & 127) || a.chunk_keys_cross * 128 > a.slot_bytes) return false;
    if (a.n_groups < 1 || a.n_groups > kStepMaxGroups) return false;
    for (int g = 0; g < (a.n_groups > 1 ? a.n_groups : 0); ++g) if (a.n_grp[g] < 1 || a.n_grp[g] > kStepMaxRows || !a.phases_grp[g]) return false;
    if (a.n < 1 || (a.n_groups == 1 && a.n > kStepMaxRows) || a.n_audio_ctx > kKeysCap || a.kv_cells > kKeysCap || (a.d & 63) || a.n_phases < 2 ||
        a.n_phases > kStepMaxPhases) return false;
    void * args[] = { (void *) &a };
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *) k_decode_step, dim3(grid), dim3(kThreads), args, smem_bytes, st);
    if (e != cudaSuccess) {
        fprintf(stderr, "whisper_b200: cooperative launch of k_decode_step failed: %s\n", cudaGetErrorString(e));
        return false;
    }
    return true;
}

}  // namespace wb200
