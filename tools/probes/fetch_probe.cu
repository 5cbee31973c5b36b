// Micro-benchmark: how fast can ONE CTA (and many CTAs) stream global memory into shared memory on sm_100a?
//   mode 0: cp.async 16 B per thread (LDGSTS), contiguous chunks          mode 1: cp.async 16 B, 128-byte rows with a 768-byte stride
//   mode 2: cp.async.bulk (TMA, 1-D) one instruction per chunk            mode 3: cp.async.bulk, one 128-byte copy per row (strided)
// Chunks of CH bytes go round a ring of NS slots; DEPTH = NS - 1 chunks are in flight while one is "consumed" (a few smem reads).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sa(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t ph) {
    uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(ph) : "memory"); return ok != 0;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void * src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }

template <int NS>
__global__ void __launch_bounds__(256, 1) k_probe(const uint8_t * src, size_t per_cta_bytes, int CH, int n_chunks, int mode, float * sink, long long * cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[NS];
    const uint8_t * base = src + (size_t) blockIdx.x * per_cta_bytes;
    if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) mbar_init(sa(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](int c) {
        uint8_t * dst = smem + (size_t) (c % NS) * CH;
        const uint8_t * s = base + (size_t) c * CH * (mode == 1 || mode == 3 ? 6 : 1);
        if (mode == 0) {
            for (int i = threadIdx.x * 16; i < CH; i += 256 * 16) cpa16(sa(dst + i), s + i);
        } else if (mode == 1) {
            for (int i = threadIdx.x * 16; i < CH; i += 256 * 16) { const int row = i >> 7, col = i & 127; cpa16(sa(dst + i), s + (size_t) row * 768 + col); }
        } else if (mode == 2) {
            if (threadIdx.x == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect(sa(&bars[c % NS]), CH); bulk(sa(dst), s, CH, sa(&bars[c % NS])); }
        } else {
            if (threadIdx.x == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect(sa(&bars[c % NS]), CH); }
            __syncthreads();
            for (int row = threadIdx.x; row < CH / 128; row += 256) bulk(sa(dst + row * 128), s + (size_t) row * 768, 128, sa(&bars[c % NS]));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float acc = 0.f;
    const long long t0 = clock64();
    for (int c = 0; c < NS - 1 && c < n_chunks; ++c) issue(c);
    for (int c = 0; c < n_chunks; ++c) {
        if (c + NS - 1 < n_chunks) issue(c + NS - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        if (mode < 2) { asm volatile("cp.async.wait_group %0;" :: "n"(NS - 1) : "memory"); }
        else { while (!mbar_try(sa(&bars[c % NS]), (c / NS) & 1)) {} }
        __syncthreads();
        const float * f = (const float *) (smem + (size_t) (c % NS) * CH);
        acc += f[threadIdx.x] + f[(CH / 4) - 1 - threadIdx.x];
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; }
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
    const size_t total = (size_t) 2 << 30;
    uint8_t * src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
    float * sink; cudaMalloc(&sink, 256 * 256 * 4);
    long long * cyc; cudaMalloc(&cyc, 256 * 8);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char * names[4] = {"cp.async contiguous", "cp.async 128B rows stride 768", "bulk 1 per chunk", "bulk 128B per row stride 768"};
    for (int ns = 3; ns <= 4; ++ns)
    for (int CH : {16384, 32768, 49152}) {
        if ((size_t) ns * CH > 200 * 1024) continue;
        for (int mode = 0; mode < 4; ++mode) {
            for (int ctas : {1, 6, 96, 148}) {
                const int n_chunks = 64;
                const size_t per_cta = (size_t) n_chunks * CH * 6 + 4096;
                if (per_cta * ctas > total) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    if (ns == 3) k_probe<3><<<ctas, 256, (size_t) ns * CH>>>(src, per_cta, CH, n_chunks, mode, sink, cyc);
                    else         k_probe<4><<<ctas, 256, (size_t) ns * CH>>>(src, per_cta, CH, n_chunks, mode, sink, cyc);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                    cudaError_t err = cudaGetLastError();
                    if (rep == 1) {
                        const double bytes = (double) n_chunks * CH;
                        printf("slots %d chunk %5d KB  %-32s ctas %3d : %7.1f us  per-CTA %6.1f GB/s  total %7.1f GB/s  %s\n", ns, CH / 1024, names[mode], ctas,
                               ms * 1e3, bytes / (ms * 1e-3) / 1e9, bytes * ctas / (ms * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
                    }
                }
            }
        }
    }
    return 0;
}
