// Shared helpers of libgsn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gsn_b200.h"

namespace gsn {

extern thread_local char g_last_error[256];
extern unsigned long long g_launches;   // kernels this library has launched (gsn_launch_count)
#define GSN_BUMP(n) (__atomic_fetch_add(&gsn::g_launches, (unsigned long long)(n), __ATOMIC_RELAXED))

inline int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return GSN_E_CUDA;
}

#define GSN_CUDA_OK(expr)                                          \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return gsn::cuda_fail(_e, #expr);   \
    } while (0)

#define GSN_LAUNCH_OK(name)                                        \
    do {                                                           \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return gsn::cuda_fail(_e, name);    \
    } while (0)

constexpr int kNumSMs = 148;   // B200

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device-side layout of the batched graph workspace (gsn_graph_build) ----
struct GraphLayout {
    size_t nbase, adj, rowptr, slot_src, slot_dst, slot_col, scan_tmp, total;
    int64_t slot_cap;
};

inline GraphLayout graph_layout(int64_t N, int64_t E, int W) {
    GraphLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.slot_cap = 2 * E;
    L.nbase = take(sizeof(int32_t) * (size_t)(N + 4));
    L.adj = take(sizeof(uint64_t) * ((size_t)N * W + 4));
    L.rowptr = take(sizeof(int32_t) * (size_t)(N + 8));
    L.slot_src = take(sizeof(int32_t) * (size_t)(L.slot_cap + 4));
    L.slot_dst = take(sizeof(int32_t) * (size_t)(L.slot_cap + 4));
    L.slot_col = take(sizeof(int32_t) * (size_t)(L.slot_cap + 4));
    L.scan_tmp = take(sizeof(int32_t) * (size_t)(ceil_div(N + 1, 1024) + 8));
    L.total = off;
    return L;
}

// ---- exclusive scan of int32 (three small kernels; n up to 2^31) -------------
int exclusive_scan_i32(const int32_t *d_in, int32_t *d_out, int64_t n, int32_t *d_tmp, cudaStream_t stream);

#if defined(__CUDACC__)
// ---- mbarrier / bulk-copy (TMA 1-D) wrappers ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace gsn
