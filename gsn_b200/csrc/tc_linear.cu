// Tensor-core version of the dense tail (gsn_linear_fwd semantics) for sm_100a:
// tcgen05.mma kind::tf32 with the accumulator in TMEM, operands staged by TMA
// (cp.async.bulk.tensor, SWIZZLE_128B) through a 3-stage mbarrier ring.
//
// fp32 parity (1e-5, BASELINE.json) rules out plain TF32 (10-bit mantissa), so
// every product is evaluated as 3xTF32:
//     a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi ,  x_hi = rn_tf32(x), x_lo = rn_tf32(x - x_hi)
// accumulated in fp32 in TMEM; every dropped term is <= 2^-22 relative and unbiased.
// Weights are split once (cached by the host side); activations are split by
// split_a_kernel into one [M, K1+K2] hi/lo pair (this also performs the
// torch.cat((x, agg)) of graph_filters/GSN_edge_sparse.py:111).
//
// Warp roles (384 threads, one 128 x BN output tile per CTA):
//   warp 0 lane 0 : TMA producer          warp 1 lane 0 : MMA issuer
//   warp 2        : TMEM alloc / dealloc  warps 4..11   : tf32 split of the landed A tile (in shared memory)
//   warps 0..7    : epilogue (tcgen05.ld -> registers -> global)
#include <cuda.h>
#include "common.cuh"

namespace gsn {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;        // 32 fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 8;     // tf32

struct TcEpilogue {
    const float *bias, *row_scale, *row_vec, *tab, *scale, *shift;
    const int32_t *tab_idx;
    float *C;
    int M, Nout, K, ldc, tab_ld, act, accumulate;
};

// elu / tanh are kept out of line: inlining them into the unrolled epilogue makes it tens of KB of straight-line
// code that every warp executes once with a cold instruction cache (measured: 20k cycles of fetch stalls per tile)
__device__ __noinline__ float tc_act_slow(float v, int act) {
    return act == 1 ? (v > 0.0f ? v : expm1f(v)) : tanhf(v);
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(const void *smem_tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);      // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                     // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                           // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                     // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                     // layout: SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// SPLIT_IN_KERNEL = false: tmA_hi / tmA_lo address pre-split activations (split_a_kernel).
// SPLIT_IN_KERNEL = true : tmA_hi / tmA_lo are the RAW fp32 sources A1 [M,K1] and A2 [M,K2] (K1 % 32 == 0 when K2 > 0);
//                          warps 4-7 split every landed tile in shared memory (hi in place, lo next to it) while the
//                          previous stage is being multiplied: no extra kernel, no extra HBM traffic.
template <int BN, int STAGES, bool SPLIT_IN_KERNEL>
__global__ void __launch_bounds__(384, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const __grid_constant__ TcEpilogue ep, int K1) {
    constexpr int A_TILE = TC_BM * TC_BK * 4;      // bytes
    constexpr int W_TILE = BN * TC_BK * 4;
    constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], conv_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int nk = (ep.K + TC_BK - 1) / TC_BK;
    // 1024-byte aligned base of the stage ring (SWIZZLE_128B atoms); derived by offsetting the __shared__ symbol so
    // that the compiler keeps the shared address space (LDS/STS instead of generic LD/ST in the epilogue)
    unsigned char *ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&conv_bar[s], 8);          // one arrival per converter warp (warps 4..11)
        }
        mbar_init(&tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            unsigned char *st = ring + (size_t)s * STAGE;
            const int k0 = kt * TC_BK;
            if (SPLIT_IN_KERNEL) {
                mbar_arrive_expect_tx(&full_bar[s], A_TILE + 2 * W_TILE);
                if (k0 < K1) tma_load_2d(st, &tmA_hi, k0, m0, &full_bar[s]);           // raw A1 tile
                else tma_load_2d(st, &tmA_lo, k0 - K1, m0, &full_bar[s]);             // raw A2 tile
            } else {
                mbar_arrive_expect_tx(&full_bar[s], STAGE);
                tma_load_2d(st, &tmA_hi, k0, m0, &full_bar[s]);
                tma_load_2d(st + A_TILE, &tmA_lo, k0, m0, &full_bar[s]);
            }
            tma_load_2d(st + 2 * A_TILE, &tmW_hi, k0, n0, &full_bar[s]);
            tma_load_2d(st + 2 * A_TILE + W_TILE, &tmW_lo, k0, n0, &full_bar[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(SPLIT_IN_KERNEL ? &conv_bar[s] : &full_bar[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned char *st = ring + (size_t)s * STAGE;
            const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + A_TILE);
            const uint64_t w_hi = umma_desc(st + 2 * A_TILE), w_lo = umma_desc(st + 2 * A_TILE + W_TILE);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);      // 32 bytes per k-step inside the swizzle row
                const uint32_t first = (kt > 0 || k > 0) ? 1u : 0u;
                // The tensor core adds into the fp32 accumulator with truncation, a bias that grows with the
                // number of accumulations: the two correction terms (2^-11 of the main one) therefore get their
                // own accumulator (columns [BN, 2BN)) and are added once, in the epilogue.
                umma_tf32(tmem_d, a_hi + adv, w_hi + adv, idesc, first);
                umma_tf32(tmem_d + BN, a_lo + adv, w_hi + adv, idesc, first);
                umma_tf32(tmem_d + BN, a_hi + adv, w_lo + adv, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);          // frees the stage when these MMAs have read it
        }
        umma_commit(&tmem_full_bar);             // accumulator complete
    } else if (SPLIT_IN_KERNEL && warp >= 4) {
        // ---------------- converters: raw fp32 tile -> (hi, lo) tf32 pair, elementwise in the swizzled layout
        const int tid = threadIdx.x - 128;        // 0..255: eight converter warps keep the split off the critical path
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            float4 *hi = reinterpret_cast<float4 *>(ring + (size_t)s * STAGE);
            float4 *lo = reinterpret_cast<float4 *>(ring + (size_t)s * STAGE + A_TILE);
            // all loads first (the tile is converted in place, so the compiler cannot hoist them itself), then the
            // conversions, then the stores: one shared-memory round trip per stage instead of eight
            constexpr int NV = A_TILE / 16 / 256;
            float4 av[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) av[i] = hi[tid + i * 256];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 a = av[i];
                float4 h, l;
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.x)); h.x = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.y)); h.y = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.z)); h.z = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.w)); h.w = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.x - h.x)); l.x = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.y - h.y)); l.y = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.z - h.z)); l.z = __uint_as_float(u);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.w - h.w)); l.w = __uint_as_float(u);
                hi[tid + i * 256] = h;
                lo[tid + i * 256] = l;
            }
            // generic-proxy writes must be visible to the tensor core (async proxy) before the MMA is issued
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
            }
        }
    }
    __syncwarp();
    if (warp < 8) {
        // ---------------- epilogue (warps 0..7): TMEM lane = output row, one row per thread.
        // Warp w may only touch TMEM lanes [32*(w%4), +32); warps 0-3 take the first half of the columns, warps
        // 4-7 the second half.  Each thread adds the two accumulators, applies the epilogue and writes its row
        // with 128-bit stores straight from registers (the L2 merges the sectors of a row).
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int wq = warp & 3;
        const int half = warp >> 2;
        constexpr int HALF_COLS = BN >= 64 ? BN / 2 : BN;      // BN = 32: one 32-column chunk, warps 0..3 only
        const int m = m0 + wq * 32 + lane;
        const bool mok = m < ep.M;
        const bool accumulate = ep.accumulate == 1;
        const int act = ep.act;
        const float *bias = ep.bias, *row_vec = ep.row_vec, *scale = ep.scale, *shift = ep.shift;
        const float rs = (ep.row_scale && mok) ? __ldg(ep.row_scale + m) : 0.f;
        const float *trow = (ep.tab && mok) ? ep.tab + (int64_t)__ldg(ep.tab_idx + m) * ep.tab_ld : nullptr;
        float *crow = ep.C + (int64_t)(mok ? m : 0) * ep.ldc;
        const bool vec = (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15) == 0);
#pragma unroll 1
        for (int c0 = half * HALF_COLS; c0 < (half + 1) * HALF_COLS && c0 < BN; c0 += 32) {
            uint32_t r[32], q[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                  "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
                  "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]),
                  "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                : "r"(taddr + (uint32_t)BN)
                : "memory");
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int nb = n0 + c0;                    // first column of this chunk
            if (mok && nb < ep.Nout) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(q[j]);
                const bool whole = nb + 32 <= ep.Nout;       // all 32 columns valid: unguarded parameter loads
                if (row_vec) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(rs, (whole || nb + j < ep.Nout) ? __ldg(row_vec + nb + j) : 0.f, v[j]);
                }
                if (trow) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += (whole || nb + j < ep.Nout) ? __ldg(trow + nb + j) : 0.f;
                }
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += (whole || nb + j < ep.Nout) ? __ldg(bias + nb + j) : 0.f;
                }
                if (scale) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= (whole || nb + j < ep.Nout) ? __ldg(scale + nb + j) : 1.f;
                }
                if (shift) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += (whole || nb + j < ep.Nout) ? __ldg(shift + nb + j) : 0.f;
                }
                if (act == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (act != 3) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = tc_act_slow(v[j], act);
                }
                if (accumulate) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (whole || nb + j < ep.Nout) v[j] += crow[nb + j];
                }
                if (vec && whole) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(crow + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < ep.Nout) crow[nb + j] = v[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant for large problems: one CTA per SM loops over output tiles; the TMA -> split -> MMA ring
// runs continuously across tiles and the accumulator is double-buffered in TMEM, so the epilogue of tile i
// overlaps the main loop of tile i+1 (the one-tile kernel above exposes prologue + pipeline fill + epilogue on
// every tile: ~18 k cycles per tile of which ~8 k are MMA work).
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-11: tf32 split   warps 12-15: epilogue
// Activations are always split in the kernel (raw A1 [M,K1] / A2 [M,K2], K1 % 32 == 0 when K2 > 0).
template <int BN, int STAGES>
__global__ void __launch_bounds__(512, 1)
tc_linear_persistent_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                            const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                            const __grid_constant__ CUtensorMap tmC,
                            const __grid_constant__ TcEpilogue ep, int K1, int m_tiles, int n_tiles) {
    constexpr int A_TILE = TC_BM * TC_BK * 4;
    constexpr int W_TILE = BN * TC_BK * 4;
    constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], conv_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (ep.K + TC_BK - 1) / TC_BK;
    const int total_tiles = m_tiles * n_tiles;
    unsigned char *ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char *stage_out = ring + (size_t)STAGES * STAGE;                        // 4 epilogue warps x 4 KB, 1024-aligned
    float *col_const = reinterpret_cast<float *>(stage_out + 4 * 4096);              // [3][BN]: scale | bias*scale+shift | vec*scale

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&conv_bar[s], 8);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)(4 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int m0 = (t / n_tiles) * TC_BM, n0 = (t % n_tiles) * BN;
            // The ring holds 3-4 k-steps (48 KB of DRAM reads in flight per SM); under load an HBM round trip is ~3 k
            // cycles, i.e. ~1.5 k cycles per k-step against 768 of MMA work.  The activation rows of this CTA's NEXT
            // tile are therefore pulled into L2 now (no shared memory needed), one whole tile ahead.
            {
                const int tn = t + (int)gridDim.x;
                if (tn < total_tiles) {
                    const int mn = (tn / n_tiles) * TC_BM;
                    for (int kt = 0; kt < nk; ++kt) {
                        const int k0 = kt * TC_BK;
                        if (k0 < K1)
                            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmA1), "r"(k0), "r"(mn) : "memory");
                        else
                            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmA2), "r"(k0 - K1), "r"(mn) : "memory");
                    }
                }
            }
            for (int kt = 0; kt < nk; ++kt, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                unsigned char *st = ring + (size_t)s * STAGE;
                const int k0 = kt * TC_BK;
                mbar_arrive_expect_tx(&full_bar[s], A_TILE + 2 * W_TILE);
                if (k0 < K1) tma_load_2d(st, &tmA1, k0, m0, &full_bar[s]);
                else tma_load_2d(st, &tmA2, k0 - K1, m0, &full_bar[s]);
                tma_load_2d(st + 2 * A_TILE, &tmW_hi, k0, n0, &full_bar[s]);
                tma_load_2d(st + 2 * A_TILE + W_TILE, &tmW_lo, k0, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        int it = 0, tc = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tc) {
            const int b = tc & 1;
            mbar_wait(&tmem_empty_bar[b], ((tc >> 1) & 1) ^ 1);          // the epilogue has drained this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc_main = tmem_d + (uint32_t)(b * 2 * BN), acc_corr = acc_main + BN;
            for (int kt = 0; kt < nk; ++kt, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&conv_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char *st = ring + (size_t)s * STAGE;
                const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + A_TILE);
                const uint64_t w_hi = umma_desc(st + 2 * A_TILE), w_lo = umma_desc(st + 2 * A_TILE + W_TILE);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                    const uint32_t first = (kt > 0 || k > 0) ? 1u : 0u;
                    umma_tf32(acc_main, a_hi + adv, w_hi + adv, idesc, first);
                    umma_tf32(acc_corr, a_lo + adv, w_hi + adv, idesc, first);
                    umma_tf32(acc_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar[b]);
        }
    } else if (warp >= 4 && warp < 12) {
        // ---------------- converters
        const int tid = threadIdx.x - 128;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            for (int kt = 0; kt < nk; ++kt, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                float4 *hi = reinterpret_cast<float4 *>(ring + (size_t)s * STAGE);
                float4 *lo = reinterpret_cast<float4 *>(ring + (size_t)s * STAGE + A_TILE);
                constexpr int NV = A_TILE / 16 / 256;
                float4 av[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) av[i] = hi[tid + i * 256];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float4 a = av[i];
                    float4 h, l;
                    uint32_t u;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.x)); h.x = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.y)); h.y = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.z)); h.z = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.w)); h.w = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.x - h.x)); l.x = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.y - h.y)); l.y = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.z - h.z)); l.z = __uint_as_float(u);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a.w - h.w)); l.w = __uint_as_float(u);
                    hi[tid + i * 256] = h;
                    lo[tid + i * 256] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
            }
        }
    } else if (warp >= 12) {
        // ---------------- epilogue.  Thread = output row (TMEM lane), 32 columns per pass.  The per-column operands are
        // folded once per n-tile into three shared-memory vectors (broadcast LDS instead of 4 predicated LDG per
        // element), the finished 32 x 32 block is written to a SWIZZLE_128B staging buffer owned by this warp and
        // leaves through a TMA store (or reduce-add for `accumulate`): full-line writes, no strided 16-byte stores,
        // ragged M / Nout clipped by the tensor map.  Measured before this rewrite: the epilogue bounded the kernel
        // (1188 us with it, 519 us without, M = 3.04 M, K = N = 128).
        const int wq = warp & 3;
        const int te = threadIdx.x - 384;                     // 0..127 among the epilogue threads
        const bool accumulate = ep.accumulate == 1;
        const int act = ep.act;
        unsigned char *stg = stage_out + wq * 4096;
        const uint32_t stg_row = smem_u32(stg) + (uint32_t)lane * 128u;
        float *cS = col_const, *cB = col_const + BN, *cV = col_const + 2 * BN;
        int tc = 0, staged_n0 = -1;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tc) {
            const int b = tc & 1;
            const int m0 = (t / n_tiles) * TC_BM, n0 = (t % n_tiles) * BN;
            if (n0 != staged_n0) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = te; c < BN; c += 128) {
                    const int col = n0 + c;
                    const bool ok = col < ep.Nout;
                    const float sc = (ok && ep.scale) ? __ldg(ep.scale + col) : 1.f;
                    const float bi = (ok && ep.bias) ? __ldg(ep.bias + col) : 0.f;
                    const float sf = (ok && ep.shift) ? __ldg(ep.shift + col) : 0.f;
                    const float rv = (ok && ep.row_vec) ? __ldg(ep.row_vec + col) : 0.f;
                    cS[c] = sc;
                    cB[c] = fmaf(bi, sc, sf);
                    cV[c] = rv * sc;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                staged_n0 = n0;
            }
            mbar_wait(&tmem_full_bar[b], (tc >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = m0 + wq * 32 + lane;
            const bool mok = m < ep.M;
            const float rs = (ep.row_scale && mok) ? __ldg(ep.row_scale + m) : 0.f;
            const float *trow = (ep.tab && mok) ? ep.tab + (int64_t)__ldg(ep.tab_idx + m) * ep.tab_ld : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int nb = n0 + c0;
                if (nb >= ep.Nout) break;            // warp-uniform
                uint32_t r[32], q[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * 2 * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                      "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
                      "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]),
                      "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                    : "r"(taddr + (uint32_t)BN)
                    : "memory");
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(q[j]);
                if (trow) {
                    if (nb + 32 <= ep.Nout && (ep.tab_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.tab) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 tv = __ldg(reinterpret_cast<const float4 *>(trow + nb + j));
                            v[j] += tv.x; v[j + 1] += tv.y; v[j + 2] += tv.z; v[j + 3] += tv.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += (nb + j < ep.Nout) ? __ldg(trow + nb + j) : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 s4 = *reinterpret_cast<const float4 *>(cS + c0 + j);
                    const float4 b4 = *reinterpret_cast<const float4 *>(cB + c0 + j);
                    v[j] = fmaf(v[j], s4.x, b4.x); v[j + 1] = fmaf(v[j + 1], s4.y, b4.y);
                    v[j + 2] = fmaf(v[j + 2], s4.z, b4.z); v[j + 3] = fmaf(v[j + 3], s4.w, b4.w);
                }
                if (ep.row_vec) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 v4 = *reinterpret_cast<const float4 *>(cV + c0 + j);
                        v[j] = fmaf(rs, v4.x, v[j]); v[j + 1] = fmaf(rs, v4.y, v[j + 1]);
                        v[j + 2] = fmaf(rs, v4.z, v[j + 2]); v[j + 3] = fmaf(rs, v4.w, v[j + 3]);
                    }
                }
                if (act == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (act != 3) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = tc_act_slow(v[j], act);
                }
                // the previous TMA store of this warp must have finished READING the staging buffer
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const uint32_t addr = stg_row + (uint32_t)((j4 ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j4]), "f"(v[4 * j4 + 1]),
                                 "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (accumulate)
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                                     ::"l"(&tmC), "r"(nb), "r"(m0 + wq * 32), "r"(smem_u32(stg)) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                                     ::"l"(&tmC), "r"(nb), "r"(m0 + wq * 32), "r"(smem_u32(stg)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            // this warp is done reading accumulator buffer b: hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[b])) : "memory");
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(4 * BN)) : "memory");
    }
}

// x -> (rn_tf32(x), x - rn_tf32(x)) ; two row-major sources concatenated along K into [M, K1+K2]
__global__ void split_a_kernel(const float *__restrict__ A1, int K1, int lda1, const float *__restrict__ A2, int K2,
                               int lda2, int64_t M, float *__restrict__ hi, float *__restrict__ lo) {
    const int K = K1 + K2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 each
    const int kq = K / 4;
    if (t >= M * kq) return;
    const int64_t m = t / kq;
    const int k = (int)(t % kq) * 4;
    float4 v = k < K1 ? __ldg(reinterpret_cast<const float4 *>(A1 + m * lda1 + k))
                      : __ldg(reinterpret_cast<const float4 *>(A2 + m * lda2 + (k - K1)));
    float a[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a[i]));
        h[i] = __uint_as_float(u);
        // the residual is rounded (not left for the tensor core to truncate): truncation is biased and its
        // error grows linearly with K, rounding keeps it a random walk
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a[i] - h[i]));
        l[i] = __uint_as_float(u);
    }
    *reinterpret_cast<float4 *>(hi + m * K + k) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(lo + m * K + k) = make_float4(l[0], l[1], l[2], l[3]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 row-major [rows, cols] (row stride ld elements), box = [box_rows, 32 cols], 128-byte swizzle, OOB -> 0
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return GSN_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled failed: %d", (int)r);
        return GSN_E_CUDA;
    }
    return GSN_OK;
}

template <int BN, int STAGES, bool SPLIT>
static int launch_tc(const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &w_hi, const CUtensorMap &w_lo,
                     const TcEpilogue &ep, int K1, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (2 * TC_BM * TC_BK * 4 + 2 * BN * TC_BK * 4) + 1024;
    // per launch: the attribute belongs to the current device (a process-wide flag would skip a second GPU)
    GSN_CUDA_OK(cudaFuncSetAttribute(tc_linear_kernel<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(ep.M, TC_BM), (unsigned)ceil_div(ep.Nout, BN));
    tc_linear_kernel<BN, STAGES, SPLIT><<<grid, 384, smem, stream>>>(a_hi, a_lo, w_hi, w_lo, ep, K1);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("tc_linear_kernel");
    return GSN_OK;
}

template <int BN, int STAGES>
static int launch_tc_persistent(const CUtensorMap &a1, const CUtensorMap &a2, const CUtensorMap &w_hi, const CUtensorMap &w_lo,
                                const CUtensorMap &c_map, const TcEpilogue &ep, int K1, cudaStream_t stream) {
    // ring | 4 x 4 KB store staging | 3 x BN column constants | alignment slack
    constexpr size_t smem = (size_t)STAGES * (2 * TC_BM * TC_BK * 4 + 2 * BN * TC_BK * 4) + 4 * 4096 + 3 * BN * 4 + 1024;
    GSN_CUDA_OK(cudaFuncSetAttribute(tc_linear_persistent_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int m_tiles = (int)ceil_div(ep.M, TC_BM), n_tiles = (int)ceil_div(ep.Nout, BN);
    const int64_t total = (int64_t)m_tiles * n_tiles;
    const unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
    tc_linear_persistent_kernel<BN, STAGES><<<grid, 512, smem, stream>>>(a1, a2, w_hi, w_lo, c_map, ep, K1, m_tiles, n_tiles);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("tc_linear_persistent_kernel");
    return GSN_OK;
}

}  // namespace gsn

using namespace gsn;

extern "C" int gsn_split_tf32(const float *d_src, int64_t rows, int32_t cols, int32_t ld, float *d_hi, float *d_lo,
                              void *stream_) {
    if (rows < 0 || cols < 4 || cols % 4 || ld % 4 || !d_src || !d_hi || !d_lo) return GSN_E_INVALID;
    if (rows == 0) return GSN_OK;
    const int64_t n4 = rows * (cols / 4);
    split_a_kernel<<<(unsigned)ceil_div(n4, 256), 256, 0, (cudaStream_t)stream_>>>(d_src, cols, ld, nullptr, 0, 0, rows, d_hi, d_lo);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("gsn_split_tf32");
    return GSN_OK;
}

extern "C" int gsn_tc_linear_workspace_bytes(int64_t M, int32_t K, size_t *bytes) {
    if (!bytes || M < 0 || K < 0) return GSN_E_INVALID;
    *bytes = 2 * align_up(sizeof(float) * (size_t)M * (size_t)K, 1024) + 1024;
    return GSN_OK;
}

extern "C" int gsn_tc_linear_fwd(const GsnLinear *h_p, const float *d_Whi, const float *d_Wlo, void *d_ws, size_t ws_bytes,
                                 void *stream_) {
    if (!h_p || !d_Whi || !d_Wlo) return GSN_E_INVALID;
    const GsnLinear &p = *h_p;
    const int K = p.K1 + p.K2;
    if (p.M < 0 || p.Nout < 1 || K < 4 || !p.C || !p.A1) return GSN_E_INVALID;
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    if (p.K1 % 4 || p.K2 % 4 || p.lda1 % 4 || (p.K2 > 0 && (p.lda2 % 4 || !p.A2)) || !al16(p.A1) || !al16(p.A2) ||
        !al16(d_Whi) || !al16(d_Wlo))
        return GSN_E_UNSUPPORTED;
    if ((p.row_scale == nullptr) != (p.row_vec == nullptr) || (p.tab == nullptr) != (p.tab_idx == nullptr)) return GSN_E_INVALID;
    if (p.M == 0) return GSN_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    // wide tiles for throughput; when the grid would leave most SMs idle, narrow tiles give more CTAs and a deeper
    // ring (5 stages of 40 KB at BN = 32) so that a small GEMM is not bound by the TMA round-trip per stage
    int BN = p.Nout > 128 ? 256 : (p.Nout > 64 ? 128 : 64);
    const int64_t m_tiles = ceil_div(p.M, TC_BM);
    if (m_tiles * ceil_div(p.Nout, BN) < kNumSMs / 2 && p.Nout >= 32) BN = 32;
    TcEpilogue ep{p.bias, p.row_scale, p.row_vec, p.tab, p.scale, p.shift, p.tab_idx, p.C,
                  p.M, p.Nout, K, p.ldc, p.tab_ld, p.act, p.accumulate};
    CUtensorMap mA_hi, mA_lo, mW_hi, mW_lo;
    int rc;
    if ((rc = make_map(&mW_hi, d_Whi, p.Nout, K, K, BN))) return rc;
    if ((rc = make_map(&mW_lo, d_Wlo, p.Nout, K, K, BN))) return rc;
    const bool split_in_kernel = (p.K2 == 0 || p.K1 % TC_BK == 0) && p.tc_path != GSN_TC_PATH_PRESPLIT;
    if (split_in_kernel) {
        // raw activations straight through TMA; the tile is split into (hi, lo) in shared memory by the kernel
        if ((rc = make_map(&mA_hi, p.A1, p.M, p.K1, p.lda1, TC_BM))) return rc;
        if (p.K2 > 0) { if ((rc = make_map(&mA_lo, p.A2, p.M, p.K2, p.lda2, TC_BM))) return rc; }
        else mA_lo = mA_hi;
        // large problems: persistent kernel (continuous ring, double-buffered TMEM accumulators)
        const int BNp = p.Nout > 64 ? 128 : 64;
        if (p.tc_path != GSN_TC_PATH_ONE_TILE && m_tiles * ceil_div(p.Nout, BNp) >= 2 * kNumSMs && p.ldc % 4 == 0 && al16(p.C)) {
            CUtensorMap mC;                       // output tile store: box = 32 rows x 32 columns (128 B), swizzled
            if ((rc = make_map(&mC, p.C, p.M, p.Nout, p.ldc, 32))) return rc;
            if (BNp != BN) {
                if ((rc = make_map(&mW_hi, d_Whi, p.Nout, K, K, BNp))) return rc;
                if ((rc = make_map(&mW_lo, d_Wlo, p.Nout, K, K, BNp))) return rc;
            }
            if (BNp == 128) return launch_tc_persistent<128, 3>(mA_hi, mA_lo, mW_hi, mW_lo, mC, ep, p.K1, stream);
            return launch_tc_persistent<64, 4>(mA_hi, mA_lo, mW_hi, mW_lo, mC, ep, p.K1, stream);
        }
        if (BN == 32) return launch_tc<32, 5, true>(mA_hi, mA_lo, mW_hi, mW_lo, ep, p.K1, stream);
        if (BN == 256) return launch_tc<256, 2, true>(mA_hi, mA_lo, mW_hi, mW_lo, ep, p.K1, stream);
        if (BN == 128) return launch_tc<128, 3, true>(mA_hi, mA_lo, mW_hi, mW_lo, ep, p.K1, stream);
        return launch_tc<64, 4, true>(mA_hi, mA_lo, mW_hi, mW_lo, ep, p.K1, stream);
    }
    size_t need = 0;
    gsn_tc_linear_workspace_bytes(p.M, K, &need);
    if (!d_ws || ws_bytes < need) return GSN_E_WORKSPACE;
    float *a_hi = (float *)(((uintptr_t)d_ws + 1023) & ~(uintptr_t)1023);
    float *a_lo = a_hi + align_up(sizeof(float) * (size_t)p.M * K, 1024) / sizeof(float);
    const int64_t n4 = (int64_t)p.M * (K / 4);
    split_a_kernel<<<(unsigned)ceil_div(n4, 256), 256, 0, stream>>>(p.A1, p.K1, p.lda1, p.A2, p.K2, p.lda2, p.M, a_hi, a_lo);
    GSN_BUMP(1);
    if ((rc = make_map(&mA_hi, a_hi, p.M, K, K, TC_BM))) return rc;
    if ((rc = make_map(&mA_lo, a_lo, p.M, K, K, TC_BM))) return rc;
    if (BN == 32) return launch_tc<32, 5, false>(mA_hi, mA_lo, mW_hi, mW_lo, ep, K, stream);
    if (BN == 256) return launch_tc<256, 2, false>(mA_hi, mA_lo, mW_hi, mW_lo, ep, K, stream);
    if (BN == 128) return launch_tc<128, 3, false>(mA_hi, mA_lo, mW_hi, mW_lo, ep, K, stream);
    return launch_tc<64, 4, false>(mA_hi, mA_lo, mW_hi, mW_lo, ep, K, stream);
}
