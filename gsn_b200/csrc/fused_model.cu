// Whole-model fused inference forward of GNNSubstructures ('general' message kind) for sm_100a: ONE persistent
// kernel runs every layer of models_graph_classification.py:204-247 for a tile of WHOLE graphs.
//
// A PyG batch is a block-diagonal graph: a row tile that holds whole graphs (<= 128 nodes) never reads
// a row outside itself, so nothing but the inputs and the per-graph readout has to touch HBM -- the per-layer
// activations x, the split first Linear P = (P_i | P_j) of msg_fn, the neighbour sums S and the hidden rows H of
// update_fn live in shared memory / tensor memory from the first layer to the readout:
//
//   per layer (GSN_edge_sparse.py:111-166 + models_misc.py:52-59, re-associated as in gsn_b200/fused.py):
//     P_j = x Wxj^T            tcgen05 -> TMEM -> smem (fp32, gathered by neighbours)
//     P_i = x Wxi^T + shift    tcgen05 -> TMEM (read in place by the row's own thread)
//     Ux  = x U1x^T            tcgen05 -> TMEM (kept until the update epilogue)
//     S_i = sum_e act(P_i[i] + P_j[nbr e] + sum_g Te[rows_g(e)])        registers, smem gathers
//     H   = act((Ux + S Wf^T + deg vf + c1 [+ Tu[x_i]]) su + tu)        tcgen05 + epilogue
//     x'  = act((H U2^T + c2) sm + tm)                                  tcgen05 + epilogue (+ per-graph readout)
//
// Tensor-core arithmetic: fp32 parity (1e-5) needs ~21 mantissa bits.  fp16 has the SAME 11-bit mantissa as tf32 at
// twice the MMA rate and half the operand bytes, so every product is 3 x fp16 (a_hi w_hi + a_lo w_hi + a_hi w_lo)
// with EXACT power-of-two scaling that removes the range problem: each activation row is scaled so that its largest
// element lies in [2^14, 2^15) (dynamic, in the epilogue that produces the operand), each weight row likewise (static,
// host side); the epilogue multiplies the two exponents back.  Elements 2^17 below their row's maximum lose low bits
// of the lo part only (absolute error <= 2^-39 of the row maximum).  The two correction terms accumulate in their own
// TMEM accumulator (the tensor core truncates when it adds into fp32, see tc_linear.cu).
//
// Warp roles (384 threads, 1 CTA / SM):  warp 0 TMA producer (weight ring), warp 1 MMA issuer, warp 2 TMEM alloc,
// warps 4..11 compute (thread = (row, half of the columns): epilogues, message phase, operand conversion).
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"

namespace gsn {

constexpr int FM_MAX_LAYERS = GSN_FUSED_MAX_LAYERS;
constexpr int FM_ROWS = 128;
constexpr int FM_NV = 9;            // per-layer column vectors
enum { V_CJ = 0, V_CI, V_SHIFT, V_CU, V_CF, V_CV, V_CB, V_C2S, V_C2B };
constexpr int FM_TE_SMALL = 4096;   // bytes of the always-available edge-table staging area
constexpr int FM_NPART = 4;                       // column parts of a row = compute warps / 4
constexpr int FM_COMPUTE_THREADS = 128 * FM_NPART;
constexpr int FM_THREADS = 128 + FM_COMPUTE_THREADS;

struct FmLayer {
    const int32_t *node_rows; const float *Tn;
    const int32_t *tu_rows; const float *Tu;
    const int32_t *edge_rows; const float *Te;
    const float *vec; float *pooled; float *x_out;
    int32_t n_node_cols, tu_stride, n_edge_cols, te_rows, has_dense, mat0, act_msg, act_upd, act_out, pool;
    const float *jk_W0T, *jk_vec, *jk_W1, *jk_b1;
    int32_t jk_kind, jk_act, jk_first;       // jk_first: the first projecting layer writes `out`, later ones add
};

struct FmParams {
    FmLayer layer[FM_MAX_LAYERS];
    int32_t n_layers;
    const int32_t *rowptr, *nbr;
    const int64_t *node_ptr;
    const float *x0;
    int32_t x0_ld, x0_d;
    int32_t G, unit, n_units;
    const int32_t *tile_plan;     // optional: { n_tiles, first graph of every tile, G }
    float *out;                   // [G, n_out] in-kernel JK head
    int32_t n_out;
    int32_t *status;
};

// Build-flag-only profiling aid (python -m gsn_b200.build --profile -> libgsn_b200_prof.so): clock64 stamps of the
// phases of the first tile of CTA 0, [layer][16].  Not compiled into the shipped library.
#ifdef GSN_PROFILE_STAMPS
__device__ long long g_fm_stamps[FM_MAX_LAYERS * 16];
#define FM_STAMP(l, i) do { if (blockIdx.x == 0 && ct == 0 && first_tile) g_fm_stamps[(l) * 16 + (i)] = clock64(); } while (0)
#else
#define FM_STAMP(l, i) do { } while (0)
#endif

__device__ __noinline__ float fm_act_slow(float v, int act) {
    return act == 1 ? (v > 0.0f ? v : expm1f(v)) : tanhf(v);
}
__device__ __forceinline__ float fm_act(float v, int act) {
    if (act == 0) return fmaxf(v, 0.f);
    if (act == 3) return v;
    return fm_act_slow(v, act);
}

__device__ __forceinline__ void fm_tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t fm_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void fm_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void fm_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __noinline__ void fm_wait_timeout(int what, uint32_t parity) {
    printf("fused_model_kernel: mbarrier wait %d timed out (block %d thread %d parity %u)\n", what, blockIdx.x, threadIdx.x, parity);
    __trap();
}

// mbarrier wait with a watchdog: a protocol error traps (the launch fails with an error the host reports) instead of
// hanging the GPU.  `what` identifies the wait site in the message.
__device__ __forceinline__ void fm_wait(uint64_t *bar, uint32_t parity, int what) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done) {
            // back off: a spinning warp would otherwise take issue slots from the compute warps of its sub-partition
            // (ncu: 30 % of all issued instructions were try_wait loops before this)
            __nanosleep(32);
            if (++spins > (1u << 22)) fm_wait_timeout(what, parity);
        }
    } while (!done);
}

__device__ __forceinline__ void fm_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// N consecutive TMEM columns of this thread's lane (N = 16 | 32)
template <int N>
__device__ __forceinline__ void fm_tmem_ld(uint32_t taddr, float (&v)[N]) {
    uint32_t r[N];
    if constexpr (N == 32) {
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
    } else {
        static_assert(N == 16, "16 or 32 columns");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = __uint_as_float(r[j]);
}

// 16 fp32 values -> 16 consecutive TMEM columns of this thread's lane (completion: tcgen05.wait::st)
__device__ __forceinline__ void fm_tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// main + correction accumulator of one column chunk, 16 columns at a time (bounded register footprint)
template <int D, int N>
__device__ __forceinline__ void fm_acc_ld(uint32_t acc_base, int col0, float (&v)[N]) {
#pragma unroll
    for (int h = 0; h < N / 16; ++h) {
        float a[16], b[16];
        fm_tmem_ld<16>(acc_base + (uint32_t)(col0 + 16 * h), a);
        fm_tmem_ld<16>(acc_base + (uint32_t)(D + col0 + 16 * h), b);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * h + j] = a[j] + b[j];
    }
}

// v[j] = fma(acc[j], rs * cvec[j], v[j]) with cvec in shared memory
template <int D, int N>
__device__ __forceinline__ void fm_acc_fma(uint32_t acc_base, int col0, float rs, const float *cvec, float (&v)[N]) {
#pragma unroll
    for (int h = 0; h < N / 16; ++h) {
        float a[16], b[16];
        fm_tmem_ld<16>(acc_base + (uint32_t)(col0 + 16 * h), a);
        fm_tmem_ld<16>(acc_base + (uint32_t)(D + col0 + 16 * h), b);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 c = *reinterpret_cast<const float4 *>(cvec + col0 + 16 * h + 4 * i);
            v[16 * h + 4 * i] = fmaf(a[4 * i] + b[4 * i], rs * c.x, v[16 * h + 4 * i]);
            v[16 * h + 4 * i + 1] = fmaf(a[4 * i + 1] + b[4 * i + 1], rs * c.y, v[16 * h + 4 * i + 1]);
            v[16 * h + 4 * i + 2] = fmaf(a[4 * i + 2] + b[4 * i + 2], rs * c.z, v[16 * h + 4 * i + 2]);
            v[16 * h + 4 * i + 3] = fmaf(a[4 * i + 3] + b[4 * i + 3], rs * c.w, v[16 * h + 4 * i + 3]);
        }
    }
}

// compile-time loop: the body sees its index as a constant, so per-iteration register arrays keep static indices even
// when the body contains a data-dependent loop the compiler will not unroll
template <int I, int N, class F>
__device__ __forceinline__ void fm_static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        fm_static_for<I + 1, N>(f);
    }
}

// byte offset of 16-byte chunk q of row `row` in an fp32 [rows, D] shared-memory matrix whose chunks are XOR-swizzled
// inside every 128-byte group (rows that differ in their low 3 bits hit different banks when read at the same column)
template <int D>
__device__ __forceinline__ uint32_t fm_swz(int row, int q) {
    return (uint32_t)row * (uint32_t)(D * 4) + (uint32_t)(((q & ~7) | ((q ^ row) & 7)) << 4);
}

// first graph after the tile that starts at graph g0 (whole graphs, <= 128 rows, <= 32 graphs, inside the unit)
__device__ __forceinline__ int fm_tile_end(const int64_t *node_ptr, int g0, int gend, int lane) {
    const int64_t base = __ldg(node_ptr + g0);
    const int g = g0 + 1 + lane;
    const bool ok = g <= gend && (__ldg(node_ptr + g) - base) <= FM_ROWS;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int cnt = (m == 0xffffffffu) ? 32 : (__ffs((int)~m) - 1);
    return g0 + cnt;
}

__device__ __forceinline__ void fm_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(FM_COMPUTE_THREADS) : "memory"); }

template <int D>
struct FmCfg {
    static constexpr int KS = D / 64;                       // 64-half (128-byte) k-slabs per operand
    static constexpr int SLAB = FM_ROWS * 128;              // bytes of one A k-slab (hi or lo)
    static constexpr int A_BYTES = 2 * KS * SLAB;           // hi slabs | lo slabs
    static constexpr int PJ_BYTES = FM_ROWS * D * 4;
    static constexpr int STAGE = D * 128;                   // one weight k-slab (hi or lo): D rows x 128 bytes
    static constexpr int NST = D == 128 ? 5 : 8;
    static constexpr int VEC_BYTES = FM_NV * D * 4;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_PJ = OFF_A + A_BYTES;
    static constexpr int OFF_RING = OFF_PJ + PJ_BYTES;
    static constexpr int OFF_TE = OFF_RING + NST * STAGE;
    static constexpr int OFF_VEC = OFF_TE + FM_TE_SMALL;
    static constexpr int OFF_RMAX = OFF_VEC + VEC_BYTES;
    static constexpr int SMEM = OFF_RMAX + 2 * FM_NPART * FM_ROWS * 4 + 1024;     // two row-max buffers + alignment slack
    static constexpr int CW = D / FM_NPART;                 // columns per compute thread (FM_NPART threads share a row)
    static constexpr uint32_t TMEM_COLS = 4 * D;            // acc0 (main|corr) | acc1 (main|corr)
};

template <int D>
__global__ void __launch_bounds__(FM_THREADS, 1)
fused_model_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                   const __grid_constant__ FmParams P) {
    using C = FmCfg<D>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t w_full[C::NST], w_empty[C::NST], a_ready, acc_full[2], acc1_free;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the grid of a planned launch is the plan's CAPACITY (fixed when the step is captured): CTAs past the tile count leave
    // before they allocate anything
    if (P.tile_plan && (int)blockIdx.x >= P.tile_plan[0]) return;
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sA = smem_u32(sm + C::OFF_A), sPJ = smem_u32(sm + C::OFF_PJ), sRING = smem_u32(sm + C::OFF_RING);
    const uint32_t sTE = smem_u32(sm + C::OFF_TE);
    float *vecs = reinterpret_cast<float *>(sm + C::OFF_VEC);
    float *rmax = reinterpret_cast<float *>(sm + C::OFF_RMAX);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::NST; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        mbar_init(&a_ready, 4 * FM_NPART);
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        mbar_init(&acc1_free, 4 * FM_NPART);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t ACC0 = tmem, ACC1 = tmem + 2 * D;
    const int nL = P.n_layers;
    // units of work: tiles of the plan (small batches: one CTA per tile) or runs of P.unit graphs cut into tiles here
    const int32_t *tplan = P.tile_plan;
    const int n_units = tplan ? tplan[0] : P.n_units;

    if (warp == 0) {
        // ------------------------------------------------------------- TMA producer: weight k-slabs in order of use
        uint32_t it = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int gend = tplan ? tplan[2 + u] : min(P.G, (u + 1) * P.unit);
            int g0 = tplan ? tplan[1 + u] : u * P.unit;
            while (g0 < gend) {
                int g1 = fm_tile_end(P.node_ptr, g0, gend, lane);
                if (g1 == g0) { g0 += 1; continue; }
                if (lane == 0) {
                    for (int l = 0; l < nL; ++l) {
                        const int nm = P.layer[l].has_dense ? 5 : 2;
                        for (int m = 0; m < nm; ++m) {
                            const int wrow = (P.layer[l].mat0 + m) * D;
                            for (int s = 0; s < C::KS; ++s) {
#pragma unroll 1
                                for (int part = 0; part < 2; ++part, ++it) {
                                    const uint32_t st = it % C::NST, ph = (it / C::NST) & 1;
                                    fm_wait(&w_empty[st], ph ^ 1, 1);
                                    mbar_arrive_expect_tx(&w_full[st], C::STAGE);
                                    fm_tma_load_2d(sm + C::OFF_RING + st * C::STAGE, part == 0 ? &tmWhi : &tmWlo, s * 64, wrow,
                                                   &w_full[st]);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                g0 = g1;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer
        // instruction descriptor: D = F32, A = B = F16, both K-major, N = D, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(FM_ROWS >> 4) << 24);
        uint32_t it = 0, n_aready = 0, n_free = 0;
        // one GEMM: acc (main | corr) = A(smem operand buffer) x W(ring)^T
        auto gemm = [&](uint32_t acc) {
            for (int s = 0; s < C::KS; ++s) {
                const uint64_t a_hi = fm_desc(sA + s * C::SLAB), a_lo = fm_desc(sA + (C::KS + s) * C::SLAB);
                {   // hi weights: main += a_hi w_hi ; corr += a_lo w_hi
                    const uint32_t st = it % C::NST, ph = (it / C::NST) & 1;
                    fm_wait(&w_full[st], ph, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t w = fm_desc(sRING + st * C::STAGE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t first = (s > 0 || k > 0) ? 1u : 0u;
                        fm_mma_f16(acc, a_hi + 2 * k, w + 2 * k, idesc, first);
                        fm_mma_f16(acc + D, a_lo + 2 * k, w + 2 * k, idesc, first);
                    }
                    fm_commit(&w_empty[st]);
                    ++it;
                }
                {   // lo weights: corr += a_hi w_lo
                    const uint32_t st = it % C::NST, ph = (it / C::NST) & 1;
                    fm_wait(&w_full[st], ph, 3);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t w = fm_desc(sRING + st * C::STAGE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) fm_mma_f16(acc + D, a_hi + 2 * k, w + 2 * k, idesc, 1u);
                    fm_commit(&w_empty[st]);
                    ++it;
                }
            }
        };
        auto wait_aready = [&]() {
            fm_wait(&a_ready, n_aready & 1, 4);
            ++n_aready;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int gend = tplan ? tplan[2 + u] : min(P.G, (u + 1) * P.unit);
            int g0 = tplan ? tplan[1 + u] : u * P.unit;
            while (g0 < gend) {
                int g1 = fm_tile_end(P.node_ptr, g0, gend, lane);
                if (g1 == g0) { g0 += 1; continue; }
                if (lane == 0) {
                    for (int l = 0; l < nL; ++l) {
                        if (P.layer[l].has_dense) {
                            wait_aready();                       // x operand in place, both accumulators drained
                            gemm(ACC1);                          // P_j
                            fm_commit(&acc_full[1]);
                            gemm(ACC0);                          // P_i
                            fm_commit(&acc_full[0]);
                            fm_wait(&acc1_free, n_free & 1, 5);   // P_j copied out of ACC1
                            ++n_free;
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            gemm(ACC1);                          // Ux = x U1x^T
                            fm_commit(&acc_full[1]);
                        }
                        wait_aready();                           // S operand in place, P_i (ACC0) consumed
                        gemm(ACC0);                              // S Wf^T
                        fm_commit(&acc_full[0]);
                        wait_aready();                           // H operand in place, ACC0 / ACC1 consumed
                        gemm(ACC0);                              // H U2^T
                        fm_commit(&acc_full[0]);
                    }
                }
                __syncwarp();
                g0 = g1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------- compute warps
        // thread = (tile row r = TMEM lane, column part): CW = D / FM_NPART consecutive columns of the row, always
        // handled in PIECES of 16 columns so that the register footprint stays ~3 x 16 values whatever D is.  A row of
        // values that becomes the next MMA operand (x, S, H, x') is first stashed, fp32, in the already consumed
        // columns of accumulator 0 (tensor memory is the only free storage of that size), because the operand's
        // power-of-two scale needs the row maximum before the first element can be converted.
        constexpr int CW = C::CW;
        constexpr int NP = CW / 16;              // pieces per thread
        const int cw = warp - 4;
        const int q = cw & 3;                    // TMEM lane quarter this warp may touch (= warp % 4)
        const int part = cw >> 2;                // which part of the columns
        const int col0 = part * CW;
        // slot r = TMEM lane = operand row = row of the P_j buffer.  Tile-local graph rows are INTERLEAVED over the four
        // lane quarters (row = lane * 4 + q): the warps of a quarter share one SM sub-partition (warp % 4), so a tile
        // with few rows (one molecule per tile at B = 128) still spreads its work over all four schedulers
        const int r = q * 32 + lane;
        const int row_l = lane * 4 + q;
        const int ct = threadIdx.x - 128;        // 0..FM_COMPUTE_THREADS-1
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t STASH = ACC0 + lane_off;
        uint32_t n_full[2] = {0, 0};
        auto wait_acc = [&](int b) {
            fm_wait(&acc_full[b], n_full[b] & 1, 6 + b);
            ++n_full[b];
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
        // stashed row (columns [col0, col0 + CW) of accumulator 0, this thread's lane) -> fp16 (hi, lo) operand row scaled
        // by the row's power of two; m = this thread's partial row maximum.  Returns the inverse scale, leaves the operand
        // visible to the tensor core and signals the MMA warp.  `live` = the warp holds at least one valid row (other
        // warps only take part in the barriers: their operand rows are never read back -- row i of a product depends on
        // row i of the operand alone).
        // The partial maxima alternate between two buffers: a warp that runs ahead into the next call writes the other
        // buffer, and the one after that is ordered behind this call's reads by the next call's barrier.
        uint32_t n_operand = 0;
        auto finish_operand = [&](float m, bool live) -> float {
            float *rm = rmax + (n_operand & 1) * (FM_NPART * FM_ROWS);
            ++n_operand;
            if (live) rm[part * FM_ROWS + r] = m;
            fm_bar_compute();
            float inv = 1.f;
            if (live) {
                m = rm[r];
#pragma unroll
                for (int pp = 1; pp < FM_NPART; ++pp) m = fmaxf(m, rm[pp * FM_ROWS + r]);
                int e = (int)((__float_as_uint(m) >> 23) & 0xFF);                 // biased exponent of the row maximum
                if (e == 0) e = 127 + 14;                                          // zero / denormal row: scale 1
                int se = 127 + 14 - (e - 127);                                     // scale = 2^(14 - (e - 127))
                se = max(1, min(254, se));
                const float scale = __uint_as_float((uint32_t)se << 23);
                inv = __uint_as_float((uint32_t)(254 - se) << 23);                // 2^-(se-127)
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll 1
                for (int pc = 0; pc < NP; ++pc) {
                    const int c0 = col0 + 16 * pc;
                    float v[16];
                    fm_tmem_ld<16>(STASH + (uint32_t)c0, v);
                    const int slab = c0 >> 6;
                    const int j0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float a = v[jj * 8 + 2 * t] * scale, b = v[jj * 8 + 2 * t + 1] * scale;
                            const __half2 h2 = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
                            hi[t] = *reinterpret_cast<const uint32_t *>(&h2);
                            lo[t] = *reinterpret_cast<const uint32_t *>(&l2);
                        }
                        const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((j0 + jj) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + slab * C::SLAB + off), "r"(hi[0]),
                                     "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + (C::KS + slab) * C::SLAB + off),
                                     "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) fm_arrive(&a_ready);
            return inv;
        };
        auto absmax16 = [](float m, const float (&v)[16]) -> float {
#pragma unroll
            for (int j = 0; j < 16; ++j) m = fmaxf(m, fabsf(v[j]));
            return m;
        };
        // v (+)= 16 columns of a row of a swizzled shared-memory matrix / of a global-memory matrix
        auto add_lds = [&](uint32_t base, int row, int c0, float (&v)[16]) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                float4 t;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                             : "r"(base + fm_swz<D>(row, c0 / 4 + i4)));
                v[4 * i4] += t.x; v[4 * i4 + 1] += t.y; v[4 * i4 + 2] += t.z; v[4 * i4 + 3] += t.w;
            }
        };
        auto add_ldg = [&](const float *rowp, int c0, float (&v)[16]) {
            const float4 *t4 = reinterpret_cast<const float4 *>(rowp + c0);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 t = __ldg(t4 + i4);
                v[4 * i4] += t.x; v[4 * i4 + 1] += t.y; v[4 * i4 + 2] += t.z; v[4 * i4 + 3] += t.w;
            }
        };
        auto sts16 = [&](uint32_t base, int row, int c0, const float (&v)[16]) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + fm_swz<D>(row, c0 / 4 + i4)), "f"(v[4 * i4]),
                             "f"(v[4 * i4 + 1]), "f"(v[4 * i4 + 2]), "f"(v[4 * i4 + 3]) : "memory");
        };
        auto set16 = [&](const float *cvec, int c0, float (&v)[16]) {       // per-column constants (shared memory)
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 t = *reinterpret_cast<const float4 *>(cvec + c0 + 4 * i4);
                v[4 * i4] = t.x; v[4 * i4 + 1] = t.y; v[4 * i4 + 2] = t.z; v[4 * i4 + 3] = t.w;
            }
        };

        // constants of the first layer of the first tile; afterwards they are prefetched one layer ahead
        constexpr int NVQ = FM_NV * D / 4;          // float4 of constants per layer
        constexpr int NVT = (NVQ + FM_COMPUTE_THREADS - 1) / FM_COMPUTE_THREADS;
        for (int i = ct; i < NVQ; i += FM_COMPUTE_THREADS)
            reinterpret_cast<float4 *>(vecs)[i] = __ldg(reinterpret_cast<const float4 *>(P.layer[0].vec) + i);

        bool first_tile = true;
        (void)first_tile;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int gend = tplan ? tplan[2 + u] : min(P.G, (u + 1) * P.unit);
            int g0 = tplan ? tplan[1 + u] : u * P.unit;
            while (g0 < gend) {
                const int g1 = fm_tile_end(P.node_ptr, g0, gend, lane);
                if (g1 == g0) {
                    if (ct == 0) atomicOr(P.status, GSN_S_GRAPH_TOO_LARGE);
                    g0 += 1;
                    continue;
                }
                const int64_t row0 = __ldg(P.node_ptr + g0);
                const int n_rows = (int)(__ldg(P.node_ptr + g1) - row0);
                const bool valid = row_l < n_rows;
                const bool live = q < n_rows;               // warp-uniform: lane 0 holds the warp's smallest row
                const int64_t grow = row0 + (valid ? row_l : 0);
                const int e_begin = valid ? __ldg(P.rowptr + grow) : 0;
                const int e_end = valid ? __ldg(P.rowptr + grow + 1) : 0;
                const float deg = (float)(e_end - e_begin);
                // neighbours of the first four in-edges, tile-local, 8 bits each (0xFF = outside the tile): fetched once
                // per tile -- with ~227 KB of shared memory the SM has no L1 left, every global load is an L2 round trip
                uint32_t nbr_pk = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (e_begin + i < e_end) {
                        const int64_t jl = (int64_t)__ldg(P.nbr + e_begin + i) - row0;
                        nbr_pk |= (uint32_t)((jl >= 0 && jl < n_rows) ? (((int)jl & 3) * 32 + ((int)jl >> 2)) : 0xFF) << (8 * i);
                    }
                }
                float rx_inv = 1.f;       // inverse row scale of the operand currently in the A buffer

                for (int l = 0; l < nL; ++l) {
                    const FmLayer &L = P.layer[l];
                    FM_STAMP(l, 0);
                    const int ng = L.n_edge_cols;
                    // table rows of the first four in-edges, 16 bits each (up to 3 column groups)
                    const bool rows_pk_ok = ng <= 3 && L.te_rows <= 65536;
                    uint64_t rows_pk[3] = {0ull, 0ull, 0ull};       // [group]: 4 x 16 bits
                    if (rows_pk_ok) {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (g < ng && e_begin + i < e_end)
                                    rows_pk[g] |= (uint64_t)(uint32_t)__ldg(L.edge_rows + (int64_t)(e_begin + i) * ng + g) << (16 * i);
                    }
                    // edge tables: 0 global, 1 small area, 2 the (still unused) operand buffer of a table-only layer
                    int te_loc = 0;
                    const int te_bytes = L.te_rows * D * 4;
                    if (ng > 0) {
                        if (te_bytes <= FM_TE_SMALL) te_loc = 1;
                        else if (!L.has_dense && te_bytes <= C::A_BYTES) te_loc = 2;
                    }
                    const uint32_t sTab = te_loc == 2 ? sA : sTE;
                    // the previous layer's readers of the small table area / the constants are past their last use
                    fm_bar_compute();
                    if (te_loc) {
                        const int nq = L.te_rows * (D / 4);
                        for (int i = ct; i < nq; i += FM_COMPUTE_THREADS) {
                            const int trow = i / (D / 4), tq = i % (D / 4);
                            const float4 v = __ldg(reinterpret_cast<const float4 *>(L.Te) + i);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sTab + fm_swz<D>(trow, tq)), "f"(v.x),
                                         "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        }
                    }
                    if (l == 0 && L.has_dense) {
                        // ---- dense input features: fp32 rows -> operand buffer (accumulator 0 is free: stash, then convert)
                        float m = 0.f;
                        if (live) {
#pragma unroll 1
                            for (int pc = 0; pc < NP; ++pc) {
                                const int c0 = col0 + 16 * pc;
                                float v[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    v[j] = (valid && c0 + j < P.x0_d) ? __ldg(P.x0 + grow * P.x0_ld + c0 + j) : 0.f;
                                m = absmax16(m, v);
                                fm_tmem_st16(STASH + (uint32_t)c0, v);
                            }
                        }
                        rx_inv = finish_operand(m, live);
                    }
                    fm_bar_compute();          // constants / tables staged
                    FM_STAMP(l, 1);

                    // ---- P_j rows into shared memory
                    if (L.has_dense) wait_acc(1);
                    if (live) {
#pragma unroll 1
                        for (int pc = 0; pc < NP; ++pc) {
                            const int c0 = col0 + 16 * pc;
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;
                            if (L.has_dense) fm_acc_fma<D, 16>(ACC1 + lane_off, c0, rx_inv, vecs + V_CJ * D, v);
                            if (valid)
                                for (int c = 0; c < L.n_node_cols; ++c)
                                    add_ldg(L.Tn + (int64_t)__ldg(L.node_rows + grow * L.n_node_cols + c) * (2 * D) + D, c0, v);
                            sts16(sPJ, r, c0, v);
                        }
                    }
                    if (L.has_dense) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) fm_arrive(&acc1_free);
                    }
                    fm_bar_compute();          // every row of P_j visible
                    FM_STAMP(l, 2);

                    // ---- message phase: S_i = sum_e act(P_i + P_j[nbr] + sum_g Te[rows]), 16 columns per pass over the
                    //      row's in-edges; each finished piece replaces the consumed P_i piece in accumulator 0
                    float m_s = 0.f;
                    if (L.has_dense) wait_acc(0);
                    FM_STAMP(l, 10);
                    if (live) {
                        const int act = L.act_msg;
#pragma unroll 1
                        for (int pc = 0; pc < NP; ++pc) {
                            const int c0 = col0 + 16 * pc;
                            float p[16], S[16];
                            set16(vecs + V_SHIFT * D, c0, p);
                            if (L.has_dense) fm_acc_fma<D, 16>(ACC0 + lane_off, c0, rx_inv, vecs + V_CI * D, p);
                            if (valid)
                                for (int c = 0; c < L.n_node_cols; ++c)
                                    add_ldg(L.Tn + (int64_t)__ldg(L.node_rows + grow * L.n_node_cols + c) * (2 * D), c0, p);
#pragma unroll
                            for (int j = 0; j < 16; ++j) S[j] = 0.f;
                            for (int k = e_begin; k < e_end; ++k) {
                                const int i = k - e_begin;
                                int jl;
                                if (i < 4) jl = (int)((nbr_pk >> (8 * i)) & 0xFF);
                                else {
                                    const int64_t t = (int64_t)__ldg(P.nbr + k) - row0;
                                    jl = (t >= 0 && t < n_rows) ? (((int)t & 3) * 32 + ((int)t >> 2)) : 0xFF;
                                }
                                if (jl == 0xFF) {
                                    atomicOr(P.status, GSN_S_CROSS_GRAPH_EDGE);
                                    continue;
                                }
                                float h[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) h[j] = p[j];
                                add_lds(sPJ, jl, c0, h);
                                if (rows_pk_ok && i < 4) {
#pragma unroll
                                    for (int g = 0; g < 3; ++g) {
                                        if (g < ng) {
                                            const int tr = (int)((rows_pk[g] >> (16 * i)) & 0xFFFF);
                                            if (te_loc) add_lds(sTab, tr, c0, h);
                                            else add_ldg(L.Te + (int64_t)tr * D, c0, h);
                                        }
                                    }
                                } else {
                                    for (int g = 0; g < ng; ++g) {
                                        const int tr = __ldg(L.edge_rows + (int64_t)k * ng + g);
                                        if (te_loc) add_lds(sTab, tr, c0, h);
                                        else add_ldg(L.Te + (int64_t)tr * D, c0, h);
                                    }
                                }
                                if (act == 0) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) S[j] += fmaxf(h[j], 0.f);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) S[j] += fm_act(h[j], act);
                                }
                            }
                            m_s = absmax16(m_s, S);
                            __syncwarp();          // lanes leave the edge loop at different trip counts; the store is warp-wide
                            fm_tmem_st16(STASH + (uint32_t)c0, S);
                        }
                    }
                    FM_STAMP(l, 3);
                    // the operand buffer may be overwritten once Ux (the last GEMM reading x) has completed; a table-only
                    // layer has no GEMM in flight, but its tables may sit in the operand buffer: finish_operand's barrier
                    // (every thread is past its message loop) orders that
                    if (L.has_dense) wait_acc(1);
                    FM_STAMP(l, 4);
                    const float rs_inv = finish_operand(m_s, live);
                    FM_STAMP(l, 5);

                    // ---- update epilogue: H = act((Ux + S Wf^T + deg vf + c1 [+ Tu]) su + tu)
                    float rh_inv;
                    {
                        // row of the one-hot input's update table: requested before the GEMM wait
                        const float *tu_row = (L.Tu && valid) ? L.Tu + (int64_t)__ldg(L.tu_rows + grow * L.tu_stride) * D : nullptr;
                        wait_acc(0);
                        FM_STAMP(l, 6);
                        float m = 0.f;
                        if (live) {
                            const int act = L.act_upd;
#pragma unroll 1
                            for (int pc = 0; pc < NP; ++pc) {
                                const int c0 = col0 + 16 * pc;
                                float H[16];
#pragma unroll
                                for (int i4 = 0; i4 < 4; ++i4) {
                                    const float4 cv = *reinterpret_cast<const float4 *>(vecs + V_CV * D + c0 + 4 * i4);
                                    const float4 cb = *reinterpret_cast<const float4 *>(vecs + V_CB * D + c0 + 4 * i4);
                                    H[4 * i4] = fmaf(deg, cv.x, cb.x); H[4 * i4 + 1] = fmaf(deg, cv.y, cb.y);
                                    H[4 * i4 + 2] = fmaf(deg, cv.z, cb.z); H[4 * i4 + 3] = fmaf(deg, cv.w, cb.w);
                                }
                                fm_acc_fma<D, 16>(ACC0 + lane_off, c0, rs_inv, vecs + V_CF * D, H);
                                if (L.has_dense) fm_acc_fma<D, 16>(ACC1 + lane_off, c0, rx_inv, vecs + V_CU * D, H);
                                if (tu_row) add_ldg(tu_row, c0, H);
                                if (act == 0) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) H[j] = valid ? fmaxf(H[j], 0.f) : 0.f;
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) H[j] = valid ? fm_act(H[j], act) : 0.f;
                                }
                                m = absmax16(m, H);
                                fm_tmem_st16(STASH + (uint32_t)c0, H);
                            }
                        }
                        rh_inv = finish_operand(m, live);
                        FM_STAMP(l, 7);
                    }

                    // ---- output epilogue: x' = act((H U2^T + c2) sm + tm), readout, next layer's operand
                    {
                        // constants of the next layer in sequence (the first layer of the next tile after the last one)
                        const FmLayer &Ln = P.layer[l + 1 < nL ? l + 1 : 0];
                        float4 vnext[NVT];
#pragma unroll
                        for (int i = 0; i < NVT; ++i)
                            if (ct + i * FM_COMPUTE_THREADS < NVQ)
                                vnext[i] = __ldg(reinterpret_cast<const float4 *>(Ln.vec) + ct + i * FM_COMPUTE_THREADS);
                        wait_acc(0);
                        FM_STAMP(l, 8);
                        float m = 0.f;
                        if (live) {
                            const int act = L.act_out;
#pragma unroll 1
                            for (int pc = 0; pc < NP; ++pc) {
                                const int c0 = col0 + 16 * pc;
                                float X[16];
                                set16(vecs + V_C2B * D, c0, X);
                                fm_acc_fma<D, 16>(ACC0 + lane_off, c0, rh_inv, vecs + V_C2S * D, X);
#pragma unroll
                                for (int j = 0; j < 16; ++j) X[j] = valid ? fm_act(X[j], act) : 0.f;
                                if (L.x_out && valid) {
                                    float4 *o4 = reinterpret_cast<float4 *>(L.x_out + grow * D + c0);
#pragma unroll
                                    for (int i4 = 0; i4 < 4; ++i4) o4[i4] = make_float4(X[4 * i4], X[4 * i4 + 1], X[4 * i4 + 2], X[4 * i4 + 3]);
                                }
                                if (L.pool) sts16(sPJ, r, c0, X);
                                if (l + 1 < nL) {
                                    m = absmax16(m, X);
                                    fm_tmem_st16(STASH + (uint32_t)c0, X);
                                }
                            }
                        }
                        if (L.pool) {
                            // per-graph readout (global_add_pool_sparse / global_mean_pool_sparse, utils_graph_learning.py:23-41):
                            // rows of a graph summed in row order -> deterministic
                            fm_bar_compute();
                            const int col = ct % D;
                            constexpr int GPP = FM_COMPUTE_THREADS / D;       // graphs per pass of the pooling loop
                            constexpr int MAXP = 32 / GPP;                    // a tile holds <= 32 graphs
                            float pv[MAXP];
#pragma unroll
                            for (int k = 0; k < MAXP; ++k) pv[k] = 0.f;
#pragma unroll
                            for (int k = 0; k < MAXP; ++k) {
                                const int g = g0 + ct / D + k * GPP;
                                if (g < g1) {
                                    const int ra = (int)(__ldg(P.node_ptr + g) - row0), rb = (int)(__ldg(P.node_ptr + g + 1) - row0);
                                    float acc = 0.f;
                                    for (int rr = ra; rr < rb; ++rr) {
                                        float t;
                                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(sPJ + fm_swz<D>((rr & 3) * 32 + (rr >> 2), col >> 2) + (uint32_t)((col & 3) << 2)));
                                        acc += t;
                                    }
                                    if (L.pool == 2 && rb > ra) acc /= (float)(rb - ra);
                                    if (L.pooled) L.pooled[(int64_t)g * D + col] = acc;
                                    pv[k] = acc;
                                }
                            }
                            if (L.jk_kind) {
                                // ---- JK head of this readout (models_graph_classification.py:236-240): the pooled rows and
                                //      the hidden rows of the head live in the (now consumed) P_j buffer, plain [graph][col]
                                float *sPool = reinterpret_cast<float *>(sm + C::OFF_PJ);
                                float *sHid = sPool + 32 * D;
                                const int ng = g1 - g0;
                                fm_bar_compute();                      // every pooling read of the buffer is done
#pragma unroll
                                for (int k = 0; k < MAXP; ++k)
                                    if (ct / D + k * GPP < ng) sPool[(ct / D + k * GPP) * D + col] = pv[k];
                                fm_bar_compute();
                                const float *sIn = sPool;
                                if (L.jk_kind == 2) {
                                    // hidden rows: the k range is split over the GPP thread groups (every weight is read once
                                    // per tile, all loads of a thread in flight together), partial sums meet in shared memory
                                    constexpr int KPT = D / GPP;              // k values per thread
                                    float *sPart = sHid + 32 * D;             // [GPP][8][D]
                                    const int kq = ct / D;
                                    const float b0 = __ldg(L.jk_vec + col), s0 = __ldg(L.jk_vec + D + col), t0 = __ldg(L.jk_vec + 2 * D + col);
                                    for (int gb = 0; gb < ng; gb += 8) {
                                        float a[8];
#pragma unroll
                                        for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll 16
                                        for (int kk = 0; kk < KPT; ++kk) {
                                            const int k = kq * KPT + kk;
                                            const float w = __ldg(L.jk_W0T + k * D + col);
#pragma unroll
                                            for (int j = 0; j < 8; ++j) a[j] = fmaf(sPool[(gb + j) * D + k], w, a[j]);
                                        }
#pragma unroll
                                        for (int j = 0; j < 8; ++j) sPart[(kq * 8 + j) * D + col] = a[j];
                                        fm_bar_compute();
                                        for (int j = kq; j < 8 && gb + j < ng; j += GPP) {
                                            float h = 0.f;
#pragma unroll
                                            for (int qq = 0; qq < GPP; ++qq) h += sPart[(qq * 8 + j) * D + col];
                                            sHid[(gb + j) * D + col] = fm_act(fmaf(h + b0, s0, t0), L.jk_act);
                                        }
                                        fm_bar_compute();
                                    }
                                    sIn = sHid;
                                }
                                // out[g, c] (+)= <row, W1[c, :]> + b1[c]: one warp per (graph, output column)
                                for (int pr = cw; pr < ng * P.n_out; pr += FM_COMPUTE_THREADS / 32) {
                                    const int gi = pr / P.n_out, c = pr % P.n_out;
                                    float a = 0.f;
                                    for (int k = lane; k < D; k += 32) a = fmaf(sIn[gi * D + k], __ldg(L.jk_W1 + c * D + k), a);
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                                    if (lane == 0) {
                                        float *dst = P.out + (int64_t)(g0 + gi) * P.n_out + c;
                                        a += __ldg(L.jk_b1 + c);
                                        *dst = L.jk_first ? a : *dst + a;
                                    }
                                }
                            }
                        }
                        if (l + 1 < nL) rx_inv = finish_operand(m, live);
                        else fm_bar_compute();          // every reader of this layer's constants is done
#pragma unroll
                        for (int i = 0; i < NVT; ++i)
                            if (ct + i * FM_COMPUTE_THREADS < NVQ) reinterpret_cast<float4 *>(vecs)[ct + i * FM_COMPUTE_THREADS] = vnext[i];
                        FM_STAMP(l, 9);
                    }
                }
                first_tile = false;
                g0 = g1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
    }
}

typedef CUresult (*FmEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static FmEncodeTiledFn fm_encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        return (FmEncodeTiledFn)p;
    return nullptr;
}

// fp16 [rows, D] row-major, box = [D rows, 64 halves], 128-byte swizzle
static int fm_make_map(CUtensorMap *map, const void *base, int64_t rows, int D) {
    FmEncodeTiledFn fn = fm_encode_fn();
    if (!fn) return GSN_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)D};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled (fused model) failed: %d", (int)r);
        return GSN_E_CUDA;
    }
    return GSN_OK;
}

template <int D>
static int fm_launch(const CUtensorMap &hi, const CUtensorMap &lo, const FmParams &p, cudaStream_t stream) {
    constexpr int smem = FmCfg<D>::SMEM;
    GSN_CUDA_OK(cudaFuncSetAttribute(fused_model_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const unsigned grid = (unsigned)(p.n_units < kNumSMs ? p.n_units : kNumSMs);      // with a tile plan: n_units = its capacity
    fused_model_kernel<D><<<grid, FM_THREADS, smem, stream>>>(hi, lo, p);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("fused_model_kernel");
    return GSN_OK;
}

// Greedy tile plan of a small batch (see gsn_tile_plan in the header): node_ptr staged in shared memory, one warp walks it
// with the same 32-graph probe the model kernel uses inside a unit (fm_tile_end).
__global__ void tile_plan_kernel(const int64_t *__restrict__ node_ptr, int G, int32_t *__restrict__ plan, int cap,
                                 int32_t *status) {
    extern __shared__ int32_t tp_ptr[];
    for (int i = threadIdx.x; i <= G; i += blockDim.x) tp_ptr[i] = (int32_t)node_ptr[i];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    int g0 = 0, t = 0;
    while (g0 < G) {
        const int g = g0 + 1 + lane;
        const bool ok = g <= G && tp_ptr[g] - tp_ptr[g0] <= FM_ROWS;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        int cnt = (m == 0xffffffffu) ? 32 : (__ffs((int)~m) - 1);
        if (cnt == 0) cnt = 1;                      // wider than a tile: alone, the model kernel reports it
        if (lane == 0 && t < cap) plan[1 + t] = g0;
        ++t;
        g0 += cnt;
    }
    if (lane == 0) {
        if (t > cap) {                               // cannot happen with the documented capacity
            atomicOr(status, GSN_S_GRAPH_TOO_LARGE);
            t = cap;
        }
        plan[0] = t;
        plan[1 + t] = G;
    }
}

}  // namespace gsn

using namespace gsn;

#ifdef GSN_PROFILE_STAMPS
extern "C" int gsn_fm_profile_read(long long *h_out) {
    GSN_CUDA_OK(cudaMemcpyFromSymbol(h_out, g_fm_stamps, sizeof(long long) * FM_MAX_LAYERS * 16));
    return GSN_OK;
}
#endif

extern "C" int gsn_fused_model_fwd(const GsnFusedModel *h_m, void *stream_) {
    if (!h_m) return GSN_E_INVALID;
    const GsnFusedModel &m = *h_m;
    if (m.n_layers < 1 || m.n_layers > GSN_FUSED_MAX_LAYERS || (m.D != 64 && m.D != 128) || m.G < 0 || m.N < 0 ||
        !m.d_rowptr || !m.d_node_ptr || !m.d_Whi || !m.d_Wlo || !m.d_status || m.graphs_per_unit < 1 || m.n_mats < 2)
        return GSN_E_INVALID;
    if (m.N + 1 >= (int64_t)1 << 31 || m.G >= (int64_t)1 << 30) return GSN_E_UNSUPPORTED;
    if (m.G == 0 || m.N == 0) return GSN_OK;
    FmParams p;
    p.n_layers = m.n_layers;
    p.rowptr = m.d_rowptr; p.nbr = m.d_nbr; p.node_ptr = m.d_node_ptr;
    p.x0 = m.d_x0; p.x0_ld = m.x0_ld; p.x0_d = m.x0_d;
    p.G = (int32_t)m.G; p.unit = m.graphs_per_unit; p.n_units = (int32_t)ceil_div(m.G, m.graphs_per_unit);
    p.tile_plan = m.d_tile_plan;
    if (m.d_tile_plan) {
        if (m.max_tiles < 1) return GSN_E_INVALID;
        p.n_units = m.max_tiles;          // the grid: CTAs past the plan's tile count exit
    }
    p.status = m.d_status;
    bool any_jk = false;
    for (int l = 0; l < m.n_layers; ++l) {
        const GsnFusedLayer &s = m.layers[l];
        FmLayer &d = p.layer[l];
        if (!s.d_vec || s.n_node_cols < 0 || s.n_edge_cols < 0 || (s.n_node_cols > 0 && (!s.d_node_rows || !s.d_Tn)) ||
            (s.n_edge_cols > 0 && (!s.d_edge_rows || !s.d_Te || s.te_rows < 1)) || (s.d_Tu && !s.d_tu_rows) ||
            s.mat0 < 0 || s.mat0 + (s.has_dense ? 5 : 2) > m.n_mats)
            return GSN_E_INVALID;
        if (s.has_dense && l == 0 && (!m.d_x0 || m.x0_d < 1 || m.x0_d > m.D)) return GSN_E_INVALID;
        if (!s.has_dense && l > 0) return GSN_E_UNSUPPORTED;      // a layer after the first always consumes dense rows
        if (m.E > 0 && !m.d_nbr) return GSN_E_INVALID;
        d.node_rows = s.d_node_rows; d.Tn = s.d_Tn; d.tu_rows = s.d_tu_rows; d.Tu = s.d_Tu; d.edge_rows = s.d_edge_rows;
        d.Te = s.d_Te; d.vec = s.d_vec; d.pooled = s.d_pooled; d.x_out = s.d_x_out;
        d.n_node_cols = s.n_node_cols; d.tu_stride = s.tu_stride; d.n_edge_cols = s.n_edge_cols; d.te_rows = s.te_rows;
        d.has_dense = s.has_dense ? 1 : 0; d.mat0 = s.mat0; d.act_msg = s.act_msg; d.act_upd = s.act_upd; d.act_out = s.act_out;
        d.pool = s.pool;
        d.jk_kind = s.jk_kind; d.jk_act = s.jk_act; d.jk_W0T = s.d_jk_W0T; d.jk_vec = s.d_jk_vec; d.jk_W1 = s.d_jk_W1;
        d.jk_b1 = s.d_jk_b1;
        d.jk_first = 0;
        if (s.jk_kind) {
            if (s.jk_kind < 0 || s.jk_kind > 2 || !s.pool || !s.d_jk_W1 || !s.d_jk_b1 || !m.d_out || m.n_out < 1 || m.n_out > 32 ||
                (s.jk_kind == 2 && (!s.d_jk_W0T || !s.d_jk_vec)))
                return GSN_E_INVALID;
            d.jk_first = any_jk ? 0 : 1;
            any_jk = true;
        }
        if (s.pool && !s.d_pooled && !s.jk_kind) return GSN_E_INVALID;
    }
    p.out = m.d_out; p.n_out = m.n_out;
    CUtensorMap hi, lo;
    int rc;
    if ((rc = fm_make_map(&hi, m.d_Whi, (int64_t)m.n_mats * m.D, m.D))) return rc;
    if ((rc = fm_make_map(&lo, m.d_Wlo, (int64_t)m.n_mats * m.D, m.D))) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m.D == 128) return fm_launch<128>(hi, lo, p, stream);
    return fm_launch<64>(hi, lo, p, stream);
}

extern "C" int gsn_tile_plan(const int64_t *d_node_ptr, int64_t G, int32_t *d_tile_plan, int32_t max_tiles, int32_t *d_status,
                             void *stream_) {
    if (!d_node_ptr || !d_tile_plan || !d_status || G < 0 || max_tiles < 1) return GSN_E_INVALID;
    if (G > 8192) return GSN_E_UNSUPPORTED;
    tile_plan_kernel<<<1, 256, sizeof(int32_t) * (size_t)(G + 1), (cudaStream_t)stream_>>>(d_node_ptr, (int)G, d_tile_plan,
                                                                                          max_tiles, d_status);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("tile_plan_kernel");
    return GSN_OK;
}
