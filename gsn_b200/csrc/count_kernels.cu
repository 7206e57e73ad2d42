// COUNT kernels: one launch enumerates one pattern (or one whole cycle / clique
// family) over every graph of the batch.
//
// Replaces the per-graph loop utils_data_gen.py:60-78 -> utils_ids.py:19-27 ->
// utils_graph_processing.py:103-179 of the reference.
//
// Mapping onto the GPU
//   * the batch is cut into chunks of ~T consecutive nodes aligned to graph
//     boundaries; ONE CTA owns a chunk, so all accumulation is CTA-private:
//     no global atomics, deterministic, and the result rows of a chunk are
//     written once, coalesced;
//   * the chunk's adjacency bitmasks and slot offsets are contiguous in HBM and
//     are staged into shared memory with two 1-D bulk copies (TMA, UBLKCP)
//     completing on an mbarrier;
//   * work items (directed edges = slots of the chunk) are handed to threads
//     through a shared-memory ticket counter; each thread runs the bitmask DFS
//     of count_core.cuh with 64-bit set operations (popc / ffs);
//   * accumulators are uint32 in shared memory; chunks that do not fit fall
//     back to global atomics on rows only this CTA touches.
#include "common.cuh"
#include "count_core.cuh"

namespace gsn {

struct CountParams {
    const uint64_t *adj;
    const int32_t *nbase;
    const int32_t *rowptr;
    const int32_t *slot_src;
    const int32_t *slot_dst;
    const int64_t *node_ptr;
    int64_t G, N;
    int32_t T;               // nodes per chunk (before rounding to graph boundaries)
    int32_t smem_adj_words;  // capacity of the adjacency stage (uint64 words), 0 = read HBM
    int32_t smem_row_words;  // capacity of the rowptr stage (int32)
    int32_t smem_acc_words;  // capacity of the accumulator stage (uint32)
    int32_t parts;           // sub-items per directed edge (small batches: shorter critical path)
    int32_t warp_items;      // 1: heavy directed edges are deferred to count_heavy_kernel (whole warps, all SMs)
    int32_t *heavy;          // global: [0] number of deferred items, [1] work ticket, [4..] their slots
    int64_t *out;            // vertex scope: [N, out_ld] ; edge scope: unused here
    int64_t out_ld;
    uint32_t *slot_acc;      // edge scope: [S, n_cols]
    int32_t *status;
    GsnPlan plan;
};

// 32-bit accumulators: a carry out of bit 31 sets GSN_S_COUNT_OVERFLOW (raise_on_status reports it) instead of wrapping
// silently; the shared-memory and the global accumulation paths detect it the same way.
struct SmemAcc {
    uint32_t *acc;     // chunk-local [rows, C]
    int C;
    int vbase;         // local vertex -> row: vbase + v
    int sbase;         // global slot  -> row: slot - sbase
    int32_t *status;
    __device__ __forceinline__ void add(uint32_t *p, uint32_t c) {
        const uint32_t old = atomicAdd(p, c);
        if (old + c < old) atomicOr(status, GSN_S_COUNT_OVERFLOW);
    }
    __device__ __forceinline__ void vertex(int lv, int col, uint32_t c) { add(&acc[(vbase + lv) * C + col], c); }
    __device__ __forceinline__ void slot(int s, int col, uint32_t c) { add(&acc[(s - sbase) * C + col], c); }
    __device__ __forceinline__ void overflow() { atomicOr(status, GSN_S_COUNT_OVERFLOW); }
};

struct GlobalAcc {
    unsigned long long *out;   // vertex scope rows (int64), row stride ld, already offset to col0
    int64_t ld;
    int64_t vrow0;             // global node id of local vertex 0
    uint32_t *slot_acc;
    int C;
    int32_t *status;
    __device__ __forceinline__ void vertex(int lv, int col, uint32_t c) {
        // int64 output rows: the value is checked against the 32-bit range of the shared-memory path so that both
        // paths report the same condition whatever the chunking
        const unsigned long long old = atomicAdd(&out[(vrow0 + lv) * ld + col], (unsigned long long)c);
        if (old + c > 0xFFFFFFFFull) atomicOr(status, GSN_S_COUNT_OVERFLOW);
    }
    __device__ __forceinline__ void slot(int s, int col, uint32_t c) {
        const uint32_t old = atomicAdd(&slot_acc[(size_t)s * C + col], c);
        if (old + c < old) atomicOr(status, GSN_S_COUNT_OVERFLOW);
    }
    __device__ __forceinline__ void overflow() { atomicOr(status, GSN_S_COUNT_OVERFLOW); }
};

template <int W, class Acc>
__device__ __forceinline__ void run_item(const GsnPlan &P, const GraphView<W> &G, int a, int b, Acc &acc, int part, int parts) {
    if (P.family == GSN_FAMILY_CYCLES) enumerate_cycles<W>(P.kmin, P.kmax, P.induced, P.scope, G, a, b, acc, part, parts);
    else if (P.family == GSN_FAMILY_CLIQUES) enumerate_cliques<W>(P.kmin, P.kmax, P.scope, G, a, b, acc, part, parts);
    else enumerate_generic<W>(P, G, a, b, acc, part, parts);
}

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t *a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;   // first index with a[i] >= key
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int kCountThreads = 128;
constexpr int kHeavyItem = 12;      // third-level candidates from which an item is deferred to count_heavy_kernel
constexpr int kHeavyClique = 8;     // same for cliques (an item counts inside its pool of common neighbours); a heavy clique item is
                                    // one warp's task, the lanes take the pool's vertices as first choices
constexpr int kHeavyBatch = 4;       // consecutive tasks per ticket
constexpr int kHeavySplit = 4;       // warps per heavy item (x 32 lanes = 128 shares of its third-level candidates)

template <int W>
__global__ void __launch_bounds__(kCountThreads) count_kernel(const __grid_constant__ CountParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int ticket;
    __shared__ int64_t sh_range[2];

    const GsnPlan &P = prm.plan;
    const int C = P.n_cols;

    if (threadIdx.x == 0) {
        int64_t lo_node = (int64_t)blockIdx.x * prm.T;
        int64_t hi_node = lo_node + prm.T;
        if (hi_node > prm.N) hi_node = prm.N;
        // graphs whose first node lies in [lo_node, hi_node)
        int64_t g_lo = lower_bound_i64(prm.node_ptr, prm.G + 1, lo_node);
        int64_t g_hi = lower_bound_i64(prm.node_ptr, prm.G + 1, hi_node);
        if (g_lo > prm.G) g_lo = prm.G;
        if (g_hi > prm.G) g_hi = prm.G;
        sh_range[0] = prm.node_ptr[g_lo];
        sh_range[1] = prm.node_ptr[g_hi];
        ticket = 0;
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t v0 = sh_range[0], v1 = sh_range[1];
    if (v1 <= v0) return;
    const int nn = (int)(v1 - v0);
    const int s0 = prm.rowptr[v0], s1 = prm.rowptr[v1];
    const int ns = s1 - s0;
    const int rows = P.scope == 0 ? nn : ns;

    // ---- carve shared memory: [adj words | rowptr | accumulators]
    uint64_t *sm_adj = (uint64_t *)smem_raw;
    int32_t *sm_row = (int32_t *)(smem_raw + (size_t)prm.smem_adj_words * 8);
    uint32_t *sm_acc = (uint32_t *)(smem_raw + (size_t)prm.smem_adj_words * 8 + (size_t)prm.smem_row_words * 4);

    // adjacency words [v0*W, v1*W) widened to 16-byte boundaries
    const int64_t aw0 = (v0 * W) & ~(int64_t)1;
    const int64_t aw1 = (v1 * W + 1) & ~(int64_t)1;
    const int64_t rw0 = v0 & ~(int64_t)3;
    const int64_t rw1 = (v1 + 1 + 3) & ~(int64_t)3;          // rowptr[v0..v1] inclusive
    const bool stage = (aw1 - aw0) <= prm.smem_adj_words && (rw1 - rw0) <= prm.smem_row_words;
    const bool acc_in_smem = (int64_t)rows * C <= prm.smem_acc_words;

    if (stage) {
        if (threadIdx.x == 0) {
            uint32_t bytes_a = (uint32_t)((aw1 - aw0) * 8), bytes_r = (uint32_t)((rw1 - rw0) * 4);
            mbar_arrive_expect_tx(&bar, bytes_a + bytes_r);
            bulk_g2s(sm_adj, prm.adj + aw0, bytes_a, &bar);
            bulk_g2s(sm_row, prm.rowptr + rw0, bytes_r, &bar);
        }
    }
    if (acc_in_smem) {
        for (int i = threadIdx.x; i < rows * C; i += kCountThreads) sm_acc[i] = 0;
    } else if (P.scope == 0) {
        for (int64_t i = threadIdx.x; i < (int64_t)nn * C; i += kCountThreads)
            prm.out[(v0 + i / C) * prm.out_ld + P.col0 + i % C] = 0;
    } else {
        for (int64_t i = threadIdx.x; i < (int64_t)ns * C; i += kCountThreads) prm.slot_acc[(size_t)s0 * C + i] = 0;
    }
    __syncthreads();
    if (stage) mbar_wait(&bar, 0);

    // views: local vertex ids are relative to each graph's first node
    const uint64_t *adj_base = stage ? sm_adj + (v0 * W - aw0) : prm.adj + v0 * W;       // row of node v0
    const int32_t *row_base = stage ? sm_row + (v0 - rw0) : prm.rowptr + v0;             // rowptr of node v0

    auto process = [&](int s, int part, int nparts) {
        const int a = prm.slot_src[s], b = prm.slot_dst[s];
        const int gb = prm.nbase[a];                      // first node of the graph
        const int off = (int)(gb - v0);
        // a graph of <= 64 vertices inside a batch laid out for wider ones (node gb + 64 is not its own): one-word sets
        const bool narrow = W > 1 && (gb + 64 >= prm.N || prm.nbase[gb + 64] != gb);
        GraphView<W> G{adj_base + (size_t)off * W, row_base + off};
        GraphView<1> G1{adj_base + (size_t)off * W, row_base + off, W};
        if (acc_in_smem) {
            SmemAcc acc{sm_acc, C, off, s0, prm.status};
            if (narrow) run_item<1>(P, G1, a - gb, b - gb, acc, part, nparts);
            else run_item<W>(P, G, a - gb, b - gb, acc, part, nparts);
        } else {
            GlobalAcc acc{(unsigned long long *)(prm.out + P.col0), prm.out_ld, (int64_t)gb, prm.slot_acc, C, prm.status};
            if (narrow) run_item<1>(P, G1, a - gb, b - gb, acc, part, nparts);
            else run_item<W>(P, G, a - gb, b - gb, acc, part, nparts);
        }
    };
    // how many third-level candidates an item has (cliques: common neighbours; otherwise the degree of b): an item
    // above the threshold is HEAVY -- a directed edge of a dense graph can root thousands of occurrences (IMDB-BINARY:
    // 5,576 K5 on one edge) and would be a serial tail on one thread
    auto weight = [&](int s) -> int {
        const int a = prm.slot_src[s], b = prm.slot_dst[s];
        const int gb = prm.nbase[a];
        const uint64_t *ra = adj_base + (size_t)(a - v0) * W, *rb = adj_base + (size_t)(b - v0) * W;
        (void)gb;
        if (P.family == GSN_FAMILY_CLIQUES && P.scope != 0 && b <= a) return 0;     // the a < b item counts for both slots
        int c = 0;
        for (int i = 0; i < W; ++i) c += __popcll(P.family == GSN_FAMILY_CLIQUES ? (ra[i] & rb[i]) : rb[i]);
        return c;
    };
    const int heavy_min = P.family == GSN_FAMILY_CLIQUES ? kHeavyClique : kHeavyItem;
    const int parts = prm.parts;
    while (true) {
        const int it = atomicAdd(&ticket, 1);
        if (it >= ns * parts) break;
        const int s = s0 + it / parts;
        if (prm.warp_items && weight(s) >= heavy_min) {
            // deferred to count_heavy_kernel: one CTA owns this chunk, but a heavy graph (IMDB-BINARY: 136 nodes, every
            // edge in hundreds of K5) needs the whole machine, not four warps
            if (it % parts == 0) prm.heavy[4 + atomicAdd(&prm.heavy[0], 1)] = s;
            continue;
        }
        process(s, it % parts, parts);
    }
    if (!acc_in_smem) return;
    __syncthreads();
    if (P.scope == 0) {
        for (int i = threadIdx.x; i < nn * C; i += kCountThreads)
            prm.out[(v0 + i / C) * prm.out_ld + P.col0 + i % C] = (int64_t)sm_acc[i];
    } else {
        uint32_t *dst = prm.slot_acc + (size_t)s0 * C;
        for (int i = threadIdx.x; i < ns * C; i += kCountThreads) dst[i] = sm_acc[i];
    }
}

// Heavy items (deferred by count_kernel): every warp takes (item, share) tasks from a global ticket; the item's
// third-level candidates are split 32 * kHeavySplit ways over lanes and warps, the graph is read from global memory (L2)
// and the counts are added with atomics to the rows count_kernel has already written.
template <int W>
__global__ void __launch_bounds__(256) count_heavy_kernel(const __grid_constant__ CountParams prm) {
    // per warp: the adjacency rows and slot offsets of the graph its current task belongs to (W <= 4: <= 256 vertices,
    // 9 KB); the search does one dependent adjacency read per search node, from L2 that was the whole cost of a task
    extern __shared__ __align__(16) unsigned char hv_smem[];
    constexpr int MAXN = 64 * W;
    constexpr bool STAGE = W <= 4;
    const GsnPlan &P = prm.plan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *wadj = (uint64_t *)hv_smem + (size_t)warp * MAXN * W;
    int32_t *wrow = (int32_t *)(hv_smem + (size_t)8 * MAXN * W * 8) + (size_t)warp * (MAXN + 4);
    const int split = P.family == GSN_FAMILY_CLIQUES ? 1 : kHeavySplit;
    const int batch = P.family == GSN_FAMILY_CLIQUES ? kHeavyBatch : 1;      // the shares of a split item go to different warps
    const int n_tasks = prm.heavy[0] * split;
    int staged = -1;
    // tickets are taken kHeavyBatch at a time: neighbouring entries of the list mostly come from one graph (one chunk
    // of count_kernel pushed them), so the staged copy is reused and the ticket's round trip is shared
    int t_next = 0, t_end = 0;
    while (true) {
        if (t_next == t_end) {
            t_next = lane == 0 ? atomicAdd(&prm.heavy[1], batch) : 0;
            t_next = __shfl_sync(0xffffffffu, t_next, 0);
            t_end = t_next + batch;
        }
        const int t = t_next++;
        if (t >= n_tasks) break;
        const int s = prm.heavy[4 + t / split];
        const int part = (t % split) * 32 + lane;
        const int a = prm.slot_src[s], b = prm.slot_dst[s];
        const int gb = prm.nbase[a];
        if (STAGE && gb != staged) {
            __syncwarp();
            // upper bound on the graph's size from 32 probes of the node -> first-node map
            constexpr int STEP = MAXN / 32;
            const int64_t probe = (int64_t)gb + STEP * (lane + 1);
            const unsigned past = __ballot_sync(0xffffffffu, probe >= prm.N || prm.nbase[probe] != gb);
            const int n = (int)min((int64_t)(past ? STEP * __ffs(past) : MAXN), prm.N - gb);
            for (int i = lane; i < n * W; i += 32) wadj[i] = prm.adj[(size_t)gb * W + i];
            for (int i = lane; i <= n; i += 32) wrow[i] = prm.rowptr[gb + i];
            staged = gb;
            __syncwarp();
        }
        const bool narrow = W > 1 && (gb + 64 >= prm.N || prm.nbase[gb + 64] != gb);       // see count_kernel
        GraphView<W> G{STAGE ? wadj : prm.adj + (size_t)gb * W, STAGE ? wrow : prm.rowptr + gb};
        GraphView<1> G1{G.adj, G.rowptr, W};
        GlobalAcc acc{(unsigned long long *)(prm.out + P.col0), prm.out_ld, (int64_t)gb, prm.slot_acc, P.n_cols, prm.status};
        if (narrow) run_item<1>(P, G1, a - gb, b - gb, acc, part, 32 * split);
        else run_item<W>(P, G, a - gb, b - gb, acc, part, 32 * split);
    }
}

// edge scope epilogue: rows follow edge_index columns (utils_graph_processing.py:142-144,173)
__global__ void k_edge_out(const int64_t *__restrict__ src, const int64_t *__restrict__ dst, int64_t E, int64_t N, int W,
                           const int32_t *__restrict__ nbase, const uint64_t *__restrict__ adj,
                           const int32_t *__restrict__ rowptr, const int32_t *__restrict__ slot_col,
                           const uint32_t *__restrict__ slot_acc, int C, int64_t *__restrict__ out, int64_t ld,
                           int col0) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t a = src[e], b = dst[e];
    int s = -1;
    if (a >= 0 && b >= 0 && a < N && b < N && a != b) {
        int base = nbase[a];
        int lb = (int)(b - base);
        if (nbase[b] == base && lb < 64 * W && a - base < 64 * W) {
            int r = 0;
            const uint64_t *p = adj + (size_t)a * W;
            for (int i = 0; i < W; ++i) {
                int lo = i * 64;
                if (lb >= lo + 64) r += __popcll(p[i]);
                else if (lb > lo) r += __popcll(p[i] & ((1ull << (lb - lo)) - 1ull));
            }
            s = rowptr[a] + r;
            if (slot_col[s] != (int32_t)e) s = -1;       // an earlier duplicate column: edge_dict forgot it
        }
    }
    for (int c = 0; c < C; ++c) out[e * ld + col0 + c] = s >= 0 ? (int64_t)slot_acc[(size_t)s * C + c] : 0;
}

// a slot that matches used but edge_index never listed: the reference's KeyError
__global__ void k_missing_check(const int32_t *__restrict__ rowptr, int64_t N, const int32_t *__restrict__ slot_col,
                                const uint32_t *__restrict__ slot_acc, int C, int32_t *status) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rowptr[N]) return;
    if (slot_col[s] >= 0) return;
    for (int c = 0; c < C; ++c)
        if (slot_acc[(size_t)s * C + c]) { atomicOr(status, GSN_S_MISSING_EDGE); return; }
}

template <int W>
int launch_count(const CountParams &prm, int64_t chunks, size_t smem, cudaStream_t stream) {
    // per launch (cheap): the attribute belongs to the current device
    GSN_CUDA_OK(cudaFuncSetAttribute(count_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (prm.warp_items) GSN_CUDA_OK(cudaMemsetAsync(prm.heavy, 0, 16, stream));
    count_kernel<W><<<(unsigned)chunks, kCountThreads, smem, stream>>>(prm);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("count_kernel");
    if (prm.warp_items) {
        const size_t hsm = W <= 4 ? (size_t)8 * (64 * W * W * 8 + (64 * W + 4) * 4) : 0;
        GSN_CUDA_OK(cudaFuncSetAttribute(count_heavy_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        count_heavy_kernel<W><<<kNumSMs * 2, 256, hsm, stream>>>(prm);
        GSN_BUMP(1);
        GSN_LAUNCH_OK("count_heavy_kernel");
    }
    return GSN_OK;
}

}  // namespace gsn

using namespace gsn;

extern "C" int gsn_count_scratch_bytes(int64_t N, int64_t E, const GsnPlan *h_plan, size_t *bytes) {
    if (!h_plan || !bytes || N < 0 || E < 0) return GSN_E_INVALID;
    // [per-slot accumulators (edge scope)] [deferred heavy items: count, ticket, 2E slots]
    const size_t acc = h_plan->scope == 1 ? align_up(sizeof(uint32_t) * (size_t)(2 * E + 4) * (size_t)h_plan->n_cols, 256) : 0;
    *bytes = acc + align_up(sizeof(int32_t) * (size_t)(2 * E + 8), 256);
    return GSN_OK;
}

extern "C" int gsn_count_pattern(const void *d_ws, int64_t N, int64_t E, int32_t W, const int64_t *d_edge_index,
                                 const int64_t *d_node_ptr, int64_t G, const GsnPlan *h_plan, int64_t *d_out,
                                 int64_t out_ld, void *d_scratch, size_t scratch_bytes, int32_t *d_status,
                                 void *stream_) {
    if (!d_ws || !h_plan || !d_status || !d_node_ptr || N < 0 || E < 0 || W < 1) return GSN_E_INVALID;
    const GsnPlan &P = *h_plan;
    if (N == 0 || (P.scope == 1 && E == 0)) return GSN_OK;       // no output rows
    if (!d_out) return GSN_E_INVALID;
    if (P.k < 2 || P.k > GSN_MAXK || P.n_cols < 1 || P.col0 < 0 || P.col0 + P.n_cols > out_ld) return GSN_E_INVALID;
    if (P.family != GSN_FAMILY_GENERIC && (P.kmin < 3 || P.kmax > GSN_MAXK || P.kmax < P.kmin)) return GSN_E_INVALID;
    if (W != 1 && W != 2 && W != 4 && W != 8 && W != 16) return GSN_E_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0 || (P.scope == 1 && E == 0)) return GSN_OK;
    GraphLayout L = graph_layout(N, E, W);
    const char *ws = (const char *)d_ws;

    CountParams prm;
    prm.adj = (const uint64_t *)(ws + L.adj);
    prm.nbase = (const int32_t *)(ws + L.nbase);
    prm.rowptr = (const int32_t *)(ws + L.rowptr);
    prm.slot_src = (const int32_t *)(ws + L.slot_src);
    prm.slot_dst = (const int32_t *)(ws + L.slot_dst);
    prm.node_ptr = d_node_ptr;
    prm.G = G;
    prm.N = N;
    prm.out = d_out;
    prm.out_ld = out_ld;
    prm.slot_acc = (uint32_t *)d_scratch;
    prm.status = d_status;
    prm.plan = P;
    {
        size_t need = 0;
        gsn_count_scratch_bytes(N, E, h_plan, &need);
        if (!d_scratch || scratch_bytes < need) return GSN_E_WORKSPACE;
        const size_t acc = P.scope == 1 ? align_up(sizeof(uint32_t) * (size_t)(2 * E + 4) * (size_t)P.n_cols, 256) : 0;
        prm.heavy = (int32_t *)((char *)d_scratch + acc);
    }

    // chunking: enough CTAs to cover the machine several times, enough items per CTA to fill it
    const double avg_deg = N > 0 ? (double)(2 * E) / (double)N : 1.0;   // upper bound on slots per node
    int64_t T = (int64_t)(1024.0 / (avg_deg > 1.0 ? avg_deg : 1.0));    // ~1024 items per chunk
    int64_t t_fill = N / (kNumSMs * 8);
    if (T > t_fill) T = t_fill;
    if (T < 16) T = 16;
    if (T > 2048) T = 2048;
    const int64_t max_n = (int64_t)64 * W;
    const int64_t chunks = ceil_div(N, T);
    // shared-memory budget (chunk covers < T + max_n nodes)
    const int64_t node_cap = T + max_n;
    int64_t adj_words = (node_cap * W + 4) & ~(int64_t)1;
    int64_t row_words = (node_cap + 8 + 3) & ~(int64_t)3;
    int64_t rows_est = P.scope == 0 ? node_cap : (int64_t)(node_cap * avg_deg) + 64;
    int64_t acc_words = rows_est * P.n_cols;
    const int64_t budget = 96 * 1024;
    if (adj_words * 8 + row_words * 4 > budget / 2) { adj_words = 0; row_words = 0; }
    int64_t left = budget - adj_words * 8 - row_words * 4;
    if (acc_words * 4 > left) acc_words = left / 4;
    prm.T = (int32_t)T;
    // few items (small batch): split every directed edge into sub-items so that more threads share the search
    prm.parts = E < (int64_t)kNumSMs * 128 ? 8 : (E < (int64_t)kNumSMs * 512 ? 4 : (E < (int64_t)kNumSMs * 2048 ? 2 : 1));
    // average degree >= 8: skewed item weights -> heavy items are deferred to whole warps
    prm.warp_items = avg_deg >= 8.0 ? 1 : 0;
    prm.smem_adj_words = (int32_t)adj_words;
    prm.smem_row_words = (int32_t)row_words;
    prm.smem_acc_words = (int32_t)acc_words;
    const size_t smem = (size_t)adj_words * 8 + (size_t)row_words * 4 + (size_t)acc_words * 4;

    int rc;
    switch (W) {
        case 1: rc = launch_count<1>(prm, chunks, smem, stream); break;
        case 2: rc = launch_count<2>(prm, chunks, smem, stream); break;
        case 4: rc = launch_count<4>(prm, chunks, smem, stream); break;
        case 8: rc = launch_count<8>(prm, chunks, smem, stream); break;
        default: rc = launch_count<16>(prm, chunks, smem, stream); break;
    }
    if (rc) return rc;
    if (P.scope == 1) {
        const int TB = 256;
        k_edge_out<<<(unsigned)ceil_div(E, TB), TB, 0, stream>>>(
            d_edge_index, d_edge_index + E, E, N, W, prm.nbase, prm.adj, prm.rowptr,
            (const int32_t *)(ws + L.slot_col), prm.slot_acc, P.n_cols, d_out, out_ld, P.col0);
        k_missing_check<<<(unsigned)ceil_div(2 * E, TB), TB, 0, stream>>>(
            prm.rowptr, N, (const int32_t *)(ws + L.slot_col), prm.slot_acc, P.n_cols, d_status);
        GSN_BUMP(2);
        GSN_LAUNCH_OK("k_edge_out");
    }
    return GSN_OK;
}
