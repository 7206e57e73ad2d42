// Batched simple-graph construction for the COUNT path.
//
// Replaces, for the whole batch at once, the per-graph
//   G_gt = gt.Graph(directed=False); add_edge_list; remove_self_loops; remove_parallel_edges
// of /root/reference/utils_graph_processing.py:110-113 / :150-153 and the
// edge_dict of :142-144.
//
// Layout in HBM (GraphLayout, common.cuh):
//   nbase[v]    int32   first node of v's graph (local id of v = v - nbase[v])
//   adj[v*W..]  uint64  adjacency bitmask over LOCAL ids: OR-ing both directions
//                       symmetrises, de-duplicates and drops self loops for free
//   rowptr[v]   int32   first "slot" of v; slots = directed edges of the simple
//                       graph, neighbours ascending (= popcount prefix of adj)
//   slot_src/dst int32  global end points of every slot (work items of the kernels)
//   slot_col    int32   LAST edge_index column (a,b) mapping to the slot, -1 none
#include "common.cuh"
#include "count_core.cuh"

namespace gsn {

thread_local char g_last_error[256] = "";
unsigned long long g_launches = 0;

// ------------------------------------------------------------------ scan
__global__ void scan_block_sums(const int32_t *__restrict__ in, int32_t *__restrict__ sums, int64_t n) {
    __shared__ int32_t warp_sums[8];
    int64_t base = (int64_t)blockIdx.x * 1024;
    int32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t idx = base + threadIdx.x + i * 256;
        if (idx < n) v += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t s = 0;
        for (int i = 0; i < 8; ++i) s += warp_sums[i];
        sums[blockIdx.x] = s;
    }
}

__global__ void scan_sums_inplace(int32_t *sums, int64_t nb) {
    // single block, exclusive scan with running carry
    __shared__ int32_t buf[1024];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        int64_t idx = base + threadIdx.x;
        int32_t v = idx < nb ? sums[idx] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int32_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        int32_t incl = buf[threadIdx.x];
        if (idx < nb) sums[idx] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

__global__ void scan_apply(const int32_t *in, int32_t *out, const int32_t *__restrict__ sums,
                           int64_t n) {
    // 1024 elements per block, thread t owns elements 4t..4t+3 of the block
    __shared__ int32_t warp_off[8];
    int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    int32_t v[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_off[threadIdx.x >> 5] = incl;
    __syncthreads();
    int32_t woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_off[w];
    int32_t run = sums[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

int exclusive_scan_i32(const int32_t *d_in, int32_t *d_out, int64_t n, int32_t *d_tmp, cudaStream_t stream) {
    if (n <= 0) return GSN_OK;
    int64_t nb = ceil_div(n, 1024);
    scan_block_sums<<<(unsigned)nb, 256, 0, stream>>>(d_in, d_tmp, n);
    scan_sums_inplace<<<1, 1024, 0, stream>>>(d_tmp, nb);
    scan_apply<<<(unsigned)nb, 256, 0, stream>>>(d_in, d_out, d_tmp, n);
    GSN_BUMP(3);
    GSN_LAUNCH_OK("exclusive_scan_i32");
    return GSN_OK;
}

// ------------------------------------------------------------ graph build
__global__ void k_node_base(const int64_t *__restrict__ node_ptr, int64_t G, int64_t N, int W,
                            int32_t *__restrict__ nbase, int32_t *status) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    // last g with node_ptr[g] <= v
    int64_t lo = 0, hi = G;   // invariant: node_ptr[lo] <= v < node_ptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (node_ptr[mid] <= v) lo = mid; else hi = mid;
    }
    int64_t b = node_ptr[lo];
    nbase[v] = (int32_t)b;
    if (v == b && node_ptr[lo + 1] - b > (int64_t)64 * W) atomicOr(status, GSN_S_GRAPH_TOO_LARGE);
}

__global__ void k_set_bits(const int64_t *__restrict__ src, const int64_t *__restrict__ dst, int64_t E, int64_t N,
                           int W, const int32_t *__restrict__ nbase, unsigned long long *adj, int32_t *status) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t a = src[e], b = dst[e];
    if (a < 0 || b < 0 || a >= N || b >= N) { atomicOr(status, GSN_S_INDEX_RANGE); return; }
    if (a == b) return;
    int32_t base = nbase[a];
    if (nbase[b] != base) { atomicOr(status, GSN_S_CROSS_GRAPH_EDGE); return; }
    int la = (int)(a - base), lb = (int)(b - base);
    if (la >= 64 * W || lb >= 64 * W) return;   // GRAPH_TOO_LARGE already flagged
    atomicOr(&adj[(size_t)a * W + (lb >> 6)], 1ull << (lb & 63));
    atomicOr(&adj[(size_t)b * W + (la >> 6)], 1ull << (la & 63));
}

__global__ void k_degree(const uint64_t *__restrict__ adj, int64_t N, int W, int32_t *__restrict__ deg) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > N) return;
    int d = 0;
    if (v < N)
        for (int w = 0; w < W; ++w) d += __popcll(adj[(size_t)v * W + w]);
    deg[v] = d;   // deg[N] = 0 so that the exclusive scan yields rowptr[N] = total
}

__global__ void k_fill_slots(const uint64_t *__restrict__ adj, const int32_t *__restrict__ nbase,
                             const int32_t *__restrict__ rowptr, int64_t N, int W, int32_t *__restrict__ slot_src,
                             int32_t *__restrict__ slot_dst, int32_t *__restrict__ slot_col) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    int s = rowptr[v];
    int base = nbase[v];
    for (int w = 0; w < W; ++w) {
        uint64_t x = adj[(size_t)v * W + w];
        while (x) {
            int b = __ffsll((long long)x) - 1;
            x &= x - 1;
            slot_src[s] = (int32_t)v;
            slot_dst[s] = base + w * 64 + b;
            slot_col[s] = -1;
            ++s;
        }
    }
}

template <int W>
__device__ __forceinline__ int slot_of(const uint64_t *adj, const int32_t *rowptr, int64_t a, int lb) {
    return rowptr[a] + rank_below<W>(adj + (size_t)a * W, lb);
}

__device__ __forceinline__ int slot_of_rt(const uint64_t *adj, const int32_t *rowptr, int W, int64_t a, int lb) {
    int r = 0;
    const uint64_t *p = adj + (size_t)a * W;
    for (int i = 0; i < W; ++i) {
        int lo = i * 64;
        if (lb >= lo + 64) r += __popcll(p[i]);
        else if (lb > lo) r += __popcll(p[i] & ((1ull << (lb - lo)) - 1ull));
    }
    return rowptr[a] + r;
}

// edge_dict of utils_graph_processing.py:142-144: the LAST column wins
__global__ void k_slot_cols(const int64_t *__restrict__ src, const int64_t *__restrict__ dst, int64_t E, int64_t N,
                            int W, const int32_t *__restrict__ nbase, const uint64_t *__restrict__ adj,
                            const int32_t *__restrict__ rowptr, int32_t *slot_col) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t a = src[e], b = dst[e];
    if (a < 0 || b < 0 || a >= N || b >= N || a == b) return;
    int base = nbase[a];
    if (nbase[b] != base) return;
    int lb = (int)(b - base);
    if (lb >= 64 * W || a - base >= 64 * W) return;
    atomicMax(&slot_col[slot_of_rt(adj, rowptr, W, a, lb)], (int32_t)e);
}

// Small batches: the whole graph build in ONE CTA (see csr_build_small_kernel): node bases, adjacency bits,
// degrees + scan, slots and the edge_dict, with the adjacency words and the scan in shared memory.
constexpr int kGraphSmallMaxWords = 8192;     // N * W adjacency words (64 KB) ...
constexpr int kGraphSmallMaxN = 8192;         // ... and N + 1 row offsets (32 KB)
__global__ void __launch_bounds__(1024) graph_build_small_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                                 int E, const int64_t *__restrict__ node_ptr, int G, int N, int W,
                                                                 int32_t *__restrict__ nbase, uint64_t *__restrict__ adj,
                                                                 int32_t *__restrict__ rowptr, int32_t *__restrict__ slot_src,
                                                                 int32_t *__restrict__ slot_dst, int32_t *__restrict__ slot_col,
                                                                 int32_t *status) {
    extern __shared__ unsigned long long smg[];          // adj[N*W] | rp[N+1] (int32) | nb[N] (int32)
    unsigned long long *sadj = smg;
    int32_t *rp = (int32_t *)(smg + (size_t)N * W);
    int32_t *nb = rp + (N + 1);
    __shared__ int32_t warp_tot[32];
    const int tid = threadIdx.x;
    for (int i = tid; i < N * W; i += 1024) sadj[i] = 0ull;
    for (int v = tid; v < N; v += 1024) {
        int lo = 0, hi = G;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (node_ptr[mid] <= v) lo = mid; else hi = mid;
        }
        const int64_t b = node_ptr[lo];
        nb[v] = (int32_t)b;
        nbase[v] = (int32_t)b;
        if (v == b && node_ptr[lo + 1] - b > (int64_t)64 * W) atomicOr(status, GSN_S_GRAPH_TOO_LARGE);
    }
    __syncthreads();
    for (int e = tid; e < E; e += 1024) {
        const int64_t a = src[e], b = dst[e];
        if (a < 0 || b < 0 || a >= N || b >= N) { atomicOr(status, GSN_S_INDEX_RANGE); continue; }
        if (a == b) continue;
        const int base = nb[a];
        if (nb[b] != base) { atomicOr(status, GSN_S_CROSS_GRAPH_EDGE); continue; }
        const int la = (int)(a - base), lb = (int)(b - base);
        if (la >= 64 * W || lb >= 64 * W) continue;
        atomicOr(&sadj[(size_t)a * W + (lb >> 6)], 1ull << (lb & 63));
        atomicOr(&sadj[(size_t)b * W + (la >> 6)], 1ull << (la & 63));
    }
    __syncthreads();
    // degrees -> exclusive scan (each thread owns a contiguous span of vertices)
    const int span = (N + 1 + 1023) / 1024;
    const int b0 = tid * span, b1 = min(b0 + span, N + 1);
    int local = 0;
    for (int v = b0; v < b1; ++v) {
        int d = 0;
        if (v < N)
            for (int w = 0; w < W; ++w) d += __popcll(sadj[(size_t)v * W + w]);
        rp[v] = d;
        local += d;
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int v = warp_tot[tid], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, iv, o);
            if (tid >= o) iv += t;
        }
        warp_tot[tid] = iv - v;
    }
    __syncthreads();
    int run = warp_tot[tid >> 5] + incl - local;
    for (int v = b0; v < b1; ++v) {
        const int d = rp[v];
        rp[v] = run;
        run += d;
    }
    __syncthreads();
    for (int v = tid; v <= N; v += 1024) rowptr[v] = rp[v];
    for (int i = tid; i < N * W; i += 1024) adj[i] = sadj[i];
    for (int v = tid; v < N; v += 1024) {
        int s = rp[v];
        const int base = nb[v];
        for (int w = 0; w < W; ++w) {
            unsigned long long x = sadj[(size_t)v * W + w];
            while (x) {
                const int b = __ffsll((long long)x) - 1;
                x &= x - 1;
                slot_src[s] = v;
                slot_dst[s] = base + w * 64 + b;
                slot_col[s] = -1;
                ++s;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < E; e += 1024) {
        const int64_t a = src[e], b = dst[e];
        if (a < 0 || b < 0 || a >= N || b >= N || a == b) continue;
        const int base = nb[a];
        if (nb[b] != base) continue;
        const int lb = (int)(b - base);
        if (lb >= 64 * W || a - base >= 64 * W) continue;
        int r = 0;
        for (int i = 0; i < W; ++i) {
            const int lo = i * 64;
            const unsigned long long word = sadj[(size_t)a * W + i];
            if (lb >= lo + 64) r += __popcll(word);
            else if (lb > lo) r += __popcll(word & ((1ull << (lb - lo)) - 1ull));
        }
        atomicMax(&slot_col[rp[a] + r], e);
    }
}

}  // namespace gsn

using namespace gsn;

extern "C" int gsn_graph_workspace_bytes(int64_t N, int64_t E, int32_t W, size_t *bytes) {
    if (!bytes || N < 0 || E < 0 || W < 1) return GSN_E_INVALID;
    if (N + 1 >= (int64_t)1 << 31 || 2 * E >= (int64_t)1 << 31) return GSN_E_UNSUPPORTED;
    *bytes = graph_layout(N, E, W).total;
    return GSN_OK;
}

extern "C" int gsn_graph_build(const int64_t *d_edge_index, int64_t E, const int64_t *d_node_ptr, int64_t G,
                               int64_t N, int32_t W, void *d_ws, size_t ws_bytes, int32_t *d_status,
                               void *stream_) {
    if (N < 0 || E < 0 || G < 0 || W < 1 || !d_ws || !d_status || (E > 0 && !d_edge_index) || !d_node_ptr)
        return GSN_E_INVALID;
    if (N + 1 >= (int64_t)1 << 31 || 2 * E >= (int64_t)1 << 31) return GSN_E_UNSUPPORTED;
    GraphLayout L = graph_layout(N, E, W);
    if (ws_bytes < L.total) return GSN_E_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)stream_;
    char *ws = (char *)d_ws;
    int32_t *nbase = (int32_t *)(ws + L.nbase);
    uint64_t *adj = (uint64_t *)(ws + L.adj);
    int32_t *rowptr = (int32_t *)(ws + L.rowptr);
    int32_t *slot_src = (int32_t *)(ws + L.slot_src);
    int32_t *slot_dst = (int32_t *)(ws + L.slot_dst);
    int32_t *slot_col = (int32_t *)(ws + L.slot_col);
    int32_t *scan_tmp = (int32_t *)(ws + L.scan_tmp);
    const int64_t *src = d_edge_index, *dst = d_edge_index + E;

    if (N > 0 && N <= kGraphSmallMaxN && (int64_t)N * W <= kGraphSmallMaxWords && E <= 65536) {
        const size_t smem = sizeof(uint64_t) * (size_t)N * W + sizeof(int32_t) * (size_t)(2 * N + 2);
        GSN_CUDA_OK(cudaFuncSetAttribute(graph_build_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(uint64_t) * kGraphSmallMaxWords + sizeof(int32_t) * (2 * kGraphSmallMaxN + 2))));
        graph_build_small_kernel<<<1, 1024, smem, stream>>>(src, dst, (int)E, d_node_ptr, (int)G, (int)N, W, nbase, adj, rowptr,
                                                            slot_src, slot_dst, slot_col, d_status);
        GSN_BUMP(1);
        GSN_LAUNCH_OK("gsn_graph_build");
        return GSN_OK;
    }
    GSN_CUDA_OK(cudaMemsetAsync(adj, 0, sizeof(uint64_t) * ((size_t)N * W + 4), stream));
    if (N == 0) {
        GSN_CUDA_OK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * 8, stream));
        return GSN_OK;
    }
    const int TB = 256;
    k_node_base<<<(unsigned)ceil_div(N, TB), TB, 0, stream>>>(d_node_ptr, G, N, W, nbase, d_status);
    GSN_BUMP(3 + (E > 0 ? 2 : 0));
    if (E > 0)
        k_set_bits<<<(unsigned)ceil_div(E, TB), TB, 0, stream>>>(src, dst, E, N, W, nbase,
                                                                (unsigned long long *)adj, d_status);
    // degrees go into rowptr and are scanned in place
    k_degree<<<(unsigned)ceil_div(N + 1, TB), TB, 0, stream>>>(adj, N, W, rowptr);
    int rc = exclusive_scan_i32(rowptr, rowptr, N + 1, scan_tmp, stream);
    if (rc) return rc;
    k_fill_slots<<<(unsigned)ceil_div(N, TB), TB, 0, stream>>>(adj, nbase, rowptr, N, W, slot_src, slot_dst,
                                                              slot_col);
    if (E > 0)
        k_slot_cols<<<(unsigned)ceil_div(E, TB), TB, 0, stream>>>(src, dst, E, N, W, nbase, adj, rowptr, slot_col);
    GSN_LAUNCH_OK("gsn_graph_build");
    return GSN_OK;
}

extern "C" int gsn_abi_version(void) { return GSN_ABI_VERSION; }
extern "C" uint64_t gsn_launch_count(void) { return __atomic_load_n(&gsn::g_launches, __ATOMIC_RELAXED); }
extern "C" const char *gsn_last_cuda_error(void) { return gsn::g_last_error; }
