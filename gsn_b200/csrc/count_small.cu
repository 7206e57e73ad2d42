// COUNT for batches of small graphs (<= 64 nodes, one 64-bit adjacency word per vertex): graph build, enumeration and
// the write-out in edge_index order in ONE launch, no global workspace.
//
// Replaces, for the whole batch, utils_data_gen.py:60-78 -> utils_ids.py:19-27 -> utils_graph_processing.py:103-179
// (per graph: gt.Graph + remove_parallel_edges, graph-tool subgraph_isomorphism, the Python loops over all maps).
//
//   * the batch is cut into chunks of ~T consecutive nodes on graph boundaries; ONE CTA owns a chunk.  edge_index of a
//     PyG batch is grouped by graph (collate concatenates the graphs), so the chunk's edge columns are one
//     contiguous segment found by a warp-wide 32-ary search; every edge checks that it really belongs to its segment's
//     chunk (GSN_S_NOT_GROUPED otherwise -- segments tile [0,E), so no stray edge escapes the check);
//   * adjacency bitmasks (= remove_self_loops + remove_parallel_edges, :112-113), slot offsets (popcount prefix), the
//     edge_dict of :142-144 (slot -> last edge_index column) and the accumulators live in shared memory;
//   * cycles (all lengths kmin..kmax in one traversal): WARP-COOPERATIVE depth-first search.  The search stack is a
//     per-warp array of 32-byte frames (path, exclusion mask, remaining candidates) in shared memory; every step the 32
//     lanes pop the top 32 frames, each takes ONE candidate of its frame (ballot/popc compaction pushes the frame back
//     and the child on top), so all lanes execute the same instruction stream whatever the shape of the search tree,
//     and work is shared inside the warp for free.  Stack depth <= 32 * (kmax - 2) + 32 frames for any graph.
//   * cliques and generic patterns: the per-thread bitmask DFS of count_core.cuh on the shared-memory graph.
#include "common.cuh"
#include "count_core.cuh"

namespace gsn {

// Build-flag-only profiling aid (python -m gsn_b200.build --profile): clock64 stamps of the phases of every CTA, [block][8]
// (first 256 blocks).  Not compiled into the shipped library.
#ifdef GSN_PROFILE_STAMPS
__device__ long long g_cs_stamps[256 * 8];
#define CS_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_cs_stamps[blockIdx.x * 8 + (i)] = clock64(); } while (0)
#else
#define CS_STAMP(i) do { } while (0)
#endif

struct CsParams {
    const int64_t *src, *dst;
    int64_t E;
    const int64_t *node_ptr;
    int64_t G, N;
    int32_t T;             // nodes per chunk before rounding to graph boundaries
    int32_t node_cap;      // shared-memory capacity in nodes (>= T + 64)
    int32_t slot_cap;      // edge scope: slots per pass (edge_dict entries)
    int32_t acc_words;     // accumulator words in shared memory
    int32_t frame_cap;     // cycles: frames per warp
    int64_t *out;
    int64_t out_ld;
    int32_t *status;
    GsnPlan plan;
};

// first index with a[idx] >= key, searched by one warp with 32 probes per round
__device__ __forceinline__ int64_t warp_lower_bound(const int64_t *__restrict__ a, int64_t n, int64_t key, int lane) {
    int64_t lo = 0, hi = n;
    while (hi - lo > 31) {
        const int64_t step = (hi - lo + 31) / 32;
        const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
        const bool ge = idx < hi ? (__ldg(a + idx) >= key) : true;
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0) return hi;              // the probes ended exactly at hi - 1 and every element is below the key
        const int j = __ffs((int)m) - 1;
        const int64_t nhi = lo + (int64_t)(j + 1) * step - 1;
        lo = lo + (int64_t)j * step;
        hi = nhi < hi ? nhi : hi;
    }
    const int64_t idx = lo + lane;
    const bool ge = idx < hi ? (__ldg(a + idx) >= key) : true;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    return lo + (__ffs((int)m) - 1);
}

// accumulators of one pass: shared memory, or (a pass too large for it) atomics straight into the output rows
struct CsAcc {
    uint32_t *acc;          // [rows, C] in shared memory, or nullptr
    const int32_t *colmap;  // edge scope: pass-local slot -> edge_index column
    int64_t *out;           // already offset to col0
    int64_t ld;
    int64_t v0;             // global id of chunk-local node 0
    int32_t *status;
    int C, sbase;           // first slot of the pass (chunk-local numbering)
    __device__ __forceinline__ void add_vertex(int cv, int col, uint32_t c) {      // cv: chunk-local node
        if (acc) {
            const uint32_t old = atomicAdd(&acc[cv * C + col], c);
            if (old + c < old) atomicOr(status, GSN_S_COUNT_OVERFLOW);
        } else {
            const unsigned long long old = atomicAdd((unsigned long long *)&out[(v0 + cv) * ld + col], (unsigned long long)c);
            if (old + c > 0xFFFFFFFFull) atomicOr(status, GSN_S_COUNT_OVERFLOW);
        }
    }
    __device__ __forceinline__ void add_slot(int s, int col, uint32_t c) {         // s: chunk-local slot
        if (acc) {
            const uint32_t old = atomicAdd(&acc[(s - sbase) * C + col], c);
            if (old + c < old) atomicOr(status, GSN_S_COUNT_OVERFLOW);
        } else {
            const int e = colmap[s - sbase];
            if (e < 0) atomicOr(status, GSN_S_MISSING_EDGE);
            else {
                const unsigned long long old = atomicAdd((unsigned long long *)&out[(int64_t)e * ld + col], (unsigned long long)c);
                if (old + c > 0xFFFFFFFFull) atomicOr(status, GSN_S_COUNT_OVERFLOW);
            }
        }
    }
};

// adaptor for the per-thread enumerators of count_core.cuh (graph-local vertex ids, chunk-local slots)
struct CsThreadAcc {
    CsAcc *a;
    int goff;
    __device__ __forceinline__ void vertex(int lv, int col, uint32_t c) { a->add_vertex(goff + lv, col, c); }
    __device__ __forceinline__ void slot(int s, int col, uint32_t c) { a->add_slot(s, col, c); }
    __device__ __forceinline__ void overflow() { atomicOr(a->status, GSN_S_COUNT_OVERFLOW); }
};

__device__ __forceinline__ uint64_t bits_gt(int v) { return ~((2ull << v) - 1ull); }     // ids > v   (v <= 63)
__device__ __forceinline__ uint64_t bits_lt(int v) { return (1ull << v) - 1ull; }        // ids < v   (v <= 63)

// ---------------------------------------------------------------------------------------------------------------
// warp-cooperative cycle enumeration over the nodes [sn0, sn1) of the chunk (roots); see the file header
// frame: q0 = {cand.lo, cand.hi, X.lo, X.hi}, q1 = {path.lo, path.hi, meta, -}
//   meta = depth p (4 bits) | graph offset in the chunk (16 bits) << 4 | f0 << 20 | f1 << 26 ; path = f[2..] 6 bits each
//   X = vertices of the path (non-induced) or the union of adj(f[q]), 1 <= q <= p (induced: chords)
// The per-warp stack is sorted by depth, deepest on top.  A step takes frames from the top until they hold 32 candidates
// (prefix sum over the top 32 frames) and every lane extends ONE (frame, candidate) pair -- a frame with c candidates is
// consumed in one step, not in c: the critical path of a root is its depth, not the sum of its branching factors (clock64
// stamps at B = 128: the slowest warp of a CTA searched 4x longer than the first one to finish with one candidate per frame
// and step).  Children are pushed in reverse lane order (the window is read deepest first, so that keeps the stack sorted
// without a pass per depth), a partially consumed frame goes back below them, new roots are inserted at the bottom.
// Every level holds <= 32 frames: a level-d frame is created only in a step whose window reaches level d-1, i.e. one that
// consumes every older level-d frame -> size <= 32 * (kmax - 2).
template <bool INDUCED, int SCOPE>
__device__ void cycles_warp(const GsnPlan &P, const uint64_t *adj, const int32_t *rp, const uint16_t *goff, int sn0, int sn1,
                            int *ticket, uint4 *stack, int frame_cap, CsAcc &acc, int nwarps) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const int kmin = P.kmin, kmax = P.kmax;
    constexpr int scope = SCOPE;
    int size = 0;
    bool seeds_left = true;
    while (true) {
        if (size < 32 && seeds_left) {
            // ---- refill with root frames (one per node); near the end of the chunk take fewer so that warps share the tail
            int base = 0, take = 0;
            if (lane == 0) {
                const int remaining = sn1 - *(volatile int *)ticket;
                take = remaining > 0 ? remaining / (4 * nwarps) + 1 : 1;
                if (take > 32 - size) take = 32 - size;
                base = atomicAdd(ticket, take);
            }
            base = __shfl_sync(FULL, base, 0);
            take = __shfl_sync(FULL, take, 0);
            if (base >= sn1) seeds_left = false;
            const int node = base + lane;
            bool ok = lane < take && node < sn1;
            uint64_t cand = 0;
            int f0 = 0, go = 0;
            if (ok) {
                go = goff[node];
                f0 = node - go;
                cand = adj[node] & bits_gt(f0);
                ok = cand != 0 && kmax >= 3;
            }
            const unsigned m = __ballot_sync(FULL, ok);
            const int nnew = __popc(m);
            if (nnew) {
                // roots are the shallowest frames: they go to the bottom, the (< 32) frames of the stack move up
                uint4 a0, a1;
                if (lane < size) { a0 = stack[2 * lane]; a1 = stack[2 * lane + 1]; }
                __syncwarp();
                if (lane < size) { stack[2 * (lane + nnew)] = a0; stack[2 * (lane + nnew) + 1] = a1; }
                if (ok) {
                    const int pos = __popc(m & lt);
                    const uint64_t X = INDUCED ? 0ull : (1ull << f0);
                    stack[2 * pos] = make_uint4((uint32_t)cand, (uint32_t)(cand >> 32), (uint32_t)X, (uint32_t)(X >> 32));
                    stack[2 * pos + 1] = make_uint4(0u, 0u, (uint32_t)(0 | (go << 4) | (f0 << 20)), 0u);
                }
                size += nnew;
                __syncwarp();
            }
            if (size == 0) {
                if (!seeds_left) break;
                continue;
            }
        }
        if (size == 0) break;
        // ---- the window: the top <= 32 frames, deepest first; prefix sum of their candidate counts
        const int n = size < 32 ? size : 32;
        const int top = size - 1;
        uint4 wq0 = make_uint4(0u, 0u, 0u, 0u), wq1 = wq0;
        int cnt = 0;
        if (lane < n) {
            wq0 = stack[2 * (top - lane)];
            wq1 = stack[2 * (top - lane) + 1];
            cnt = __popc(wq0.x) + __popc(wq0.y);
        }
        int pre = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, pre, o);
            if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(FULL, pre, 31);
        const int excl = pre - cnt;
        const int m_taken = __popc(__ballot_sync(FULL, lane < n && excl < 32));      // frames touched by this step
        const int work = total < 32 ? total : 32;
        // the last touched frame may keep candidates: it returns to the stack without the ones handed out now
        bool keep = false;
        uint4 kq0 = wq0;
        if (lane == m_taken - 1 && pre > 32) {
            uint64_t c = (uint64_t)wq0.x | ((uint64_t)wq0.y << 32);
            for (int i = excl; i < 32; ++i) c &= c - 1;
            kq0.x = (uint32_t)c;
            kq0.y = (uint32_t)(c >> 32);
            keep = true;
        }
        // ---- lane t extends candidate t of the window: frame f = number of frames with pre <= t, k-th candidate of it
        int f = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int pv = __shfl_sync(FULL, pre, f + step - 1);
            if (pv <= lane) f += step;
        }
        f &= 31;
        const int pf = __shfl_sync(FULL, pre, f), cf = __shfl_sync(FULL, cnt, f);
        const bool active = lane < work;
        bool child = false;
        uint4 cq0, cq1;
        if (active) {
            const int k = lane - (pf - cf);
            const uint4 q0 = stack[2 * (top - f)], q1 = stack[2 * (top - f) + 1];
            const int c_lo = __popc(q0.x);
            const int j = k < c_lo ? (int)__fns(q0.x, 0, k + 1) : 32 + (int)__fns(q0.y, 0, k - c_lo + 1);
            const uint64_t X = (uint64_t)q0.z | ((uint64_t)q0.w << 32);
            uint64_t path = (uint64_t)q1.x | ((uint64_t)q1.y << 32);
            uint32_t meta = q1.z;
            const int p = meta & 15, go = (meta >> 4) & 0xFFFF, f0 = (meta >> 20) & 63;
            // ---- child: the path extended by j
            const int pc = p + 1;
            if (pc == 1) meta = (meta & ~(63u << 26)) | ((uint32_t)j << 26);
            else path |= (uint64_t)j << (6 * (pc - 2));
            meta = (meta & ~15u) | (uint32_t)pc;
            const int f1 = (meta >> 26) & 63;
            const uint64_t rowj = adj[go + j], rowa = adj[go + f0];
            uint64_t ext, Xc;
            if (!INDUCED) {
                Xc = X | (1ull << j);
                ext = rowj & ~Xc & bits_gt(f0);
            } else {
                ext = rowj & ~X & ~((1ull << f0) | (1ull << f1)) & bits_gt(f0);
                Xc = X | rowj;
            }
            const int len = pc + 2;                      // length of the cycle a closer would make
            if (len >= kmin) {
                uint64_t closers = ext & rowa & bits_gt(f1);      // canonical form: f[1] < last vertex
                if (closers) {
                    const uint32_t c = (uint32_t)__popcll(closers);
                    const int col = len - kmin;
                    if (scope == 0) {
                        acc.add_vertex(go + f0, col, c);
                        acc.add_vertex(go + f1, col, c);
                        for (int q = 2; q <= pc; ++q) acc.add_vertex(go + (int)((path >> (6 * (q - 2))) & 63), col, c);
                        while (closers) {
                            const int t = __ffsll((long long)closers) - 1;
                            closers &= closers - 1;
                            acc.add_vertex(go + t, col, 1u);
                        }
                    } else {
                        auto add2 = [&](int u, int v, uint32_t cc) {
                            acc.add_slot(rp[go + u] + __popcll(adj[go + u] & bits_lt(v)), col, cc);
                            acc.add_slot(rp[go + v] + __popcll(adj[go + v] & bits_lt(u)), col, cc);
                        };
                        int prev = f0;
                        for (int q = 1; q <= pc; ++q) {
                            const int cur = q == 1 ? f1 : (int)((path >> (6 * (q - 2))) & 63);
                            add2(prev, cur, c);
                            prev = cur;
                        }
                        while (closers) {
                            const int t = __ffsll((long long)closers) - 1;
                            closers &= closers - 1;
                            add2(prev, t, 1u);
                            add2(t, f0, 1u);
                        }
                    }
                }
            }
            if (len < kmax) {
                const uint64_t candc = INDUCED ? (ext & ~rowa) : ext;     // induced: a neighbour of the root would be a chord
                child = candc != 0;
                cq0 = make_uint4((uint32_t)candc, (uint32_t)(candc >> 32), (uint32_t)Xc, (uint32_t)(Xc >> 32));
                cq1 = make_uint4((uint32_t)path, (uint32_t)(path >> 32), meta, 0u);
            }
        }
        // ---- pop the touched frames, push [kept frame] [children, deepest on top]
        __syncwarp();                                   // every read of the window is done
        size -= m_taken;
        const unsigned mk = __ballot_sync(FULL, keep), mc = __ballot_sync(FULL, child);
        const int nkeep = mk ? 1 : 0, nchild = __popc(mc);
        if (size + nkeep + nchild > frame_cap) {         // cannot happen (bound above); never write outside the stack
            if (lane == 0) atomicOr(acc.status, GSN_S_GRAPH_TOO_LARGE);
            break;
        }
        if (keep) { stack[2 * size] = kq0; stack[2 * size + 1] = wq1; }
        if (child) {
            const int at = size + nkeep + (nchild - 1 - __popc(mc & lt));
            stack[2 * at] = cq0;
            stack[2 * at + 1] = cq1;
        }
        size += nkeep + nchild;
        __syncwarp();
    }
}

// One instantiation per (search kind, scope): MODE 0 = cycles, 1 = induced cycles (warp-cooperative search), 2 = cliques /
// generic patterns (per-thread search).  A CTA runs every phase once, with a cold instruction cache (clock64 stamps: the
// 50-edge write-out took 16 k cycles in the all-in-one kernel of 127 KB of SASS, `stall_no_instruction` 3.2 per issue): a
// launch should touch as little code as possible.
template <int NT, int MODE, int SCOPE>
__global__ void __launch_bounds__(NT) count_small_kernel(const __grid_constant__ CsParams prm) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    __shared__ int64_t sh_g[2], sh_e[2];
    __shared__ int sh_ticket, sh_pass[4], sh_warp_tot[NT / 32];
    const GsnPlan &P = prm.plan;
    const int C = P.n_cols;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    CS_STAMP(0);

    // ---- chunk: graphs whose first node lies in [blockIdx * T, (blockIdx + 1) * T)
    {
        const int64_t lo_node = (int64_t)blockIdx.x * prm.T;
        int64_t hi_node = lo_node + prm.T;
        if (hi_node > prm.N) hi_node = prm.N;
        if (warp == 0) {
            int64_t g = warp_lower_bound(prm.node_ptr, prm.G + 1, lo_node, lane);
            if (lane == 0) sh_g[0] = g > prm.G ? prm.G : g;
        } else if (warp == 1) {
            int64_t g = warp_lower_bound(prm.node_ptr, prm.G + 1, hi_node, lane);
            if (lane == 0) sh_g[1] = g > prm.G ? prm.G : g;
        }
    }
    __syncthreads();
    const int64_t g_lo = sh_g[0], g_hi = sh_g[1];
    if (g_hi <= g_lo) return;
    const int64_t v0 = prm.node_ptr[g_lo], v1 = prm.node_ptr[g_hi];
    if (warp == 0) {
        int64_t e = g_lo == 0 ? 0 : warp_lower_bound(prm.src, prm.E, v0, lane);
        if (lane == 0) sh_e[0] = e;
    } else if (warp == 1) {
        int64_t e = g_hi == prm.G ? prm.E : warp_lower_bound(prm.src, prm.E, v1, lane);
        if (lane == 0) sh_e[1] = e;
    }
    const int nn = (int)(v1 - v0);
    // ---- carve shared memory
    uint64_t *adj = (uint64_t *)cs_smem;
    int32_t *rp = (int32_t *)(adj + prm.node_cap);
    int32_t *colmap = rp + (prm.node_cap + 4);
    uint32_t *sacc = (uint32_t *)(colmap + prm.slot_cap);
    uint4 *stacks = (uint4 *)(sacc + prm.acc_words);
    uint16_t *goff = (uint16_t *)(stacks + (size_t)(MODE != 2 ? NW * prm.frame_cap * 2 : 0));
    if (nn > prm.node_cap) {          // cannot happen when every graph has <= 64 nodes (node_cap >= T + 64)
        if (tid == 0) atomicOr(prm.status, GSN_S_GRAPH_TOO_LARGE);
        return;
    }
    for (int i = tid; i < nn; i += NT) adj[i] = 0ull;
    for (int64_t g = g_lo + tid; g < g_hi; g += NT) {
        const int a = (int)(prm.node_ptr[g] - v0), b = (int)(prm.node_ptr[g + 1] - v0);
        if (b - a > 64) atomicOr(prm.status, GSN_S_GRAPH_TOO_LARGE);
        for (int v = a; v < b; ++v) goff[v] = (uint16_t)a;
    }
    __syncthreads();
    CS_STAMP(1);          // searches + graph offsets done
    const int64_t e0 = sh_e[0], e1 = sh_e[1];

    // edge (a, b) of the segment -> chunk-local a and graph-local b, or -1 when it contributes nothing
    auto decode = [&](int64_t e, int &ca, int &lb, bool flag) -> bool {
        const int64_t a = __ldg(prm.src + e), b = __ldg(prm.dst + e);
        if (a < 0 || b < 0 || a >= prm.N || b >= prm.N) { if (flag) atomicOr(prm.status, GSN_S_INDEX_RANGE); return false; }
        if (a < v0 || a >= v1) { if (flag) atomicOr(prm.status, GSN_S_NOT_GROUPED); return false; }
        if (b < v0 || b >= v1) { if (flag) atomicOr(prm.status, GSN_S_CROSS_GRAPH_EDGE); return false; }
        ca = (int)(a - v0);
        const int cb = (int)(b - v0);
        const int go = goff[ca];
        if (goff[cb] != go) { if (flag) atomicOr(prm.status, GSN_S_CROSS_GRAPH_EDGE); return false; }
        if (ca == cb) return false;
        lb = cb - go;
        return lb < 64 && ca - go < 64;
    };
    for (int64_t e = e0 + tid; e < e1; e += NT) {
        int ca, lb;
        if (!decode(e, ca, lb, true)) continue;
        const int go = goff[ca];
        atomicOr((unsigned long long *)&adj[ca], 1ull << lb);
        atomicOr((unsigned long long *)&adj[go + lb], 1ull << (ca - go));
    }
    __syncthreads();
    // ---- slot offsets: exclusive scan of the degrees (thread t owns a contiguous span of nodes)
    {
        const int span = (nn + 1 + NT - 1) / NT;
        const int b0 = tid * span, b1 = min(b0 + span, nn + 1);
        int local = 0;
        for (int v = b0; v < b1; ++v) {
            const int d = v < nn ? __popcll(adj[v]) : 0;
            rp[v] = d;
            local += d;
        }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sh_warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += sh_warp_tot[w];
        int run = woff + incl - local;
        for (int v = b0; v < b1; ++v) {
            const int d = rp[v];
            rp[v] = run;
            run += d;
        }
    }
    __syncthreads();
    CS_STAMP(2);          // adjacency + slot offsets built
    int64_t *outc = prm.out + P.col0;

    // ---- passes: vertex scope = the whole chunk; edge scope = runs of graphs whose slots fit the edge_dict capacity
    int pn0 = 0;                         // first chunk-local node of the pass
    int64_t pg = g_lo;                   // first graph of the pass
    while (pg < g_hi) {
        if (tid == 0) {
            int64_t g = pg;
            int n1 = pn0;
            if (SCOPE == 0) {
                g = g_hi;
                n1 = nn;
            } else {
                while (g < g_hi) {
                    const int nb = (int)(prm.node_ptr[g + 1] - v0);
                    if (rp[nb] - rp[pn0] > prm.slot_cap && g > pg) break;
                    n1 = nb;
                    ++g;
                }
            }
            sh_pass[0] = n1;
            sh_pass[1] = (int)(g - pg);
            sh_ticket = pn0;
        }
        __syncthreads();
        const int pn1 = sh_pass[0];
        const int64_t pg1 = pg + sh_pass[1];
        const int ps0 = rp[pn0], ps1 = rp[pn1];
        const int rows = SCOPE == 0 ? (pn1 - pn0) : (ps1 - ps0);
        const bool in_smem = (int64_t)rows * C <= prm.acc_words && (SCOPE == 0 || ps1 - ps0 <= prm.slot_cap);
        if (SCOPE == 1 && ps1 - ps0 > prm.slot_cap) {       // one graph denser than the edge_dict capacity: impossible
            if (tid == 0) atomicOr(prm.status, GSN_S_GRAPH_TOO_LARGE);     // (<= 64 * 63 slots <= slot_cap), kept as a guard
            return;
        }
        if (in_smem) {
            for (int i = tid; i < rows * C; i += NT) sacc[i] = 0u;
        } else if (SCOPE == 0) {
            for (int i = tid; i < rows * C; i += NT) outc[(v0 + pn0 + i / C) * prm.out_ld + i % C] = 0;
        }
        if (SCOPE == 1) {
            for (int i = tid; i < ps1 - ps0; i += NT) colmap[i] = -1;
            __syncthreads();
            // edge_dict (:142-144): the LAST edge_index column of a pair wins
            for (int64_t e = e0 + tid; e < e1; e += NT) {
                int ca, lb;
                if (!decode(e, ca, lb, false)) {
                    if (pn0 == 0)                    // rows of columns that cannot match (self loops, bad ids): zeros, written once
                        for (int c = 0; c < C; ++c) outc[e * prm.out_ld + c] = 0;
                    continue;
                }
                if (ca < pn0 || ca >= pn1) continue;
                atomicMax(&colmap[rp[ca] + __popcll(adj[ca] & bits_lt(lb)) - ps0], (int32_t)e);
                if (!in_smem)
                    for (int c = 0; c < C; ++c) outc[e * prm.out_ld + c] = 0;
            }
        }
        __syncthreads();

        CS_STAMP(3);      // pass set-up + edge_dict done
        CsAcc acc{in_smem ? sacc : nullptr, colmap, outc, prm.out_ld, v0, prm.status, C, ps0};
        if (SCOPE == 0) acc.sbase = 0;
        // vertex-scope rows are chunk-local nodes relative to the pass start
        CsAcc vacc = acc;
        if (SCOPE == 0 && in_smem) vacc.acc = sacc - (size_t)pn0 * C;
        if (MODE != 2) {
            uint4 *st = stacks + (size_t)warp * prm.frame_cap * 2;
            cycles_warp<MODE == 1, SCOPE>(P, adj, rp, goff, pn0, pn1, &sh_ticket, st, prm.frame_cap, vacc, NW);
        } else {
            // per-thread DFS: one root vertex at a time through a shared ticket
            while (true) {
                const int node = atomicAdd(&sh_ticket, 1);
                if (node >= pn1) break;
                const int go = goff[node];
                GraphView<1> Gv{adj + go, rp + go};
                CsThreadAcc ta{&vacc, go};
                uint64_t nb = adj[node];
                const int a = node - go;
                while (nb) {
                    const int b = __ffsll((long long)nb) - 1;
                    nb &= nb - 1;
                    if (P.family == GSN_FAMILY_CLIQUES) enumerate_cliques<1>(P.kmin, P.kmax, SCOPE, Gv, a, b, ta);
                    else enumerate_generic<1>(P, Gv, a, b, ta);
                }
            }
        }
        __syncthreads();
        CS_STAMP(4);      // search done (all warps)
        // ---- write-out
        if (in_smem) {
            if (SCOPE == 0) {
                for (int i = tid; i < rows * C; i += NT) outc[(v0 + pn0 + i / C) * prm.out_ld + i % C] = (int64_t)sacc[i];
            } else {
                for (int64_t e = e0 + tid; e < e1; e += NT) {
                    int ca, lb;
                    if (!decode(e, ca, lb, false)) continue;           // zero rows were written above
                    if (ca < pn0 || ca >= pn1) continue;
                    const int s = rp[ca] + __popcll(adj[ca] & bits_lt(lb)) - ps0;
                    const bool last = colmap[s] == (int32_t)e;         // an earlier duplicate column: edge_dict forgot it
                    for (int c = 0; c < C; ++c) outc[e * prm.out_ld + c] = last ? (int64_t)sacc[s * C + c] : 0;
                }
                // a slot that matches used but edge_index never listed: the reference's KeyError (:173)
                for (int s = tid; s < ps1 - ps0; s += NT) {
                    if (colmap[s] >= 0) continue;
                    for (int c = 0; c < C; ++c)
                        if (sacc[s * C + c]) { atomicOr(prm.status, GSN_S_MISSING_EDGE); break; }
                }
            }
        }
        __syncthreads();
        CS_STAMP(5);      // write-out done
        pn0 = pn1;
        pg = pg1;
    }
}

template <int NT, int MODE, int SCOPE>
static int cs_launch_one(const CsParams &prm, int64_t chunks, size_t smem, cudaStream_t stream) {
    GSN_CUDA_OK(cudaFuncSetAttribute(count_small_kernel<NT, MODE, SCOPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    count_small_kernel<NT, MODE, SCOPE><<<(unsigned)chunks, NT, smem, stream>>>(prm);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("count_small_kernel");
    return GSN_OK;
}

template <int NT>
static int cs_launch(const CsParams &prm, int64_t chunks, size_t smem, cudaStream_t stream) {
    const GsnPlan &P = prm.plan;
    const int mode = P.family == GSN_FAMILY_CYCLES ? (P.induced ? 1 : 0) : 2;
    if (P.scope == 0) {
        if (mode == 0) return cs_launch_one<NT, 0, 0>(prm, chunks, smem, stream);
        if (mode == 1) return cs_launch_one<NT, 1, 0>(prm, chunks, smem, stream);
        return cs_launch_one<NT, 2, 0>(prm, chunks, smem, stream);
    }
    if (mode == 0) return cs_launch_one<NT, 0, 1>(prm, chunks, smem, stream);
    if (mode == 1) return cs_launch_one<NT, 1, 1>(prm, chunks, smem, stream);
    return cs_launch_one<NT, 2, 1>(prm, chunks, smem, stream);
}

}  // namespace gsn

using namespace gsn;

#ifdef GSN_PROFILE_STAMPS
extern "C" int gsn_cs_profile_read(long long *h_out) {
    GSN_CUDA_OK(cudaMemcpyFromSymbol(h_out, g_cs_stamps, sizeof(long long) * 256 * 8));
    return GSN_OK;
}
#endif

extern "C" int gsn_count_small(const int64_t *d_edge_index, int64_t E, const int64_t *d_node_ptr, int64_t G, int64_t N,
                               const GsnPlan *h_plan, int64_t *d_out, int64_t out_ld, int32_t *d_status, void *stream_) {
    if (!h_plan || !d_status || !d_node_ptr || N < 0 || E < 0 || G < 0 || (E > 0 && !d_edge_index)) return GSN_E_INVALID;
    const GsnPlan &P = *h_plan;
    if (N == 0 || G == 0 || (P.scope == 1 && E == 0)) return GSN_OK;
    if (!d_out) return GSN_E_INVALID;
    if (P.k < 2 || P.k > GSN_MAXK || P.n_cols < 1 || P.col0 < 0 || P.col0 + P.n_cols > out_ld) return GSN_E_INVALID;
    if (P.family != GSN_FAMILY_GENERIC && (P.kmin < 3 || P.kmax > GSN_MAXK || P.kmax < P.kmin)) return GSN_E_INVALID;
    if (P.family == GSN_FAMILY_CYCLES && P.kmax > 12) return GSN_E_UNSUPPORTED;       // path bits of a frame: f[2..10]
    if (N + 1 >= (int64_t)1 << 31 || E >= (int64_t)1 << 31) return GSN_E_UNSUPPORTED;
    const int C = P.n_cols;
    // chunking: ~768 slots per chunk, at least ~4 chunks per SM when the batch is large enough, >= 16 nodes
    const double avg_deg = N > 0 ? (double)E / (double)N : 1.0;
    int64_t T = (int64_t)(768.0 / (avg_deg > 1.0 ? avg_deg : 1.0));
    const int64_t t_fill = N / (kNumSMs * 4);
    if (T > t_fill) T = t_fill;
    if (T < 48) T = 48;              // >= ~2 molecules per CTA: a chunk pays its set-up (searches, scan) once
    if (T > 1024) T = 1024;
    // a batch that fits one wave of CTAs at ~one molecule each (B = 128): the step waits for the slowest CTA, so the
    // roots of a molecule are spread over eight warps instead of two molecules over four
    const bool one_wave = ceil_div(N, 24) <= kNumSMs;
    if (one_wave) T = 24;
    const bool small_batch = !one_wave && ceil_div(N, T) <= 2 * kNumSMs;
    const int NT = small_batch ? 128 : 256;
    CsParams prm;
    prm.src = d_edge_index; prm.dst = d_edge_index ? d_edge_index + E : nullptr; prm.E = E;
    prm.node_ptr = d_node_ptr; prm.G = G; prm.N = N;
    prm.T = (int32_t)T;
    prm.node_cap = (int32_t)((T + 64 + 3) & ~3);
    prm.slot_cap = P.scope == 1 ? 4096 : 0;
    int64_t acc_words;
    if (P.scope == 0) acc_words = (int64_t)prm.node_cap * C;
    else {
        acc_words = (int64_t)((double)prm.node_cap * avg_deg * 1.5) * C;
        if (acc_words > 12288) acc_words = 12288;
        if (acc_words < 64 * C) acc_words = 64 * C;
    }
    prm.acc_words = (int32_t)((acc_words + 3) & ~3);
    // frames exist at depths 0 .. kmax - 3 and a level never holds more than 32 (see cycles_warp): the bound is tight, and
    // shared memory is what limits the resident CTAs
    prm.frame_cap = P.family == GSN_FAMILY_CYCLES ? 32 * (P.kmax > 3 ? P.kmax - 2 : 1) : 0;
    prm.out = d_out; prm.out_ld = out_ld; prm.status = d_status; prm.plan = P;
    const size_t smem = sizeof(uint64_t) * prm.node_cap + sizeof(int32_t) * (prm.node_cap + 4) + sizeof(int32_t) * prm.slot_cap +
                        sizeof(uint32_t) * prm.acc_words + (size_t)(NT / 32) * prm.frame_cap * 32 + sizeof(uint16_t) * prm.node_cap + 16;
    if (smem > 200 * 1024) return GSN_E_UNSUPPORTED;
    const int64_t chunks = ceil_div(N, T);
    cudaStream_t stream = (cudaStream_t)stream_;
    return small_batch ? cs_launch<128>(prm, chunks, smem, stream) : cs_launch<256>(prm, chunks, smem, stream);
}
