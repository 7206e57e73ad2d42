// Enumeration cores of the COUNT path (one work item = one directed edge (a,b)
// of one graph, matched onto pattern positions 0 and 1).
//
// Replaces, per item, graph-tool's VF2 search plus the Python accumulation loops
// of /root/reference/utils_graph_processing.py:116-126 (vertex scope) and
// :156-173 (edge scope).  Instead of visiting all |Aut(H)| maps of every
// occurrence and dividing by aut_count (:127, :175) the plan's symmetry-breaking
// constraints make every occurrence appear exactly once.
//
// The functions are __host__ __device__ so that tests/host_sim can run the very
// same code on the CPU against the oracle when no GPU is present.  They are NOT
// a CPU fallback: gsn_b200/ only ever launches them from the CUDA kernels in
// count_kernels.cu.
#pragma once
#include <stdint.h>
#include "../../include/gsn_b200.h"

#if defined(__CUDACC__)
#define GSN_HD __host__ __device__ __forceinline__
#else
#define GSN_HD inline
#endif

namespace gsn {

GSN_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
GSN_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
GSN_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// ---------------------------------------------------------------- vertex sets
template <int W>
struct VSet {
    uint64_t w[W];
    GSN_HD void load(const uint64_t *p) {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = p[i];
    }
    GSN_HD void clear() {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = 0;
    }
    GSN_HD bool empty() const {
        uint64_t o = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) o |= w[i];
        return o == 0;
    }
    GSN_HD int count() const {
        int c = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) c += popc64(w[i]);
        return c;
    }
    GSN_HD int count_and(const uint64_t *p) const {          // |this & p|
        int c = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) c += popc64(w[i] & p[i]);
        return c;
    }
    GSN_HD void and_with(const uint64_t *p) {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] &= p[i];
    }
    GSN_HD void andnot_with(const uint64_t *p) {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] &= ~p[i];
    }
    GSN_HD void or_with(const uint64_t *p) {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] |= p[i];
    }
    GSN_HD void and_set(const VSet &o) { and_with(o.w); }
    GSN_HD void andnot_set(const VSet &o) { andnot_with(o.w); }
    GSN_HD void or_set(const VSet &o) { or_with(o.w); }
    GSN_HD void set_bit(int j) { w[j >> 6] |= 1ull << (j & 63); }
    GSN_HD void clear_bit(int j) { w[j >> 6] &= ~(1ull << (j & 63)); }
    GSN_HD bool test(int j) const { return (w[j >> 6] >> (j & 63)) & 1ull; }
    // keep only vertices with id > j
    GSN_HD void keep_gt(int j) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            int lo = i * 64;
            if (j >= lo + 63) w[i] = 0;
            else if (j >= lo) w[i] &= ~((2ull << (j - lo)) - 1ull);
        }
    }
    // remove and return the smallest vertex (set must be non-empty)
    GSN_HD int pop_lowest() {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (w[i]) {
                int b = ctz64(w[i]);
                w[i] &= w[i] - 1;
                return i * 64 + b;
            }
        }
        return -1;
    }
};

// Work splitting for small batches: keep the elements of `s` whose ordinal inside the set is congruent to
// `part` modulo `parts` (the sub-items of one directed edge partition its third-level candidates).
template <int W>
GSN_HD void keep_part(VSet<W> &s, int part, int parts) {
    if (parts <= 1) return;
    VSet<W> in = s;
    s.clear();
    int ord = 0;
    while (!in.empty()) {
        int v = in.pop_lowest();
        if (ord % parts == part) s.set_bit(v);
        ++ord;
    }
}

// number of neighbours of the row `p` (W words) with id < b  == position of b
// in the ascending neighbour list == slot offset inside the row
template <int W>
GSN_HD int rank_below(const uint64_t *p, int b) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        int lo = i * 64;
        if (b >= lo + 64) r += popc64(p[i]);
        else if (b > lo) r += popc64(p[i] & ((1ull << (b - lo)) - 1ull));
    }
    return r;
}

// One graph as the kernels see it: adjacency rows of W words per LOCAL vertex id
// and the slot offset of every local vertex (CSR of the simple graph).
// `stride` (words between rows) may exceed W: a graph of <= 64 * W vertices inside a batch laid out for wider graphs
// is searched with W-word sets (its rows are zero beyond word W).
template <int W>
struct GraphView {
    const uint64_t *adj;     // adj + v*stride
    const int32_t *rowptr;   // rowptr[v] = first slot of local vertex v (batch-global slot ids)
    int32_t stride = W;
    GSN_HD const uint64_t *row(int v) const { return adj + (size_t)v * stride; }
    GSN_HD int slot(int a, int b) const { return rowptr[a] + rank_below<W>(row(a), b); }
};

// Accumulator concept:
//   void vertex(int local_v, int col, uint32_t c);   // counts[v, col] += c
//   void slot(int slot, int col, uint32_t c);        // edge counts by simple-graph slot
//   void overflow();                                 // a count left the 32-bit range (GSN_S_COUNT_OVERFLOW)
// DFS sub-totals are 64-bit; clamp32 hands them to the 32-bit accumulator interface and reports values that do not fit
// (the reference computes the same counts exactly in float64, utils_graph_processing.py:118-127).
template <class Acc>
GSN_HD uint32_t clamp32(uint64_t c, Acc &acc) {
    if (c > 0xFFFFFFFFull) {
        acc.overflow();
        return 0xFFFFFFFFu;
    }
    return (uint32_t)c;
}

// ------------------------------------------------------------- generic pattern
template <int W, class Acc>
GSN_HD void flush_position(const GsnPlan &P, const GraphView<W> &G, const int *f, int p, uint64_t c64, Acc &acc) {
    if (c64 == 0) return;
    const uint32_t c = clamp32(c64, acc);
    if (P.scope == 0) {
        acc.vertex(f[p], P.vorbit[p], c);
    } else {
        uint32_t m = P.nbr_mask[p];
        while (m) {
            int q = ctz32(m);
            m &= m - 1;
            acc.slot(G.slot(f[q], f[p]), P.e_fwd[p][q], c);
            acc.slot(G.slot(f[p], f[q]), P.e_bwd[p][q], c);
        }
    }
}

template <int W>
GSN_HD void candidates(const GsnPlan &P, const GraphView<W> &G, const int *f, int p, VSet<W> &out) {
    uint32_t m = P.nbr_mask[p];
    int q0 = ctz32(m);
    m &= m - 1;
    out.load(G.row(f[q0]));
    while (m) {
        int q = ctz32(m);
        m &= m - 1;
        out.and_with(G.row(f[q]));
    }
    if (P.induced) {
        uint32_t nm = P.non_mask[p];
        while (nm) {
            int q = ctz32(nm);
            nm &= nm - 1;
            out.andnot_with(G.row(f[q]));
        }
    }
    int lo = -1;
    uint32_t gm = P.gt_mask[p];
    for (int q = 0; q < p; ++q) {
        out.clear_bit(f[q]);
        if ((gm >> q) & 1u) lo = f[q] > lo ? f[q] : lo;
    }
    if (lo >= 0) out.keep_gt(lo);
}

// Enumerates every occurrence of P whose positions 0,1 sit on the directed edge
// (a,b) and accumulates one unit per occurrence for every pattern vertex / edge.
// Sub-totals are pushed up the search stack so each search node costs O(1)
// accumulator updates.
template <int W, class Acc>
GSN_HD void enumerate_generic(const GsnPlan &P, const GraphView<W> &G, int a, int b, Acc &acc, int part = 0, int parts = 1) {
    const int k = P.k;
    if ((P.gt_mask[1] & 1u) && !(a < b)) return;
    if (k == 2 && part != 0) return;
    int f[GSN_MAXK];
    uint64_t cnt[GSN_MAXK];
    VSet<W> cand[GSN_MAXK];
    f[0] = a;
    f[1] = b;
    cnt[1] = 0;
    if (k == 2) {
        cnt[1] = 1;
    } else {
        int p = 2;
        candidates<W>(P, G, f, 2, cand[2]);
        keep_part<W>(cand[2], part, parts);
        while (true) {
            if (cand[p].empty()) {
                if (p == 2) break;
                --p;
                uint64_t c = cnt[p];
                flush_position<W>(P, G, f, p, c, acc);
                cnt[p - 1] += c;
                continue;
            }
            if (p == k - 1) {
                uint64_t c = 0;
                while (!cand[p].empty()) {
                    f[p] = cand[p].pop_lowest();
                    flush_position<W>(P, G, f, p, 1u, acc);
                    ++c;
                }
                cnt[p - 1] += c;
                continue;
            }
            f[p] = cand[p].pop_lowest();
            cnt[p] = 0;
            ++p;
            candidates<W>(P, G, f, p, cand[p]);
        }
    }
    flush_position<W>(P, G, f, 1, cnt[1], acc);
    if (P.scope == 0 && cnt[1]) acc.vertex(f[0], P.vorbit[0], clamp32(cnt[1], acc));
}

// ------------------------------------------------------------------- cycles
// All cycle lengths kmin..kmax in ONE traversal (column = length - kmin).
// Canonical form of a cycle: f[0] = its smallest vertex, f[1] < f[last]
// (one representative of the 2k maps of utils_graph_processing.py:116).
// induced=1: chordless cycles only (graph-tool induced=True).
template <int W, class Acc>
GSN_HD void record_cycle_set(int scope, const GraphView<W> &G, const int *f, int p, VSet<W> closers, int col, Acc &acc) {
    uint32_t c = (uint32_t)closers.count();
    if (c == 0) return;
    if (scope == 0) {
        for (int q = 0; q <= p; ++q) acc.vertex(f[q], col, c);
        while (!closers.empty()) acc.vertex(closers.pop_lowest(), col, 1u);
    } else {
        for (int q = 0; q < p; ++q) {
            acc.slot(G.slot(f[q], f[q + 1]), col, c);
            acc.slot(G.slot(f[q + 1], f[q]), col, c);
        }
        while (!closers.empty()) {
            int j = closers.pop_lowest();
            acc.slot(G.slot(f[p], j), col, 1u);
            acc.slot(G.slot(j, f[p]), col, 1u);
            acc.slot(G.slot(j, f[0]), col, 1u);
            acc.slot(G.slot(f[0], j), col, 1u);
        }
    }
}

template <int W, class Acc>
GSN_HD void enumerate_cycles(int kmin, int kmax, int induced, int scope, const GraphView<W> &G, int a, int b, Acc &acc,
                             int part = 0, int parts = 1) {
    if (b <= a) return;
    int f[GSN_MAXK];
    VSet<W> cand[GSN_MAXK];   // cand[p]: remaining choices for f[p]
    VSet<W> used[GSN_MAXK];   // used[p]: {f[0..p]}
    VSet<W> forb[GSN_MAXK];   // induced: forb[p] = union of adj(f[q]), 1 <= q < p
    f[0] = a;
    f[1] = b;
    used[1].clear();
    used[1].set_bit(a);
    used[1].set_bit(b);
    forb[1].clear();
    int p = 1;
    while (true) {
        // visit the path f[0..p]: close it (cycle of length p+2) and/or extend it
        VSet<W> ext;
        ext.load(G.row(f[p]));
        ext.andnot_set(used[p]);
        ext.keep_gt(a);
        if (induced) ext.andnot_set(forb[p]);
        const int len = p + 2;
        if (len >= kmin && (p > 1 || part == 0)) {          // triangles on (a,b) are recorded by sub-item 0 only
            VSet<W> closers = ext;
            closers.and_with(G.row(a));
            closers.keep_gt(b);
            record_cycle_set<W>(scope, G, f, p, closers, len - kmin, acc);
        }
        if (len < kmax) {
            if (induced) {
                ext.andnot_with(G.row(a));        // a neighbour of the root would be a chord later
                forb[p + 1] = forb[p];
                forb[p + 1].or_with(G.row(f[p]));
            }
            cand[p + 1] = ext;
            if (p == 1) keep_part<W>(cand[2], part, parts);
            ++p;
        }
        while (p >= 2 && cand[p].empty()) --p;
        if (p < 2) break;
        f[p] = cand[p].pop_lowest();
        used[p] = used[p - 1];
        used[p].set_bit(f[p]);
    }
}

// ------------------------------------------------------------------- cliques
// All clique sizes kmin..kmax of one item, counted LOCALLY: the number of k-cliques through the edge {a,b} is the
// number of (k-2)-cliques inside the common neighbourhood N(a) & N(b), so an item owns its result and adds it once --
// no update per occurrence, no slot lookup per occurrence.  The search runs over increasing vertex tuples inside that
// set (one representative of the (k-2)! orders) and the last level is a population count.  column = size - kmin.
//   edge scope:   item (a,b), a < b, pool = N(a) & N(b); both slots of the edge receive the totals.
//   vertex scope: item (a,b), every direction, pool = N(a) & N(b) & {> b}: the cliques through a whose smallest other
//                 member is b; vertex a receives the totals (summed over b: every clique through a exactly once).
template <int W, class Acc>
GSN_HD void enumerate_cliques(int kmin, int kmax, int scope, const GraphView<W> &G, int a, int b, Acc &acc,
                              int part = 0, int parts = 1) {
    if (scope != 0 && b <= a) return;
    VSet<W> cand[GSN_MAXK];
    uint64_t total[GSN_MAXK + 1];
#pragma unroll
    for (int i = 0; i <= GSN_MAXK; ++i) total[i] = 0;
    cand[2].load(G.row(a));
    cand[2].and_with(G.row(b));
    if (scope == 0) cand[2].keep_gt(b);
    const VSet<W> pool = cand[2];            // deeper levels draw from the whole pool
    keep_part<W>(cand[2], part, parts);      // each sub-item counts and extends its own share of the triangles
    int p = 2;
    bool fresh = true;   // cand[p] was just computed: it closes cand[p].count() cliques of size p+1
    while (true) {
        if (fresh) {
            fresh = false;
            const int size = p + 1;
            if (size >= kmin && size <= kmax) total[size] += (uint64_t)cand[p].count();
            if (size >= kmax) {
                cand[p].clear();                 // no deeper level needed
            } else if (size + 1 == kmax) {
                // the next level is the last: every j of cand[p] closes |candidates after j that are adjacent to j|
                // cliques of size kmax -- counted here, without a visit per (j, candidate set)
                uint64_t c = 0;
                VSet<W> t = cand[p];
                if (p == 2 && parts > 1) {
                    while (!t.empty()) {
                        const int j = t.pop_lowest();
                        VSet<W> u = pool;
                        u.keep_gt(j);
                        c += (uint64_t)u.count_and(G.row(j));
                    }
                } else {
                    while (!t.empty()) {
                        const int j = t.pop_lowest();    // ascending: what is left in t is > j
                        c += (uint64_t)t.count_and(G.row(j));
                    }
                }
                total[kmax] += c;
                cand[p].clear();
            }
        }
        if (cand[p].empty()) {
            if (p == 2) break;
            --p;
            continue;
        }
        const int j = cand[p].pop_lowest();
        if (p == 2) {                   // the sub-item filter applies to the first choice only
            cand[3] = pool;
            cand[3].keep_gt(j);
        } else {
            cand[p + 1] = cand[p];      // remaining candidates are all > j already (ascending pop)
        }
        cand[p + 1].and_with(G.row(j));
        ++p;
        fresh = true;
    }
    int s_ab = -1, s_ba = -1;
    for (int size = kmin; size <= kmax; ++size) {
        if (total[size] == 0) continue;
        const uint32_t c = clamp32(total[size], acc);
        if (scope == 0) {
            acc.vertex(a, size - kmin, c);
        } else {
            if (s_ab < 0) {
                s_ab = G.slot(a, b);
                s_ba = G.slot(b, a);
            }
            acc.slot(s_ab, size - kmin, c);
            acc.slot(s_ba, size - kmin, c);
        }
    }
}

}  // namespace gsn
