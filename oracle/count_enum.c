/*
 * COUNT oracle, C restatement.  TEST INFRASTRUCTURE ONLY: nothing under
 * gsn_b200/ links, loads or calls this file.  Users: tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference arm.
 *
 * It restates /root/reference/utils_graph_processing.py:103-179 the way the
 * reference (through graph-tool's VF2) computes it:
 *   1. build the simple undirected graph of edge_index, vertices 0..max id
 *      (:110-113 / :150-153),
 *   2. enumerate EVERY injective map f: V(H) -> V(G) that preserves edges
 *      (induced=0) or edges and non-edges (induced=1)            (:116 / :156),
 *   3. for every map bump counts[f(i), orbit(i)] (vertex scope, :123-126) or
 *      counts[edge_dict[(f(u),f(v))], edge_orbit(u,v)] for every directed
 *      pattern edge (edge scope, :161-173),
 *   4. divide by |Aut(H)|                                          (:127 / :175).
 * No symmetry breaking, no occurrence de-duplication: it does occurrences x
 * |Aut(H)| work exactly like the reference, which makes it (a) independent of
 * the CUDA path's enumerate-once scheme and (b) an honest "graph-tool
 * equivalent" CPU baseline.
 *
 * Parity pin: tests/test_oracle_count.py checks it against the reference's
 * shipped graph-tool fixture (tests/golden/imdb_k5_edge_counts.npz) and against
 * oracle/count_vf2.py (networkx VF2).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle/count_enum.c -o oracle/_build/libgsn_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 16

typedef struct {
    int n;            /* vertices 0..n-1 */
    int W;            /* 64-bit words per adjacency row */
    uint64_t *bits;   /* n*W */
    int *rowptr;      /* n+1 */
    int *col;         /* sorted neighbours */
} sgraph;

static inline int sg_adj(const sgraph *g, int a, int b) {
    return (int)((g->bits[(size_t)a * g->W + (b >> 6)] >> (b & 63)) & 1ull);
}

static int sg_build(sgraph *g, int n, int64_t E, const int64_t *src, const int64_t *dst, int64_t base) {
    g->n = n;
    g->W = (n + 63) / 64;
    if (g->W == 0) g->W = 1;
    g->bits = (uint64_t *)calloc((size_t)(n > 0 ? n : 1) * g->W, sizeof(uint64_t));
    g->rowptr = (int *)calloc((size_t)n + 1, sizeof(int));
    if (!g->bits || !g->rowptr) return -2;
    for (int64_t e = 0; e < E; ++e) {
        int64_t a = src[e] - base, b = dst[e] - base;
        if (a < 0 || b < 0 || a >= n || b >= n) return -3;
        if (a == b) continue;                       /* remove_self_loops */
        g->bits[(size_t)a * g->W + (b >> 6)] |= 1ull << (b & 63);   /* undirected + */
        g->bits[(size_t)b * g->W + (a >> 6)] |= 1ull << (a & 63);   /* remove_parallel_edges */
    }
    int tot = 0;
    for (int v = 0; v < n; ++v) {
        g->rowptr[v] = tot;
        for (int w = 0; w < g->W; ++w) tot += __builtin_popcountll(g->bits[(size_t)v * g->W + w]);
    }
    g->rowptr[n] = tot;
    g->col = (int *)malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
    if (!g->col) return -2;
    int p = 0;
    for (int v = 0; v < n; ++v)
        for (int w = 0; w < g->W; ++w) {
            uint64_t x = g->bits[(size_t)v * g->W + w];
            while (x) { g->col[p++] = w * 64 + __builtin_ctzll(x); x &= x - 1; }
        }
    return 0;
}

static void sg_free(sgraph *g) { free(g->bits); free(g->rowptr); free(g->col); }

typedef struct {
    int k, m2;
    int pe[MAXK * MAXK][2];       /* directed pattern edges, coalesced order */
    uint8_t padj[MAXK][MAXK];
    int order[MAXK];              /* matching order: position -> pattern vertex */
    int parent[MAXK];             /* an earlier-matched pattern neighbour, or -1 */
} pattern;

static int pat_build(pattern *P, int k, int m2, const int32_t *pat_edges) {
    if (k < 1 || k > MAXK || m2 > MAXK * MAXK) return -4;
    P->k = k; P->m2 = m2;
    memset(P->padj, 0, sizeof(P->padj));
    for (int i = 0; i < m2; ++i) {
        int u = pat_edges[2 * i], v = pat_edges[2 * i + 1];
        if (u < 0 || v < 0 || u >= k || v >= k) return -4;
        P->pe[i][0] = u; P->pe[i][1] = v;
        if (u != v) { P->padj[u][v] = 1; P->padj[v][u] = 1; }
    }
    /* BFS-like order: always extend with a vertex adjacent to the matched set if any */
    uint8_t done[MAXK] = {0};
    for (int p = 0; p < k; ++p) {
        int pick = -1, par = -1;
        for (int u = 0; u < k && pick < 0; ++u) {
            if (done[u]) continue;
            for (int q = 0; q < p; ++q) if (P->padj[u][P->order[q]]) { pick = u; par = P->order[q]; break; }
        }
        if (pick < 0) for (int u = 0; u < k; ++u) if (!done[u]) { pick = u; break; }
        P->order[p] = pick; P->parent[p] = par; done[pick] = 1;
    }
    return 0;
}

typedef void (*map_cb)(const int *f, void *ctx);

static void enum_maps(const pattern *P, const sgraph *G, int induced, map_cb cb, void *ctx) {
    int k = P->k, n = G->n;
    if (n < k) return;
    int f[MAXK];            /* f[pattern vertex] = target vertex */
    int it[MAXK];           /* candidate cursor per position */
    uint8_t *used = (uint8_t *)calloc((size_t)n, 1);
    int p = 0; it[0] = 0;
    while (p >= 0) {
        int u = P->order[p], par = P->parent[p];
        int lo, hi;
        if (par >= 0) { lo = G->rowptr[f[par]]; hi = G->rowptr[f[par] + 1]; }
        else { lo = 0; hi = n; }
        int found = 0;
        while (lo + it[p] < hi) {
            int c = (par >= 0) ? G->col[lo + it[p]] : it[p];
            it[p]++;
            if (used[c]) continue;
            int ok = 1;
            for (int q = 0; q < p && ok; ++q) {
                int w = P->order[q];
                int a = sg_adj(G, c, f[w]);
                if (P->padj[u][w]) { if (!a) ok = 0; }
                else if (induced && a) ok = 0;
            }
            if (!ok) continue;
            f[u] = c; found = 1; break;
        }
        if (!found) { --p; if (p >= 0) used[f[P->order[p]]] = 0; continue; }
        if (p == k - 1) { cb(f, ctx); continue; }     /* stay on this level, next candidate */
        used[f[u]] = 1; ++p; it[p] = 0;
    }
    free(used);
}

/* ---- automorphism count: number of maps H -> H (utils_graph_processing.py:22,48) ---- */
static void cb_count(const int *f, void *ctx) { (void)f; ++*(int64_t *)ctx; }

int64_t gsn_oracle_aut_count(int k, int m2, const int32_t *pat_edges) {
    pattern P; if (pat_build(&P, k, m2, pat_edges)) return -4;
    int64_t *src = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m2 > 0 ? m2 : 1));
    int64_t *dst = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m2 > 0 ? m2 : 1));
    for (int i = 0; i < m2; ++i) { src[i] = P.pe[i][0]; dst[i] = P.pe[i][1]; }
    sgraph H; int rc = sg_build(&H, k, m2, src, dst, 0);
    int64_t cnt = 0;
    if (!rc) enum_maps(&P, &H, 0, cb_count, &cnt);
    sg_free(&H); free(src); free(dst);
    return rc ? rc : cnt;
}

typedef struct {
    const pattern *P; const sgraph *G;
    const int32_t *orbit; int n_orbits; int scope;
    double *out;              /* vertex scope: [rows, n_orbits]; edge scope: per unique slot */
    int missing;              /* edge scope: mapped edge absent from edge_dict */
    const int64_t *slot2col;  /* edge scope: simple-graph slot -> last edge_index column, -1 if absent */
} acc_ctx;

static void cb_vertex(const int *f, void *vctx) {
    acc_ctx *c = (acc_ctx *)vctx;
    for (int i = 0; i < c->P->k; ++i) c->out[(size_t)f[i] * c->n_orbits + c->orbit[i]] += 1.0;
}

static inline int sg_slot(const sgraph *G, int a, int b) {
    /* index of b inside a's sorted neighbour list */
    int lo = G->rowptr[a], hi = G->rowptr[a + 1];
    while (lo < hi) { int mid = (lo + hi) >> 1; if (G->col[mid] < b) lo = mid + 1; else hi = mid; }
    return lo;
}

static void cb_edge(const int *f, void *vctx) {
    acc_ctx *c = (acc_ctx *)vctx;
    for (int i = 0; i < c->P->m2; ++i) {
        int a = f[c->P->pe[i][0]], b = f[c->P->pe[i][1]];
        int64_t colidx = c->slot2col[sg_slot(c->G, a, b)];
        if (colidx < 0) { c->missing = 1; continue; }      /* reference: KeyError */
        c->out[(size_t)colidx * c->n_orbits + c->orbit[i]] += 1.0;
    }
}

/*
 * One graph.  edge_index is [2,E] row-major int64 holding vertex ids in
 * [base, base+num_nodes).  scope 0 = vertex (out [num_nodes, n_orbits]),
 * scope 1 = edge (out [E, n_orbits], rows in edge_index column order).
 * Returns 0, or -1 when an edge-scope map hits a (src,dst) pair that is not a
 * column of edge_index (asymmetric input; the reference raises KeyError).
 */
int gsn_oracle_count_graph(int64_t num_nodes, int64_t base, int64_t E,
                           const int64_t *src, const int64_t *dst,
                           int k, int m2, const int32_t *pat_edges,
                           const int32_t *orbit, int n_orbits,
                           int induced, int scope, int64_t aut_count, double *out) {
    pattern P; int rc = pat_build(&P, k, m2, pat_edges); if (rc) return rc;
    int64_t maxid = -1;
    for (int64_t e = 0; e < E; ++e) {
        if (src[e] - base > maxid) maxid = src[e] - base;
        if (dst[e] - base > maxid) maxid = dst[e] - base;
    }
    /* matching runs over vertices 0..max id; rows are num_nodes (:118-122) */
    int n = (int)(maxid + 1);
    if (n > num_nodes && scope == 0) return -3;
    sgraph G; rc = sg_build(&G, n, E, src, dst, base); if (rc) { sg_free(&G); return rc; }
    int64_t rows = scope == 0 ? num_nodes : E;
    memset(out, 0, sizeof(double) * (size_t)rows * n_orbits);
    acc_ctx c = {&P, &G, orbit, n_orbits, scope, out, 0, NULL};
    int64_t *slot2col = NULL;
    if (scope == 1) {
        int S = G.rowptr[n];
        slot2col = (int64_t *)malloc(sizeof(int64_t) * (size_t)(S > 0 ? S : 1));
        for (int s = 0; s < S; ++s) slot2col[s] = -1;
        for (int64_t e = 0; e < E; ++e) {           /* edge_dict: last column wins (:142-144) */
            int a = (int)(src[e] - base), b = (int)(dst[e] - base);
            if (a == b) continue;
            slot2col[sg_slot(&G, a, b)] = e;
        }
        c.slot2col = slot2col;
        enum_maps(&P, &G, induced, cb_edge, &c);
    } else {
        enum_maps(&P, &G, induced, cb_vertex, &c);
    }
    double inv = (double)aut_count;
    for (int64_t i = 0; i < rows * n_orbits; ++i) out[i] /= inv;
    free(slot2col); sg_free(&G);
    return c.missing ? -1 : 0;
}

/*
 * Batch of graphs = the reference's per-graph loop (utils_data_gen.py:74-78, or
 * the joblib pool :60-71 when nthreads > 1).  Graph g owns nodes
 * [node_ptr[g], node_ptr[g+1]) and edge_index columns [edge_ptr[g], edge_ptr[g+1]).
 * out is [N, n_orbits] or [E, n_orbits] for the whole batch.
 */
int gsn_oracle_count_batch(int64_t num_graphs, const int64_t *node_ptr, const int64_t *edge_ptr,
                           const int64_t *src, const int64_t *dst,
                           int k, int m2, const int32_t *pat_edges,
                           const int32_t *orbit, int n_orbits,
                           int induced, int scope, int64_t aut_count, double *out, int nthreads) {
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t g = 0; g < num_graphs; ++g) {
        int64_t n0 = node_ptr[g], n1 = node_ptr[g + 1], e0 = edge_ptr[g], e1 = edge_ptr[g + 1];
        double *o = out + (size_t)(scope == 0 ? n0 : e0) * n_orbits;
        if (scope == 1 && e1 == e0) continue;
        int rc = gsn_oracle_count_graph(n1 - n0, n0, e1 - e0, src + e0, dst + e0, k, m2, pat_edges,
                                        orbit, n_orbits, induced, scope, aut_count, o);
        if (rc) {
#pragma omp critical
            err = rc;
        }
    }
    return err;
}

int gsn_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
