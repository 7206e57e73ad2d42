/*
 * gsn_b200 -- C ABI of libgsn_b200.so (sm_100a).
 *
 * The reference (gbouritsas/GSN) is pure Python and has no FFI; these entry
 * points are what a ctypes binding added to the reference would call for its
 * two data-parallel hot paths (INTEGRATION.md shows the stubs):
 *
 *   COUNT  utils_graph_processing.py:103-179 + utils_ids.py:7-29
 *          (graph-tool subgraph_isomorphism + the Python per-map loops)
 *   MP     graph_filters/GSN_sparse.py:122-176, GSN_edge_sparse.py:119-170,
 *          GSN_edge_sparse_ogb.py:86-129 and the MPNN_* twins
 *          (gather -> concat structural ids -> [transform] -> scatter-add)
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; h_* is host
 *     memory.  The library never allocates, frees or retains caller memory.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises unless its comment says so.
 *   - return value: 0 = enqueued / ok, <0 = GSN_E_* (argument or CUDA error
 *     detected on the host).  Data-dependent errors (a graph larger than the
 *     bitmask width, an edge joining two graphs, an asymmetric edge hit by a
 *     match -- the reference's KeyError) are written to the caller's
 *     `d_status` word (int32, device) as GSN_S_* bits; the caller reads it when
 *     it next synchronises.
 *   - no C++ exceptions cross this boundary; reentrant; no global mutable state.
 *   - edge_index is the reference's layout: int64 [2,E] row-major, row 0 =
 *     source, row 1 = target, GLOBAL (batched) node ids.
 */
#ifndef GSN_B200_H_
#define GSN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSN_ABI_VERSION 1

/* host-side return codes */
#define GSN_OK 0
#define GSN_E_INVALID (-1)      /* bad argument */
#define GSN_E_UNSUPPORTED (-2)  /* shape outside the built kernels */
#define GSN_E_WORKSPACE (-3)    /* workspace too small */
#define GSN_E_CUDA (-4)         /* a CUDA runtime call failed; see gsn_last_cuda_error */

/* device-side status bits (OR-ed into *d_status) */
#define GSN_S_GRAPH_TOO_LARGE 1 /* a graph has more than 64*W vertices */
#define GSN_S_CROSS_GRAPH_EDGE 2/* an edge joins two graphs of the batch */
#define GSN_S_MISSING_EDGE 4    /* edge scope: a match used (a,b) but edge_index has no column (a,b)
                                   (utils_graph_processing.py:173 raises KeyError) */
#define GSN_S_INDEX_RANGE 8     /* a node id outside [0,N) */
#define GSN_S_COUNT_OVERFLOW 16 /* a per-vertex / per-edge count exceeded 2^32 - 1 inside a CTA-private accumulator */
#define GSN_S_NOT_GROUPED 32    /* gsn_count_small: edge_index columns are not grouped by graph (PyG collate order) */
#define GSN_S_UNSEEN_VALUE 64   /* gsn_encode_rows: a value that is not in the one_hot_unique vocabulary (encoded as the next
                                   larger known value; informative, the reference cannot encode such a value at all) */

#define GSN_MAXK 16             /* max pattern vertices */

/* ------------------------------------------------------------------ */
/* COUNT                                                               */
/* ------------------------------------------------------------------ */

/*
 * Matching plan of one pattern H, compiled on the host (gsn_b200/patterns.py)
 * from the output of automorphism_orbits / induced_edge_automorphism_orbits
 * (utils_graph_processing.py:10-100).  Positions 0..k-1 are pattern vertices in
 * matching order; position 1 is adjacent to position 0; every position p>=1 has
 * at least one earlier neighbour.  The gt constraints break Aut(H) so that each
 * occurrence (class of |Aut(H)| maps) is enumerated exactly once, which equals
 * the reference's "count every map, divide by aut_count" (:127, :175).
 */
typedef struct GsnPlan {
    int32_t k;                       /* pattern vertices, 2..GSN_MAXK */
    int32_t induced;                 /* 0: monomorphisms, 1: induced  (:116 / :156) */
    int32_t scope;                   /* 0: vertex counts (:103), 1: edge counts (:134) */
    int32_t n_cols;                  /* orbits of this pattern = output columns */
    int32_t col0;                    /* first output column of this pattern */
    int32_t family;                  /* GSN_FAMILY_* */
    int32_t kmin, kmax;              /* GSN_FAMILY_CYCLES/CLIQUES: sizes kmin..kmax, one column each */
    uint32_t nbr_mask[GSN_MAXK];     /* bit q (q<p): positions q,p adjacent in H */
    uint32_t non_mask[GSN_MAXK];     /* bit q (q<p): positions q,p NOT adjacent in H */
    uint32_t gt_mask[GSN_MAXK];      /* bit q (q<p): require f(q) < f(p) */
    int8_t vorbit[GSN_MAXK];         /* vertex scope: orbit (column) of position p */
    int8_t e_fwd[GSN_MAXK][GSN_MAXK];/* edge scope: [p][q], q<p: orbit of directed pattern edge q->p, -1 none */
    int8_t e_bwd[GSN_MAXK][GSN_MAXK];/* edge scope: [p][q], q<p: orbit of directed pattern edge p->q, -1 none */
} GsnPlan;

#define GSN_FAMILY_GENERIC 0
#define GSN_FAMILY_CYCLES 1      /* all cycle lengths kmin..kmax in one traversal */
#define GSN_FAMILY_CLIQUES 2     /* all clique sizes kmin..kmax in one traversal */

/*
 * Bytes of device workspace gsn_graph_build needs for a batch with N nodes, E
 * edge_index columns and W 64-bit adjacency words per vertex
 * (W >= ceil(max graph size / 64)).
 */
int gsn_graph_workspace_bytes(int64_t N, int64_t E, int32_t W, size_t *bytes);

/*
 * Builds the batched simple undirected graph the reference builds per graph with
 *   gt.Graph(directed=False); add_edge_list; remove_self_loops; remove_parallel_edges
 * (utils_graph_processing.py:110-113, :150-153): per-vertex adjacency bitmasks,
 * slot offsets (CSR of the simple graph, neighbours ascending) and the
 * edge_dict of :142-144 (slot -> LAST edge_index column holding that pair).
 *   d_node_ptr  int64 [G+1]  first node of every graph (PyG batch.ptr)
 */
int gsn_graph_build(const int64_t *d_edge_index, int64_t E, const int64_t *d_node_ptr, int64_t G,
                    int64_t N, int32_t W, void *d_ws, size_t ws_bytes, int32_t *d_status, void *stream);

/* Bytes of scratch gsn_count_pattern needs (edge scope: per-slot accumulators). */
int gsn_count_scratch_bytes(int64_t N, int64_t E, const GsnPlan *h_plan, size_t *bytes);

/*
 * Replaces count_fn(edge_index, subgraph_dict=, induced=, num_nodes=) of
 * utils_ids.py:24 for the whole batch in one launch:
 *   scope 0: d_out[v, col0+o] = #occurrences of H in which vertex v plays orbit o   (:103-131)
 *   scope 1: d_out[e, col0+o] = #occurrences in which edge_index column e plays
 *            edge orbit o; rows follow edge_index columns, duplicate columns ->
 *            last one wins, self loops -> 0                                       (:134-179)
 * d_out is int64 [rows, out_ld] (identifiers after .long(), utils_ids.py:27);
 * the columns [col0, col0+n_cols) are overwritten.
 * d_ws must come from gsn_graph_build on the same (edge_index, node_ptr).
 */
int gsn_count_pattern(const void *d_ws, int64_t N, int64_t E, int32_t W, const int64_t *d_edge_index,
                      const int64_t *d_node_ptr, int64_t G, const GsnPlan *h_plan, int64_t *d_out,
                      int64_t out_ld, void *d_scratch, size_t scratch_bytes, int32_t *d_status,
                      void *stream);

/*
 * COUNT for batches of SMALL graphs (every graph <= 64 nodes) in ONE launch and without a workspace: graph build
 * (gsn_graph_build), enumeration (gsn_count_pattern) and the write-out in edge_index order happen inside one kernel,
 * CTA-private in shared memory.  Same outputs as gsn_graph_build + gsn_count_pattern.  Cycle families run a
 * warp-cooperative depth-first search (shared frame stack, ballot / popc compaction; kmax <= 12), cliques and
 * generic patterns the per-thread bitmask search.
 * Requires the columns of edge_index to be grouped by graph in batch order (what PyG's collate produces):
 * the columns of a run of graphs are located by binary search on the source row; a batch that violates this sets
 * GSN_S_NOT_GROUPED (the caller falls back to gsn_graph_build + gsn_count_pattern).  A graph with more than 64 nodes
 * sets GSN_S_GRAPH_TOO_LARGE.  Returns GSN_E_UNSUPPORTED for plans outside the kernel (cycles with kmax > 12).
 */
int gsn_count_small(const int64_t *d_edge_index, int64_t E, const int64_t *d_node_ptr, int64_t G, int64_t N,
                    const GsnPlan *h_plan, int64_t *d_out, int64_t out_ld, int32_t *d_status, void *stream);

/* ------------------------------------------------------------------ */
/* MP                                                                  */
/* ------------------------------------------------------------------ */

/* Bytes of workspace for gsn_csr_build. */
int gsn_csr_workspace_bytes(int64_t N, int64_t E, size_t *bytes);

/*
 * Groups edge_index columns by d_key[e] (the aggregation index: edge_index[1]
 * for flow='source_to_target', GSN_sparse.py:125-129) into a CSR:
 *   d_rowptr int32 [N+1], d_eid int32 [E] (edge_index column ids, ASCENDING
 *   inside every row -> fixed summation order, deterministic results),
 *   d_nbr int32 [E] = d_other[d_eid[k]] (the gathered endpoint x_j).
 * Replaces the COO construction + torch.sparse.sum coalesce of
 * GSN_sparse.py:140-143.
 */
int gsn_csr_build(const int64_t *d_key, const int64_t *d_other, int64_t E, int64_t N, int32_t *d_rowptr,
                  int32_t *d_eid, int32_t *d_nbr, void *d_ws, size_t ws_bytes, int32_t *d_status,
                  void *stream);

/*
 * One column segment of the concatenated 'gin' message / self term.
 *   neighbour part: src[index(e) * src_ld + c]   index = d_nbr (mode 1), d_eid (mode 2), none (mode 0)
 *   self part     : self[i * self_ld + c]  (self_ld = 0 broadcasts one row; NULL = 0)  + self_const
 * Segments express cat(x_j, identifiers, edge_features) as well as the extra
 * "self-loop" column / embedding of central_encoder (utils_graph_learning.py:211-260)
 * without materialising the [E, d+1] tensors the reference builds.
 */
typedef struct GsnSegment {
    const float *src;
    const float *self;
    int32_t width;
    int32_t src_ld;
    int32_t self_ld;
    int32_t index_mode;   /* 0: no neighbour contribution, 1: gather at neighbour, 2: per-edge row */
    float self_const;
    int32_t _pad;
} GsnSegment;

#define GSN_MAX_SEGMENTS 8

/*
 * Fused gather -> concat -> scatter-add of the 'gin' message kind
 * (GSN_sparse.py:103-111,160-164; GSN_edge_sparse.py:95-109,155-159;
 *  MPNN_* with no identifier segment):
 *   out[i, :] = (1+eps) * cat_s(self_s[i] + self_const_s) + sum_{e in row i} cat_s(src_s[index_s(e)])
 * h_segs: HOST array of n_segs segments (device pointers inside).  d_eps: device
 * float (trainable eps) or NULL (= 0).  fp32 row-major.  out is [N, sum widths].
 */
int gsn_mp_gin_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N, int64_t E,
                   const GsnSegment *h_segs, int32_t n_segs, const float *d_eps, float *d_out, void *stream);

/*
 * 'ogb' message kind (GSN_edge_sparse_ogb.py:75-84,119-126; MPNN_edge_sparse_ogb):
 *   out[i,:] = (1+eps) * (x[i] + [id[i] if ids are per node])
 *            + sum_{e in row i} relu(x[nbr(e)] + (id[nbr(e)] | id[e]) + ef[e])
 * d_id may be NULL (MPNN).  All operands [.,d].
 */
int gsn_mp_ogb_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N, int64_t E,
                   const float *d_x, const float *d_id, int32_t id_per_edge, const float *d_ef, int32_t d,
                   const float *d_eps, float *d_out, void *stream);

/*
 * Plain segment sum  out[i,:] = sum_{e in row i} msgs[eid(e), :]   (the
 * scatter-add of GSN_sparse.py:140-143 applied to already-computed messages;
 * also the backward of a row gather).  gather=1 reads msgs[nbr(e)] instead
 * (out[i] = sum of neighbour rows).
 */
int gsn_mp_segment_sum(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N,
                       int64_t E, const float *d_msgs, int32_t d, int32_t gather, float *d_out, void *stream);

/*
 * 'general' message kind with the first Linear of msg_fn split algebraically
 * (models_misc.py:52-59 applied to cat(x_i, x_j, ids, ef),
 *  GSN_edge_sparse.py:161-166):
 *   h_e   = P[i, 0:dh] + P[nbr(e), dh:2dh] + Q[e, :]            (pre-activation of fc[0], bias inside Q or P)
 *   S[i]  = sum_{e in row i} act(h_e * scale + shift)            (BatchNorm folded into scale/shift; NULL = identity;
 *                                                                 act: 0 relu, 1 elu, 2 tanh, 3 identity = models_misc.py:5-15)
 * P is [N, 2*dh] (x_i half | x_j half), Q is [E, dh] or NULL.  The second
 * Linear is applied by the caller on N rows: sum_e (W2 h + b2) = W2 S + deg b2.
 * stats != NULL: instead of S, accumulate per-channel sum and sum of squares of
 * h_e over all edges into d_stats [2, dh] (double) for training-mode BatchNorm.
 */
int gsn_mp_general_edge_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N,
                            int64_t E, const float *d_P, const float *d_Q, int32_t dh, const float *d_scale,
                            const float *d_shift, int32_t act, float *d_S, double *d_stats, void *stream);

/* ------------------------------------------------------------------ */
/* dense tail + index ("one-hot free") fast path                      */
/* ------------------------------------------------------------------ */

/*
 * Fused Linear of models_misc.py:52-59 (update_fn, JK projections, the N-row halves
 * of the split msg_fn):
 *   C[m,n] = act( ( sum_k A1[m,k] W[n,k] + sum_k A2[m,k] W[n,K1+k]        cat(A1,A2) @ W^T, no cat tensor
 *                   + row_scale[m]*row_vec[n]                              deg_i * (W b2): bias of the summed messages
 *                   + tab[tab_idx[m], n]                                   one-hot input block folded into a table
 *                   + bias[n] ) * scale[n] + shift[n] )                    BatchNorm1d (eval) as scale/shift
 * W: nn.Linear layout [Nout, K1+K2], row stride ldw.  Optional pointers may be NULL
 * (row_scale/row_vec and tab/tab_idx come in pairs).  act: 0 relu, 1 elu, 2 tanh, 3 identity.
 * fp32 FFMA accumulation.  `vec_ok` is filled in by the library.  The library keeps no state between calls (no global
 * switches, no environment variables): everything that selects a code path is an argument.
 */
typedef struct GsnLinear {
    const float *A1; const float *A2; const float *W;
    const float *bias; const float *row_scale; const float *row_vec;
    const int32_t *tab_idx; const float *tab;
    const float *scale; const float *shift;
    float *C;
    int32_t M, Nout, K1, K2, lda1, lda2, ldw, ldc, tab_ld, act, vec_ok;
    int32_t accumulate;      /* 1: C += result (sum of JK projections, models_graph_classification.py:236-240) */
    int32_t tc_path;         /* gsn_tc_linear_fwd only, GSN_TC_PATH_*: which of its kernels runs (0 = chosen from the shape) */
} GsnLinear;
#define GSN_TC_PATH_AUTO 0
#define GSN_TC_PATH_PRESPLIT 1   /* activations split into (hi, lo) by a pre-pass into d_ws instead of inside the GEMM kernel */
#define GSN_TC_PATH_ONE_TILE 2   /* one output tile per CTA even where the persistent kernel would be chosen */

int gsn_linear_fwd(const GsnLinear *h_p, void *stream);

/*
 * Tensor-core (tcgen05, TMEM accumulator, TMA-staged operands) version of gsn_linear_fwd with
 * fp32-equivalent accuracy through 3xTF32 (a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi).
 *   gsn_split_tf32           x -> (rn_tf32(x), x - rn_tf32(x)), [rows, cols] with row stride ld -> dense [rows, cols];
 *                            used once per weight matrix (W in nn.Linear layout [Nout, K1+K2])
 *   gsn_tc_linear_fwd        same meaning as gsn_linear_fwd; d_Whi / d_Wlo are the split weights.  Activations are
 *                            split inside the kernel (shared memory) when K2 == 0 or K1 % 32 == 0; otherwise by a
 *                            small pre-pass into d_ws (gsn_tc_linear_workspace_bytes; d_ws may be NULL otherwise).  Needs K1, K2, lda1, lda2 % 4 == 0
 *                            and 16-byte aligned operands, otherwise GSN_E_UNSUPPORTED (use gsn_linear_fwd).
 */
int gsn_split_tf32(const float *d_src, int64_t rows, int32_t cols, int32_t ld, float *d_hi, float *d_lo, void *stream);
int gsn_tc_linear_workspace_bytes(int64_t M, int32_t K, size_t *bytes);
int gsn_tc_linear_fwd(const GsnLinear *h_p, const float *d_Whi, const float *d_Wlo, void *d_ws, size_t ws_bytes,
                      void *stream);
/*
 * out[g,:] = sum (mean=1: average) of the rows x[ptr[g] .. ptr[g+1]) : the readouts
 * global_add_pool_sparse / global_mean_pool_sparse (utils_graph_learning.py:23-41) for a
 * PyG batch whose nodes are grouped by graph (ptr = batch.ptr).
 */
int gsn_pool_ptr(const float *d_x, const int64_t *d_ptr, int64_t G, int32_t d, int32_t ldx, int32_t mean,
                 float *d_out, void *stream);

/*
 * Categorical columns -> rows of one concatenated embedding table (replaces
 * one_hot_unique + one_hot_encoder + the first Linear's one-hot block,
 * utils_encoding.py:37-59 / utils_graph_learning.py:170-187):
 *   out[r, c] = table_off[c] + rank_c(src_c[r * stride_c])
 * rank_c = position of the value among the sorted distinct values vocab[vocab_ptr[c] .. vocab_ptr[c+1])
 * (one_hot_unique; clamped to the last entry), or the value itself when the range is empty.
 */
typedef struct GsnEncodeCol {
    const int64_t *src;      /* device */
    int64_t stride;          /* elements between consecutive rows */
    int32_t vocab_begin, vocab_end;   /* into d_vocab; begin == end: identity */
    int32_t table_off;
    int32_t rows;            /* identity columns: number of categories (0 = unchecked); a value outside [0, rows) sets
                                GSN_S_INDEX_RANGE and is clamped (the reference's one-hot scatter_ / nn.Embedding raises) */
} GsnEncodeCol;

#define GSN_MAX_ENCODE_COLS 16
int gsn_encode_rows(const GsnEncodeCol *h_cols, int32_t n_cols, const int64_t *d_vocab, const int32_t *d_perm, int64_t R,
                    int32_t *d_out, int32_t *d_status, void *stream);
/* d_perm (optional): out row r encodes source row d_perm[r];  d_status (optional): GSN_S_INDEX_RANGE / GSN_S_UNSEEN_VALUE */

/* Grouped form: columns with the same h_group[c] are folded into ONE output column as a mixed-radix number,
 *   out[r, g] = sum_{c in g} (table_off_c + rank_c * h_mult[c]),
 * which addresses a pre-summed table of the group's joint vocabulary (gsn_b200/fused.py builds it): the message
 * kernel then gathers one table row per group instead of one per column. */
int gsn_encode_rows_grouped(const GsnEncodeCol *h_cols, int32_t n_cols, const int32_t *h_group, const int32_t *h_mult,
                            int32_t n_groups, const int64_t *d_vocab, const int32_t *d_perm, int64_t R, int32_t *d_out,
                            int32_t *d_status, void *stream);

/*
 * 'general' message kind with categorical inputs kept as indices (layer 0 of the ZINC /
 * IMDB recipes: x, edge features and identifiers are one-hot, so the first Linear of msg_fn
 * is a sum of table rows):
 *   h_e = [P[i,0:dh] + P[nbr,dh:2dh]]  +  sum_c Tn[node_rows[i,c], 0:dh] + sum_c Tn[node_rows[nbr,c], dh:2dh]
 *       + [Q[e,:]] + sum_c Te[edge_rows[e,c], :]
 *   S[i] = sum_{e in row i} act(h_e * scale + shift)
 * Any of the dense (P, Q) and indexed (node_rows/Tn, edge_rows/Te) parts may be NULL.
 */
int gsn_mp_general_edge_idx_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N,
                                int64_t E, const float *d_P, const float *d_Q, const int32_t *d_node_rows,
                                int32_t n_node_cols, const float *d_Tn, const int32_t *d_edge_rows, int32_t n_edge_cols,
                                const float *d_Te, int32_t te_rows, int32_t edge_rows_csr, int32_t dh, const float *d_scale,
                                const float *d_shift, int32_t act, float *d_S, void *stream);
/* te_rows: number of rows of d_Te (<= 4 rows with a single edge column selects a register-resident variant).
 * edge_rows_csr = 1: d_edge_rows is ordered like d_eid (row k belongs to CSR position k; gsn_encode_rows with
 * d_perm = d_eid), which removes one dependent load per edge; 0: indexed by edge_index column. */

/* ------------------------------------------------------------------ */
/* training step (ogbg-molhiv recipe)                                  */
/* ------------------------------------------------------------------ */
/*
 * Embedding bag over categorical columns: out[r,:] = sum_c table_c[idx[r*ld + c], :]  -- multi_embedding with aggr 'sum'
 * (utils_graph_learning.py:134-167) and ogb's AtomEncoder / BondEncoder in one launch; backward adds
 * sum_{r: idx[r,c]=v} grad_out[r,:] into grad_table_c[v,:] (caller zero-initialises the gradient tables; accumulation
 * order is not fixed, like torch's embedding backward).  Values outside a table set GSN_S_INDEX_RANGE (nn.Embedding
 * raises IndexError) and are skipped.
 */
#define GSN_MAX_BAG_COLS 96
typedef struct GsnBagCol {
    const float *d_table;   /* [rows, d] fp32 (forward: weights; backward: gradient of the weights) */
    int32_t rows, _pad;
} GsnBagCol;
int gsn_embedding_bag_fwd(const GsnBagCol *h_cols, int32_t n_cols, const int64_t *d_idx, int64_t ld, int64_t R, int32_t d,
                          float *d_out, int32_t *d_status, void *stream);
int gsn_embedding_bag_bwd(const GsnBagCol *h_grad_cols, int32_t n_cols, const int64_t *d_idx, int64_t ld, int64_t R, int32_t d,
                          const float *d_grad_out, int32_t *d_status, void *stream);

/*
 * Backward of gsn_mp_ogb_fwd (autograd through GSN_edge_sparse_ogb.py:86-129) w.r.t. x, identifiers and edge features.
 * The CSR is the TRANSPOSED grouping (gsn_csr_build keyed by the gathered endpoint j; nbr = aggregation node i):
 *   grad_x[j]  = (1+eps) g[j] + sum_{e: j->i} [x[j] + id + ef[e] > 0] g[i]     (also the gradient of per-node identifiers)
 *   grad_ef[e] = [x[j] + id + ef[e] > 0] g[i]                                  (also the gradient of per-edge identifiers)
 * Nothing of size [E, d] has to be saved by the forward: the relu mask is recomputed from the layer inputs.
 */
int gsn_mp_ogb_bwd(const int32_t *d_rowptr_src, const int32_t *d_eid_src, const int32_t *d_nbr_src, int64_t N, int64_t E,
                   const float *d_x, const float *d_id, int32_t id_per_edge, const float *d_ef, int32_t d, const float *d_eps,
                   const float *d_grad_out, float *d_grad_x, float *d_grad_ef, void *stream);

/* ------------------------------------------------------------------ */
/* whole-model fused forward                                           */
/* ------------------------------------------------------------------ */
/*
 * GNNSubstructures.forward (models_graph_classification.py:204-247, eval mode, 'general' message kind) for a whole
 * batch in ONE launch.  A PyG batch is block-diagonal, so a tile of whole graphs (<= 128 nodes) is self-contained:
 * each CTA carries its tile through every layer (split first Linear of msg_fn, neighbour gather + activation + sum,
 * update_fn with the second message Linear folded in, model BatchNorm + activation, GSN_edge_sparse.py:111-166,
 * models_misc.py:52-59) with all activations in shared / tensor memory and writes only the per-graph readout
 * (utils_graph_learning.py:23-41).  GEMMs: tcgen05, 3 x fp16 with exact power-of-two row scaling (fp32-equivalent).
 *
 * All matrices are padded to D x D (D = 64 or 128 >= every hidden width; padding rows / columns are zero):
 *   d_Whi / d_Wlo  fp16 [n_mats * D, D]: rows scaled per output row by a power of two; layer l owns matrices
 *                  mat0 .. mat0+4 = Wxj, Wxi, U1x, Wf, U2 (has_dense) or mat0, mat0+1 = Wf, U2 (table-only layer 0)
 *   d_vec          fp32 [9, D] per layer: cJ, cI, shiftI, cU, cF, cV, cB, c2s, c2b  (inverse weight scales, BatchNorm
 *                  scale / shift and biases folded per output column; gsn_b200/fused_model.py builds them)
 *   d_Tn [rows, 2D] node-side table rows (P_i half | P_j half), d_Te [te_rows, D] edge-side rows, d_Tu [rows, D]
 *   d_node_rows int32 [N, n_node_cols], d_edge_rows int32 [E, n_edge_cols] in CSR order, d_tu_rows int32 (stride tu_stride)
 *   pool: 0 none, 1 sum, 2 mean -> d_pooled fp32 [G, D];  d_x_out (optional) fp32 [N, D] = the layer's output rows
 * CSR (d_rowptr, d_nbr) from gsn_csr_build.  Graphs must have <= 128 nodes (GSN_S_GRAPH_TOO_LARGE otherwise) and
 * edges must stay inside their graph (GSN_S_CROSS_GRAPH_EDGE).  graphs_per_unit: consecutive graphs handed to a CTA
 * at a time (work granularity; tiles are packed greedily inside a unit).
 */
#define GSN_FUSED_MAX_LAYERS 8
typedef struct GsnFusedLayer {
    const int32_t *d_node_rows; const float *d_Tn;
    const int32_t *d_tu_rows; const float *d_Tu;
    const int32_t *d_edge_rows; const float *d_Te;
    const float *d_vec; float *d_pooled; float *d_x_out;
    int32_t n_node_cols, tu_stride, n_edge_cols, te_rows, has_dense, mat0, act_msg, act_upd, act_out, pool;
    /* JK head of this layer's readout, evaluated in the kernel (models_graph_classification.py:236-240; jk_kind 0 = not
     * here, the caller projects d_pooled): 1 = Linear: out += pooled W1^T + b1;  2 = mlp (models_misc.py:52-59, eval-mode
     * BatchNorm folded): out += act((pooled W0^T + b0) s + t) W1^T + b1.   d_jk_W0T fp32 [D, D] = W0 transposed (k-major),
     * d_jk_vec fp32 [3, D] = b0, s, t; d_jk_W1 fp32 [n_out, D]; d_jk_b1 fp32 [n_out]; zero padding as for the layers. */
    const float *d_jk_W0T; const float *d_jk_vec; const float *d_jk_W1; const float *d_jk_b1;
    int32_t jk_kind, jk_act;
} GsnFusedLayer;
typedef struct GsnFusedModel {
    GsnFusedLayer layers[GSN_FUSED_MAX_LAYERS];
    int32_t n_layers, D, n_mats, graphs_per_unit;
    const void *d_Whi; const void *d_Wlo;
    const int32_t *d_rowptr; const int32_t *d_nbr; const int64_t *d_node_ptr;
    const float *d_x0; int32_t x0_ld, x0_d;
    int64_t N, E, G;
    int32_t *d_status;
    const int32_t *d_tile_plan;   /* optional (gsn_tile_plan): one CTA per tile instead of graphs_per_unit graphs per CTA */
    int32_t max_tiles;            /* capacity the plan was built with */
    int32_t n_out;                /* columns of d_out (<= 32) when any layer has jk_kind != 0 */
    float *d_out;                 /* fp32 [G, n_out]: sum of the in-kernel JK projections (every row is written) */
} GsnFusedModel;
int gsn_fused_model_fwd(const GsnFusedModel *h_m, void *stream);
/*
 * Tiles of the one-kernel forward for a small batch: consecutive whole graphs packed greedily into tiles of <= 128 rows and
 * <= 32 graphs (the collation of main.py:243-258 decides which graphs share a batch; this decides which share an SM).
 * d_tile_plan: int32 [max_tiles + 2] = { n_tiles, first graph of tile 0, ..., first graph of tile n_tiles - 1, G }.
 * A graph with more than 128 rows becomes a tile of its own (gsn_fused_model_fwd then reports GSN_S_GRAPH_TOO_LARGE).
 * max_tiles >= min(G, 2 * N / 128 + G / 32 + 2) always suffices; G <= 8192 (one CTA scans node_ptr from shared memory).
 */
int gsn_tile_plan(const int64_t *d_node_ptr, int64_t G, int32_t *d_tile_plan, int32_t max_tiles, int32_t *d_status,
                  void *stream);

/*
 * Submission of one captured step on `stream` (host side only; see csrc/runtime.cu): copy in_bytes from src (pinned host
 * or device) into the captured step's static input buffer d_in, launch graph_exec (a cudaGraphExec_t, e.g.
 * torch.cuda.CUDAGraph.raw_cuda_graph_exec()), then copy out_bytes of the static output d_out to pinned host memory h_out.
 * in_bytes == 0 / out_bytes == 0 skip the copies.  Replaces the per-step body of the reference's evaluation loop
 * (train_test_funcs.py:190-215: `data.to(device)`, `model(data)`, `.cpu()`) for a captured step.
 */
int gsn_submit_step(void *graph_exec, void *d_in, const void *src, size_t in_bytes, void *h_out, const void *d_out,
                    size_t out_bytes, void *stream);

/* ------------------------------------------------------------------ */
/* DGN consumer of COUNT (directional_gsn/)                            */
/* ------------------------------------------------------------------ */
/*
 * Replaces DGNLayerSimple.pretrans_edges / message_func / reduce_func (directional_gsn/nets/dgn_layer.py:28-54), the
 * aggregators of nets/aggregators.py:8-69 and the scalers of nets/scalers.py:7-20 (DGL mailboxes per in-degree bucket,
 * one chain of torch ops per aggregator).  CSR from gsn_csr_build over key = edge_index[1] (messages flow src -> dst).
 *   vector_field(e = j->i) = [ node_field[j,:] - node_field[i,:]  |  edge_field[e,:] ]      (dgn_layer.py:28-35)
 *   out[i, s*(A*d) + a*d + c] = scaler_s( aggregator_a( {h[j,c]}, vector_field, h[i,c] ) )
 * Nodes without in-edges get zeros (DGL leaves rows that receive no message zero-filled).  With n_scalers == 1 no
 * scaling is applied whatever the scaler is (dgn_layer.py:50-51).  avg_log = avg_d["log"] of the training set.
 * d_scratch: 4 * E floats (per-edge weights of up to four directional aggregators at a time; may be NULL when the list
 * holds none).  Launches: per group of four directional aggregators one weight kernel (thread per node) + one aggregation
 * kernel (thread per node and 4-channel chunk, one pass over the in-edges).
 */
#define GSN_DGN_MEAN 0
#define GSN_DGN_SUM 1
#define GSN_DGN_MAX 2
#define GSN_DGN_MIN 3
#define GSN_DGN_STD 4
#define GSN_DGN_VAR 5
#define GSN_DGN_DIR_AV 6            /* dir{k}-av        aggregators.py:35-39 */
#define GSN_DGN_DIR_SOFTMAX 7       /* dir{k}-{alpha}   aggregators.py:42-45 */
#define GSN_DGN_DIR_DX 8            /* dir{k}-dx        aggregators.py:48-52 */
#define GSN_DGN_DIR_DX_NO_ABS 9     /* aggregators.py:55-59 */
#define GSN_DGN_DIR_DX_BALANCED 10  /* aggregators.py:62-71 */
#define GSN_DGN_SCALE_IDENTITY 0
#define GSN_DGN_SCALE_AMPLIFICATION 1
#define GSN_DGN_SCALE_ATTENUATION 2
#define GSN_DGN_MAX_AGGR 16
#define GSN_DGN_MAX_SCALERS 3
typedef struct GsnDgnAggr {
    int32_t kind;    /* GSN_DGN_* */
    int32_t field;   /* eig_idx: component of the vector field (directional kinds) */
    float alpha;     /* softmax temperature (GSN_DGN_DIR_SOFTMAX) */
    int32_t _pad;
} GsnDgnAggr;
int gsn_dgn_aggregate_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N, int64_t E,
                          const float *d_h, int32_t d, const float *d_node_field, int32_t Fn, const float *d_edge_field,
                          int32_t Fe, const GsnDgnAggr *h_aggr, int32_t n_aggr, const int32_t *h_scalers,
                          int32_t n_scalers, float avg_log, float *d_out, float *d_scratch, void *stream);
/* Backward w.r.t. h (autograd of aggregators.py:8-69; the fields are data).  Per in-edge k (j -> i) the gradient that
 * flows to h[j] is written to d_M[eid[k], 0:d] (edge-id order, [E, d]) and the gradient to the node's own row (dx kinds)
 * to d_SG [N, d]:  grad_h = d_SG + gsn_mp_segment_sum(plan grouped by edge_index[0], d_M).  Deterministic. */
int gsn_dgn_aggregate_bwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N, int64_t E,
                          const float *d_h, int32_t d, const float *d_node_field, int32_t Fn, const float *d_edge_field,
                          int32_t Fe, const GsnDgnAggr *h_aggr, int32_t n_aggr, const int32_t *h_scalers,
                          int32_t n_scalers, float avg_log, const float *d_grad_out, float *d_M, float *d_SG, void *stream);

/* ------------------------------------------------------------------ */
/* misc                                                                */
/* ------------------------------------------------------------------ */
int gsn_abi_version(void);
/* Number of CUDA kernels this library has launched (or recorded into a capturing stream) so far. */
uint64_t gsn_launch_count(void);
/* Last CUDA error string seen by this thread inside the library (static storage). */
const char *gsn_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GSN_B200_H_ */
