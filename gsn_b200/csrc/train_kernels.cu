// Training-step kernels of the ogbg-molhiv recipe (BASELINE config 3):
//
//   * embedding bag over categorical columns, forward and backward: the sum of per-column nn.Embedding lookups of
//     multi_embedding (aggr 'sum', utils_graph_learning.py:134-167 -- 72 identifier columns in the molhiv recipe) and of
//     ogb's AtomEncoder / BondEncoder (9 / 3 columns) in ONE launch each way instead of one lookup + one add (forward)
//     and one scatter per column (backward);
//   * backward of the 'ogb' message kind (GSN_edge_sparse_ogb.py:86-129 through autograd): one pass over the edges
//     grouped by SOURCE node recomputes the relu mask from the layer inputs and produces the gradients w.r.t. x
//     (identifiers) and the per-edge features -- no [E, d] tensor is saved by the forward.
#include "common.cuh"

namespace gsn {

struct BagParams {
    const float *tab[GSN_MAX_BAG_COLS];     // forward: tables; backward: gradient tables (zero-initialised by the caller)
    int32_t rows_of[GSN_MAX_BAG_COLS];      // rows of every table
    const int64_t *idx;                     // [R, ld] categorical values, column c at idx[r * ld + c]
    int64_t ld;
    int64_t R;
    int32_t n_cols, d;
    float *out;                             // forward [R, d]
    const float *grad_out;                  // backward [R, d]
    int32_t rows_per_cta;
    int32_t *status;
};

// out[r, :] = sum_c tab_c[idx[r, c], :]
template <int VEC>
__global__ void __launch_bounds__(256) bag_fwd_kernel(const __grid_constant__ BagParams p) {
    const int cpr = p.d / VEC;
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= p.R * cpr) return;
    const int64_t r = t / cpr;
    const int c0 = (int)(t % cpr) * VEC;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    const int64_t *row = p.idx + r * p.ld;
    for (int c = 0; c < p.n_cols; ++c) {
        const int64_t v = __ldg(row + c);
        if (v < 0 || v >= p.rows_of[c]) {           // nn.Embedding raises IndexError; here: status bit, row skipped
            atomicOr(p.status, GSN_S_INDEX_RANGE);
            continue;
        }
        const float *src = p.tab[c] + v * p.d + c0;
        if (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(src));
            acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += __ldg(src + i);
        }
    }
    float *dst = p.out + r * p.d + c0;
    if (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
#pragma unroll
        for (int i = 0; i < VEC; ++i) dst[i] = acc[i];
}

// grad_tab_c[v, :] += sum_{r : idx[r, c] = v} grad_out[r, :].  One CTA = (block of 1024 rows, column c); the column's values
// of the block are staged in shared memory.  No shared-memory fp32 atomics (a compare-and-swap loop that serialises on the
// few distinct values of a categorical column); the only atomics are one global fp32 add per (table row seen by the
// block, channel).  Two mappings:
//   many values (V > 8): every warp OWNS table rows v = warp, warp + 8, ...: it finds the block's rows holding v with
//                        ballots over the staged values and sums their gradient rows in registers (lanes across d), four
//                        rows in flight at a time;
//   few values (V <= 8, e.g. the bond features): every warp takes a slice of the rows and adds each gradient row into its
//                        private [V, d] buffer in shared memory (plain read-modify-write: one owner per address), four
//                        rows in flight; the eight buffers are then summed.
constexpr int kBagRows = 1024;          // rows per CTA
constexpr int kBagMaxD = 1024;          // channels a lane can hold (32 per lane)
constexpr int kBagSmallV = 8;
template <int PER>          // channels per lane: d <= 32 * PER
__global__ void __launch_bounds__(256) bag_bwd_kernel(const __grid_constant__ BagParams p) {
    extern __shared__ float bag_buf[];                  // few values: [8 warps][V * d]
    __shared__ int32_t vals[kBagRows];
    const int c = blockIdx.y;
    const int V = p.rows_of[c], d = p.d;
    float *gt = const_cast<float *>(p.tab[c]);
    const int64_t r0 = (int64_t)blockIdx.x * kBagRows;
    const int n = (int)((r0 + kBagRows < p.R ? r0 + kBagRows : p.R) - r0);
    const bool small = V <= kBagSmallV && p.rows_per_cta != 0;       // rows_per_cta != 0: the launch provided the buffers
    for (int i = threadIdx.x; i < n; i += 256) {
        const int64_t v = __ldg(p.idx + (r0 + i) * p.ld + c);
        vals[i] = (v >= 0 && v < V) ? (int32_t)v : -1;           // out of range: flagged by the forward, skipped here
    }
    if (small)
        for (int i = threadIdx.x; i < 8 * V * d; i += 256) bag_buf[i] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *gbase = p.grad_out + r0 * d;
    if (small) {
        float *mine = bag_buf + (size_t)warp * V * d;
        const int per_warp = (n + 7) / 8;
        const int a = warp * per_warp, e = min(n, a + per_warp);
        for (int r = a; r < e; r += 4) {
            float g[4][PER];
            int v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = r + q < e ? vals[r + q] : -1;
#pragma unroll
                for (int j = 0; j < PER; ++j)
                    g[q][j] = (v[q] >= 0 && lane + 32 * j < d) ? __ldg(gbase + (int64_t)(r + q) * d + lane + 32 * j) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (v[q] < 0) continue;
#pragma unroll
                for (int j = 0; j < PER; ++j)
                    if (lane + 32 * j < d) mine[v[q] * d + lane + 32 * j] += g[q][j];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < V * d; i += 256) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += bag_buf[(size_t)w * V * d + i];
            if (t != 0.f) atomicAdd(gt + i, t);
        }
        return;
    }
    __shared__ uint16_t match[8][kBagRows];          // per warp: the block's rows that hold the warp's current value
    for (int v = warp; v < V; v += 8) {
        int cnt = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool hit = i < n && vals[i] == v;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) match[warp][cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
            cnt += __popc(m);
        }
        if (cnt == 0) continue;
        __syncwarp();
        float acc[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) acc[j] = 0.f;
        for (int k = 0; k < cnt; k += 4) {                   // four gradient rows in flight
            float g[4][PER];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int row = k + q < cnt ? (int)match[warp][k + q] : -1;
#pragma unroll
                for (int j = 0; j < PER; ++j)
                    g[q][j] = (row >= 0 && lane + 32 * j < d) ? __ldg(gbase + (int64_t)row * d + lane + 32 * j) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < PER; ++j) acc[j] += g[q][j];
        }
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (lane + 32 * j < d) atomicAdd(gt + (int64_t)v * d + lane + 32 * j, acc[j]);
        __syncwarp();
    }
}

struct OgbBwdParams {
    const int32_t *rowptr, *eid, *nbr;      // edges grouped by SOURCE node j: eid = edge_index column, nbr = aggregation node i
    int64_t N;
    const float *x, *id, *ef, *eps, *g;
    int32_t d, id_per_edge;
    float *grad_x, *grad_ef;
};

// forward: out[i] = (1+eps) (x[i] [+ id[i]]) + sum_{e: j -> i} relu(x[j] + (id[j] | id[e]) + ef[e])
// backward for node j:  grad_x[j] = (1+eps) g[j] + sum_{e: j -> i} [pre_e > 0] g[i]   ( = grad_id[j] for per-node ids)
//                       grad_ef[e] = [pre_e > 0] g[i]                                  ( = grad_id[e] for per-edge ids)
template <int VEC>
__global__ void __launch_bounds__(256) ogb_bwd_kernel(const __grid_constant__ OgbBwdParams p) {
    const int cpr = p.d / VEC;
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= p.N * cpr) return;
    const int64_t j = t / cpr;
    const int c0 = (int)(t % cpr) * VEC;
    auto ld = [&](const float *ptr, float (&v)[VEC]) {
        if (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(ptr));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = __ldg(ptr + i);
        }
    };
    auto st = [&](float *ptr, const float (&v)[VEC]) {
        if (VEC == 4) *reinterpret_cast<float4 *>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
        else
#pragma unroll
            for (int i = 0; i < VEC; ++i) ptr[i] = v[i];
    };
    float base[VEC], acc[VEC], gj[VEC];
    ld(p.x + j * p.d + c0, base);
    if (p.id && !p.id_per_edge) {
        float idv[VEC];
        ld(p.id + j * p.d + c0, idv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) base[i] += idv[i];
    }
    ld(p.g + j * p.d + c0, gj);
    const float one_eps = 1.0f + (p.eps ? __ldg(p.eps) : 0.0f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = one_eps * gj[i];
    const int kend = __ldg(p.rowptr + j + 1);
    for (int k = __ldg(p.rowptr + j); k < kend; ++k) {
        const int64_t e = __ldg(p.eid + k), i_node = __ldg(p.nbr + k);
        float pre[VEC], gi[VEC], m[VEC];
        ld(p.ef + e * p.d + c0, pre);
        if (p.id && p.id_per_edge) {
            float idv[VEC];
            ld(p.id + e * p.d + c0, idv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) pre[i] += idv[i];
        }
        ld(p.g + i_node * p.d + c0, gi);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            m[i] = (pre[i] + base[i] > 0.0f) ? gi[i] : 0.0f;
            acc[i] += m[i];
        }
        st(p.grad_ef + e * p.d + c0, m);
    }
    st(p.grad_x + j * p.d + c0, acc);
}

inline bool tk_aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace gsn

using namespace gsn;

static int fill_bag(BagParams &p, const GsnBagCol *h_cols, int32_t n_cols, const int64_t *d_idx, int64_t ld, int64_t R, int32_t d,
                    int32_t *d_status) {
    if (!h_cols || n_cols < 1 || n_cols > GSN_MAX_BAG_COLS || !d_idx || R < 0 || d < 1 || ld < n_cols || !d_status) return GSN_E_INVALID;
    for (int c = 0; c < n_cols; ++c) {
        if (!h_cols[c].d_table || h_cols[c].rows < 1) return GSN_E_INVALID;
        p.tab[c] = h_cols[c].d_table;
        p.rows_of[c] = h_cols[c].rows;
    }
    p.idx = d_idx; p.ld = ld; p.R = R; p.n_cols = n_cols; p.d = d; p.status = d_status;
    p.out = nullptr; p.grad_out = nullptr; p.rows_per_cta = 0;
    return GSN_OK;
}

extern "C" int gsn_embedding_bag_fwd(const GsnBagCol *h_cols, int32_t n_cols, const int64_t *d_idx, int64_t ld, int64_t R, int32_t d,
                                     float *d_out, int32_t *d_status, void *stream_) {
    BagParams p;
    int rc = fill_bag(p, h_cols, n_cols, d_idx, ld, R, d, d_status);
    if (rc) return rc;
    if (!d_out) return GSN_E_INVALID;
    if (R == 0) return GSN_OK;
    p.out = d_out;
    bool v4 = d % 4 == 0 && tk_aligned16(d_out);
    for (int c = 0; c < n_cols; ++c) v4 = v4 && tk_aligned16(h_cols[c].d_table);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (v4) bag_fwd_kernel<4><<<(unsigned)ceil_div(R * (d / 4), 256), 256, 0, stream>>>(p);
    else bag_fwd_kernel<1><<<(unsigned)ceil_div(R * (int64_t)d, 256), 256, 0, stream>>>(p);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("gsn_embedding_bag_fwd");
    return GSN_OK;
}

extern "C" int gsn_embedding_bag_bwd(const GsnBagCol *h_grad_cols, int32_t n_cols, const int64_t *d_idx, int64_t ld, int64_t R,
                                     int32_t d, const float *d_grad_out, int32_t *d_status, void *stream_) {
    BagParams p;
    int rc = fill_bag(p, h_grad_cols, n_cols, d_idx, ld, R, d, d_status);
    if (rc) return rc;
    if (!d_grad_out) return GSN_E_INVALID;
    if (R == 0) return GSN_OK;
    p.grad_out = d_grad_out;
    if (d > kBagMaxD) return GSN_E_UNSUPPORTED;
    dim3 grid((unsigned)ceil_div(R, kBagRows), (unsigned)n_cols);
    cudaStream_t stream = (cudaStream_t)stream_;
    // few-valued columns: eight private [V, d] buffers per CTA (<= 8 * 8 * d floats)
    int vmin = 1 << 30;
    for (int c = 0; c < n_cols; ++c) vmin = h_grad_cols[c].rows < vmin ? h_grad_cols[c].rows : vmin;
    size_t smem = 0;
    if (vmin <= kBagSmallV && (size_t)8 * kBagSmallV * d * sizeof(float) <= 160 * 1024) {
        smem = (size_t)8 * kBagSmallV * d * sizeof(float);
        p.rows_per_cta = 1;
    }
    if (d <= 128) {
        GSN_CUDA_OK(cudaFuncSetAttribute(bag_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        bag_bwd_kernel<4><<<grid, 256, smem, stream>>>(p);
    } else if (d <= 320) {
        GSN_CUDA_OK(cudaFuncSetAttribute(bag_bwd_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        bag_bwd_kernel<10><<<grid, 256, smem, stream>>>(p);
    } else if (d <= 512) {
        GSN_CUDA_OK(cudaFuncSetAttribute(bag_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        bag_bwd_kernel<16><<<grid, 256, smem, stream>>>(p);
    } else {
        GSN_CUDA_OK(cudaFuncSetAttribute(bag_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        bag_bwd_kernel<32><<<grid, 256, smem, stream>>>(p);
    }
    GSN_BUMP(1);
    GSN_LAUNCH_OK("gsn_embedding_bag_bwd");
    return GSN_OK;
}

extern "C" int gsn_mp_ogb_bwd(const int32_t *d_rowptr_src, const int32_t *d_eid_src, const int32_t *d_nbr_src, int64_t N, int64_t E,
                              const float *d_x, const float *d_id, int32_t id_per_edge, const float *d_ef, int32_t d,
                              const float *d_eps, const float *d_grad_out, float *d_grad_x, float *d_grad_ef, void *stream_) {
    if (N < 0 || E < 0 || d < 1 || !d_rowptr_src || !d_x || !d_grad_out || !d_grad_x || (E > 0 && (!d_ef || !d_grad_ef || !d_eid_src || !d_nbr_src)))
        return GSN_E_INVALID;
    if (N == 0) return GSN_OK;
    OgbBwdParams p{d_rowptr_src, d_eid_src, d_nbr_src, N, d_x, d_id, d_ef, d_eps, d_grad_out, d, id_per_edge, d_grad_x, d_grad_ef};
    const bool v4 = d % 4 == 0 && tk_aligned16(d_x) && tk_aligned16(d_id) && tk_aligned16(d_ef) && tk_aligned16(d_grad_out) &&
                    tk_aligned16(d_grad_x) && tk_aligned16(d_grad_ef);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (v4) ogb_bwd_kernel<4><<<(unsigned)ceil_div(N * (d / 4), 256), 256, 0, stream>>>(p);
    else ogb_bwd_kernel<1><<<(unsigned)ceil_div(N * (int64_t)d, 256), 256, 0, stream>>>(p);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("gsn_mp_ogb_bwd");
    return GSN_OK;
}
