// Directional aggregation of the DGN consumer of COUNT.
//
// Replaces DGNLayerSimple.pretrans_edges / message_func / reduce_func of the reference
// (directional_gsn/nets/dgn_layer.py:28-54) together with the aggregator and scaler
// functions it dispatches to (directional_gsn/nets/aggregators.py:8-69, nets/scalers.py:7-20):
// DGL builds a mailbox [n, D, d] per in-degree bucket D, evaluates every aggregator on it
// as a separate chain of torch ops and concatenates; the "vector field" of an edge is
// eig[src] - eig[dst] (node fields, i.e. vertex-scope substructure counts) followed by the
// edge fields (edge-scope counts) -- data/HIV.py:91-98.
//
// Here: the in-edges of a node come from the same CSR the message-passing kernels use
// (ascending edge id inside a row = DGL's mailbox order), one thread owns one (node,
// 4-channel chunk) and produces every aggregator x scaler block of the output row.  The
// per-edge weights of the directional aggregators depend only on the fields, so they are
// computed once per aggregator and applied while h_j streams through L1; nothing of size
// [E, d] or [n, D, d] is materialised.
#include "common.cuh"

namespace gsn {

constexpr float kDgnEps = 1e-8f;      // aggregators.py:5

struct DgnParams {
    const int32_t *rowptr, *eid, *nbr;
    int64_t N;
    const float *h, *node_field, *edge_field;
    int d, Fn, Fe, n_aggr, n_scalers;
    int aggr_kind[GSN_DGN_MAX_AGGR], aggr_idx[GSN_DGN_MAX_AGGR];
    float aggr_alpha[GSN_DGN_MAX_AGGR];
    int scaler_kind[GSN_DGN_MAX_SCALERS];
    float avg_log;
    float *out;
};

template <int VEC> __device__ __forceinline__ void ld_vec(const float *src, float (&dst)[VEC]) {
    if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(src));
        dst[0] = t.x; dst[1 % VEC] = t.y; dst[2 % VEC] = t.z; dst[3 % VEC] = t.w;
    } else {
        dst[0] = __ldg(src);
    }
}

// component `f` of the vector field of CSR position k (in-edge j -> i): dgn_layer.py:28-35
__device__ __forceinline__ float dgn_field(const DgnParams &p, int f, int64_t i, int j, int k) {
    if (f < p.Fn) return __ldg(p.node_field + (int64_t)j * p.Fn + f) - __ldg(p.node_field + i * p.Fn + f);
    return __ldg(p.edge_field + (int64_t)__ldg(p.eid + k) * p.Fe + (f - p.Fn));
}

// ---- forward, step 1: the per-edge weights of up to four directional aggregators (one group), once per NODE.
// The weights depend on the fields only, not on the channel: the 15-75 threads that own the channel chunks of a node used
// to recompute them each (ncu: the old one-kernel forward issued 2,960 instructions per thread and ran at 74 % issue
// utilisation -- instruction bound at 19 % of the HBM peak).  W[k, q] (CSR position k, aggregator q of the group).
struct DgnGroup {
    int n;                 // directional aggregators in this group (<= 4)
    int kind[4], field[4], out_block[4];
    float alpha[4];
};

__global__ void __launch_bounds__(256) dgn_weights_kernel(const __grid_constant__ DgnParams p, const __grid_constant__ DgnGroup g,
                                                          float *__restrict__ W) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= p.N) return;
    const int k0 = __ldg(p.rowptr + i), k1 = __ldg(p.rowptr + i + 1);
    if (k1 == k0) return;
    // pass 1: the norms of all aggregators of the group in one walk over the in-edges
    float n_abs[4], n_pos[4], n_neg[4], mxs[4], se[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { n_abs[q] = 0.f; n_pos[q] = 0.f; n_neg[q] = 0.f; mxs[q] = -INFINITY; se[q] = 0.f; }
    bool softmax = false;
    for (int k = k0; k < k1; ++k) {
        const int j = __ldg(p.nbr + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < g.n) {
                const float F = dgn_field(p, g.field[q], i, j, k);
                n_abs[q] += fabsf(F); n_pos[q] += fmaxf(F, 0.f); n_neg[q] += fmaxf(-F, 0.f);
                mxs[q] = fmaxf(mxs[q], g.alpha[q] * fabsf(F));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) softmax = softmax || (q < g.n && g.kind[q] == GSN_DGN_DIR_SOFTMAX);
    if (softmax) {
        for (int k = k0; k < k1; ++k) {
            const int j = __ldg(p.nbr + k);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < g.n && g.kind[q] == GSN_DGN_DIR_SOFTMAX)
                    se[q] += expf(g.alpha[q] * fabsf(dgn_field(p, g.field[q], i, j, k)) - mxs[q]);
        }
    }
    // pass 2: the weights, one 16-byte store per edge
    for (int k = k0; k < k1; ++k) {
        const int j = __ldg(p.nbr + k);
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            w[q] = 0.f;
            if (q < g.n) {
                const int kind = g.kind[q];
                const float F = dgn_field(p, g.field[q], i, j, k);
                if (kind == GSN_DGN_DIR_AV) w[q] = fabsf(F) / (n_abs[q] + kDgnEps);
                else if (kind == GSN_DGN_DIR_SOFTMAX) w[q] = expf(g.alpha[q] * fabsf(F) - mxs[q]) / se[q];
                else if (kind == GSN_DGN_DIR_DX_BALANCED)
                    w[q] = (fmaxf(F, 0.f) / (n_pos[q] + kDgnEps) + fmaxf(-F, 0.f) / (n_neg[q] + kDgnEps)) / 2.0f;
                else w[q] = F / (n_abs[q] + kDgnEps);
            }
        }
        reinterpret_cast<float4 *>(W)[k] = make_float4(w[0], w[1], w[2], w[3]);
    }
}

// ---- forward, step 2: one thread per (node, VEC-channel chunk), ONE pass over the node's in-edges: every h_j chunk is
// loaded once and feeds the statistics (mean / sum / max / min / std / var aggregators, when `stats`) and the NDIR
// directional accumulators of the group (weights from step 1, one 16-byte load per edge).
template <int VEC, int NDIR, bool STATS>
__global__ void __launch_bounds__(256, 4) dgn_aggregate_kernel(const __grid_constant__ DgnParams p, const __grid_constant__ DgnGroup g,
                                                            const float *__restrict__ W, int cpr, int npb) {
    constexpr bool stats = STATS;
    const int ln = (int)threadIdx.x / cpr;
    if (ln >= npb) return;
    const int64_t i = (int64_t)blockIdx.x * npb + ln;
    if (i >= p.N) return;
    const int d = p.d;
    const int c = ((int)threadIdx.x - ln * cpr) * VEC;
    const int k0 = __ldg(p.rowptr + i), k1 = __ldg(p.rowptr + i + 1);
    const int D = k1 - k0;
    const int AD = p.n_aggr * d;
    float *orow = p.out + i * (int64_t)(AD * p.n_scalers) + c;
    // scaler factors (scalers.py:7-20); applied only when more than one scaler is listed (dgn_layer.py:50-51)
    float sfac[GSN_DGN_MAX_SCALERS];
#pragma unroll
    for (int s = 0; s < GSN_DGN_MAX_SCALERS; ++s) {
        float f = 1.0f;
        if (s < p.n_scalers && p.n_scalers > 1 && D > 0) {
            const double lg = log((double)D + 1.0);
            if (p.scaler_kind[s] == 1) f = (float)(lg / (double)p.avg_log);
            else if (p.scaler_kind[s] == 2) f = (float)((double)p.avg_log / lg);
        }
        sfac[s] = f;
    }
    float s1[VEC], s2[VEC], mx[VEC], mn[VEC], r[NDIR > 0 ? NDIR : 1][VEC], wsum[NDIR > 0 ? NDIR : 1];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s1[v] = 0.f; s2[v] = 0.f; mx[v] = -INFINITY; mn[v] = INFINITY; }
#pragma unroll
    for (int q = 0; q < (NDIR > 0 ? NDIR : 1); ++q) {
        wsum[q] = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[q][v] = 0.f;
    }
    const float *hc = p.h + c;
    for (int k = k0; k < k1; ++k) {
        float xs[VEC];
        ld_vec<VEC>(hc + (int64_t)__ldg(p.nbr + k) * d, xs);
        if (stats) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float x = xs[v];
                s1[v] += x; s2[v] = __fadd_rn(s2[v], __fmul_rn(x, x)); mx[v] = fmaxf(mx[v], x); mn[v] = fminf(mn[v], x);
            }
        }
        if (NDIR > 0) {
            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(W) + k);
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int q = 0; q < NDIR; ++q) {
                wsum[q] += w[q];
#pragma unroll
                for (int v = 0; v < VEC; ++v) r[q][v] += xs[v] * w[q];
            }
        }
    }
    auto store = [&](int block, const float (&val)[VEC]) {
        for (int s = 0; s < p.n_scalers; ++s) {
            float *o = orow + s * AD + block * d;
            // the output is written once and never read by this call: streaming stores keep it from displacing h in L2
            if (VEC == 4) __stcs(reinterpret_cast<float4 *>(o), make_float4(val[0] * sfac[s], val[1 % VEC] * sfac[s], val[2 % VEC] * sfac[s], val[3 % VEC] * sfac[s]));
            else __stcs(o, val[0] * sfac[s]);
        }
    };
    if (stats) {
        for (int a = 0; a < p.n_aggr; ++a) {
            const int kind = p.aggr_kind[a];
            if (kind > GSN_DGN_VAR) continue;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // separately rounded products, as the reference's chain of torch ops (aggregators.py:24-28): an
                // fma here would turn the exact zero of a one-message mailbox into the rounding error of h*h,
                // which sqrt(var + 1e-8) amplifies to ~1e-4
                const float mean = __fdiv_rn(s1[v], (float)D);
                const float var = fmaxf(__fsub_rn(__fdiv_rn(s2[v], (float)D), __fmul_rn(mean, mean)), 0.f);
                o[v] = D == 0 ? 0.f : kind == GSN_DGN_MEAN ? mean : kind == GSN_DGN_SUM ? s1[v] : kind == GSN_DGN_MAX ? mx[v]
                     : kind == GSN_DGN_MIN ? mn[v] : kind == GSN_DGN_STD ? sqrtf(var + kDgnEps) : var;
            }
            store(a, o);
        }
    }
    if (NDIR > 0) {
        float hin[VEC];
        ld_vec<VEC>(hc + i * d, hin);
#pragma unroll
        for (int q = 0; q < NDIR; ++q) {
            const int kind = g.kind[q];
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float t = r[q][v];
                if (kind == GSN_DGN_DIR_DX || kind == GSN_DGN_DIR_DX_BALANCED) t = fabsf(t - wsum[q] * hin[v]);
                else if (kind == GSN_DGN_DIR_DX_NO_ABS) t = t - wsum[q] * hin[v];
                o[v] = D == 0 ? 0.f : t;
            }
            store(g.out_block[q], o);
        }
    }
}

// ------------------------------------------------------------------ backward
// d(out)/d(h) of dgn_aggregate_kernel.  Same ownership as the forward (one thread per destination node and 4-channel
// chunk): it recomputes the node's statistics / directional weights and writes, for every in-edge k (j -> i), the
// gradient that flows to the SOURCE row h[j] into M[eid[k], :] (edge-id order) plus the gradient to its own row
// (the -(sum w) h_i term of the dx aggregators) into SG[i, :].  grad_h = SG + segment-sum of M over the grouping by
// source (gsn_mp_segment_sum with the transposed plan): deterministic, no float atomics.  Formulas = autograd of
// aggregators.py:8-69 (relu'(0) = 0, sign(0) = 0, max / min route to the first extremum in mailbox order).
template <int VEC>
__global__ void __launch_bounds__(256) dgn_aggregate_bwd_kernel(const __grid_constant__ DgnParams p, const float *__restrict__ gout,
                                                                float *__restrict__ M, float *__restrict__ SG) {
    const int d = p.d, cpr = d / VEC;
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= p.N * cpr) return;
    const int64_t i = t / cpr;
    const int c = (int)(t % cpr) * VEC;
    const int k0 = __ldg(p.rowptr + i), k1 = __ldg(p.rowptr + i + 1);
    const int D = k1 - k0;
    const int AD = p.n_aggr * d;
    float sg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) sg[v] = 0.f;
    if (D == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) SG[i * d + c + v] = 0.f;
        return;
    }
    float sfac[GSN_DGN_MAX_SCALERS];
    for (int s = 0; s < p.n_scalers; ++s) {
        float f = 1.0f;
        if (p.n_scalers > 1) {
            const double lg = log((double)D + 1.0);
            if (p.scaler_kind[s] == 1) f = (float)(lg / (double)p.avg_log);
            else if (p.scaler_kind[s] == 2) f = (float)((double)p.avg_log / lg);
        }
        sfac[s] = f;
    }
    const float *grow = gout + i * (int64_t)(AD * p.n_scalers);
    auto upstream = [&](int a, float (&G)[VEC]) {            // sum over the scaler blocks of aggregator a
#pragma unroll
        for (int v = 0; v < VEC; ++v) G[v] = 0.f;
        for (int s = 0; s < p.n_scalers; ++s) {
            float gv[VEC];
            ld_vec<VEC>(grow + s * AD + a * d + c, gv);
#pragma unroll
            for (int v = 0; v < VEC; ++v) G[v] = fmaf(sfac[s], gv[v], G[v]);
        }
    };
    float hin[VEC];
    ld_vec<VEC>(p.h + i * d + c, hin);
    // ---- basic aggregators: upstream gradients summed per kind, one statistics pass, one emit pass
    float Gk[6][VEC];
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int v = 0; v < VEC; ++v) Gk[q][v] = 0.f;
    bool any_basic = false;
    for (int a = 0; a < p.n_aggr; ++a) {
        const int kind = p.aggr_kind[a];
        if (kind > GSN_DGN_VAR) continue;
        any_basic = true;
        float G[VEC];
        upstream(a, G);
#pragma unroll
        for (int q = 0; q < 6; ++q)
            if (q == kind)
#pragma unroll
                for (int v = 0; v < VEC; ++v) Gk[q][v] += G[v];
    }
    float s1[VEC], s2[VEC], mx[VEC], mn[VEC];
    int kmx[VEC], kmn[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s1[v] = 0.f; s2[v] = 0.f; mx[v] = -INFINITY; mn[v] = INFINITY; kmx[v] = k0; kmn[v] = k0; }
    if (any_basic) {
        for (int k = k0; k < k1; ++k) {
            float xs[VEC];
            ld_vec<VEC>(p.h + (int64_t)__ldg(p.nbr + k) * d + c, xs);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float x = xs[v];
                s1[v] += x; s2[v] = __fadd_rn(s2[v], __fmul_rn(x, x));
                if (x > mx[v]) { mx[v] = x; kmx[v] = k; }
                if (x < mn[v]) { mn[v] = x; kmn[v] = k; }
            }
        }
    }
    float m1[VEC], cvar[VEC];          // mean, and d(loss)/d(var) incl. the std branch, zero where relu clipped
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        m1[v] = __fdiv_rn(s1[v], (float)D);
        const float raw = __fsub_rn(__fdiv_rn(s2[v], (float)D), __fmul_rn(m1[v], m1[v]));
        const float var = fmaxf(raw, 0.f);
        cvar[v] = raw > 0.f ? Gk[GSN_DGN_VAR][v] + Gk[GSN_DGN_STD][v] / (2.0f * sqrtf(var + kDgnEps)) : 0.f;
    }
    for (int k = k0; k < k1; ++k) {                       // first write of every M row of this node
        float xs[VEC], m[VEC];
        ld_vec<VEC>(p.h + (int64_t)__ldg(p.nbr + k) * d + c, xs);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float acc = 0.f;
            if (any_basic) {
                acc = Gk[GSN_DGN_MEAN][v] / (float)D + Gk[GSN_DGN_SUM][v];
                if (k == kmx[v]) acc += Gk[GSN_DGN_MAX][v];
                if (k == kmn[v]) acc += Gk[GSN_DGN_MIN][v];
                acc += cvar[v] * (2.0f * (xs[v] - m1[v]) / (float)D);
            }
            m[v] = acc;
        }
        float *mrow = M + (int64_t)__ldg(p.eid + k) * d + c;
#pragma unroll
        for (int v = 0; v < VEC; ++v) mrow[v] = m[v];
    }
    // ---- directional aggregators, one at a time
    for (int a = 0; a < p.n_aggr; ++a) {
        const int kind = p.aggr_kind[a], fi = p.aggr_idx[a];
        if (kind <= GSN_DGN_VAR) continue;
        float G[VEC];
        upstream(a, G);
        float n_abs = 0.f, n_pos = 0.f, n_neg = 0.f, mxs = -INFINITY;
        for (int k = k0; k < k1; ++k) {
            const float F = dgn_field(p, fi, i, __ldg(p.nbr + k), k);
            n_abs += fabsf(F); n_pos += fmaxf(F, 0.f); n_neg += fmaxf(-F, 0.f);
            mxs = fmaxf(mxs, p.aggr_alpha[a] * fabsf(F));
        }
        float se = 0.f;
        if (kind == GSN_DGN_DIR_SOFTMAX)
            for (int k = k0; k < k1; ++k)
                se += expf(p.aggr_alpha[a] * fabsf(dgn_field(p, fi, i, __ldg(p.nbr + k), k)) - mxs);
        auto weight = [&](float F) {
            if (kind == GSN_DGN_DIR_AV) return fabsf(F) / (n_abs + kDgnEps);
            if (kind == GSN_DGN_DIR_SOFTMAX) return expf(p.aggr_alpha[a] * fabsf(F) - mxs) / se;
            if (kind == GSN_DGN_DIR_DX_BALANCED)
                return (fmaxf(F, 0.f) / (n_pos + kDgnEps) + fmaxf(-F, 0.f) / (n_neg + kDgnEps)) / 2.0f;
            return F / (n_abs + kDgnEps);
        };
        float sgn[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) sgn[v] = 1.0f;
        float wsum = 0.f;
        const bool centred = kind >= GSN_DGN_DIR_DX;          // dx, dx-no-abs, dx-balanced subtract (sum w) h_i
        if (centred) {
            float u[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) u[v] = 0.f;
            for (int k = k0; k < k1; ++k) {
                const int j = __ldg(p.nbr + k);
                const float w = weight(dgn_field(p, fi, i, j, k));
                wsum += w;
                float xs[VEC];
                ld_vec<VEC>(p.h + (int64_t)j * d + c, xs);
#pragma unroll
                for (int v = 0; v < VEC; ++v) u[v] += xs[v] * w;
            }
            if (kind != GSN_DGN_DIR_DX_NO_ABS) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float uu = u[v] - wsum * hin[v];
                    sgn[v] = uu > 0.f ? 1.0f : (uu < 0.f ? -1.0f : 0.f);
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) sg[v] -= sgn[v] * G[v] * wsum;
        }
        for (int k = k0; k < k1; ++k) {
            const int j = __ldg(p.nbr + k);
            const float w = weight(dgn_field(p, fi, i, j, k));
            float *mrow = M + (int64_t)__ldg(p.eid + k) * d + c;
#pragma unroll
            for (int v = 0; v < VEC; ++v) mrow[v] += sgn[v] * G[v] * w;
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) SG[i * d + c + v] = sg[v];
}

}  // namespace gsn

using namespace gsn;

static int dgn_fill(DgnParams &p, const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N, int64_t E,
                    const float *d_h, int32_t d, const float *d_node_field, int32_t Fn, const float *d_edge_field, int32_t Fe,
                    const GsnDgnAggr *h_aggr, int32_t n_aggr, const int32_t *h_scalers, int32_t n_scalers, float avg_log) {
    if (N < 0 || E < 0 || d < 1 || !d_rowptr || !d_h || !h_aggr || n_aggr < 1 || n_aggr > GSN_DGN_MAX_AGGR ||
        n_scalers < 1 || n_scalers > GSN_DGN_MAX_SCALERS || !h_scalers || Fn < 0 || Fe < 0)
        return GSN_E_INVALID;
    if ((Fn > 0 && !d_node_field) || (Fe > 0 && (!d_edge_field || !d_eid)) || (E > 0 && !d_nbr)) return GSN_E_INVALID;
    p.rowptr = d_rowptr; p.eid = d_eid; p.nbr = d_nbr; p.N = N; p.h = d_h; p.node_field = d_node_field;
    p.edge_field = d_edge_field; p.d = d; p.Fn = Fn; p.Fe = Fe; p.n_aggr = n_aggr; p.n_scalers = n_scalers;
    p.avg_log = avg_log; p.out = nullptr;
    for (int a = 0; a < n_aggr; ++a) {
        const GsnDgnAggr &g = h_aggr[a];
        if (g.kind < GSN_DGN_MEAN || g.kind > GSN_DGN_DIR_DX_BALANCED) return GSN_E_INVALID;
        if (g.kind >= GSN_DGN_DIR_AV && (g.field < 0 || g.field >= Fn + Fe)) return GSN_E_INVALID;   // reference: IndexError
        p.aggr_kind[a] = g.kind; p.aggr_idx[a] = g.field; p.aggr_alpha[a] = g.alpha;
    }
    for (int s = 0; s < n_scalers; ++s) {
        if (h_scalers[s] < 0 || h_scalers[s] > 2) return GSN_E_INVALID;
        p.scaler_kind[s] = h_scalers[s];
    }
    return GSN_OK;
}

extern "C" int gsn_dgn_aggregate_bwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N,
                                     int64_t E, const float *d_h, int32_t d, const float *d_node_field, int32_t Fn,
                                     const float *d_edge_field, int32_t Fe, const GsnDgnAggr *h_aggr, int32_t n_aggr,
                                     const int32_t *h_scalers, int32_t n_scalers, float avg_log, const float *d_grad_out,
                                     float *d_M, float *d_SG, void *stream_) {
    DgnParams p;
    const int rc = dgn_fill(p, d_rowptr, d_eid, d_nbr, N, E, d_h, d, d_node_field, Fn, d_edge_field, Fe, h_aggr, n_aggr,
                            h_scalers, n_scalers, avg_log);
    if (rc) return rc;
    if (!d_grad_out || !d_SG || (E > 0 && (!d_M || !d_eid))) return GSN_E_INVALID;
    if (N == 0) return GSN_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d % 4 == 0) dgn_aggregate_bwd_kernel<4><<<(unsigned)ceil_div(N * (d / 4), 256), 256, 0, stream>>>(p, d_grad_out, d_M, d_SG);
    else dgn_aggregate_bwd_kernel<1><<<(unsigned)ceil_div(N * (int64_t)d, 256), 256, 0, stream>>>(p, d_grad_out, d_M, d_SG);
    GSN_BUMP(1);
    GSN_LAUNCH_OK("gsn_dgn_aggregate_bwd");
    return GSN_OK;
}

template <int VEC, bool STATS>
static void dgn_launch_group(const DgnParams &p, const DgnGroup &g, const float *W, cudaStream_t stream) {
    const int cpr = p.d / VEC;
    const int npb = cpr >= 256 ? 1 : 256 / cpr;
    const unsigned grid = (unsigned)ceil_div(p.N, (int64_t)npb);
    // one CTA = npb nodes x cpr chunks (rows wider than 256 chunks: d > 1024 is outside what the layer is used with)
    switch (g.n) {
        case 0: dgn_aggregate_kernel<VEC, 0, STATS><<<grid, 256, 0, stream>>>(p, g, W, cpr, npb); break;
        case 1: dgn_aggregate_kernel<VEC, 1, STATS><<<grid, 256, 0, stream>>>(p, g, W, cpr, npb); break;
        case 2: dgn_aggregate_kernel<VEC, 2, STATS><<<grid, 256, 0, stream>>>(p, g, W, cpr, npb); break;
        case 3: dgn_aggregate_kernel<VEC, 3, STATS><<<grid, 256, 0, stream>>>(p, g, W, cpr, npb); break;
        default: dgn_aggregate_kernel<VEC, 4, STATS><<<grid, 256, 0, stream>>>(p, g, W, cpr, npb); break;
    }
}

extern "C" int gsn_dgn_aggregate_fwd(const int32_t *d_rowptr, const int32_t *d_eid, const int32_t *d_nbr, int64_t N,
                                     int64_t E, const float *d_h, int32_t d, const float *d_node_field, int32_t Fn,
                                     const float *d_edge_field, int32_t Fe, const GsnDgnAggr *h_aggr, int32_t n_aggr,
                                     const int32_t *h_scalers, int32_t n_scalers, float avg_log, float *d_out,
                                     float *d_scratch, void *stream_) {
    DgnParams p;
    const int rc = dgn_fill(p, d_rowptr, d_eid, d_nbr, N, E, d_h, d, d_node_field, Fn, d_edge_field, Fe, h_aggr, n_aggr,
                            h_scalers, n_scalers, avg_log);
    if (rc) return rc;
    if (!d_out) return GSN_E_INVALID;
    p.out = d_out;
    if (N == 0) return GSN_OK;
    const int VEC = d % 4 == 0 ? 4 : 1;
    if (d / VEC > 256) return GSN_E_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    // statistics aggregators ride along with the first group of directional ones; groups of four share the scratch
    // (stream order: a group's weights are consumed before the next group overwrites them)
    bool any_stats = false;
    int dirs[GSN_DGN_MAX_AGGR], n_dir = 0;
    for (int a = 0; a < n_aggr; ++a) {
        if (h_aggr[a].kind <= GSN_DGN_VAR) any_stats = true;
        else dirs[n_dir++] = a;
    }
    if (n_dir > 0 && E > 0 && !d_scratch) return GSN_E_WORKSPACE;
    int done = 0;
    bool first = true;
    while (first || done < n_dir) {
        DgnGroup g;
        g.n = n_dir - done < 4 ? n_dir - done : 4;
        for (int q = 0; q < 4; ++q) {
            const int a = q < g.n ? dirs[done + q] : 0;
            g.kind[q] = q < g.n ? h_aggr[a].kind : 0; g.field[q] = q < g.n ? h_aggr[a].field : 0;
            g.alpha[q] = q < g.n ? h_aggr[a].alpha : 0.f; g.out_block[q] = a;
        }
        if (g.n > 0 && E > 0) {
            dgn_weights_kernel<<<(unsigned)ceil_div(N, (int64_t)256), 256, 0, stream>>>(p, g, d_scratch);
            GSN_BUMP(1);
        }
        const int stats = first && any_stats ? 1 : 0;
        if (stats || g.n > 0) {
            if (VEC == 4 && stats) dgn_launch_group<4, true>(p, g, d_scratch, stream);
            else if (VEC == 4) dgn_launch_group<4, false>(p, g, d_scratch, stream);
            else if (stats) dgn_launch_group<1, true>(p, g, d_scratch, stream);
            else dgn_launch_group<1, false>(p, g, d_scratch, stream);
            GSN_BUMP(1);
        }
        done += g.n;
        first = false;
    }
    GSN_LAUNCH_OK("gsn_dgn_aggregate_fwd");
    return GSN_OK;
}
