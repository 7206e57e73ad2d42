// TEST HARNESS ONLY (built by tests/, never by gsn_b200/): runs the
// __host__ __device__ enumeration cores of gsn_b200/csrc/count_core.cuh on the
// CPU for ONE graph so that the matching logic can be checked against the oracle
// in the GPU-less container.  The product path launches the same cores from
// CUDA kernels (count_kernels.cu); nothing in gsn_b200/ loads this file.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../gsn_b200/csrc/count_core.cuh"

static int g_parts = 1;
namespace {
struct HostAcc {
    int64_t *v;      // [n, ld]
    int64_t *s;      // [slots, ld]
    int ld, col0;
    void vertex(int lv, int col, uint32_t c) { v[(size_t)lv * ld + col0 + col] += c; }
    void slot(int sl, int col, uint32_t c) { s[(size_t)sl * ld + col0 + col] += c; }
    void overflow() {}
};

template <int W>
int run(int n, int64_t E, const int64_t *src, const int64_t *dst, const GsnPlan *P, int64_t *out, int ld) {
    std::vector<uint64_t> adj((size_t)n * W, 0);
    for (int64_t e = 0; e < E; ++e) {
        int a = (int)src[e], b = (int)dst[e];
        if (a == b) continue;
        adj[(size_t)a * W + (b >> 6)] |= 1ull << (b & 63);
        adj[(size_t)b * W + (a >> 6)] |= 1ull << (a & 63);
    }
    std::vector<int32_t> rowptr(n + 1, 0);
    for (int v = 0; v < n; ++v) {
        int d = 0;
        for (int w = 0; w < W; ++w) d += __builtin_popcountll(adj[(size_t)v * W + w]);
        rowptr[v + 1] = rowptr[v] + d;
    }
    int S = rowptr[n];
    gsn::GraphView<W> G{adj.data(), rowptr.data()};
    std::vector<int64_t> sacc((size_t)(S > 0 ? S : 1) * ld, 0);
    std::vector<int64_t> vacc((size_t)(n > 0 ? n : 1) * ld, 0);
    HostAcc acc{vacc.data(), sacc.data(), ld, P->col0};
    for (int a = 0; a < n; ++a) {
        gsn::VSet<W> row;
        row.load(G.row(a));
        while (!row.empty()) {
            int b = row.pop_lowest();
            for (int part = 0; part < g_parts; ++part) {
                if (P->family == GSN_FAMILY_CYCLES)
                    gsn::enumerate_cycles<W>(P->kmin, P->kmax, P->induced, P->scope, G, a, b, acc, part, g_parts);
                else if (P->family == GSN_FAMILY_CLIQUES)
                    gsn::enumerate_cliques<W>(P->kmin, P->kmax, P->scope, G, a, b, acc, part, g_parts);
                else
                    gsn::enumerate_generic<W>(*P, G, a, b, acc, part, g_parts);
            }
        }
    }
    if (P->scope == 0) {
        for (int v = 0; v < n; ++v)
            for (int c = 0; c < P->n_cols; ++c) out[(size_t)v * ld + P->col0 + c] = vacc[(size_t)v * ld + P->col0 + c];
        return 0;
    }
    // edge scope: edge_dict semantics (last column wins), utils_graph_processing.py:142-144
    std::vector<int64_t> slot_col(S > 0 ? S : 1, -1);
    for (int64_t e = 0; e < E; ++e) {
        int a = (int)src[e], b = (int)dst[e];
        if (a == b) continue;
        slot_col[G.slot(a, b)] = e;
    }
    int missing = 0;
    for (int64_t e = 0; e < E; ++e)
        for (int c = 0; c < P->n_cols; ++c) out[(size_t)e * ld + P->col0 + c] = 0;
    for (int s = 0; s < S; ++s)
        for (int c = 0; c < P->n_cols; ++c) {
            int64_t val = sacc[(size_t)s * ld + P->col0 + c];
            if (slot_col[s] < 0) { if (val) missing = 1; continue; }
            out[(size_t)slot_col[s] * ld + P->col0 + c] = val;
        }
    return missing ? -1 : 0;
}
}  // namespace

extern "C" void gsn_host_sim_set_parts(int parts) { g_parts = parts < 1 ? 1 : parts; }

extern "C" int gsn_host_sim_count(int n, int64_t E, const int64_t *src, const int64_t *dst, const GsnPlan *P,
                                  int64_t *out, int ld) {
    int W = (n + 63) / 64;
    if (W <= 1) return run<1>(n, E, src, dst, P, out, ld);
    if (W <= 2) return run<2>(n, E, src, dst, P, out, ld);
    if (W <= 4) return run<4>(n, E, src, dst, P, out, ld);
    if (W <= 8) return run<8>(n, E, src, dst, P, out, ld);
    return -2;
}
