// ms_plan_test.cpp -- C wrapper around the host-side min-sum layout planner (ms_plan.h) for the CPU unit test
// tests/test_ms_plan.py (built with g++ by the test; not part of libqldpc_b200.so).
#include "ms_plan.h"

#include <cstring>

extern "C" int ms_plan_probe(int m, int n, int E, const int *row_ptr, const int *col_idx, int nl, const int *layer_ptr,
                             const int *layer_chk, int dc_inst, int dv_inst, int dmin, int search,
                             int *perm_out /*[n]*/, int *slot_edge_out /*[m*dc_inst]*/, int *lvar_ptr_out /*[nl+1]*/,
                             unsigned *lvar_out, int lvar_cap, long long *stats /*[6]: wavefronts, ideal, evals, c2v_words, lvar_len, baseline*/,
                             int packed)
{
    std::vector<int> cw(n, 0), col_ptr(n + 1, 0), row_idx(E), fill(n, 0);
    for (int x = 0; x < E; ++x) cw[col_idx[x]]++;
    for (int j = 0; j < n; ++j) col_ptr[j + 1] = col_ptr[j] + cw[j];
    for (int i = 0; i < m; ++i)
        for (int x = row_ptr[i]; x < row_ptr[i + 1]; ++x) row_idx[col_ptr[col_idx[x]] + fill[col_idx[x]]++] = i;
    qldpc::MsGraphView g{m, n, E, row_ptr, col_idx, col_ptr.data(), row_idx.data(), nl, layer_ptr, layer_chk};
    qldpc::MsPlanLayout base, L;
    qldpc::ms_plan_layout(g, dc_inst, dv_inst, dmin, false, base, 1, false, packed != 0);
    qldpc::ms_plan_layout(g, dc_inst, dv_inst, dmin, search != 0, L, 1, false, packed != 0);
    std::memcpy(perm_out, L.perm.data(), sizeof(int) * n);
    std::memcpy(slot_edge_out, L.slot_edge.data(), sizeof(int) * (size_t)m * dc_inst);
    std::memcpy(lvar_ptr_out, L.lvar_ptr.data(), sizeof(int) * (nl + 1));
    if ((int)L.lvar.size() > lvar_cap) return -1;
    std::memcpy(lvar_out, L.lvar.data(), sizeof(unsigned) * L.lvar.size());
    stats[0] = L.wavefronts; stats[1] = L.ideal; stats[2] = L.search_evals; stats[3] = L.c2v_words; stats[4] = (long long)L.lvar.size();
    stats[5] = base.wavefronts;
    return 0;
}

// eight-lane kernel layout (every layer one check): renumbering searched under the mod-8 model, cells dealt by sub8_deal
extern "C" int ms_plan_probe_sub8(int m, int n, int E, const int *row_ptr, const int *col_idx, int dcs, int dv_inst, int dmin, int search,
                                  int *perm_out /*[n]*/, int *cell_out /*[m*dcs]*/, long long *stats /*[4]: wavefronts, ideal, accepted moves, baseline*/)
{
    std::vector<int> cw(n, 0), col_ptr(n + 1, 0), row_idx(E), fill(n, 0), layer_ptr(m + 1), layer_chk(m);
    for (int x = 0; x < E; ++x) cw[col_idx[x]]++;
    for (int j = 0; j < n; ++j) col_ptr[j + 1] = col_ptr[j] + cw[j];
    for (int i = 0; i < m; ++i)
        for (int x = row_ptr[i]; x < row_ptr[i + 1]; ++x) row_idx[col_ptr[col_idx[x]] + fill[col_idx[x]]++] = i;
    for (int i = 0; i <= m; ++i) layer_ptr[i] = i;
    for (int i = 0; i < m; ++i) layer_chk[i] = i;
    qldpc::MsGraphView g{m, n, E, row_ptr, col_idx, col_ptr.data(), row_idx.data(), m, layer_ptr.data(), layer_chk.data()};
    qldpc::MsPlanLayout base, L;
    qldpc::ms_plan_layout(g, dcs, dv_inst, dmin, false, base, 1, true);
    qldpc::ms_plan_layout(g, dcs, dv_inst, dmin, search != 0, L, 1, true);
    std::memcpy(perm_out, L.perm.data(), sizeof(int) * n);
    std::memcpy(cell_out, L.sub_cell.data(), sizeof(int) * (size_t)m * dcs);
    stats[0] = L.wavefronts; stats[1] = L.ideal; stats[2] = L.search_evals; stats[3] = base.wavefronts;
    return 0;
}
