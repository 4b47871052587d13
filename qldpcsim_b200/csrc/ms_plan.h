// ms_plan.h -- host-side layout planner of the min-sum kernel (no CUDA in this file; also compiled by the CPU unit test
// tests/test_ms_plan.py through csrc/ms_plan_test.cpp).
//
// Decides, for one (Tanner graph, layer list):
//   * the variable renumbering j -> j' (descending column weight; inside a weight class the order of 16-variable units is
//     searched so that every layer touches the two halves of the 32 shared-memory banks evenly),
//   * the region layout of the variable-major c2v array (see ms_kernel.cuh),
//   * which edge of a check goes to which (lane, step) cell of the check phase,
//   * the per-layer variable groups of the variable phase (two 32-variable sub-groups per trip),
// and evaluates the resulting number of shared-memory wavefronts per iteration (every access of a lane to S_j' or to a c2v
// word falls on bank j' mod 32 because the region bases are multiples of 32 words).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace qldpc {

struct MsGraphView {
    int m, n, E;
    const int *row_ptr, *col_idx;     // CSR, ascending variable per check
    const int *col_ptr, *row_idx;     // CSC, ascending check per variable
    int nl;
    const int *layer_ptr, *layer_chk;
};

struct MsPlanLayout {
    int dc_inst = 0, dv = 0, dv_inst = 0, dmin = 0;
    std::vector<int> order, perm;          // order[j'] = j, perm[j] = j'
    std::vector<int> edge_rank;            // [E] CSR edge -> rank of its check among the checks of its variable
    int cnt[16] = {0}, coff[16] = {0};     // variables of degree > x; word offset of region x
    int c2v_words = 0;
    std::vector<int> lpc;                  // [nl] lanes per check
    std::vector<int> slot_edge;            // [m][dc_inst] CSR edge in slot k of check i, -1 = padding
    std::vector<int> lvar_ptr;             // [nl+1] in packed entries (32 per trip)
    std::vector<uint32_t> lvar;            // packed (4*j'_a) | (4*j'_b << 16); dummy = 4*n
    int sub_total = 0, sub_min = 0;        // sub-groups listed (whole pair-trips) with / without the [low, low, any, any] quad pattern
    long long wavefronts = 0, ideal = 0;   // per full iteration over all layers (check phase: S load + c2v load + c2v store;
                                           // variable phase: S load + S store + dv_inst c2v loads per sub-group)
    int search_evals = 0;
    // eight-lane kernel (ms_sub_kernel.cuh) only: CSR edge of cell (trip, lane-in-group) of every check, -1 = none
    int sub_spl = 0;
    std::vector<int> sub_cell;             // [m][sub_spl * 8]
};

namespace msplan {

// lanes per check of a layer: the largest power of two (<= 8, dividing the row-weight class) that keeps the layer within one
// pass of the `team_warps` warps that share a shot
inline int lanes_per_check(int layer_checks, int dc_inst, int team_warps)
{
    const int lc = std::max(1, layer_checks);
    int lpc = 1;
    while (lpc < 8 && lpc * 2 * lc <= 32 * team_warps && lpc * 2 <= std::max(1, dc_inst)) lpc *= 2;
    return lpc;
}

inline int max_mult(const int *bank_cnt)
{
    int mx = 0;
    for (int b = 0; b < 32; ++b) mx = std::max(mx, bank_cnt[b]);
    return mx;
}

// Cells of the check phase: slot k of a check is handled by lane group h = k / SPL at step s = k % SPL.  All lanes of a
// pass (CPP = 32 / LPC checks) execute step s together, so the accesses of one step are {cell (h, s) of every check of the
// pass}.  Strategy 0 keeps the edges in ascending variable order.  Strategy 1 gives the cell of "virtual lane"
// v = h * CPP + (check's position in the pass) the edge whose j' has bit 4 equal to bit 4 of v, as far as the check has such
// edges: for 16-variable circulant blocks this places the two half-warps on different halves of the banks.
struct Evaluator {
    const MsGraphView &g;
    MsPlanLayout &L;
    std::vector<int> tmp_slot;     // [dc_inst] scratch

    Evaluator(const MsGraphView &g_, MsPlanLayout &L_) : g(g_), L(L_), tmp_slot(64) {}

    void assign_check(int i, int pos_in_pass, int lpc, int strategy, int *slots /*[dc_inst]*/) const
    {
        const int dc = L.dc_inst, spl = (dc + lpc - 1) / lpc, cpp = 32 / lpc;
        const int a = g.row_ptr[i], b = g.row_ptr[i + 1];
        for (int k = 0; k < dc; ++k) slots[k] = -1;
        if (strategy == 0) {
            for (int x = a; x < b; ++x) slots[x - a] = x;
            return;
        }
        int pool[2][32], np[2] = {0, 0}, take[2] = {0, 0};
        for (int x = a; x < b; ++x) {
            const int par = ((L.coff[L.edge_rank[x]] + L.perm[g.col_idx[x]]) >> 4) & 1;      // bank half of the edge's c2v word
            pool[par][np[par]++] = x;
        }
        int left = b - a;
        for (int k = 0; k < dc && left > 0; ++k) {
            const int h = k / spl;
            const int want = ((h * cpp + pos_in_pass) >> 4) & 1;
            int par = want;
            if (take[par] >= np[par]) par ^= 1;
            // keep enough cells for the remaining edges: never skip a cell while edges remain
            slots[k] = pool[par][take[par]++];
            --left;
        }
    }

    // wavefronts of the check phase of layer l with the given strategy; optionally stores the slot assignment
    long long check_cost(int l, int strategy, bool store, long long *ideal) const
    {
        const int dc = L.dc_inst, lpc = L.lpc[l], spl = (dc + lpc - 1) / lpc, cpp = 32 / lpc;
        const int qb = g.layer_ptr[l], qe = g.layer_ptr[l + 1];
        long long cost = 0;
        std::vector<int> slots((size_t)cpp * dc);
        for (int q0 = qb; q0 < qe; q0 += cpp) {
            const int nq = std::min(cpp, qe - q0);
            for (int q = 0; q < nq; ++q) {
                const int i = g.layer_chk[q0 + q];
                assign_check(i, q, lpc, strategy, &slots[(size_t)q * dc]);
                if (store) for (int k = 0; k < dc; ++k) L.slot_edge[(size_t)i * dc + k] = slots[(size_t)q * dc + k];
            }
            for (int s = 0; s < spl; ++s) {
                // S_j' sits on bank j' mod 32 (plus a constant), the edge's c2v word on (coff[rank] + j') mod 32: the same bank when
                // the regions start on multiples of 32 words, a per-region rotation when they are packed
                int bank_s[32] = {0}, bank_c[32] = {0}, any = 0;
                for (int q = 0; q < nq; ++q)
                    for (int h = 0; h < lpc; ++h) {
                        const int k = h * spl + s;
                        if (k >= dc) continue;
                        const int e = slots[(size_t)q * dc + k];
                        if (e < 0) continue;
                        const int jp = L.perm[g.col_idx[e]];
                        bank_s[jp & 31]++;
                        bank_c[(L.coff[L.edge_rank[e]] + jp) & 31]++;
                        any = 1;
                    }
                cost += max_mult(bank_s) + 2ll * max_mult(bank_c);
                if (ideal) *ideal += 3ll * any;
            }
        }
        return cost;
    }

    // variable groups of layer l: the variables of each residue class (j' mod 32) are dealt to the sub-groups holding the
    // fewest of that class.  QUAD PATTERN [low, low, any, any]: variables whose degree does not exceed `dmin` (the number of
    // leading regions that hold every variable) only have terms in those regions, so a sub-group made of such variables alone
    // is summed with dmin loads and no degree guards.  When 0 < dmin < dv_inst the kernel treats the first two sub-groups of
    // EVERY quad trip (two consecutive pair-trips) that way -- statically, without a per-trip test -- so the plan must put
    // low-degree variables (or dummies) only there; the last pair-trip of an odd count stays generic.  On the lifted-product
    // codes (column weights 3 and 5, five low and three high variables per check) this costs no extra trips.
    long long var_cost(int l, bool store, long long *ideal) const
    {
        std::vector<int> vs;
        for (int q = g.layer_ptr[l]; q < g.layer_ptr[l + 1]; ++q) {
            const int i = g.layer_chk[q];
            for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) vs.push_back(L.perm[g.col_idx[x]]);
        }
        std::sort(vs.begin(), vs.end());
        vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
        // as few sub-groups as possible (an instruction stream per sub-group costs more than a two-way bank conflict); the list
        // is padded to whole pair-trips with an empty sub-group, which the kernel skips (single-variable trip)
        const int V = (int)vs.size();
        const bool pattern = L.dmin > 0 && L.dmin < L.dv_inst;
        const int low_from = pattern ? L.cnt[L.dmin] : 0;               // j' >= low_from: degree <= dmin (descending-degree numbering)
        std::vector<int> low, rest;
        for (int jp : vs) ((pattern && jp >= low_from) ? low : rest).push_back(jp);
        int G = (((V + 31) / 32) + 1) & ~1;                             // sub-groups, whole pair-trips
        auto low_only = [&](int g2, int G2) { return pattern && g2 < 4 * (G2 / 4) && (g2 & 3) < 2; };   // position in the layer's list
        if (pattern) {
            for (;; G += 2) {                                           // enough room for the high-degree variables outside the low-only slots
                const int nlow_groups = 2 * (G / 4);
                const int placed_low = std::min((int)low.size(), 32 * nlow_groups);
                if (V - placed_low <= 32 * (G - nlow_groups)) break;
            }
        }
        if (store) { const_cast<MsPlanLayout &>(L).sub_total += G; const_cast<MsPlanLayout &>(L).sub_min += (((V + 31) / 32) + 1) & ~1; }
        std::vector<std::vector<int>> sub(G);
        std::vector<int> res_cnt((size_t)G * 32, 0);
        // Sub-groups are filled one after the other.  Each takes one variable of every residue class that still has some
        // (largest classes first), which is conflict-free; only when the sub-groups behind it could not hold the rest does
        // it take second, third ... members of the largest classes.  Unavoidable conflicts thus end up concentrated in few
        // sub-groups instead of being spread over all of them.  Returns the variables it could not place.
        auto deal = [&](const std::vector<int> &vars, const std::vector<int> &groups) {
            std::vector<std::vector<int>> by_res(32);
            for (int jp : vars) by_res[jp & 31].push_back(jp);
            int remaining = (int)vars.size();
            const int gc = (int)groups.size();
            for (int gi = 0; gi < gc; ++gi) {
                const int need = std::min(32, std::max(0, remaining - 32 * (gc - gi - 1)));
                int taken = 0;
                for (int round = 0; round < 32 && (round == 0 || taken < need); ++round) {
                    int rs[32];
                    for (int r = 0; r < 32; ++r) rs[r] = r;
                    std::stable_sort(rs, rs + 32, [&](int a2, int b2) { return by_res[a2].size() > by_res[b2].size(); });
                    for (int ri = 0; ri < 32 && taken < 32; ++ri) {
                        const int r = rs[ri];
                        if (by_res[r].empty() || (round > 0 && taken >= need)) continue;
                        sub[groups[gi]].push_back(by_res[r].back());
                        by_res[r].pop_back();
                        res_cnt[(size_t)groups[gi] * 32 + r]++;
                        ++taken;
                    }
                }
                remaining -= taken;
            }
            std::vector<int> left;
            for (int r = 0; r < 32; ++r) for (int jp : by_res[r]) left.push_back(jp);
            return left;
        };
        {
            std::vector<int> g_low, g_any;
            for (int g2 = 0; g2 < G; ++g2) (low_only(g2, G) ? g_low : g_any).push_back(g2);
            std::vector<int> left = deal(low, g_low);
            rest.insert(rest.end(), left.begin(), left.end());
            std::sort(rest.begin(), rest.end());
            // the generic sub-groups are filled front to back; trailing ones may stay empty (skipped as a single-variable trip)
            const int need_any = ((int)rest.size() + 31) / 32;
            g_any.resize(std::max(need_any, 0));
            deal(rest, g_any);
        }
        long long cost = 0;
        for (int g2 = 0; g2 < G; ++g2) {
            if (sub[g2].empty() && !low_only(g2, G)) continue;
            const int per = 2 + (low_only(g2, G) ? L.dmin : L.dv_inst);
            int mx = 0;
            for (int r = 0; r < 32; ++r) mx = std::max(mx, res_cnt[(size_t)g2 * 32 + r]);
            cost += (long long)per * std::max(mx, 1);
            if (ideal) *ideal += per;
        }
        if (store) {
            for (int g2 = 0; g2 < G; g2 += 2) {
                std::sort(sub[g2].begin(), sub[g2].end());
                std::sort(sub[g2 + 1].begin(), sub[g2 + 1].end());
                for (int ln = 0; ln < 32; ++ln) {
                    const uint32_t ja = ln < (int)sub[g2].size() ? (uint32_t)sub[g2][ln] : (uint32_t)g.n;
                    const uint32_t jb = ln < (int)sub[g2 + 1].size() ? (uint32_t)sub[g2 + 1][ln] : (uint32_t)g.n;
                    L.lvar.push_back((4u * ja) | ((4u * jb) << 16));
                }
            }
            L.lvar_ptr[l + 1] = (int)L.lvar.size();
        }
        return cost;
    }

    long long total(bool store)
    {
        long long cost = 0, ideal = 0;
        if (store) {
            L.slot_edge.assign((size_t)g.m * L.dc_inst, -1);
            L.lvar.clear();
            L.lvar_ptr.assign(g.nl + 1, 0);
            L.sub_total = L.sub_min = 0;
            // checks outside every layer keep the ascending order (never executed, but the table stays well defined)
            for (int i = 0; i < g.m; ++i)
                for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) L.slot_edge[(size_t)i * L.dc_inst + (x - g.row_ptr[i])] = x;
        }
        // a check listed in several layers keeps the assignment of its LAST layer: process layers in reverse when storing so
        // that the first layer wins
        for (int l = g.nl - 1; l >= 0; --l) {
            long long id0 = 0;
            const long long c0 = check_cost(l, 0, false, &id0), c1 = check_cost(l, 1, false, nullptr);
            const int strat = c1 < c0 ? 1 : 0;
            cost += std::min(c0, c1);
            ideal += id0;
            if (store) check_cost(l, strat, true, nullptr);
        }
        for (int l = 0; l < g.nl; ++l) cost += var_cost(l, store, &ideal);
        if (store) { L.wavefronts = cost; L.ideal = ideal; }
        return cost;
    }
};

}  // namespace msplan

// ---------------------------------------------------------------------------------------------------------------------------
// Eight-lane kernel (ms_sub_kernel.cuh): the four shots of a warp are interleaved word by word, so inside a group of 8 lanes an
// access to variable j' falls on bank 4 (j' mod 8) + group.  Lane h of a group handles cell (s, h) of the layer's check in step s
// of the check phase AND the same edge's variable in trip s of the variable phase, so one assignment of the check's edges to
// spl x 8 cells decides the conflicts of both phases: a step is conflict-free when its 8 variables are distinct mod 8, which a
// check allows when no residue class holds more than spl of its variables.
namespace msplan {

// modelled cost of check i: 64 * (wavefronts of one access kind = max(spl, largest residue class)) + excess members (guides the
// search towards feasibility)
inline int sub8_check_cost(const MsGraphView &g, const std::vector<int> &perm, int i, int spl)
{
    int cnt[8] = {0}, mx = 0, over = 0;
    for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) cnt[perm[g.col_idx[x]] & 7]++;
    for (int r = 0; r < 8; ++r) { mx = std::max(mx, cnt[r]); over += std::max(0, cnt[r] - spl); }
    return 64 * std::max(spl, mx) + over;
}

// Renumbering search: swaps of two variables of the same degree class whose j' differ mod 8 (simulated annealing with a fixed
// seed, bounded, stops at the conflict-free bound).  Only the checks that are layers of their own matter, i.e. all of them.
inline int sub8_search(const MsGraphView &g, MsPlanLayout &L, int spl, const std::vector<int> &deg, int max_moves)
{
    const int n = g.n, m = g.m;
    if (n < 16) return 0;
    long long cur = 0;
    for (int i = 0; i < m; ++i) cur += sub8_check_cost(g, L.perm, i, spl);
    const long long bound = 64ll * spl * m;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
    double T = 20.0;
    int accepted = 0;
    std::vector<int> touched;
    std::vector<char> mark(m, 0);
    for (int mv = 0; mv < max_moves && cur > bound; ++mv) {
        const int a = (int)(next() % (uint32_t)n), b = (int)(next() % (uint32_t)n);      // positions j'
        T = std::max(0.5, T * 0.99995);
        if (((a ^ b) & 7) == 0) continue;
        const int ja = L.order[a], jb = L.order[b];
        if (deg[ja] != deg[jb]) continue;
        touched.clear();
        for (int j : {ja, jb})
            for (int x = g.col_ptr[j]; x < g.col_ptr[j + 1]; ++x)
                if (!mark[g.row_idx[x]]) { mark[g.row_idx[x]] = 1; touched.push_back(g.row_idx[x]); }
        long long before = 0, after = 0;
        for (int i : touched) before += sub8_check_cost(g, L.perm, i, spl);
        std::swap(L.order[a], L.order[b]);
        L.perm[ja] = b; L.perm[jb] = a;
        for (int i : touched) { after += sub8_check_cost(g, L.perm, i, spl); mark[i] = 0; }
        const long long d = after - before;
        bool keep = d <= 0;
        if (!keep) {      // exp(-d / T) against a uniform number, in integers: accept with probability 2^(-d / (T ln 2))
            const double pr = std::exp(-(double)d / T);
            keep = (double)(next() & 0xFFFFFF) < pr * 16777216.0;
        }
        if (keep) { cur += d; ++accepted; }
        else { std::swap(L.order[a], L.order[b]); L.perm[ja] = a; L.perm[jb] = b; }
    }
    return accepted;
}

// Deals the edges of every check to spl trips of 8 cells: each trip takes one variable of every residue class that still has some
// (largest classes first); only what the later trips could not hold is added as second, third ... members.  Fills L.sub_cell,
// L.slot_edge (slot h * spl + s of the check table = cell (s, h)) and the modelled wavefronts (four access kinds per step in the
// check phase -- S load, c2v load, c2v store -- and 1 + dv_inst in the variable phase, as in Evaluator).
inline void sub8_deal(const MsGraphView &g, MsPlanLayout &L, int spl)
{
    const int m = g.m, dcs = spl * 8;
    L.sub_spl = spl;
    L.sub_cell.assign((size_t)m * dcs, -1);
    L.slot_edge.assign((size_t)m * dcs, -1);
    long long wf = 0;
    for (int i = 0; i < m; ++i) {
        std::vector<std::vector<int>> by_res(8);
        for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) by_res[L.perm[g.col_idx[x]] & 7].push_back(x);
        int *cell = &L.sub_cell[(size_t)i * dcs];
        int left = g.row_ptr[i + 1] - g.row_ptr[i];
        for (int tr = 0; tr < spl; ++tr) {
            const int need = std::min(8, std::max(0, left - 8 * (spl - tr - 1)));   // what the later trips cannot hold
            int taken = 0, mult[8] = {0};
            for (int round = 0; round < 8 && left > 0 && (round == 0 || taken < need); ++round) {
                int rs[8];
                for (int r = 0; r < 8; ++r) rs[r] = r;
                std::stable_sort(rs, rs + 8, [&](int a2, int b2) { return by_res[a2].size() > by_res[b2].size(); });
                for (int ri = 0; ri < 8 && taken < 8; ++ri) {
                    const int r = rs[ri];
                    if (by_res[r].empty() || (round > 0 && taken >= need)) continue;
                    cell[tr * 8 + taken] = by_res[r].back();
                    by_res[r].pop_back();
                    mult[r]++;
                    ++taken;
                }
            }
            left -= taken;
            int mx = 1;
            for (int r = 0; r < 8; ++r) mx = std::max(mx, mult[r]);
            wf += (long long)mx * (3 + 1 + L.dv_inst);
        }
        for (int tr = 0; tr < spl; ++tr)
            for (int h = 0; h < 8; ++h) L.slot_edge[(size_t)i * dcs + h * spl + tr] = cell[tr * 8 + h];
    }
    L.wavefronts = wf;
    L.ideal = (long long)m * spl * (3 + 1 + L.dv_inst);
}

}  // namespace msplan

// dc_inst / dv_inst / dmin are the shape of the kernel instance (see qldpc_api.cu: ms_select); `search` enables the unit
// order search (bounded number of evaluations).
// sub8 = true: layout for the eight-lane kernel (every layer one check; dc_inst a multiple of 8): the renumbering is searched under
// the mod-8 model above and the cells are dealt by sub8_deal; the regions of the c2v array then only need to start on multiples
// of 8 words.
// packed = true: the regions of the c2v array follow each other without rounding to 32 words (the planner's bank model accounts
// for the rotation); saves up to 31 words per region -- on LP118_0 (n = 544 = 17 * 32, so the dummy variable's word costs a
// whole row in each of the three unguarded regions) 496 bytes per shot, which is the 21st resident shot per SM.
inline void ms_plan_layout(const MsGraphView &g, int dc_inst, int dv_inst, int dmin, bool search, MsPlanLayout &L, int team_warps = 1, bool sub8 = false,
                           bool packed = false)
{
    const int n = g.n;
    L.dc_inst = dc_inst; L.dv_inst = dv_inst; L.dmin = dmin;
    std::vector<int> deg(n);
    int dv = 0;
    for (int j = 0; j < n; ++j) { deg[j] = g.col_ptr[j + 1] - g.col_ptr[j]; dv = std::max(dv, deg[j]); }
    L.dv = dv;
    L.order.resize(n);
    for (int j = 0; j < n; ++j) L.order[j] = j;
    std::stable_sort(L.order.begin(), L.order.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    L.perm.resize(n);
    auto set_perm = [&]() { for (int jp = 0; jp < n; ++jp) L.perm[L.order[jp]] = jp; };
    set_perm();
    // regions: multiples of 32 words so that bank(c2v word of j') == bank(S_j') == j' mod 32; an unguarded region (x < dmin)
    // has room for the always-zero word of the dummy variable n.  (Eight-lane kernel: multiples of 8 words, no dummy variable.)
    int words = 0;
    for (int x = 0; x < 16; ++x) {
        L.cnt[x] = 0;
        if (x < dv) for (int j = 0; j < n; ++j) L.cnt[x] += deg[j] > x;
        L.coff[x] = words;
        if (x < dv_inst) words += sub8 ? ((L.cnt[x] + 7) & ~7) : (packed ? ((L.cnt[x] + (x < dmin ? 1 : 0) + 15) & ~15) : ((L.cnt[x] + (x < dmin ? 1 : 0) + 31) & ~31));
    }
    L.c2v_words = (words + 3) & ~3;
    L.edge_rank.assign(g.E, 0);
    {
        std::vector<int> fill(n, 0);
        for (int i = 0; i < g.m; ++i)
            for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) L.edge_rank[x] = fill[g.col_idx[x]]++;
    }
    L.lpc.resize(g.nl);
    for (int l = 0; l < g.nl; ++l) L.lpc[l] = msplan::lanes_per_check(g.layer_ptr[l + 1] - g.layer_ptr[l], dc_inst, team_warps);

    if (sub8) {
        if (search) L.search_evals = msplan::sub8_search(g, L, dc_inst / 8, deg, 300000);
        msplan::sub8_deal(g, L, dc_inst / 8);
        L.lvar_ptr.assign(g.nl + 1, 0);
        return;
    }
    msplan::Evaluator ev(g, L);
    long long best = ev.total(false);
    L.search_evals = 1;
    if (search && n >= 64) {
        // units = runs of 16 consecutive j' that lie inside one degree class and start on a multiple of 16
        std::vector<int> unit_start, unit_class;
        for (int s = 0; s + 16 <= n; s += 16)
            if (deg[L.order[s]] == deg[L.order[s + 15]]) { unit_start.push_back(s); unit_class.push_back(deg[L.order[s]]); }
        const int U = (int)unit_start.size();
        long long ideal = 0;
        {   // lower bound: every access group conflict-free
            MsPlanLayout dummy;
            (void)dummy;
            for (int l = 0; l < g.nl; ++l) { long long id = 0; ev.check_cost(l, 0, false, &id); ideal += id; ev.var_cost(l, false, &ideal); }
        }
        const int max_evals = 400;
        bool improved = true;
        for (int sweep = 0; sweep < 4 && improved && best > ideal && U <= 48; ++sweep) {
            improved = false;
            for (int u = 0; u < U && best > ideal; ++u)
                for (int v = u + 1; v < U && best > ideal; ++v) {
                    if (unit_class[u] != unit_class[v]) continue;
                    if (((unit_start[u] ^ unit_start[v]) & 16) == 0) continue;     // same bank half: the swap changes nothing mod 32
                    if (L.search_evals >= max_evals) goto done;
                    std::swap_ranges(L.order.begin() + unit_start[u], L.order.begin() + unit_start[u] + 16, L.order.begin() + unit_start[v]);
                    set_perm();
                    const long long c = ev.total(false);
                    ++L.search_evals;
                    if (c < best) { best = c; improved = true; }
                    else {
                        std::swap_ranges(L.order.begin() + unit_start[u], L.order.begin() + unit_start[u] + 16, L.order.begin() + unit_start[v]);
                        set_perm();
                    }
                }
        }
    }
done:
    set_perm();
    ev.total(true);
}

}  // namespace qldpc
