// qldpc_api.cu -- C-ABI of libqldpc_b200.so (see include/qldpc_b200.h): plan builder and launchers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "bp_kernel.cuh"
#include "common.cuh"
#include "hard_kernels.cuh"
#include "ms_kernel.cuh"
#include "ms_plan.h"
#include "ms_sub_kernel.cuh"
#include "osd_kernel.cuh"
#include "sampler_kernel.cuh"

using namespace qldpc;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(QLDPC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    CU_TRY(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) CU_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int ensure_scratch(qldpc_plan *p, int slot, size_t bytes)
{
    if (p->scratch_bytes[slot] >= bytes) return 0;
    if (p->scratch[slot]) cudaFree(p->scratch[slot]);
    p->scratch[slot] = nullptr;
    p->scratch_bytes[slot] = 0;
    CU_TRY(cudaMalloc(&p->scratch[slot], bytes));
    p->scratch_bytes[slot] = bytes;
    return 0;
}

GraphDev graph_dev(const qldpc_plan *p)
{
    GraphDev g;
    g.m = p->tab.m; g.n = p->tab.n; g.mw = p->tab.mw; g.nw = p->tab.nw;
    g.row_ptr = p->d_row_ptr; g.col_idx = p->d_col_idx; g.col_ptr = p->d_col_ptr; g.row_idx = p->d_row_idx;
    return g;
}

// ---- min-sum kernel dispatch on (row weight, regular rows, column weight, unguarded regions)
typedef void (*ms_kernel_t)(MsTables, const uint16_t *, MsConst, DecodeIO);

// DMIN is either 0 (every region guarded by the degree test) or the fast value of the shape: 3 for column weights <= 5
// (the lifted-product / Tanner codes have column weights 3..5), DV for the others (column-regular codes such as bicycle).
constexpr int ms_fast_dmin(int dv_inst) { return dv_inst <= 5 ? 3 : (dv_inst <= 9 ? dv_inst : 0); }

template <int DC, int MAXW, int W, bool SPEC>
ms_kernel_t ms_pick_dv(int dv_inst, bool fast)
{
    switch (dv_inst) {
    case 4: return fast ? (ms_kernel_t)ms_decode_kernel<DC, 4, ms_fast_dmin(4), MAXW, W, SPEC> : (ms_kernel_t)ms_decode_kernel<DC, 4, 0, MAXW, W, SPEC>;
    case 5: return fast ? (ms_kernel_t)ms_decode_kernel<DC, 5, ms_fast_dmin(5), MAXW, W, SPEC> : (ms_kernel_t)ms_decode_kernel<DC, 5, 0, MAXW, W, SPEC>;
    case 9: return fast ? (ms_kernel_t)ms_decode_kernel<DC, 9, ms_fast_dmin(9), MAXW, W, SPEC> : (ms_kernel_t)ms_decode_kernel<DC, 9, 0, MAXW, W, SPEC>;
    case 16: return (ms_kernel_t)ms_decode_kernel<DC, 16, 0, MAXW, W, SPEC>;
    }
    return nullptr;
}

// merged steps (SPEC instances, see ms_kernel.cuh) exist for the row-weight classes 4, 8 and 16
inline bool ms_spec_available(int dc_inst) { return dc_inst == 4 || dc_inst == 8 || dc_inst == 16; }

// Instantiated shapes: row weight <= 4 / 8 / 16 / 24 / 32 (multiples of every lane split; shorter rows get padding edges),
// column weight <= 4 / 5 / 9 / 16.  `full_regions` = number of leading regions that hold every variable; *dmin receives the
// DMIN of the chosen kernel.
// variant: 0 = 24 warps per CTA, 1 = 32 warps per CTA (row-weight classes 4 and 8), 2 = teams of two warps per shot (class 8);
// spec: the plan holds merged steps (variants 0 and 2 only)
ms_kernel_t ms_select(int dc, int dv, int full_regions, int variant, bool spec, int *dc_inst, int *dv_inst, int *dmin)
{
    static const int dcs[] = {4, 8, 16, 24, 32}, dvs[] = {4, 5, 9, 16};
    int pc = 0, pv = 0;
    for (int s : dcs) if (!pc && s >= dc) pc = s;
    for (int s : dvs) if (!pv && s >= dv) pv = s;
    *dc_inst = pc; *dv_inst = pv; *dmin = 0;
    if (!pc || !pv) return nullptr;
    const int fd = ms_fast_dmin(pv);
    const bool fast = fd > 0 && full_regions >= fd;
    *dmin = fast ? fd : 0;
    if (spec) {
        switch (pc) {
        case 4: return ms_pick_dv<4, kMsWarps, 1, true>(pv, fast);
        case 8: return variant == 2 ? ms_pick_dv<8, kMsWarps, 2, true>(pv, fast) : ms_pick_dv<8, kMsWarps, 1, true>(pv, fast);
        case 16: return ms_pick_dv<16, kMsWarps, 1, true>(pv, fast);
        }
        return nullptr;
    }
    switch (pc) {
    case 4: return variant == 1 ? ms_pick_dv<4, kMsWarpsBig, 1, false>(pv, fast) : ms_pick_dv<4, kMsWarps, 1, false>(pv, fast);
    case 8: return variant == 1 ? ms_pick_dv<8, kMsWarpsBig, 1, false>(pv, fast) : (variant == 2 ? ms_pick_dv<8, kMsWarps, 2, false>(pv, fast) : ms_pick_dv<8, kMsWarps, 1, false>(pv, fast));
    case 16: return ms_pick_dv<16, kMsWarps, 1, false>(pv, fast);
    case 24: return ms_pick_dv<24, kMsWarps, 1, false>(pv, fast);
    case 32: return ms_pick_dv<32, kMsWarps, 1, false>(pv, fast);
    }
    return nullptr;
}

typedef void (*bp_kernel_t)(Tables, const uint16_t *, BpConst, DecodeIO);

struct Geometry {
    int grid, threads;
    size_t smem;
    int shots_per_cta;
};

}  // namespace

// Extra per-plan state that needs the kernel types.
typedef void (*ms_sub_kernel_t)(MsTables, MsSubTables, const uint16_t *, MsConst, DecodeIO);
struct PlanKernels {
    ms_sub_kernel_t ms_sub = nullptr;   // eight-lanes-per-shot kernel (every layer a single check, nothing to merge), or null
    MsSubTables sub_tab{};
    ms_kernel_t ms = nullptr;
    MsTables ms_tab{};
    int ms_full_regions = 0;
    int ms_team = 1;           // warps per shot of the warp-per-shot min-sum kernel
    bool ms_spec = false;      // the plan holds merged steps (SPEC kernel instance)
    bp_kernel_t bp = nullptr;
    const void *bf = nullptr;  // bit-flipping kernel of the plan (dense or sparse formulation)
    const void *ng = nullptr;  // naive-greedy kernel of the plan (tables in shared memory or through L1)
    int bp_team = 1;           // warps per shot of the sum-product kernel
    bool regular = false;
};
static PlanKernels *kernels_of(qldpc_plan *p) { return reinterpret_cast<PlanKernels *>(p->scratch[7]); }

extern "C" {

int qldpc_abi_version(void) { return QLDPC_ABI_VERSION; }
const char *qldpc_last_error(void) { return g_err.c_str(); }
int qldpc_words(int32_t nbits) { return (nbits + 31) / 32; }
int64_t qldpc_launch_count(void) { return g_launches.load(); }

int qldpc_plan_create(const qldpc_graph *g, const qldpc_opts *o, int device, qldpc_plan **out)
{
    if (!g || !o || !out) return fail(QLDPC_EINVAL, "null argument");
    *out = nullptr;
    if (g->m <= 0 || g->n <= 0 || g->nnz < 0 || !g->row_ptr || (g->nnz && !g->col_idx)) return fail(QLDPC_EINVAL, "bad graph");
    if (o->dec_type < QLDPC_NG || o->dec_type > QLDPC_BP) return fail(QLDPC_EINVAL, "Unrecognized decoder type.");
    if (o->max_iter < 0) return fail(QLDPC_EINVAL, "max_iter < 0");
    const int m = g->m, n = g->n, E = g->nnz;
    if (g->row_ptr[0] != 0 || g->row_ptr[m] != E) return fail(QLDPC_EINVAL, "row_ptr does not span nnz");
    const bool iterative = (o->dec_type == QLDPC_MS || o->dec_type == QLDPC_BP);
    if (iterative && (g->n_layers <= 0 || !g->layer_ptr || !g->layer_chk))
        return fail(QLDPC_EINVAL, "MS/BP need a layer list (the reference's layers=None default is unusable, decoders.py:144)");

    qldpc_plan *p = new (std::nothrow) qldpc_plan();
    if (!p) return fail(QLDPC_ENOMEM, "out of host memory");
    p->device = device;
    p->opts = *o;
    p->row_ptr.assign(g->row_ptr, g->row_ptr + m + 1);
    p->col_idx.assign(g->col_idx, g->col_idx + E);
    // every early return below (validation, QLDPC_ETOOBIG, any failing CUDA call) releases the half-built plan
    struct PlanGuard { qldpc_plan *p; ~PlanGuard() { if (p) qldpc_plan_destroy(p); } } guard{p};
    auto bail = [&](int code, const std::string &msg) { return fail(code, msg); };

    // ---- validate CSR, build CSC (ascending check per variable because rows are visited in order)
    int dc = 0;
    std::vector<int> cw(n, 0);
    for (int i = 0; i < m; ++i) {
        const int a = p->row_ptr[i], b = p->row_ptr[i + 1];
        if (b < a) return bail(QLDPC_EINVAL, "row_ptr not monotone");
        dc = std::max(dc, b - a);
        for (int x = a; x < b; ++x) {
            const int j = p->col_idx[x];
            if (j < 0 || j >= n || (x > a && j <= p->col_idx[x - 1])) return bail(QLDPC_EINVAL, "col_idx must be ascending within a row and < n");
            cw[j]++;
        }
    }
    p->col_ptr.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) p->col_ptr[j + 1] = p->col_ptr[j] + cw[j];
    const int dv = n ? *std::max_element(cw.begin(), cw.end()) : 0;
    p->row_idx.assign(E, 0);
    std::vector<int> col_slot(E, 0);   // slot (position in its row) of each CSC entry
    {
        std::vector<int> fill(n, 0);
        for (int i = 0; i < m; ++i)
            for (int x = p->row_ptr[i]; x < p->row_ptr[i + 1]; ++x) {
                const int j = p->col_idx[x];
                const int t = p->col_ptr[j] + fill[j]++;
                p->row_idx[t] = i;
                col_slot[t] = x - p->row_ptr[i];
            }
    }
    bool regular = true;
    for (int i = 0; i < m; ++i) regular = regular && (p->row_ptr[i + 1] - p->row_ptr[i] == dc);

    // ---- layers
    int nl = 0;
    if (iterative) {
        nl = g->n_layers;
        p->layer_ptr.assign(g->layer_ptr, g->layer_ptr + nl + 1);
        if (p->layer_ptr[0] != 0) return bail(QLDPC_EINVAL, "layer_ptr[0] != 0");
        for (int l = 0; l < nl; ++l) if (p->layer_ptr[l + 1] < p->layer_ptr[l]) return bail(QLDPC_EINVAL, "layer_ptr not monotone");
        p->layer_chk.assign(g->layer_chk, g->layer_chk + p->layer_ptr[nl]);
        for (int c : p->layer_chk) if (c < 0 || c >= m) return bail(QLDPC_EINVAL, "layer check index out of range (the reference raises IndexError)");
    } else {
        p->layer_ptr.assign(1, 0);
    }

    Tables &t = p->tab;
    t.m = m; t.n = n; t.E = E; t.dc = dc; t.dv = dv; t.nl = nl;
    p->row_w = dc;
    t.mw = (m + 31) / 32; t.nw = (n + 31) / 32;

    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;

    // ---- device copies of CSR / CSC and bit-packed rows
    int rc = 0;
    if ((rc = upload(&p->d_row_ptr, p->row_ptr)) || (rc = upload(&p->d_col_idx, p->col_idx)) ||
        (rc = upload(&p->d_col_ptr, p->col_ptr)) || (rc = upload(&p->d_row_idx, p->row_idx))) return rc;
    {
        std::vector<uint32_t> hb((size_t)(m + 1) * t.nw, 0u);     // row m = OR of all rows (column mask)
        for (int i = 0; i < m; ++i)
            for (int x = p->row_ptr[i]; x < p->row_ptr[i + 1]; ++x) {
                const int j = p->col_idx[x];
                hb[(size_t)i * t.nw + (j >> 5)] |= 1u << (j & 31);
                hb[(size_t)m * t.nw + (j >> 5)] |= 1u << (j & 31);
            }
        if ((rc = upload(&p->d_hbits, hb))) return rc;
        const int cwd = kColStride;
        std::vector<uint32_t> hc((size_t)n * cwd, 0u);
        if (t.mw <= 32)
            for (int i = 0; i < m; ++i)
                for (int x = p->row_ptr[i]; x < p->row_ptr[i + 1]; ++x) hc[(size_t)p->col_idx[x] * cwd + (i >> 5)] |= 1u << (i & 31);
        if ((rc = upload(&p->d_hcol, hc))) return rc;
        // GF(2) rank of H (gf2math.py:91-135) by bit-packed elimination on the host; OSD stops its column walk there
        std::vector<uint32_t> w(hb.begin(), hb.begin() + (size_t)m * t.nw);
        int r = 0;
        for (int col = 0; col < n && r < m; ++col) {
            const int cw = col >> 5; const uint32_t cb = 1u << (col & 31);
            int piv = -1;
            for (int i = r; i < m; ++i) if (w[(size_t)i * t.nw + cw] & cb) { piv = i; break; }
            if (piv < 0) continue;
            if (piv != r) for (int x = 0; x < t.nw; ++x) std::swap(w[(size_t)r * t.nw + x], w[(size_t)piv * t.nw + x]);
            for (int i = r + 1; i < m; ++i)
                if (w[(size_t)i * t.nw + cw] & cb) for (int x = 0; x < t.nw; ++x) w[(size_t)i * t.nw + x] ^= w[(size_t)r * t.nw + x];
            ++r;
        }
        p->rank_h = r;
    }
    CU_TRY(cudaMalloc((void **)&p->d_work, 4 * sizeof(unsigned long long)));
    CU_TRY(cudaMalloc((void **)&p->d_fail_count, 4 * sizeof(int)));
    CU_TRY(cudaMalloc((void **)&p->d_work_done, sizeof(unsigned long long)));
    CU_TRY(cudaMemset(p->d_work_done, 0, sizeof(unsigned long long)));
    for (int s = 0; s < 3; ++s) CU_TRY(cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking));
    for (int e = 0; e < 12; ++e) CU_TRY(cudaEventCreateWithFlags(&p->events[e], cudaEventDisableTiming));
    PlanKernels *pk = new PlanKernels();
    pk->regular = regular;
    p->scratch[7] = pk;   // host object, slot 7 is never cudaFree'd (scratch_bytes[7] stays 0)

    // ---- launch geometry
    if (iterative) {
        if ((long long)dc * m > 65535 || n >= 65535 || p->layer_ptr[nl] > 65535 || dc > 32)
            return bail(QLDPC_ETOOBIG, "code too large for the on-chip decoder tables (need m*row_weight <= 65535, n < 65535, row weight <= 32)");
        const bool is_ms = o->dec_type == QLDPC_MS;
        std::vector<uint16_t> &b = p->h_blob;
        auto put = [&](int count) { int off = (int)b.size(); b.resize(b.size() + count, 0); return off; };
        auto put32 = [&](int count) { if (b.size() & 1) b.push_back(0); int off = (int)b.size(); b.resize(b.size() + 2 * (size_t)count, 0); return off; };
        auto set32 = [&](int off, int idx, uint32_t v) { b[off + 2 * idx] = (uint16_t)(v & 0xffffu); b[off + 2 * idx + 1] = (uint16_t)(v >> 16); };
        // sorted distinct variables adjacent to the checks of layer l
        auto layer_vars = [&](int l) {
            std::vector<int> vs;
            for (int q = p->layer_ptr[l]; q < p->layer_ptr[l + 1]; ++q) {
                const int i = p->layer_chk[q];
                for (int x = p->row_ptr[i]; x < p->row_ptr[i + 1]; ++x) vs.push_back(p->col_idx[x]);
            }
            std::sort(vs.begin(), vs.end());
            vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
            return vs;
        };
        size_t state = 0;
        const void *fn = nullptr;
        if (is_ms) {
          // Two passes at most: the tables are built with the regions of the c2v array on multiples of 32 words and the shot states
          // 128 bytes apart; if the packed variant (16 words / 16 bytes) would hold more shots per SM, they are built again that way.
          static const int packed_env = [] { const char *ev = getenv("QLDPC_MS_PACKED"); return ev ? atoi(ev) : -1; }();   // tuning knob: 0 / 1 force
          bool packed = packed_env == 1;
          for (int pass = 0; pass < 2; ++pass) {
            b.clear();
            // ================= min-sum tables (layout described in ms_kernel.cuh, planned by ms_plan.h) =================
            MsTables &mt = pk->ms_tab;
            mt = MsTables{};
            pk->ms_sub = nullptr;
            if (dv > kMsMaxDv) return bail(QLDPC_ETOOBIG, "min-sum kernels are instantiated for row weight <= 32 and column weight <= 16");
            int full_regions = 0;      // leading regions that hold every variable: x < min column weight
            {
                int dmin_true = n ? *std::min_element(cw.begin(), cw.end()) : 0;
                full_regions = std::min(dmin_true, dv);
            }
            pk->ms_full_regions = full_regions;
            int dc_inst = 0, dv_inst = 0, dmin = 0;
            pk->ms = ms_select(dc, dv, full_regions, 0, false, &dc_inst, &dv_inst, &dmin);
            if (!pk->ms) return bail(QLDPC_ETOOBIG, "min-sum kernels are instantiated for row weight <= 32 and column weight <= 16");
            // ---- merged steps: runs of consecutive layers with pairwise disjoint variable sets (the single-check layers of the
            // serial schedule inside a circulant block row) become one step of the kernel, which commits them sub-layer by
            // sub-layer only when the convergence test could fire (ms_kernel.cuh).  Layer 0 stays alone (its first pass uses the
            // binary32-rounded prior); a run holds at most 32 layers and `max_vars` variables (one quad trip per warp).
            auto group_layers = [&](int max_vars, std::vector<int> &grp) {
                grp.assign(1, 0);
                std::vector<int> mark(n, -1);
                int gid = 0, cur_vars = 0, cur_sub = 0;
                for (int l = 0; l < nl; ++l) {
                    const std::vector<int> vs = layer_vars(l);
                    bool ok = l > 1 && cur_sub < 32 && cur_vars + (int)vs.size() <= max_vars;
                    for (size_t x = 0; ok && x < vs.size(); ++x) ok = mark[vs[x]] != gid;
                    if (!ok && l > 0) { grp.push_back(l); ++gid; cur_vars = 0; cur_sub = 0; }
                    for (int v : vs) mark[v] = gid;
                    cur_vars += (int)vs.size();
                    ++cur_sub;
                }
                grp.push_back(nl);
            };
            auto max_group_checks = [&](const std::vector<int> &grp) {
                int mx = 0;
                for (size_t g2 = 0; g2 + 1 < grp.size(); ++g2) mx = std::max(mx, p->layer_ptr[grp[g2 + 1]] - p->layer_ptr[grp[g2]]);
                return mx;
            };
            static const bool merge_env = [] { const char *ev = getenv("QLDPC_MS_MERGE"); return !ev || atoi(ev) != 0; }();   // tuning knob
            const bool can_merge = merge_env && ms_spec_available(dc_inst) && !(o->reserved & 1);
            std::vector<int> grp;                       // step -> first layer
            // Teams of two warps per shot: when the shot state allows only few shots per SM (LP118_2, Tanner: 10), one warp per
            // shot leaves the schedulers idle.  Decided on the state size (tables are ~10-35 KB), row-weight class 8 only.
            int W = 1;
            {
                MsPlanLayout probe_pl;
                ms_plan_layout(MsGraphView{m, n, E, p->row_ptr.data(), p->col_idx.data(), p->col_ptr.data(), p->row_idx.data(), 0, p->layer_ptr.data(), p->layer_chk.data()},
                               dc_inst, dv_inst, dmin, /*search=*/false, probe_pl);
                MsTables probe{};
                probe.n = n; probe.mw = t.mw; probe.c2v_words = probe_pl.c2v_words;
                const size_t st = ms_layout(probe).bytes;
                if (can_merge) group_layers(256, grp);
                else { grp.resize(nl + 1); for (int l = 0; l <= nl; ++l) grp[l] = l; }
                W = (dc_inst == 8 && st * 13 > (size_t)kMaxSmemPerCta - 24 * 1024 && max_group_checks(grp) >= 16) ? 2 : 1;
                if (const char *ev = getenv("QLDPC_MS_TEAM")) { const int w2 = atoi(ev); if (w2 == 1 || (w2 == 2 && dc_inst == 8)) W = w2; }   // tuning knob
                if (W == 1 && can_merge) group_layers(128, grp);
            }
            pk->ms_team = W;
            // merging pays when it removes most of the steps (serial schedules: 450 -> 16); a layered schedule whose cross-wired
            // partition happens to hold a few disjoint neighbours keeps its layers (and the 32-warp instances)
            if ((int)grp.size() - 1 < nl && 2 * ((int)grp.size() - 1) > nl) { grp.resize(nl + 1); for (int l = 0; l <= nl; ++l) grp[l] = l; }
            const int nsteps = (int)grp.size() - 1;
            const bool spec = nsteps < nl;
            std::vector<int> step_ptr(nsteps + 1);       // step -> range in layer_chk
            for (int g2 = 0; g2 <= nsteps; ++g2) step_ptr[g2] = p->layer_ptr[grp[g2]];
            pk->ms_spec = spec;
            // ---- eight lanes per shot (ms_sub_kernel.cuh): every layer is one check and nothing could be merged (bicycle)
            int sub_mw = 0;      // > 0: the eight-lane instance that toggles parities through per-variable masks of sub_mw words
            bool use_sub = !spec && nl >= 1 && !(o->reserved & 2) && dc <= 32;
            for (int l = 0; use_sub && l < nl; ++l) use_sub = p->layer_ptr[l + 1] - p->layer_ptr[l] == 1;
            static const bool sub_env = [] { const char *ev = getenv("QLDPC_MS_SUB"); return !ev || atoi(ev) != 0; }();   // tuning knob
            use_sub = use_sub && sub_env;
            if (use_sub) {      // its records hold byte offsets of the four-shot interleaved layout in 16 bits
                const bool fast9 = ((dc + 7) & ~7) == 24 && dv_inst == 9 && full_regions >= 9 && t.mw == 3;
                const long long words = (long long)(fast9 ? 9 : 16) * ((n + 8) & ~7) + n + 3;
                use_sub = 16 * words <= 65535;
            }
            if (use_sub) {
                const int dcs = std::max(8, (dc + 7) & ~7);
                dc_inst = dcs;
                W = 1; pk->ms_team = 1;
                if (dcs == 24 && dv_inst == 9 && full_regions >= 9 && t.mw == 3) { pk->ms_sub = ms_sub_kernel<24, 9, 9, 3>; dmin = 9; sub_mw = 3; }
                else {
                    dv_inst = 16; dmin = 0; full_regions = 0; pk->ms_full_regions = 0;
                    pk->ms_sub = dcs == 8 ? ms_sub_kernel<8, 16, 0, 0> : (dcs == 16 ? ms_sub_kernel<16, 16, 0, 0> : (dcs == 24 ? ms_sub_kernel<24, 16, 0, 0> : ms_sub_kernel<32, 16, 0, 0>));
                }
            }
            if (spec || W == 2) {
                int a1, a2, a3;
                pk->ms = ms_select(dc, dv, full_regions, W == 2 ? 2 : 0, spec, &a1, &a2, &a3);
            }
            MsGraphView gv{m, n, E, p->row_ptr.data(), p->col_idx.data(), p->col_ptr.data(), p->row_idx.data(), nsteps, step_ptr.data(), p->layer_chk.data()};
            MsPlanLayout pl;
            const bool pk_regions = packed && !use_sub;
            ms_plan_layout(gv, dc_inst, dv_inst, dmin, /*search=*/true, pl, W, /*sub8=*/use_sub, pk_regions);
            if (!use_sub && dmin > 0 && dmin < dv_inst && 100ll * pl.sub_total > 115ll * pl.sub_min) {
                // the [low, low, any, any] quad pattern of the partially guarded instances needs too many padding sub-groups on this
                // graph (few variables of degree <= dmin): use the instance that guards every region instead
                full_regions = 0;
                pk->ms_full_regions = 0;
                int a1, a2;
                pk->ms = ms_select(dc, dv, 0, W == 2 ? 2 : 0, spec, &a1, &a2, &dmin);
                pl = MsPlanLayout();
                ms_plan_layout(gv, dc_inst, dv_inst, dmin, /*search=*/true, pl, W, false, pk_regions);
            }
            p->plan_wavefronts = pl.wavefronts; p->plan_wavefronts_ideal = pl.ideal;
            mt.m = m; mt.n = n; mt.E = E; mt.dc = dc_inst; mt.dv = dv; mt.nl = nsteps; mt.mw = t.mw; mt.nw = t.nw;
            mt.n_pad = (n + 63) & ~63;
            for (int x = 0; x < kMsMaxDv; ++x) { mt.cnt4[x] = 4 * pl.cnt[x]; mt.coff4[x] = 4 * pl.coff[x]; }
            mt.c2v_words = pl.c2v_words;
            if (4ll * mt.c2v_words + 4ll * (n + 3) > 65535)
                return bail(QLDPC_ETOOBIG, "code too large for the 16-bit shared-memory offset tables (need 4*edges < 65536, 4*n < 65536)");
            // slot stride: with LPC lanes per check, lane h starts at slot h*SPL, i.e. SPL*ms words further; ms = 4 (mod 8)
            // puts the lane groups of a split check on disjoint banks
            int ms = m;
            while (ms % 8 != 4) ++ms;
            mt.ms = ms;
            // padding edge: S entry n+1 (+inf) and the scratch word S[n+2], addressed relative to the c2v array like every c2v word
            const uint32_t pad_edge = (uint32_t)(4 * (n + 1)) | ((uint32_t)(4 * mt.c2v_words + 4 * (n + 2)) << 16);
            auto chk_entry = [&](int e) {
                if (e < 0) return pad_edge;
                const int jp = pl.perm[p->col_idx[e]];
                return (uint32_t)(4 * jp) | ((uint32_t)(mt.coff4[pl.edge_rank[e]] + 4 * jp) << 16);
            };
            if (use_sub) {
                // the eight-lane kernel reads one record per (layer, lane) instead of the check table, the step records and the
                // variable lists (ms_sub_kernel.cuh)
                mt.off_chk = put32(0);
                const int spl = dc_inst / 8;
                const uint32_t pad16 = (uint32_t)(16 * (n + 1)) | ((uint32_t)(16 * mt.c2v_words + 16 * (n + 2)) << 16);
                pk->sub_tab.off_srec = put32(nl * 8 * spl);
                for (int x = 0; x < kMsMaxDv; ++x) pk->sub_tab.c16[x] = 4 * mt.coff4[x];
                for (int l = 0; l < nl; ++l) {
                    const int i = p->layer_chk[p->layer_ptr[l]];
                    for (int h2 = 0; h2 < 8; ++h2)
                        for (int s2 = 0; s2 < spl; ++s2) {
                            const int e = pl.sub_cell[(size_t)i * dc_inst + s2 * 8 + h2];
                            uint32_t v = pad16;
                            if (e >= 0) {
                                const int jp = pl.perm[p->col_idx[e]];
                                v = (uint32_t)(16 * jp) | ((uint32_t)(4 * (mt.coff4[pl.edge_rank[e]] + 4 * jp)) << 16);
                            }
                            set32(pk->sub_tab.off_srec, (l * 8 + h2) * spl + s2, v);
                        }
                }
                pl.lvar.clear();
            } else {
                mt.off_chk = put32(dc_inst * ms);
                for (int x = 0; x < dc_inst * ms; ++x) set32(mt.off_chk, x, pad_edge);
                for (int i = 0; i < m; ++i)
                    for (int k = 0; k < dc_inst; ++k) set32(mt.off_chk, k * ms + i, chk_entry(pl.slot_edge[(size_t)i * dc_inst + k]));
            }
            if (pl.lvar.size() > 65535) return bail(QLDPC_ETOOBIG, "per-layer variable lists exceed 65535 entries");
            b.resize((b.size() + 7) & ~size_t(7), 0);              // 16-byte aligned records
            mt.off_layer = put(use_sub ? 0 : 8 * nsteps);
            for (int l = 0; l < nsteps && !use_sub; ++l) {
                uint16_t *r = &b[mt.off_layer + 8 * l];
                r[0] = (uint16_t)step_ptr[l]; r[1] = (uint16_t)step_ptr[l + 1]; r[2] = (uint16_t)pl.lpc[l];
                r[3] = (uint16_t)pl.lvar_ptr[l]; r[4] = (uint16_t)pl.lvar_ptr[l + 1];
                bool single = pl.lvar_ptr[l + 1] > pl.lvar_ptr[l];
                for (int x = pl.lvar_ptr[l + 1] - 32; single && x < pl.lvar_ptr[l + 1]; ++x) single = ((pl.lvar[x] >> 16) & 0xfffcu) == 4u * (uint32_t)n;
                r[5] = single ? 1 : 0;
                r[6] = (uint16_t)(grp[l + 1] - grp[l] > 1 ? grp[l + 1] - grp[l] : 0);
                { int ed = 0; for (int q = step_ptr[l]; q < step_ptr[l + 1]; ++q) ed += p->row_ptr[p->layer_chk[q] + 1] - p->row_ptr[p->layer_chk[q]]; r[7] = (uint16_t)std::min(ed, 65535); }
                if (r[6] && pl.lvar_ptr[l + 1] - pl.lvar_ptr[l] > 64 * W) return bail(QLDPC_EINVAL, "internal: merged step exceeds one quad trip per warp");
            }
            mt.off_layer_chk = put((int)p->layer_chk.size());
            for (size_t x = 0; x < p->layer_chk.size(); ++x) b[mt.off_layer_chk + x] = (uint16_t)p->layer_chk[x];
            mt.off_lvar = put32((int)pl.lvar.size());
            for (size_t x = 0; x < pl.lvar.size(); ++x) set32(mt.off_lvar, (int)x, pl.lvar[x]);
            // sub-layer (within its merged step) of every listed variable; a variable belongs to one layer of a run
            mt.off_lsub = put(spec ? (int)pl.lvar.size() : 0);
            if (spec) {
                std::vector<int> vsub(n + 1, 0);
                for (int l = 0; l < nsteps; ++l) {
                    for (int ll = grp[l]; ll < grp[l + 1]; ++ll)
                        for (int q = p->layer_ptr[ll]; q < p->layer_ptr[ll + 1]; ++q)
                            for (int x = p->row_ptr[p->layer_chk[q]]; x < p->row_ptr[p->layer_chk[q] + 1]; ++x) vsub[p->col_idx[x]] = ll - grp[l];
                    for (int x = pl.lvar_ptr[l]; x < pl.lvar_ptr[l + 1]; ++x) {
                        const int ja = (int)(pl.lvar[x] & 0xffffu) / 4, jb = (int)(pl.lvar[x] >> 16) / 4;
                        const int sa = ja < n ? vsub[pl.order[ja]] : 0, sb2 = jb < n ? vsub[pl.order[jb]] : 0;
                        b[mt.off_lsub + x] = (uint16_t)(sa | (sb2 << 8));
                    }
                }
            }
            // checks of every variable in the renumbering, fixed stride (flip handling)
            if (sub_mw > 0) {
                mt.off_col_chk = put(0);
                pk->sub_tab.off_colmask = put32(n * sub_mw);
                for (int jp = 0; jp < n; ++jp) {
                    const int j = pl.order[jp];
                    std::vector<uint32_t> mk(sub_mw, 0u);
                    for (int x = p->col_ptr[j]; x < p->col_ptr[j + 1]; ++x) mk[p->row_idx[x] >> 5] ^= 1u << (p->row_idx[x] & 31);
                    for (int w2 = 0; w2 < sub_mw; ++w2) set32(pk->sub_tab.off_colmask, jp * sub_mw + w2, mk[w2]);
                }
            } else {
                mt.off_col_chk = put((n + 1) * dv_inst);
                std::fill(b.begin() + mt.off_col_chk, b.begin() + mt.off_col_chk + (n + 1) * dv_inst, (uint16_t)0xFFFF);
                for (int jp = 0; jp < n; ++jp) {
                    const int j = pl.order[jp];
                    for (int x = p->col_ptr[j]; x < p->col_ptr[j + 1]; ++x) b[mt.off_col_chk + jp * dv_inst + (x - p->col_ptr[j])] = (uint16_t)p->row_idx[x];
                }
            }
            mt.off_rowpar = put32(t.mw);
            {
                std::vector<uint32_t> rp(t.mw, 0u);
                for (int i = 0; i < m; ++i) if ((p->row_ptr[i + 1] - p->row_ptr[i]) & 1) rp[i >> 5] |= 1u << (i & 31);
                for (int w2 = 0; w2 < t.mw; ++w2) set32(mt.off_rowpar, w2, rp[w2]);
            }
            mt.off_unperm = put(32 * t.nw);
            for (int j = 0; j < 32 * t.nw; ++j) b[mt.off_unperm + j] = (uint16_t)(4 * (j < n ? pl.perm[j] : n));
            b.resize((b.size() + 7) & ~size_t(7), 0);
            mt.len = (int)b.size();
            t.len = mt.len;
            mt.packed = pk_regions ? 1 : 0;
            mt.team = W;
            state = pk_regions ? ms_layout(mt).bytes16 : ms_layout(mt).bytes;
            fn = (const void *)pk->ms;
            if (use_sub) { state = 4 * (size_t)ms_layout(mt).bytes16; fn = (const void *)pk->ms_sub; }      // a warp holds four interleaved shots
            if (pass == 1 || packed || use_sub || packed_env == 0) break;
            {   // would the packed variant hold more shots?  (same tables, hence the same blob size)
                MsPlanLayout probe_pl;
                ms_plan_layout(gv, dc_inst, dv_inst, dmin, /*search=*/false, probe_pl, W, false, true);
                MsTables probe = mt;
                probe.c2v_words = probe_pl.c2v_words;
                const size_t st_packed = ms_layout(probe).bytes16, bb = (size_t)ms_table_bytes(mt);
                if (bb + state > (size_t)kMaxSmemPerCta) break;
                const size_t fit_now = ((size_t)kMaxSmemPerCta - bb) / state, fit_packed = ((size_t)kMaxSmemPerCta - bb) / st_packed;
                const size_t cap = (size_t)(W > 1 ? std::min(kMsWarps / W, 15) : kMsWarpsBig);
                if (std::min(fit_packed, cap) <= std::min(fit_now, cap)) break;
                packed = true;
            }
          }
        } else {
            // ================= sum-product tables (slot-major edge layout, see common.cuh) =================
            // Slot stride of the binary64 c2v array.  lane = (check of the pass, slot): a stride of 4 (mod 16) puts the slots of the
            // consecutive checks of a pass on distinct 8-byte bank pairs, a stride that is a multiple of 16 (m = 240) puts all 8
            // slots on the same one (measured: 53 % of the shared-memory wavefronts of the kernel were bank conflicts).  The padding
            // costs shared memory, and resident warps are what this kernel lives on, so it is only used when it does not lower the
            // number of resident warps the team selection below arrives at.
            int max_layer = 0;
            for (int l = 0; l < nl; ++l) max_layer = std::max(max_layer, p->layer_ptr[l + 1] - p->layer_ptr[l]);
            const int lpc_bp = dc <= 4 ? 4 : (dc <= 8 ? 8 : (dc <= 16 ? 16 : 32));
            const int passes = (max_layer + 32 / lpc_bp - 1) / (32 / lpc_bp);
            auto build_bp = [&](int ms) -> int {          // fills b / t for slot stride ms; returns 0 or an error code
                b.clear();
                t.ms = ms;
                if ((long long)dc * ms > 65535) return QLDPC_ETOOBIG;
                t.off_var = put(dc * ms);
                std::fill(b.begin() + t.off_var, b.begin() + t.off_var + dc * ms, kPad);
                for (int i = 0; i < m; ++i)
                    for (int x = p->row_ptr[i]; x < p->row_ptr[i + 1]; ++x) b[t.off_var + (x - p->row_ptr[i]) * ms + i] = (uint16_t)(4 * p->col_idx[x]);
                t.off_col_ptr = put(n + 1);
                for (int j = 0; j <= n; ++j) b[t.off_col_ptr + j] = (uint16_t)p->col_ptr[j];
                t.off_col_pos = put(E);
                t.off_col_chk = put(E);
                for (int x = 0; x < E; ++x) {
                    b[t.off_col_pos + x] = (uint16_t)(col_slot[x] * ms + p->row_idx[x]);
                    b[t.off_col_chk + x] = (uint16_t)p->row_idx[x];
                }
                t.off_colf = put(dv < 8 ? (n + 1) * dv : 0);
                if (dv < 8) {
                    std::fill(b.begin() + t.off_colf, b.begin() + t.off_colf + (n + 1) * dv, (uint16_t)(dc * ms));
                    for (int j = 0; j < n; ++j)
                        for (int x = p->col_ptr[j]; x < p->col_ptr[j + 1]; ++x) b[t.off_colf + j * dv + (x - p->col_ptr[j])] = (uint16_t)(col_slot[x] * ms + p->row_idx[x]);
                }
                t.n_pad = (n + 31) & ~31;
                t.off_rowpar = put(2 * t.mw);
                for (int i = 0; i < m; ++i)
                    if ((p->row_ptr[i + 1] - p->row_ptr[i]) & 1) b[t.off_rowpar + 2 * (i >> 5) + ((i & 31) >> 4)] |= (uint16_t)(1u << (i & 15));
                t.off_layer_ptr = put(nl + 1);
                for (int l = 0; l <= nl; ++l) b[t.off_layer_ptr + l] = (uint16_t)p->layer_ptr[l];
                t.off_layer_chk = put((int)p->layer_chk.size());
                for (size_t x = 0; x < p->layer_chk.size(); ++x) b[t.off_layer_chk + x] = (uint16_t)p->layer_chk[x];
                std::vector<int> lvar_ptr(nl + 1, 0);
                std::vector<uint16_t> lvar;
                for (int l = 0; l < nl; ++l) {
                    for (int v : layer_vars(l)) lvar.push_back((uint16_t)v);
                    while (lvar.size() % 32) lvar.push_back((uint16_t)n);   // dummy variable n: uniform trip count per lane
                    lvar_ptr[l + 1] = (int)lvar.size();
                }
                if (lvar.size() > 65535) return QLDPC_ETOOBIG;
                t.off_lvar_ptr = put(nl + 1);
                for (int l = 0; l <= nl; ++l) b[t.off_lvar_ptr + l] = (uint16_t)lvar_ptr[l];
                t.off_lvar_idx = put((int)lvar.size());
                std::copy(lvar.begin(), lvar.end(), b.begin() + t.off_lvar_idx);
                b.resize((b.size() + 7) & ~size_t(7), 0);
                t.len = (int)b.size();
                return 0;
            };
            // warps per shot ("team") for the current b / t: as many resident warps as possible, at most 4 per team and at most 15
            // teams (named barriers 1..15); a team only pays off when a layer has work for all of its warps.  Measured on LP118_0
            // (11 shots fit): W = 1 / 2 / 3 / 4 -> 4.3 / 7.1 / 7.6 / 8.0e10 edge-iterations/s, i.e. the total number of resident
            // warps is what counts, not the number of resident shots
            auto pick_team = [&](int *W_out) -> int {     // returns the number of resident warps
                const size_t st = bp_layout(t).bytes;
                const size_t bb = (size_t)bp_table_bytes(t);
                const int teams_fit = (int)std::min<size_t>(32, bb + st <= (size_t)kMaxSmemPerCta ? ((size_t)kMaxSmemPerCta - bb) / st : 0);
                int W = 1, best_warps = std::min(teams_fit, 32);
                for (int w2 = 2; w2 <= 4 && w2 <= passes; ++w2) {
                    const int teams = std::min(teams_fit, std::min(32 / w2, 15));
                    if (teams * w2 > best_warps) { best_warps = teams * w2; W = w2; }
                }
                *W_out = W;
                return best_warps;
            };
            int W_plain = 1, W_pad = 1, ms_pad = m;
            while (ms_pad % 16 != 4) ++ms_pad;
            if (int e2 = build_bp(m)) return bail(e2, "code too large for the on-chip decoder tables (need m*row_weight <= 65535, n < 65535, row weight <= 32, per-layer variable lists <= 65535 entries)");
            const int warps_plain = pick_team(&W_plain);
            int W = W_plain;
            if (build_bp(ms_pad) == 0 && pick_team(&W_pad) >= warps_plain) W = W_pad;
            else { build_bp(m); W = W_plain; }
            state = bp_layout(t).bytes;
            // lanes per check = smallest power of two >= row weight (one lane per edge)
            {
                if (const char *ev = getenv("QLDPC_BP_TEAM")) { const int w2 = atoi(ev); if (w2 >= 1 && w2 <= 4) W = w2; }   // tuning knob
                pk->bp_team = W;
#define QLDPC_BP_PICK(L) (W == 4 ? (bp_kernel_t)bp_decode_kernel<L, 4> : (W == 3 ? (bp_kernel_t)bp_decode_kernel<L, 3> : (W == 2 ? (bp_kernel_t)bp_decode_kernel<L, 2> : (bp_kernel_t)bp_decode_kernel<L, 1>)))
                if (dc <= 4) pk->bp = QLDPC_BP_PICK(4);
                else if (dc <= 8) pk->bp = QLDPC_BP_PICK(8);
                else if (dc <= 16) pk->bp = QLDPC_BP_PICK(16);
                else pk->bp = QLDPC_BP_PICK(32);
#undef QLDPC_BP_PICK
            }
            fn = (const void *)pk->bp;
        }
        if ((rc = upload(&p->d_blob, b))) return rc;
        const size_t blob_bytes = is_ms ? (size_t)ms_table_bytes(pk->ms_tab) : (size_t)bp_table_bytes(t);
        if (!fn) return bail(QLDPC_ETOOBIG, "row weight not supported");
        if (blob_bytes + state > (size_t)kMaxSmemPerCta)
            return bail(QLDPC_ETOOBIG, "decoder state of one shot does not fit in 227 KB of shared memory");
        const int warps_fit = (int)std::min<size_t>(64, ((size_t)kMaxSmemPerCta - blob_bytes) / state);
        int warps = std::min(warps_fit, is_ms ? kMsWarps : 32);
        if (is_ms && pk->ms_team == 1 && !pk->ms_spec && !pk->ms_sub && warps_fit > kMsWarps && pk->ms_tab.dc <= 8) {           // small shot state: the 32-warp instance
            int a1, a2, a3;
            pk->ms = ms_select(dc, dv, pk->ms_full_regions, 1, false, &a1, &a2, &a3);
            fn = (const void *)pk->ms;
            warps = std::min(warps_fit, kMsWarpsBig);
        }
        const bool sub_k = is_ms && pk->ms_sub != nullptr;
        if (sub_k) warps = std::min(warps_fit, kMsSubWarps);
        if (const char *ev = getenv("QLDPC_SHOTS_CAP")) { const int cap = atoi(ev); if (cap > 0) warps = std::min(warps, cap); }   // measuring knob: resident shots per CTA
        const int team = is_ms ? pk->ms_team : pk->bp_team;
        if (is_ms && team > 1) warps = std::min(warps, std::min(kMsWarps / team, 15));          // shots per CTA (named barriers 1..15)
        if (!is_ms) warps = std::min(warps, std::min(32 / team, team > 1 ? 15 : 32));        // shots per CTA (named barriers 1..15)
        p->state_bytes = state;
        p->threads = warps * team * kWarp;
        p->shots_per_cta = sub_k ? 4 * warps : warps;
        p->smem_bytes = blob_bytes + (size_t)warps * state;
        p->grid = p->sm_count;
        CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));   // per function, shared by all plans
    } else {
        const int warps = 8;
        const bool bf_sparse = o->dec_type == QLDPC_BF && t.mw <= 32;
        size_t per = (o->dec_type == QLDPC_BF) ? (size_t)(t.nw + 2 * t.mw) * 4 : (size_t)(t.nw + t.mw + n) * 4;
        if (bf_sparse) per = (size_t)(t.nw + 2 * t.mw + 32 * t.nw + (m + 1) / 2) * 4;
        p->state_bytes = per;
        p->threads = warps * kWarp;
        p->shots_per_cta = warps;
        p->smem_bytes = per * warps + (bf_sparse ? (((size_t)2 * n + 15) & ~size_t(15)) : 0);
        if (p->smem_bytes > (size_t)kMaxSmemPerCta) return bail(QLDPC_ETOOBIG, "code too large");
        const bool ng_tab16 = o->dec_type == QLDPC_NG && E < 65536 && m < 65535 && n < 65535;
        const void *fn = ng_tab16 ? (const void *)ng_decode_kernel<true> : (const void *)ng_decode_kernel<false>;
        if (ng_tab16) p->smem_bytes = ((p->smem_bytes + 3) & ~size_t(3)) + (size_t)(m + 1 + n + 1 + 2 * E) * 2;
        if (p->smem_bytes > (size_t)kMaxSmemPerCta) return bail(QLDPC_ETOOBIG, "code too large");
        pk->ng = fn;
        if (o->dec_type == QLDPC_BF) {
            fn = (const void *)bf_decode_kernel;
            if (bf_sparse) switch (col_words(t.mw) / 4) {
                case 1: fn = (const void *)bf_sparse_kernel<1>; break;
                case 2: fn = (const void *)bf_sparse_kernel<2>; break;
                case 4: fn = (const void *)bf_sparse_kernel<4>; break;
                default: fn = (const void *)bf_sparse_kernel<8>; break;
            }
            pk->bf = fn;
        }
        CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta));
        int per_sm = 1;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, p->threads, p->smem_bytes));
        p->grid = p->sm_count * std::max(1, per_sm);
    }
    guard.p = nullptr;
    *out = p;
    return QLDPC_OK;
}

int qldpc_plan_destroy(qldpc_plan *p)
{
    if (!p) return QLDPC_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_blob); cudaFree(p->d_row_ptr); cudaFree(p->d_col_idx); cudaFree(p->d_col_ptr); cudaFree(p->d_row_idx);
    cudaFree(p->d_hbits); cudaFree(p->d_hcol); cudaFree(p->d_lcol); cudaFree(p->d_work); cudaFree(p->d_fail_count); cudaFree(p->d_work_done);
    for (int s = 0; s < 7; ++s) if (p->scratch[s]) cudaFree(p->scratch[s]);
    delete kernels_of(p);
    for (int s = 0; s < 3; ++s) if (p->streams[s]) cudaStreamDestroy(p->streams[s]);
    for (int e = 0; e < 12; ++e) if (p->events[e]) cudaEventDestroy(p->events[e]);
    for (int s = 0; s < 4; ++s) if (p->pinned[s]) cudaFreeHost(p->pinned[s]);
    delete p;
    return QLDPC_OK;
}

int qldpc_plan_set_logicals(qldpc_plan *p, const uint32_t *rows, int32_t k)
{
    if (!p || k < 0 || (k > 0 && !rows)) return fail(QLDPC_EINVAL, "null argument");
    if (k > 1024) return fail(QLDPC_ETOOBIG, "at most 1024 logical operators per plan");
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaDeviceSynchronize());          // a classification using the old basis may still be in flight
    cudaFree(p->d_lcol);
    p->d_lcol = nullptr; p->logical_k = 0; p->lkw = 0;
    if (k == 0) return QLDPC_OK;
    const int n = p->tab.n, nw = p->tab.nw, kw = (k + 31) / 32, cwd = kColStride;
    std::vector<uint32_t> cols((size_t)n * cwd, 0u);
    for (int r = 0; r < k; ++r)
        for (int j = 0; j < n; ++j)
            if ((rows[(size_t)r * nw + (j >> 5)] >> (j & 31)) & 1u) cols[(size_t)j * cwd + (r >> 5)] |= 1u << (r & 31);
    int rc = upload(&p->d_lcol, cols);
    if (rc) return rc;
    p->logical_k = k; p->lkw = kw;
    return QLDPC_OK;
}

int64_t qldpc_plan_work(qldpc_plan *p, int reset)
{
    if (!p) return -1;
    if (cudaSetDevice(p->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, p->d_work_done, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset && cudaMemset(p->d_work_done, 0, sizeof(v)) != cudaSuccess) return -1;
    return (int64_t)v;
}

int64_t qldpc_plan_info(const qldpc_plan *p, int what)
{
    if (!p) return -1;
    switch (what) {
    case 0: return p->tab.m;
    case 1: return p->tab.n;
    case 2: return p->tab.E;
    case 3: return p->tab.nl;
    case 4: return p->grid;
    case 5: return p->threads;
    case 6: return (int64_t)p->smem_bytes;
    case 7: return p->shots_per_cta;
    case 8: return p->row_w;
    case 9: return p->tab.dv;
    case 10: return p->rank_h;
    case 11: return 0;   // reserved
    case 12: return p->plan_wavefronts;
    case 13: return p->plan_wavefronts_ideal;
    case 14: return (int64_t)p->state_bytes;
    case 15: return p->logical_k;
    case 16: return p->opts.dec_type == QLDPC_MS ? reinterpret_cast<PlanKernels *>(p->scratch[7])->ms_tab.nl : p->tab.nl;   // steps per iteration
    case 17: return p->opts.dec_type == QLDPC_MS ? reinterpret_cast<PlanKernels *>(p->scratch[7])->ms_team : 1;
    }
    return -1;
}

// Launch one decode pass on `stream` using work-counter slot `slot` (0..3).
static int launch_decode(qldpc_plan *p, const uint32_t *syn, int64_t shots, uint32_t *ehat, int32_t *iters, uint8_t *conv,
                         double *llr, int *fail_count, int *fail_shot, double *fail_llr, int fail_cap, int slot, cudaStream_t st)
{
    if (shots == 0) return QLDPC_OK;
    DecodeIO io;
    io.syn = syn; io.ehat = ehat; io.iters = iters; io.conv = conv; io.llr = llr; io.shots = shots;
    io.work_counter = p->d_work + slot;
    io.fail_count = fail_count; io.fail_shot = fail_shot; io.fail_llr = fail_llr; io.fail_cap = fail_cap;
    io.work_done = p->d_work_done;
    CU_TRY(cudaMemsetAsync(p->d_work + slot, 0, sizeof(unsigned long long), st));
    const int grid = (int)std::min<int64_t>(p->grid, (shots + p->shots_per_cta - 1) / p->shots_per_cta);
    const qldpc_opts &o = p->opts;
    switch (o.dec_type) {
    case QLDPC_MS: {
        MsConst c;
        c.L = o.prior_llr; c.Lf = (double)(float)o.prior_llr; c.beta = o.beta; c.abeta = std::fabs(o.beta); c.sgn = std::signbit(o.beta) ? 0x80000000u : 0u; c.max_iter = o.max_iter;
        {   // smallest binary32 >= -L
            const double T = -o.prior_llr;
            float f = (float)T;
            if ((double)f < T) f = std::nextafterf(f, INFINITY);
            c.Tf = f;
        }
        if (kernels_of(p)->ms_sub) kernels_of(p)->ms_sub<<<grid, p->threads, p->smem_bytes, st>>>(kernels_of(p)->ms_tab, kernels_of(p)->sub_tab, p->d_blob, c, io);
        else kernels_of(p)->ms<<<grid, p->threads, p->smem_bytes, st>>>(kernels_of(p)->ms_tab, p->d_blob, c, io);
        break;
    }
    case QLDPC_BP: {
        BpConst c;
        c.L0 = o.prior_llr; c.eps = o.eps; c.max_iter = o.max_iter;
        kernels_of(p)->bp<<<grid, p->threads, p->smem_bytes, st>>>(p->tab, p->d_blob, c, io);
        break;
    }
    case QLDPC_BF:
        if (kernels_of(p)->bf == (const void *)bf_decode_kernel) {
            bf_decode_kernel<<<grid, p->threads, p->smem_bytes, st>>>(graph_dev(p), o.max_iter, io);
        } else {
            GraphDev gd = graph_dev(p);
            const uint32_t *hc = p->d_hcol;
            int mi = o.max_iter;
            void *args[] = {&gd, &hc, &mi, &io};
            CU_TRY(cudaLaunchKernel(kernels_of(p)->bf, dim3(grid), dim3(p->threads), args, p->smem_bytes, st));
        }
        break;
    case QLDPC_NG:
        {
            GraphDev gd = graph_dev(p);
            void *args[] = {&gd, &io};
            CU_TRY(cudaLaunchKernel(kernels_of(p)->ng, dim3(grid), dim3(p->threads), args, p->smem_bytes, st));
        }
        break;
    }
    g_launches++;
    CU_TRY(cudaGetLastError());
    return QLDPC_OK;
}

static int osd_on_failures(qldpc_plan *p, uint32_t *ehat, const uint32_t *syn, int64_t shots, cudaStream_t st);

int qldpc_decode(qldpc_plan *p, const uint32_t *syn, int64_t shots, uint32_t *ehat, int32_t *iters, uint8_t *conv,
                 double *llr, void *stream)
{
    if (!p || shots < 0 || (shots && (!syn || !ehat || !iters))) return fail(QLDPC_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool osd = p->opts.osd_order >= 0 && p->opts.dec_type == QLDPC_MS;   // the driver never passes OSDorder to BP (simulator.py:281-282); BP plans honour it too
    const bool osd_bp = p->opts.osd_order >= 0 && p->opts.dec_type == QLDPC_BP;
    if (!(osd || osd_bp)) return launch_decode(p, syn, shots, ehat, iters, conv, llr, nullptr, nullptr, nullptr, 0, 0, st);
    // With OSD the batch is processed in chunks so that the compacted LLR buffer of the unconverged shots
    // stays bounded: chunk * n * 8 bytes.
    // the compacted LLR buffer holds one row per shot of the chunk (a chunk in which every shot fails cannot overflow it);
    // up to 2 GiB of the 180 GB are spent on it so that a chunk stays many waves deep (LP118_2: 263k shots per chunk)
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(shots, (int64_t)(2048ll << 20) / ((int64_t)p->tab.n * 8)));
    int rc;
    if ((rc = ensure_scratch(p, 0, (size_t)chunk * sizeof(int)))) return rc;
    if ((rc = ensure_scratch(p, 1, (size_t)chunk * p->tab.n * sizeof(double)))) return rc;
    for (int64_t s0 = 0; s0 < shots; s0 += chunk) {
        const int64_t ns = std::min(chunk, shots - s0);
        CU_TRY(cudaMemsetAsync(p->d_fail_count, 0, sizeof(int), st));
        rc = launch_decode(p, syn + s0 * p->tab.mw, ns, ehat + s0 * p->tab.nw, iters + s0, conv ? conv + s0 : nullptr,
                           llr ? llr + s0 * p->tab.n : nullptr, p->d_fail_count, (int *)p->scratch[0], (double *)p->scratch[1], (int)chunk, 0, st);
        if (rc) return rc;
        if ((rc = osd_on_failures(p, ehat + s0 * p->tab.nw, syn + s0 * p->tab.mw, ns, st))) return rc;
    }
    return QLDPC_OK;
}

// OSD over the compacted failure list left by the last launch_decode (scratch[0] = shot ids, scratch[1] = LLRs).
static int osd_on_failures(qldpc_plan *p, uint32_t *ehat, const uint32_t *syn, int64_t shots, cudaStream_t st)
{
    OsdArgs a;
    a.m = p->tab.m; a.n = p->tab.n; a.mw = p->tab.mw; a.nw = p->tab.nw;
    a.hbits = p->d_hbits;
    a.row_ptr = p->d_row_ptr; a.col_idx = p->d_col_idx;
    a.ehat = ehat; a.syn = syn;
    a.llr = (const double *)p->scratch[1];
    a.perm = nullptr;
    a.shot_ids = (const int *)p->scratch[0];
    a.count_dev = p->d_fail_count;
    a.count = (int)shots;           // upper bound; the kernel reads the true count from count_dev
    a.order = p->opts.osd_order;
    a.rank_h = p->rank_h;
    int rc = osd_launch(a, p->sm_count, st);
    if (rc) return fail(QLDPC_ECUDA, std::string("osd launch: ") + cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    return QLDPC_OK;
}

int qldpc_osd(qldpc_plan *p, uint32_t *ehat, const uint32_t *syn, const double *llr, const int32_t *perm, int64_t shots,
              int32_t order, void *stream)
{
    if (!p || shots < 0 || (shots && (!ehat || !syn || !llr))) return fail(QLDPC_EINVAL, "null argument");
    if (order < 0) return QLDPC_OK;
    CU_TRY(cudaSetDevice(p->device));
    OsdArgs a;
    a.m = p->tab.m; a.n = p->tab.n; a.mw = p->tab.mw; a.nw = p->tab.nw;
    a.hbits = p->d_hbits;
    a.row_ptr = p->d_row_ptr; a.col_idx = p->d_col_idx;
    a.ehat = ehat; a.syn = syn; a.llr = llr; a.perm = perm;
    a.shot_ids = nullptr; a.count_dev = nullptr; a.count = (int)shots; a.order = order; a.rank_h = p->rank_h;
    if (shots == 0) return QLDPC_OK;
    int rc = osd_launch(a, p->sm_count, (cudaStream_t)stream);
    if (rc) return fail(QLDPC_ECUDA, std::string("osd launch: ") + cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    return QLDPC_OK;
}

int qldpc_decode_host(qldpc_plan *p, const uint32_t *syn, int64_t shots, uint32_t *ehat, int32_t *iters, uint8_t *conv,
                      double *llr)
{
    if (!p || shots < 0 || (shots && (!syn || !ehat || !iters))) return fail(QLDPC_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    const Tables &t = p->tab;
    // Two pipeline slots; slot s owns streams[s], work counter s, and one quarter-open set of device buffers.
    static const int64_t chunk_env = [] { const char *ev = getenv("QLDPC_HOST_CHUNK"); return ev ? atoll(ev) : 0ll; }();   // tuning knob
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(shots, chunk_env > 0 ? chunk_env : (llr ? (1 << 15) : (1 << 18))));
    const size_t b_syn = (size_t)chunk * t.mw * 4, b_e = (size_t)chunk * t.nw * 4, b_it = (size_t)chunk * 4, b_cv = (size_t)chunk,
                 b_llr = llr ? (size_t)chunk * t.n * 8 : 0;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t per_slot = al(b_syn) + al(b_e) + al(b_it) + al(b_cv) + al(b_llr);
    int rc;
    if ((rc = ensure_scratch(p, 2, per_slot * 2))) return rc;
    const bool osd = p->opts.osd_order >= 0 && (p->opts.dec_type == QLDPC_MS || p->opts.dec_type == QLDPC_BP);
    int64_t k = 0;
    for (int64_t s0 = 0; s0 < shots; s0 += chunk, ++k) {
        const int slot = (int)(k & 1);
        const int64_t ns = std::min(chunk, shots - s0);
        cudaStream_t st = p->streams[slot];
        unsigned char *base = (unsigned char *)p->scratch[2] + per_slot * slot;
        uint32_t *d_syn = (uint32_t *)base;
        uint32_t *d_e = (uint32_t *)(base + al(b_syn));
        int32_t *d_it = (int32_t *)(base + al(b_syn) + al(b_e));
        uint8_t *d_cv = (uint8_t *)(base + al(b_syn) + al(b_e) + al(b_it));
        double *d_llr = llr ? (double *)(base + al(b_syn) + al(b_e) + al(b_it) + al(b_cv)) : nullptr;
        // the slot's buffers are free once its previous chunk has been copied back (stream order guarantees it)
        CU_TRY(cudaMemcpyAsync(d_syn, syn + s0 * t.mw, (size_t)ns * t.mw * 4, cudaMemcpyHostToDevice, st));
        if (osd) {
            // OSD path shares scratch[0..1] and the failure counter: run those chunks through the synchronous API
            CU_TRY(cudaStreamSynchronize(p->streams[slot ^ 1]));
            if ((rc = qldpc_decode(p, d_syn, ns, d_e, d_it, d_cv, d_llr, st))) return rc;
        } else {
            if ((rc = launch_decode(p, d_syn, ns, d_e, d_it, d_cv, d_llr, nullptr, nullptr, nullptr, 0, slot, st))) return rc;
        }
        CU_TRY(cudaMemcpyAsync(ehat + s0 * t.nw, d_e, (size_t)ns * t.nw * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(iters + s0, d_it, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        if (conv) CU_TRY(cudaMemcpyAsync(conv + s0, d_cv, (size_t)ns, cudaMemcpyDeviceToHost, st));
        if (llr) CU_TRY(cudaMemcpyAsync(llr + s0 * t.n, d_llr, (size_t)ns * t.n * 8, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(p->streams[0]));
    CU_TRY(cudaStreamSynchronize(p->streams[1]));
    return QLDPC_OK;
}

int qldpc_classify(const qldpc_plan *px, const qldpc_plan *pz, const uint32_t *errx, const uint32_t *errz, const uint32_t *ehx,
                   const uint32_t *ehz, const uint32_t *synz, const uint32_t *synx, const int32_t *itx, const int32_t *itz,
                   int64_t shots, int64_t *counters, void *stream)
{
    if (!px || !pz || !counters || shots < 0) return fail(QLDPC_EINVAL, "null argument");
    if (px->tab.n != pz->tab.n) return fail(QLDPC_EINVAL, "Hx and Hz must have the same number of columns (physical qubits).");
    if (shots == 0) return QLDPC_OK;
    if (!errx || !errz || !ehx || !ehz || !synz || !synx || !itx || !itz) return fail(QLDPC_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(px->device));
    ClassifyArgs a;
    a.gz = graph_dev(px); a.gx = graph_dev(pz);
    a.colmask_z = px->d_hbits + (size_t)px->tab.m * px->tab.nw;
    a.colmask_x = pz->d_hbits + (size_t)pz->tab.m * pz->tab.nw;
    a.hcol_z = px->d_hcol; a.hcol_x = pz->d_hcol;
    const bool lg = px->d_lcol && pz->d_lcol;
    a.lcol_z = lg ? px->d_lcol : nullptr; a.lcol_x = lg ? pz->d_lcol : nullptr;
    a.errx = errx; a.errz = errz; a.ehx = ehx; a.ehz = ehz; a.synz = synz; a.synx = synx; a.itx = itx; a.itz = itz;
    a.shots = shots;
    a.counters = reinterpret_cast<unsigned long long *>(counters);
    const int threads = 256;
    const int grid = (int)std::min<int64_t>((int64_t)px->sm_count * 8, (shots + 7) / 8);
    // VH: 128-bit loads per column of H (both matrices use the wider one), VL: per column of the logical bases (2 or 8)
    const int vh = std::max(px->tab.mw, pz->tab.mw) > 32 ? 0 : col_words(std::max(px->tab.mw, pz->tab.mw)) / 4;   // 0: row-wise (CSR)
    const bool wide = lg && (std::max(px->lkw, pz->lkw) > 8);
    cudaStream_t cst = (cudaStream_t)stream;
#define QLDPC_CLS(VH) do { if (wide) classify_kernel<VH, 8><<<grid, threads, 0, cst>>>(a); else classify_kernel<VH, 2><<<grid, threads, 0, cst>>>(a); } while (0)
    switch (vh) {
    case 0: QLDPC_CLS(0); break;
    case 1: QLDPC_CLS(1); break;
    case 2: QLDPC_CLS(2); break;
    case 4: QLDPC_CLS(4); break;
    default: QLDPC_CLS(8); break;
    }
#undef QLDPC_CLS
    g_launches++;
    CU_TRY(cudaGetLastError());
    return QLDPC_OK;
}

// The whole shot loop of simulator.py:244-304 on HOST buffers: the measurement record (bit-packed [sy_z | sy_x | errX | errZ], as
// four arrays) in, the outcome counters out.  Chunks are double buffered on two compute streams of plan_x and fed by a third,
// copy-only stream: H2D of chunk k+1 overlaps the decodes and the classification of chunk k, and the first decode waits for
// its own syndromes only; the counters accumulate on the device and are read back once (80 bytes).
int qldpc_simulate_host(qldpc_plan *px, qldpc_plan *pz, const uint32_t *synz, const uint32_t *synx, const uint32_t *errx,
                        const uint32_t *errz, int64_t shots, int64_t *counters)
{
    if (!px || !pz || !counters || shots < 0) return fail(QLDPC_EINVAL, "null argument");
    if (px->tab.n != pz->tab.n) return fail(QLDPC_EINVAL, "Hx and Hz must have the same number of columns (physical qubits).");
    if (px->device != pz->device) return fail(QLDPC_EINVAL, "both plans must live on the same device");
    if (shots && (!synz || !synx || !errx || !errz)) return fail(QLDPC_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(px->device));
    const int nw = px->tab.nw, mzw = px->tab.mw, mxw = pz->tab.mw;
    static const int64_t chunk_env = [] { const char *ev = getenv("QLDPC_HOST_CHUNK"); return ev ? atoll(ev) : 0ll; }();   // tuning knob
    // 2^19 shots per chunk: measured 27.9 / 29.4 / 30.0 / 29.8 M shots/s end to end for 2^17 / 2^18 / 2^19 / 2^20 on the headline
    // configuration (every chunk costs two kernel tails, a single chunk has nothing to overlap its copies with)
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(shots, 1), chunk_env > 0 ? chunk_env : (1 << 19)));
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t b_sz = al((size_t)chunk * mzw * 4), b_sx = al((size_t)chunk * mxw * 4), b_n = al((size_t)chunk * nw * 4), b_it = al((size_t)chunk * 4);
    const size_t per_slot = b_sz + b_sx + 4 * b_n + 2 * b_it;
    int rc;
    if ((rc = ensure_scratch(px, 6, per_slot * 2 + 256))) return rc;
    unsigned long long *d_cnt = (unsigned long long *)((unsigned char *)px->scratch[6] + per_slot * 2);
    CU_TRY(cudaMemsetAsync(d_cnt, 0, QLDPC_NUM_COUNTERS * sizeof(unsigned long long), px->streams[0]));
    CU_TRY(cudaEventRecord(px->events[0], px->streams[0]));
    CU_TRY(cudaStreamWaitEvent(px->streams[1], px->events[0], 0));
    const bool osd = (px->opts.osd_order >= 0 && px->opts.dec_type >= QLDPC_MS) || (pz->opts.osd_order >= 0 && pz->opts.dec_type >= QLDPC_MS);
    // All host->device copies go through a third stream in the order they are needed -- Z syndromes, X syndromes, the two error
    // words -- and every consumer waits for its own input only: the X decode of the first chunk starts as soon as its
    // syndromes (1/6 of the chunk's bytes) have arrived, and the copies of chunk k+1 run under the kernels of chunk k.
    // events: [1+slot] Z syndromes in, [3+slot] X syndromes in, [5+slot] error words in, [7+slot] slot's buffers free again
    cudaStream_t cp = px->streams[2];
    CU_TRY(cudaStreamWaitEvent(cp, px->events[0], 0));
    int64_t k = 0;
    for (int64_t s0 = 0; s0 < shots; s0 += chunk, ++k) {
        const int slot = (int)(k & 1);
        const int64_t ns = std::min(chunk, shots - s0);
        cudaStream_t st = px->streams[slot];
        unsigned char *base = (unsigned char *)px->scratch[6] + per_slot * slot;
        uint32_t *d_sz = (uint32_t *)base, *d_sx = (uint32_t *)(base + b_sz);
        uint32_t *d_ex = (uint32_t *)(base + b_sz + b_sx), *d_ez = d_ex + b_n / 4, *d_hx = d_ez + b_n / 4, *d_hz = d_hx + b_n / 4;
        int32_t *d_ix = (int32_t *)(d_hz + b_n / 4), *d_iz = d_ix + b_it / 4;
        if (k >= 2) CU_TRY(cudaStreamWaitEvent(cp, px->events[7 + slot], 0));      // chunk k-2 has been classified
        CU_TRY(cudaMemcpyAsync(d_sz, synz + s0 * mzw, (size_t)ns * mzw * 4, cudaMemcpyHostToDevice, cp));
        CU_TRY(cudaEventRecord(px->events[1 + slot], cp));
        CU_TRY(cudaMemcpyAsync(d_sx, synx + s0 * mxw, (size_t)ns * mxw * 4, cudaMemcpyHostToDevice, cp));
        CU_TRY(cudaEventRecord(px->events[3 + slot], cp));
        CU_TRY(cudaMemcpyAsync(d_ex, errx + s0 * nw, (size_t)ns * nw * 4, cudaMemcpyHostToDevice, cp));
        CU_TRY(cudaMemcpyAsync(d_ez, errz + s0 * nw, (size_t)ns * nw * 4, cudaMemcpyHostToDevice, cp));
        CU_TRY(cudaEventRecord(px->events[5 + slot], cp));
        CU_TRY(cudaStreamWaitEvent(st, px->events[1 + slot], 0));
        if (osd) {
            // OSD plans share their failure scratch between chunks: one chunk at a time
            CU_TRY(cudaStreamSynchronize(px->streams[slot ^ 1]));
            if ((rc = qldpc_decode(px, d_sz, ns, d_hx, d_ix, nullptr, nullptr, st))) return rc;
            CU_TRY(cudaStreamWaitEvent(st, px->events[3 + slot], 0));
            if ((rc = qldpc_decode(pz, d_sx, ns, d_hz, d_iz, nullptr, nullptr, st))) return rc;
        } else {
            if ((rc = launch_decode(px, d_sz, ns, d_hx, d_ix, nullptr, nullptr, nullptr, nullptr, nullptr, 0, slot, st))) return rc;
            CU_TRY(cudaStreamWaitEvent(st, px->events[3 + slot], 0));
            if ((rc = launch_decode(pz, d_sx, ns, d_hz, d_iz, nullptr, nullptr, nullptr, nullptr, nullptr, 0, slot, st))) return rc;
        }
        CU_TRY(cudaStreamWaitEvent(st, px->events[5 + slot], 0));
        if ((rc = qldpc_classify(px, pz, d_ex, d_ez, d_hx, d_hz, d_sz, d_sx, d_ix, d_iz, ns, (int64_t *)d_cnt, st))) return rc;
        CU_TRY(cudaEventRecord(px->events[7 + slot], st));
    }
    CU_TRY(cudaStreamSynchronize(cp));
    CU_TRY(cudaStreamSynchronize(px->streams[1]));
    CU_TRY(cudaMemcpyAsync(counters, d_cnt, QLDPC_NUM_COUNTERS * sizeof(int64_t), cudaMemcpyDeviceToHost, px->streams[0]));
    CU_TRY(cudaStreamSynchronize(px->streams[0]));
    return QLDPC_OK;
}

int qldpc_sample(const qldpc_plan *px, const qldpc_plan *pz, double prob, uint64_t seed, int64_t first_shot, int64_t shots,
                 uint32_t *errx, uint32_t *errz, uint32_t *synz, uint32_t *synx, void *stream)
{
    if (!px || !pz || shots < 0) return fail(QLDPC_EINVAL, "null argument");
    if (px->tab.n != pz->tab.n) return fail(QLDPC_EINVAL, "Hx and Hz must have the same number of columns (physical qubits).");
    if (!(prob >= 0.0 && prob <= 1.0)) return fail(QLDPC_EINVAL, "p must lie in [0, 1]");
    if (shots == 0) return QLDPC_OK;
    if (!errx || !errz || !synz || !synx) return fail(QLDPC_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(px->device));
    SampleArgs a;
    a.gz = graph_dev(px); a.gx = graph_dev(pz);
    a.p = prob; a.seed = seed; a.first_shot = first_shot; a.shots = shots;
    a.errx = errx; a.errz = errz; a.synz = synz; a.synx = synx;
    const int threads = 256;
    const int grid = (int)std::min<int64_t>((int64_t)px->sm_count * 8, (shots + 7) / 8);
    const size_t smem = (size_t)(threads / 32) * 2 * a.gz.nw * 4;
    a.hcol_z = px->d_hcol; a.hcol_x = pz->d_hcol;
    const int mwmax = std::max(px->tab.mw, pz->tab.mw);
    cudaStream_t cst = (cudaStream_t)stream;
    if (mwmax > 32) sample_kernel<0><<<grid, threads, smem, cst>>>(a);
    else switch (col_words(mwmax) / 4) {
    case 1: sample_kernel<1><<<grid, threads, smem, cst>>>(a); break;
    case 2: sample_kernel<2><<<grid, threads, smem, cst>>>(a); break;
    case 4: sample_kernel<4><<<grid, threads, smem, cst>>>(a); break;
    default: sample_kernel<8><<<grid, threads, smem, cst>>>(a); break;
    }
    g_launches++;
    CU_TRY(cudaGetLastError());
    return QLDPC_OK;
}

}  // extern "C"
