// bp_kernel.cuh -- sum-product (tanh / atanh, division form) syndrome decoder, binary64 throughout.
//
// Semantics: decoders.py:189-290 (CPU restatement oracle/qldpc_oracle.c:bp_decode_one).
// Same execution shape as the min-sum kernel (one warp per shot, persistent CTAs, state in shared memory,
// slot-major edges, warp-uniform control flow).  v2c is rebuilt as T_j - c2v_e with T_j = L0 + sum_j
// (decoders.py:269), which is the value the reference stored after the previous layer step.
//   * check phase: lane <-> EDGE (LPC = 4, 8, 16 or 32 lanes per check).  tanh, the division and atanh -- the
//     expensive part -- run on all lanes in parallel; the product prod = ((t0 t1) t2) ... is formed in the
//     reference's order (np.prod is sequential, ascending variable, decoders.py:253-254) by a uniform loop of
//     shuffles, every lane of the check computing the same value (padding lanes contribute the factor 1.0).
//   * variable phase: lane <-> variable; the column sum follows NumPy's np.sum order on the gathered vector:
//     sequential below 8 terms, the 8-accumulator pairwise pattern from 8 terms on (oracle: numpy_sum_f64).
//   * hard decision T_j < 0, previous decision = sign of the old T_j; flips update the residual parity words
//     cooperatively and maintain the count of unsatisfied checks (decoders.py:280-285).
#pragma once
#include "common.cuh"
#include "npymath.cuh"

namespace qldpc {

// tanh / atanh are NumPy's own routines (npymath.cuh): with them the kernel is bit-identical to the reference on all 4800 decodes
// of the three large reference goldens (tests/golden/big_LP118_0_BP_F_p{02,05,10}_X.npz).  With the CUDA math library's
// functions (round 1) the agreement was 99.69 % at p = 0.05 -- NumPy's SIMD tanh differs from any libm in the last bit on a
// quarter of all arguments.
struct BpConst {
    double L0;     // prior LLR (decoders.py:232)
    double eps;    // decoders.py:195
    int max_iter;
};

struct BpSmemLayout {
    int off_c2v;   // double [dc*ms + 1]: the last word stays 0.0 (padding entries of the fixed-stride column table)
    int off_T;     // double [n + 1]
    int off_par, off_syn;
    int off_team;  // 16 bytes: shot index mailbox (int64), unsatisfied-check count (int32)
    int off_tk;    // double [4][32]: the tanh factors of the pass a warp is working on (one row per warp of the team)
    int bytes;
};

// graph tables + the NumPy math tables, in front of the per-team state
__host__ __device__ inline int bp_table_bytes(const Tables &t) { return ((t.len * 2 + 15) & ~15) + (int)((sizeof(NpymTables) + 15) & ~size_t(15)); }

__host__ __device__ inline BpSmemLayout bp_layout(const Tables &t)
{
    BpSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 8 * (t.dc * t.ms + 1);
    l.off_T = o;   o += 8 * (t.n + 1);
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    o = (o + 7) & ~7;
    l.off_team = o; o += 16;
    o = (o + 15) & ~15;
    l.off_tk = o;  o += 4 * 32 * 8;
    l.bytes = (o + 15) & ~15;
    return l;
}

// np.sum over fewer than 8 terms is sequential from 0.0.  Columns are read through the fixed-stride table: DVF entries per
// variable, entries past the degree point at a word that is always 0.0 (s + 0.0 == s; a sum that ends as -0.0 can only become
// +0.0, which the following L0 + s and the sign tests cannot tell apart), so the loop is straight-line code without guards.
template <int DVF>
__device__ __forceinline__ double bp_colsum_fixed(const double *c2v, const uint16_t *row)
{
    double term[DVF];
#pragma unroll
    for (int x = 0; x < DVF; ++x) term[x] = c2v[row[x]];
    double s = 0.0;
#pragma unroll
    for (int x = 0; x < DVF; ++x) s = __dadd_rn(s, term[x]);
    return s;
}

// np.sum(c2v[edges of j]) for columns of 8 and more checks (any column weight)
__device__ __forceinline__ double bp_colsum(const double *c2v, const uint16_t *col_pos, int t0, int cnt)
{
    if (cnt < 8) {                                   // mixed degrees around 8: rare, lane-divergent but correct
        double s = 0.0;
        for (int x = 0; x < cnt; ++x) s = __dadd_rn(s, c2v[col_pos[t0 + x]]);
        return s;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = c2v[col_pos[t0 + k]];
    int i = 8;
    for (; i < cnt - (cnt % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], c2v[col_pos[t0 + i + k]]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < cnt; ++i) res = __dadd_rn(res, c2v[col_pos[t0 + i]]);
    return res;
}

// W warps ("team") share one shot: the binary64 state limits a CTA to ~11 shots (LP118_0), and 11 warps cannot keep the
// issue slots busy through the long dependent instruction streams of tanh / atanh (measured 49 % issue-active).  The team
// splits the passes of the check phase and the trips of the variable phase between its warps and meets at a named barrier
// (bar.sync id, 32*W) between the phases; the unsatisfied-check count lives in shared memory.  Results are bit-identical
// for any W (every edge and every variable is still computed by exactly one lane, in the same arithmetic order).
template <int LPC, int W>
__global__ void __launch_bounds__(1024, 1) bp_decode_kernel(Tables t, const uint16_t *__restrict__ blob, BpConst c, DecodeIO io)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < t.len / 8; i += blockDim.x) dst[i] = src[i];
    }
    NpymTables *npym_tab = reinterpret_cast<NpymTables *>(smem + ((t.len * 2 + 15) & ~15));
    npym_stage_tables(npym_tab);
    __syncthreads();
    const NpymTables &nt = *npym_tab;
    const uint16_t *var_tab = tab + t.off_var;          // byte offsets 4*j
    const uint16_t *col_ptr = tab + t.off_col_ptr;
    const uint16_t *col_pos = tab + t.off_col_pos;
    const uint16_t *colf = tab + t.off_colf;
    const uint16_t *col_chk = tab + t.off_col_chk;
    const uint16_t *layer_ptr = tab + t.off_layer_ptr;
    const uint16_t *layer_chk = tab + t.off_layer_chk;
    const uint16_t *lvar_ptr = tab + t.off_lvar_ptr;
    const uint16_t *lvar_idx = tab + t.off_lvar_idx;
    const uint32_t *rowpar = reinterpret_cast<const uint32_t *>(tab + t.off_rowpar);

    const BpSmemLayout lay = bp_layout(t);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(full, (int)(threadIdx.x >> 5), 0);
    const int team = warp / W, sub = warp % W;
    const int tl = sub * 32 + lane;                            // thread index within the team
    constexpr int TT = 32 * W;                                 // threads per team
    unsigned char *base = smem + bp_table_bytes(t) + (size_t)team * lay.bytes;
    double *c2v = reinterpret_cast<double *>(base + lay.off_c2v);
    double *T = reinterpret_cast<double *>(base + lay.off_T);
    uint32_t *par = reinterpret_cast<uint32_t *>(base + lay.off_par);
    uint32_t *syn = reinterpret_cast<uint32_t *>(base + lay.off_syn);
    volatile long long *team_shot = reinterpret_cast<volatile long long *>(base + lay.off_team);
    int *team_unsat = reinterpret_cast<int *>(base + lay.off_team + 8);
    double *tkw = reinterpret_cast<double *>(base + lay.off_tk) + 32 * sub;     // this warp's row
    const int m = t.ms, n = t.n, dc = t.dc;   // m: slot stride
    const double one_m_eps = 1.0 - c.eps;                     // `1-eps` of decoders.py:257
    const bool init_bit = c.L0 < 0.0;
    constexpr int CPP = 32 / LPC;
    const int k = lane % LPC;                                  // slot of this lane
    const int grp = lane & ~(LPC - 1);                         // first lane of the check's group
    auto team_sync = [&]() {
        if (W == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(TT) : "memory");
    };

    for (;;) {
        if (tl == 0) *team_shot = (long long)atomicAdd(io.work_counter, 1ull);
        team_sync();
        const long long shot = *team_shot;
        if (shot >= io.shots) break;
        for (int i = tl; i <= dc * m; i += TT) c2v[i] = 0.0;        // :236 (and the always-zero word behind the array)
        for (int i = tl; i <= n; i += TT) T[i] = c.L0;              // v2c = L0 (:235)
        if (sub == 0) {
            int u = 0;
            for (int i = lane; i < t.mw; i += 32) {
                const uint32_t w = io.syn[shot * t.mw + i];
                const uint32_t p0 = init_bit ? (w ^ rowpar[i]) : w;
                syn[i] = w; par[i] = p0;
                u += __popc(p0);
            }
            u = __reduce_add_sync(full, u);
            if (lane == 0) *team_unsat = u;
        }
        team_sync();

        bool converged = false, first = true;
        int it = 0;
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = 0; l < t.nl; ++l) {
                // ---------------- check-node phase (decoders.py:249-262), one lane per edge
                const int qb = layer_ptr[l], qe = layer_ptr[l + 1];
                for (int q0 = qb + sub * CPP; q0 < qe; q0 += W * CPP) {
                    const int q = q0 + lane / LPC;
                    const bool act = q < qe;
                    const int i = layer_chk[act ? q : qb];
                    const int pos = (k < dc ? k : 0) * m + i;
                    const uint32_t joff = var_tab[pos];
                    const bool valid = act && k < dc && joff != kPad;
                    double tk = 1.0;
                    if (valid) tk = npym_tanh(__dsub_rn(T[joff >> 2], c2v[pos]) / 2.0, nt);     // np.tanh(v2c/2) (:254, :256)
                    // np.prod: sequential (:253-254), 1.0 * t0 = t0 exactly; the lanes past the row weight hold the exact factor 1.0.
                    // The factors of a check travel through shared memory: one store and LPC/2 broadcast 128-bit loads per lane
                    // instead of 2*LPC shuffles.
                    tkw[lane] = tk;
                    __syncwarp();
                    double prod;
                    {
                        const double2 *row = reinterpret_cast<const double2 *>(tkw + grp);
                        double2 f = row[0];
                        prod = __dmul_rn(f.x, f.y);
#pragma unroll
                        for (int x = 1; x < LPC / 2; ++x) {
                            f = row[x];
                            prod = __dmul_rn(__dmul_rn(prod, f.x), f.y);
                        }
                    }
                    if (valid) {
                        double th2 = prod / tk;                                        // :256
                        if (fabs(th2) >= one_m_eps) {                                  // :257-258
                            const double sg = (th2 > 0.0) ? 1.0 : ((th2 < 0.0) ? -1.0 : 0.0);
                            th2 = __dsub_rn(th2, __dmul_rn(c.eps, sg));
                        }
                        double val = 2.0 * npym_arctanh(th2, nt);                      // np.arctanh (:259)
                        if ((syn[i >> 5] >> (i & 31)) & 1u) val = -val;                // :260-261
                        c2v[pos] = val;                                                // :262
                    }
                    __syncwarp();
                }
                team_sync();
                // ---------------- variable-node phase (decoders.py:265-280) on the variables whose messages changed
                const int vb = first ? 0 : lvar_ptr[l], ve = first ? t.n_pad : lvar_ptr[l + 1];
                int delta = 0;
                for (int q = vb + tl; q < ve; q += TT) {
                    const int j = first ? (q < n ? q : n) : lvar_idx[q];
                    double sum;
                    if (t.dv < 8) {                                                   // warp-uniform
                        const uint16_t *row = colf + j * t.dv;
                        switch (t.dv) {
                        case 0: sum = 0.0; break;
                        case 1: sum = bp_colsum_fixed<1>(c2v, row); break;
                        case 2: sum = bp_colsum_fixed<2>(c2v, row); break;
                        case 3: sum = bp_colsum_fixed<3>(c2v, row); break;
                        case 4: sum = bp_colsum_fixed<4>(c2v, row); break;
                        case 5: sum = bp_colsum_fixed<5>(c2v, row); break;
                        case 6: sum = bp_colsum_fixed<6>(c2v, row); break;
                        default: sum = bp_colsum_fixed<7>(c2v, row); break;
                        }
                    } else {
                        const int t0 = (j < n) ? col_ptr[j] : 0, cnt = (j < n) ? col_ptr[j + 1] - t0 : 0;
                        sum = bp_colsum(c2v, col_pos, t0, cnt);
                    }
                    const double tot = __dadd_rn(c.L0, sum);                          // :269 / :275
                    const double t_old = T[j];
                    T[j] = tot;
                    uint32_t flips = __ballot_sync(full, (tot < 0.0) != (t_old < 0.0));           // :280
                    while (flips) {
                        const int src = __ffs(flips) - 1;
                        flips &= flips - 1;
                        const int jf = __shfl_sync(full, j, src);
                        for (int x = col_ptr[jf] + lane; x < col_ptr[jf + 1]; x += 32) {     // columns may hold more than 32 checks
                            const int ch = col_chk[x];
                            const uint32_t bit = 1u << (ch & 31);
                            const uint32_t old = atomicXor(&par[ch >> 5], bit);
                            delta += (old & bit) ? -1 : 1;
                        }
                    }
                }
                delta = __reduce_add_sync(full, delta);
                if (lane == 0 && delta != 0) atomicAdd(team_unsat, delta);
                team_sync();
                first = false;
                // every warp of the team reads the count before any of them can pass the next barrier, and the count is
                // only modified after that barrier
                if (*reinterpret_cast<volatile int *>(team_unsat) == 0) { converged = true; break; }   // :283-285
            }
        }
        const int iters = it;   // ++it has run after a converging break: it+1 of :285, else max_iter
        for (int w = sub; w < t.nw; w += W) {
            const int j = w * 32 + lane;
            const uint32_t bits = __ballot_sync(full, j < n && c.max_iter > 0 && T[j < n ? j : n] < 0.0);
            if (lane == 0) io.ehat[shot * t.nw + w] = bits;
        }
        if (tl == 0) { io.iters[shot] = iters; if (io.conv) io.conv[shot] = converged ? 1 : 0; }
        if (io.llr) {
            double *dst = io.llr + shot * (long long)n;
            for (int j = tl; j < n; j += TT) dst[j] = T[j];
        }
        if (!converged && io.fail_count) {
            if (tl == 0) {
                const int s2 = atomicAdd(io.fail_count, 1);
                *team_shot = s2;                           // reuse the mailbox: the shot index has been consumed by every warp
                if (s2 < io.fail_cap) io.fail_shot[s2] = (int)shot;
            }
            // the mailbox write must not overtake the other warps' read of the shot index above: they read it right after the
            // first barrier of this trip, long before this point, and cannot start the next trip before the barrier below
            team_sync();
            const int slot = (int)*team_shot;
            if (slot < io.fail_cap) {
                double *dst = io.fail_llr + (long long)slot * n;
                for (int j = tl; j < n; j += TT) dst[j] = T[j];
            }
        }
        team_sync();
    }
}

}  // namespace qldpc
