// bp_kernel.cuh -- sum-product (tanh / atanh, division form) syndrome decoder, binary64 throughout.
//
// Semantics: decoders.py:189-290 (CPU restatement oracle/qldpc_oracle.c:bp_decode_one).
// Same execution shape as the min-sum kernel (one warp per shot, persistent CTAs, state in shared memory,
// slot-major edges).  v2c is rebuilt as T_j - c2v_e with T_j = L0 + sum_j (decoders.py:269), which is the value
// the reference stored after the previous layer step.  The column sum follows NumPy's np.sum order on the
// gathered vector: sequential below 8 terms, the 8-accumulator pairwise pattern from 8 terms on.
#pragma once
#include "common.cuh"

namespace qldpc {

struct BpConst {
    double L0;     // prior LLR (decoders.py:232)
    double eps;    // decoders.py:195
    int max_iter;
};

struct BpSmemLayout {
    int off_c2v;   // double [dc*m]
    int off_T;     // double [n]
    int off_e, off_par, off_syn;
    int bytes;
};

__host__ __device__ inline BpSmemLayout bp_layout(const Tables &t)
{
    BpSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 8 * t.dc * t.m;
    l.off_T = o;   o += 8 * t.n;
    l.off_e = o;   o += 4 * t.nw;
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    l.bytes = (o + 15) & ~15;
    return l;
}

// np.sum(c2v[edges of j]) -- see numpy_sum_f64 in the oracle.
__device__ __forceinline__ double bp_colsum(const double *c2v, const uint16_t *col_pos, int t0, int t1)
{
    const int cnt = t1 - t0;
    if (cnt < 8) {
        double s = 0.0;
        for (int x = t0; x < t1; ++x) s = __dadd_rn(s, c2v[col_pos[x]]);
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

__global__ void __launch_bounds__(1024, 1) bp_decode_kernel(Tables t, const uint16_t *__restrict__ blob, BpConst c, DecodeIO io)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < t.len / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint16_t *var_tab = tab + t.off_var;
    const uint16_t *col_ptr = tab + t.off_col_ptr;
    const uint16_t *col_pos = tab + t.off_col_pos;
    const uint16_t *col_chk = tab + t.off_col_chk;
    const uint16_t *layer_ptr = tab + t.off_layer_ptr;
    const uint16_t *layer_chk = tab + t.off_layer_chk;
    const uint16_t *lvar_ptr = tab + t.off_lvar_ptr;
    const uint16_t *lvar_idx = tab + t.off_lvar_idx;

    const BpSmemLayout lay = bp_layout(t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem + ((t.len * 2 + 15) & ~15) + (size_t)warp * lay.bytes;
    double *c2v = reinterpret_cast<double *>(base + lay.off_c2v);
    double *T = reinterpret_cast<double *>(base + lay.off_T);
    uint32_t *eb = reinterpret_cast<uint32_t *>(base + lay.off_e);
    uint32_t *par = reinterpret_cast<uint32_t *>(base + lay.off_par);
    uint32_t *syn = reinterpret_cast<uint32_t *>(base + lay.off_syn);
    const int m = t.m, n = t.n, dc = t.dc;
    const unsigned full = 0xffffffffu;
    const double one_m_eps = 1.0 - c.eps;                     // `1-eps` of decoders.py:257

    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;
        for (int i = lane; i < dc * m; i += 32) c2v[i] = 0.0;       // :236
        for (int i = lane; i < n; i += 32) T[i] = c.L0;             // v2c = L0 (:235)
        for (int i = lane; i < t.nw; i += 32) eb[i] = 0u;
        for (int i = lane; i < t.mw; i += 32) { uint32_t w = io.syn[shot * t.mw + i]; syn[i] = w; par[i] = w; }
        __syncwarp();

        bool converged = false, first = true;
        int it = 0;
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = 0; l < t.nl; ++l) {
                // ---------------- check-node phase (decoders.py:249-262)
                const int qb = layer_ptr[l], qe = layer_ptr[l + 1];
                for (int q = qb + lane; q < qe; q += 32) {
                    const int i = layer_chk[q];
                    double prod = 1.0;
                    int deg = 0;
                    for (int k = 0; k < dc; ++k) {
                        const int pos = k * m + i;
                        const uint32_t joff = var_tab[pos];                      // byte offset 4*j
                        if (joff == kPad) break;
                        const double v = __dsub_rn(T[joff >> 2], c2v[pos]);
                        prod = __dmul_rn(prod, tanh(v / 2.0));                     // :253-254 (np.prod is sequential)
                        ++deg;
                    }
                    const bool neg = (syn[i >> 5] >> (i & 31)) & 1u;
                    // second pass: the v2c values are still intact (c2v of this row is only overwritten below,
                    // position by position, after its own v2c has been rebuilt)
                    for (int k = 0; k < deg; ++k) {
                        const int pos = k * m + i;
                        const double v = __dsub_rn(T[var_tab[pos] >> 2], c2v[pos]);
                        double th2 = prod / tanh(v / 2.0);                         // :256
                        if (fabs(th2) >= one_m_eps) {                              // :257-258
                            const double sg = (th2 > 0.0) ? 1.0 : ((th2 < 0.0) ? -1.0 : 0.0);
                            th2 = __dsub_rn(th2, __dmul_rn(c.eps, sg));
                        }
                        double val = 2.0 * atanh(th2);                             // :259
                        if (neg) val = -val;                                       // :260-261
                        c2v[pos] = val;                                            // :262
                    }
                }
                __syncwarp();
                // ---------------- variable-node phase (decoders.py:265-280) on the variables whose messages changed
                const int vb = first ? 0 : lvar_ptr[l], ve = first ? n : lvar_ptr[l + 1];
                for (int q = vb + lane; q < ve; q += 32) {
                    const int j = first ? q : lvar_idx[q];
                    const int t0 = col_ptr[j], t1 = col_ptr[j + 1];
                    const double tot = __dadd_rn(c.L0, bp_colsum(c2v, col_pos, t0, t1));   // :269 / :275
                    T[j] = tot;
                    const uint32_t bit = tot < 0.0 ? 1u : 0u;                      // :280
                    const uint32_t old = (eb[j >> 5] >> (j & 31)) & 1u;
                    if (bit != old) {
                        atomicXor(&eb[j >> 5], 1u << (j & 31));
                        for (int x = t0; x < t1; ++x) {
                            const int ch = col_chk[x];
                            atomicXor(&par[ch >> 5], 1u << (ch & 31));
                        }
                    }
                }
                __syncwarp();
                first = false;
                uint32_t nz = 0;
                for (int w = lane; w < t.mw; w += 32) nz |= par[w];
                if (!__any_sync(full, nz != 0)) { converged = true; break; }       // :283-285
            }
        }
        const int iters = it;
        for (int w = lane; w < t.nw; w += 32) io.ehat[shot * t.nw + w] = eb[w];
        if (lane == 0) { io.iters[shot] = iters; if (io.conv) io.conv[shot] = converged ? 1 : 0; }
        if (io.llr) {
            double *dst = io.llr + shot * (long long)n;
            for (int j = lane; j < n; j += 32) dst[j] = T[j];
        }
        if (!converged && io.fail_count) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(io.fail_count, 1);
            slot = __shfl_sync(full, slot, 0);
            if (slot < io.fail_cap) {
                if (lane == 0) io.fail_shot[slot] = (int)shot;
                double *dst = io.fail_llr + (long long)slot * n;
                for (int j = lane; j < n; j += 32) dst[j] = T[j];
            }
        }
        __syncwarp();
    }
}

}  // namespace qldpc
