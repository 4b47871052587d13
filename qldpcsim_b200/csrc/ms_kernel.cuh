// ms_kernel.cuh -- normalised min-sum syndrome decoder, flooding / layered / serial, one persistent launch.
//
// Semantics: decoders.py:110-182 of the reference (bit-level spec in SURVEY.md App. A.1, restated on the CPU
// in oracle/qldpc_oracle.c:ms_decode_one).  Mapping:
//   * one WARP owns one shot from its first layer step to its exit and then pulls the next shot from a
//     global dispenser (shots converge after 1..max_iter iterations, so there are no lock-step batches);
//   * the whole message state of the shot lives in shared memory: c2v as binary32 per edge (slot-major),
//     the binary32 column sums S_j, the hard decision and the residual syndrome H e + s as bit words;
//   * v2c is never stored: v2c_e = fl64(fl64(prior + S_j) - c2v_e) is rebuilt from S_j and c2v_e, which is
//     exactly what decoders.py:173,:177 computes (prior = binary32-rounded L during the very first layer
//     step, decoders.py:148-149, L afterwards);
//   * check phase: lane <-> check of the layer, serial over the row (first-argmin / second-min scan);
//     variable phase: lane <-> variable adjacent to the layer, re-summing ALL its c2v in ascending check
//     order in binary32 (decoders.py:172) -- layers need not be column-disjoint (simulator.py:230-234 hands
//     the decoder the partition of the OTHER matrix), so the phases are separated by a warp barrier;
//   * convergence is tested after every layer step (decoders.py:175-176) on the incrementally maintained
//     residual: a flipped hard decision toggles the parity bits of its checks.
// No fused multiply-add may be formed in this file (compile with -fmad=false); every operation below that
// the spec rounds individually uses an explicit _rn intrinsic anyway.
#pragma once
#include "common.cuh"

namespace qldpc {

struct MsSmemLayout {
    // per-shot state, offsets in bytes from the warp's base
    int off_c2v;   // float [dc*m]
    int off_S;     // float [n]
    int off_e;     // uint32 [nw]
    int off_par;   // uint32 [mw]
    int off_syn;   // uint32 [mw]
    int bytes;     // multiple of 16
};

__host__ __device__ inline MsSmemLayout ms_layout(const Tables &t)
{
    MsSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 4 * t.dc * t.m;
    l.off_S = o;   o += 4 * t.n;
    l.off_e = o;   o += 4 * t.nw;
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    l.bytes = (o + 15) & ~15;
    return l;
}

template <int DC, bool REGULAR>
__global__ void __launch_bounds__(1024, 1) ms_decode_kernel(Tables t, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    {   // graph tables -> shared memory (once per CTA), 16 B per thread per trip
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

    const MsSmemLayout lay = ms_layout(t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem + ((t.len * 2 + 15) & ~15) + (size_t)warp * lay.bytes;
    float *c2v = reinterpret_cast<float *>(base + lay.off_c2v);
    float *S = reinterpret_cast<float *>(base + lay.off_S);
    uint32_t *eb = reinterpret_cast<uint32_t *>(base + lay.off_e);
    uint32_t *par = reinterpret_cast<uint32_t *>(base + lay.off_par);
    uint32_t *syn = reinterpret_cast<uint32_t *>(base + lay.off_syn);
    const int m = t.m, n = t.n;
    const unsigned full = 0xffffffffu;

    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;

        // ---- initial state: c2v = 0 (decoders.py:150), S = 0, e = 0, residual = syndrome
        for (int i = lane; i < t.dc * m; i += 32) c2v[i] = 0.0f;
        for (int i = lane; i < n; i += 32) S[i] = 0.0f;
        for (int i = lane; i < t.nw; i += 32) eb[i] = 0u;
        for (int i = lane; i < t.mw; i += 32) {
            uint32_t w = io.syn[shot * t.mw + i];
            syn[i] = w;
            par[i] = w;
        }
        __syncwarp();

        bool converged = false;
        bool first = true;
        int it = 0;
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = 0; l < t.nl; ++l) {
                const double prior = first ? c.Lf : c.L;
                // ---------------- check-node phase (decoders.py:156-169)
                const int qb = layer_ptr[l], qe = layer_ptr[l + 1];
                for (int q = qb + lane; q < qe; q += 32) {
                    const int i = layer_chk[q];
                    double m1 = __longlong_as_double(0x7ff0000000000000ll), m2 = m1;
                    int k1 = 0, deg = 0;
                    uint32_t sb = 0;
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        const int pos = k * m + i;
                        const uint16_t j = var_tab[pos];
                        if (REGULAR || j != kPad) {
                            const double post = __dadd_rn(prior, (double)S[j]);          // :173
                            const double v = __dsub_rn(post, (double)c2v[pos]);          // :177
                            const double av = fabs(v);
                            sb |= (v < 0.0 ? 1u : 0u) << k;                              // :157-158 (0 -> +1)
                            if (av < m1) { m2 = m1; m1 = av; k1 = k; }                   // first argmin (:161)
                            else if (av < m2) m2 = av;                                   // min over the others (:162-164)
                            ++deg;
                        }
                    }
                    if (deg) {
                        if (isinf(m1)) m1 = 0.0;                                         // :165
                        if (isinf(m2)) m2 = 0.0;                                         // :166
                        float r1 = __double2float_rn(__dmul_rn(c.beta, m1));             // f64 product, f32 store (:167)
                        float r2 = __double2float_rn(__dmul_rn(c.beta, m2));             // (:168)
                        if (isinf(r1)) r1 = 0.0f;                                        // :169
                        if (isinf(r2)) r2 = 0.0f;
                        const uint32_t P = (__popc(sb) & 1u) ^ ((syn[i >> 5] >> (i & 31)) & 1u);   // sign product x syndrome sign (:151,:159)
#pragma unroll
                        for (int k = 0; k < DC; ++k) {
                            if (REGULAR || k < deg) {
                                const float mag = (k == k1) ? r2 : r1;
                                c2v[k * m + i] = (((sb >> k) & 1u) ^ P) ? -mag : mag;
                            }
                        }
                    }
                }
                __syncwarp();
                // ---------------- variable-node phase (decoders.py:172-174) on the variables whose sums changed;
                // the very first step visits every variable (the reference recomputes all posteriors, and a
                // variable outside layer 0 has posterior L, which may be negative for p > 1/2).
                const int vb = first ? 0 : lvar_ptr[l], ve = first ? n : lvar_ptr[l + 1];
                for (int q = vb + lane; q < ve; q += 32) {
                    const int j = first ? q : lvar_idx[q];
                    const int t0 = col_ptr[j], t1 = col_ptr[j + 1];
                    float s = 0.0f;
                    for (int x = t0; x < t1; ++x) s = __fadd_rn(s, c2v[col_pos[x]]);    // sequential f32, ascending check (:172)
                    S[j] = s;
                    const uint32_t bit = __dadd_rn(c.L, (double)s) < 0.0 ? 1u : 0u;     // :173-174
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
                // ---------------- H e == syndrome ?  (decoders.py:175-176)
                uint32_t nz = 0;
                for (int w = lane; w < t.mw; w += 32) nz |= par[w];
                if (!__any_sync(full, nz != 0)) { converged = true; break; }
            }
        }
        const int iters = it;   // the outer ++it has already run after a converging break: it+1 of decoders.py:176, else max_iter (:182)
        // ---- outputs
        for (int w = lane; w < t.nw; w += 32) io.ehat[shot * t.nw + w] = eb[w];
        if (lane == 0) {
            io.iters[shot] = iters;
            if (io.conv) io.conv[shot] = converged ? 1 : 0;
        }
        if (io.llr) {
            double *dst = io.llr + shot * (long long)n;
            for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)S[j]);
        }
        if (!converged && io.fail_count) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(io.fail_count, 1);
            slot = __shfl_sync(full, slot, 0);
            if (slot < io.fail_cap) {
                if (lane == 0) io.fail_shot[slot] = (int)shot;
                double *dst = io.fail_llr + (long long)slot * n;
                for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)S[j]);
            }
        }
        __syncwarp();
    }
}

}  // namespace qldpc
