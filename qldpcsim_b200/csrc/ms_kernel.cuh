// ms_kernel.cuh -- normalised min-sum syndrome decoder, flooding / layered / serial, one persistent launch.
//
// Semantics: decoders.py:110-182 of the reference (bit-level spec in SURVEY.md App. A.1, restated on the CPU
// in oracle/qldpc_oracle.c:ms_decode_one).  Mapping:
//   * one WARP owns one shot from its first layer step to its exit and then pulls the next shot from a
//     global dispenser (shots converge after 1..max_iter iterations, so there are no lock-step batches);
//   * the whole message state of the shot lives in shared memory: c2v as binary32 per edge (slot-major: the
//     k-th edge of check i at k*m+i), the binary32 column sums S_j, the residual syndrome H e + s as bit words;
//   * v2c is never stored: v2c_e = fl64(fl64(prior + S_j) - c2v_e) is rebuilt from S_j and c2v_e, which is
//     exactly what decoders.py:173,:177 computes (prior = binary32-rounded L during the very first layer
//     step, decoders.py:148-149, L afterwards);
//   * check phase: LPC = 1, 2, 4 or 8 lanes share one check (chosen per layer so that the layer fills the
//     warp: 16-check layers use 2 lanes x 4 edges, single-check layers 8 lanes x 1 edge); each lane scans its
//     slots for (min1, first argmin, min2, sign parity) and the partial results are merged with xor-shuffles;
//   * variable phase: lane <-> variable adjacent to the layer, re-summing ALL its c2v in ascending check
//     order in binary32 (decoders.py:172) through a padded table of shared-memory offsets (padding entries
//     point at a slot that always holds +0.0f, and s + 0.0f == s).  Layers need not be column-disjoint
//     (simulator.py:230-234 hands the decoder the partition of the OTHER matrix), so the two phases are
//     separated by a warp barrier;
//   * hard decision: fl64(L + S_j) < 0  <=>  S_j < Tf with Tf = L negated and rounded UP to binary32 (the
//     double sum of two doubles is exact whenever it is tiny, so its sign is the sign of the real sum), which
//     needs one binary32 compare and no stored decision bits: the previous decision is S_j(old) < Tf;
//   * convergence is tested after every layer step (decoders.py:175-176) on an incrementally maintained count
//     of unsatisfied checks: a flipped decision toggles the parity bits of its checks (shared-memory atomics,
//     rare) and adds +-1 per toggled bit.
// No fused multiply-add may be formed in this file (compile with -fmad=false); every operation that the spec
// rounds individually uses an explicit _rn intrinsic anyway.
#pragma once
#include "common.cuh"

namespace qldpc {

struct MsSmemLayout {
    // per-shot state, offsets in bytes from the warp's base
    int off_c2v;   // float [dc*m + 4]  (entry dc*m is the always-zero slot)
    int off_S;     // float [n + 1] (entry n: dummy variable)
    int off_par;   // uint32 [mw]
    int off_syn;   // uint32 [mw]
    int bytes;     // multiple of 16
};

__host__ __device__ inline MsSmemLayout ms_layout(const Tables &t)
{
    MsSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 4 * (t.dc * t.m + 4);
    o = (o + 15) & ~15;
    l.off_S = o;   o += 4 * (t.n + 1);
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    l.bytes = (o + 15) & ~15;
    return l;
}

__device__ __forceinline__ float lds_f32(const unsigned char *base, uint32_t byte_off)
{
    return *reinterpret_cast<const float *>(base + byte_off);
}

struct CnPartial {
    double m1, m2;   // smallest / second smallest |v2c| over the lane's slots (inf if none)
    int k1;          // slot of the first minimum
    uint32_t par;    // parity of the negative signs
};

// Merge two partial scans of disjoint slot sets.  Symmetric and branch-free on purpose (a lane-parity branch here
// splits the warp in two halves that then run the rest of the decode separately): the winner is the smaller first
// minimum, ties go to the lower slot, which preserves np.argmin's "first minimum" (decoders.py:161); the new second
// minimum is the smaller of the loser's first and the winner's second minimum.
__device__ __forceinline__ void cn_merge(CnPartial &a, const CnPartial &b)
{
    const bool lt = (b.m1 < a.m1) | ((b.m1 == a.m1) & (b.k1 < a.k1));
    const double loser = lt ? a.m1 : b.m1;
    const double keep = lt ? b.m2 : a.m2;
    a.m2 = (loser < keep) ? loser : keep;
    a.m1 = lt ? b.m1 : a.m1;
    a.k1 = lt ? b.k1 : a.k1;
    a.par ^= b.par;
}

// Check-node phase of one layer with LPC lanes per check (decoders.py:156-169).
template <int DC, bool REGULAR, int LPC>
__device__ __forceinline__ void ms_check_phase(int qb, int qe, int lane, int m, const uint16_t *__restrict__ layer_chk,
                                               const uint16_t *__restrict__ var_tab, unsigned char *c2v_b,
                                               const unsigned char *S_b, const uint32_t *syn, double prior, double beta)
{
    constexpr int SPL = (DC + LPC - 1) / LPC;       // slots per lane
    constexpr int CPP = 32 / LPC;                   // checks per pass
    const int h = lane % LPC;                       // which slice of the row
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int q0 = qb; q0 < qe; q0 += CPP) {
        const int q = q0 + lane / LPC;
        const bool act = q < qe;
        const int i = layer_chk[act ? q : qb];
        CnPartial pr;
        pr.m1 = inf; pr.m2 = inf; pr.k1 = 0; pr.par = 0u;
        uint32_t sb = 0;                            // sign bits of the lane's own slots
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int k = h * SPL + s;
            if (k < DC) {
                const int pos = k * m + i;
                const uint32_t joff = var_tab[pos];                         // byte offset of S_j, kPad past a short row
                if (REGULAR || joff != kPad) {
                    const double post = __dadd_rn(prior, (double)lds_f32(S_b, joff));          // :173
                    const double v = __dsub_rn(post, (double)lds_f32(c2v_b, 4u * pos));      // :177
                    const double av = fabs(v);
                    const uint32_t neg = v < 0.0 ? 1u : 0u;                                   // :157-158 (0 -> +1)
                    sb |= neg << s;
                    pr.par ^= neg;
                    const bool lt1 = av < pr.m1, lt2 = av < pr.m2;                            // selects, not branches
                    pr.m2 = lt1 ? pr.m1 : (lt2 ? av : pr.m2);                                 // min over the others (:162-164)
                    pr.m1 = lt1 ? av : pr.m1;                                                 // first argmin (:161)
                    pr.k1 = lt1 ? k : pr.k1;
                }
            }
        }
        // butterfly over the LPC lanes of the check
#pragma unroll
        for (int d = 1; d < LPC; d <<= 1) {
            CnPartial o;
            o.m1 = __shfl_xor_sync(0xffffffffu, pr.m1, d);
            o.m2 = __shfl_xor_sync(0xffffffffu, pr.m2, d);
            const uint32_t pk = __shfl_xor_sync(0xffffffffu, (uint32_t)pr.k1 | (pr.par << 8), d);
            o.k1 = (int)(pk & 0xffu);
            o.par = pk >> 8;
            cn_merge(pr, o);
        }
        if (act) {
            double m1 = pr.m1, m2 = pr.m2;
            if (isinf(m1)) m1 = 0.0;                                                          // :165
            if (isinf(m2)) m2 = 0.0;                                                          // :166
            float r1 = __double2float_rn(__dmul_rn(beta, m1));                                // f64 product, f32 store (:167)
            float r2 = __double2float_rn(__dmul_rn(beta, m2));                                // (:168)
            if (isinf(r1)) r1 = 0.0f;                                                         // :169
            if (isinf(r2)) r2 = 0.0f;
            const uint32_t P = pr.par ^ ((syn[i >> 5] >> (i & 31)) & 1u);                     // sign product x syndrome sign (:151,:159)
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const int k = h * SPL + s;
                if (k < DC) {
                    const int pos = k * m + i;
                    if (REGULAR || var_tab[pos] != kPad) {
                        const float mag = (k == pr.k1) ? r2 : r1;
                        *reinterpret_cast<float *>(c2v_b + 4u * pos) = (((sb >> s) & 1u) ^ P) ? -mag : mag;
                    }
                }
            }
        }
    }
}

// DV: number of summed terms per variable (>= max column weight); the offset table has DVS = 4, 8 or 16 entries
// per variable so that one row is one or two vector loads.
template <int DV>
struct VnRow {
    static constexpr int DVS = DV <= 4 ? 4 : (DV <= 8 ? 8 : 16);
};

template <int DV>
__device__ __forceinline__ float ms_colsum(const unsigned char *c2v_b, const uint16_t *vrow)
{
    constexpr int DVS = VnRow<DV>::DVS;
    uint32_t w[DVS / 2];
    if (DVS == 4) {
        const uint2 a = *reinterpret_cast<const uint2 *>(vrow);
        w[0] = a.x; w[1] = a.y;
    } else {
#pragma unroll
        for (int x = 0; x < DVS / 8; ++x) {
            const uint4 a = *reinterpret_cast<const uint4 *>(vrow + 8 * x);
            w[4 * x + 0] = a.x; w[4 * x + 1] = a.y; w[4 * x + 2] = a.z; w[4 * x + 3] = a.w;
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int x = 0; x < DV; ++x) {
        const uint32_t off = (x & 1) ? (w[x >> 1] >> 16) : (w[x >> 1] & 0xffffu);
        s = __fadd_rn(s, lds_f32(c2v_b, off));                       // sequential f32, ascending check (:172)
    }
    return s;
}

template <int DC, bool REGULAR, int DV>
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
    const uint16_t *var_tab = tab + t.off_var;          // byte offsets into S
    const uint16_t *col_ptr = tab + t.off_col_ptr;
    const uint16_t *col_chk = tab + t.off_col_chk;
    const uint16_t *layer_ptr = tab + t.off_layer_ptr;
    const uint16_t *layer_chk = tab + t.off_layer_chk;
    const uint16_t *layer_lpc = tab + t.off_layer_lpc;
    const uint16_t *lvar_ptr = tab + t.off_lvar_ptr;
    const uint16_t *lvar_idx = tab + t.off_lvar_idx;
    const uint16_t *vn_tab = tab + t.off_vn;            // [n][DVS] byte offsets into c2v
    constexpr int DVS = VnRow<DV>::DVS;

    const MsSmemLayout lay = ms_layout(t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem + ((t.len * 2 + 15) & ~15) + (size_t)warp * lay.bytes;
    unsigned char *c2v_b = base + lay.off_c2v;
    unsigned char *S_b = base + lay.off_S;
    float *c2v = reinterpret_cast<float *>(c2v_b);
    float *S = reinterpret_cast<float *>(S_b);   // [n + 1]: entry n is the dummy variable of the padded lists
    uint32_t *par = reinterpret_cast<uint32_t *>(base + lay.off_par);
    uint32_t *syn = reinterpret_cast<uint32_t *>(base + lay.off_syn);
    const int m = t.m, n = t.n;
    const unsigned full = 0xffffffffu;
    const float Tf = c.Tf;
    const bool init_bit = 0.0f < Tf;                    // decision of a variable whose sum is still 0 (only if L < 0)

    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;

        // ---- initial state: c2v = 0 (decoders.py:150), S = 0, residual = syndrome (+ H.1 if the all-zero sums decide 1)
        {
            float4 *z = reinterpret_cast<float4 *>(c2v);
            const int n4 = (t.dc * m + 4) / 4;
            for (int i = lane; i < n4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = n4 * 4 + lane; i < t.dc * m + 4; i += 32) c2v[i] = 0.0f;
        }
        for (int i = lane; i <= n; i += 32) S[i] = 0.0f;
        int unsat = 0;
        for (int i = lane; i < t.mw; i += 32) {
            const uint32_t w = io.syn[shot * t.mw + i];
            const uint32_t p0 = init_bit ? (w ^ (tab + t.off_rowpar)[2 * i] ^ ((uint32_t)(tab + t.off_rowpar)[2 * i + 1] << 16)) : w;
            syn[i] = w;
            par[i] = p0;
            unsat += __popc(p0);
        }
        unsat = __reduce_add_sync(full, unsat);
        __syncwarp();

        bool converged = false;
        bool first = true;
        int it = 0;
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = 0; l < t.nl; ++l) {
                const double prior = first ? c.Lf : c.L;
                // ---------------- check-node phase (decoders.py:156-169)
                const int qb = layer_ptr[l], qe = layer_ptr[l + 1];
                const int lpc = layer_lpc[l];
                if (lpc == 1) ms_check_phase<DC, REGULAR, 1>(qb, qe, lane, m, layer_chk, var_tab, c2v_b, S_b, syn, prior, c.beta);
                else if (lpc == 2) ms_check_phase<DC, REGULAR, 2>(qb, qe, lane, m, layer_chk, var_tab, c2v_b, S_b, syn, prior, c.beta);
                else if (lpc == 4) ms_check_phase<DC, REGULAR, 4>(qb, qe, lane, m, layer_chk, var_tab, c2v_b, S_b, syn, prior, c.beta);
                else ms_check_phase<DC, REGULAR, 8>(qb, qe, lane, m, layer_chk, var_tab, c2v_b, S_b, syn, prior, c.beta);
                __syncwarp();
                // ---------------- variable-node phase (decoders.py:172-174) on the variables whose sums changed;
                // the very first step visits every variable (the reference recomputes all posteriors, and a
                // variable outside layer 0 has posterior L, which may be negative for p > 1/2).
                // Control flow below is warp-uniform on purpose: every lane runs the same number of trips (the lists are
                // padded to a multiple of 32 with the dummy variable n, whose table row points at the zero slot) and
                // flips are handled cooperatively after a ballot.  A lane-divergent loop here was measured to split
                // the warp into two halves that never reconverged (BSSY/BSYNC membership is per split group).
                const int vb = first ? 0 : lvar_ptr[l], ve = first ? t.n_pad : lvar_ptr[l + 1];
                int delta = 0;
                for (int q = vb + lane; q < ve; q += 32) {
                    const int j = first ? (q < n ? q : n) : lvar_idx[q];
                    const float s = ms_colsum<DV>(c2v_b, vn_tab + j * DVS);
                    const float s_old = S[j];
                    S[j] = s;
                    uint32_t flips = __ballot_sync(full, (s < Tf) != (s_old < Tf));     // hard decision flipped (:173-174)
                    while (flips) {                                                      // rare; one flipped variable per trip
                        const int src = __ffs(flips) - 1;
                        flips &= flips - 1;
                        const int jf = __shfl_sync(full, j, src);
                        const int x = col_ptr[jf] + lane;
                        if (x < col_ptr[jf + 1]) {                                       // lane <-> check of the flipped variable
                            const int ch = col_chk[x];
                            const uint32_t bit = 1u << (ch & 31);
                            const uint32_t old = atomicXor(&par[ch >> 5], bit);
                            delta += (old & bit) ? -1 : 1;
                        }
                    }
                }
                unsat += __reduce_add_sync(full, delta);
                __syncwarp();
                first = false;
                // ---------------- H e == syndrome ?  (decoders.py:175-176)
                if (unsat == 0) { converged = true; break; }
            }
        }
        const int iters = it;   // the outer ++it has already run after a converging break: it+1 of decoders.py:176, else max_iter (:182)
        // ---- outputs: e_j = (S_j < Tf)
        for (int w = 0; w < t.nw; ++w) {
            const int j = w * 32 + lane;
            const uint32_t bits = __ballot_sync(full, j < n && (c.max_iter > 0 ? S[j] < Tf : false));
            if (lane == 0) io.ehat[shot * t.nw + w] = bits;
        }
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
