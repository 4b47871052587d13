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
//     rare) and adds +-1 per toggled bit;
//   * control flow is kept warp-uniform everywhere (padded lists, selects instead of branches, cooperative flip
//     handling): a lane-divergent loop was measured to split warps into halves that never reconverged.
// Shared memory is addressed with explicit 32-bit shared-window addresses (ld.shared / st.shared) so that the
// hot loops carry one integer add per access and no generic-pointer arithmetic.
// No fused multiply-add may be formed in this file (compile with -fmad=false); every operation that the spec
// rounds individually uses an explicit _rn intrinsic anyway.
#pragma once
#include "common.cuh"

namespace qldpc {

constexpr int kMsMaxWarps = 24;   // 768 threads per CTA -> up to 85 registers per thread

struct MsSmemLayout {
    // per-shot state, offsets in bytes from the warp's base
    int off_c2v;   // float [dc*ms + 4]  (ms = padded slot stride; entry dc*ms is the always-zero slot)
    int off_S;     // float [n + 1] (entry n: dummy variable of the padded lists)
    int off_par;   // uint32 [mw]
    int off_syn;   // uint32 [mw]
    int bytes;     // multiple of 16
};

__host__ __device__ inline MsSmemLayout ms_layout(const Tables &t)
{
    MsSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 4 * (t.dc * t.ms + 4);
    o = (o + 15) & ~15;
    l.off_S = o;   o += 4 * (t.n + 1);
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    l.bytes = (o + 15) & ~15;
    return l;
}

// ---- explicit shared-window accessors (addresses are 32-bit shared addresses)
__device__ __forceinline__ float sld_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t sld_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t sld_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sst_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sst_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 sld_v2(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 sld_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t satom_xor(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.xor.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

struct MsAddr {          // shared-window byte addresses, warp-uniform
    uint32_t var_tab;    // uint16 [dc*m]   (CTA tables)
    uint32_t layer_chk;  // uint16 [...]
    uint32_t c2v, S, par, syn;   // per-warp state
    uint32_t m2;         // 2*ms : byte stride of one slot row in var_tab
    uint32_t m4;         // 4*ms : byte stride of one slot row in c2v
};

// Check-node phase of one layer with LPC lanes per check (decoders.py:156-169).
//
// The reference takes min1 / min2 of a_k = |v2c_k| in binary64 and stores c2v_k = +-fl32(fl64(beta * (a_k == min1 ? min2 : min1))).
// g(a) = fl32(fl64(beta * a)) is monotone non-decreasing for beta >= 0, so with b_k = g(a_k):  g(min1) = min_k b_k,  and the
// value handed to an edge with a_k == min1, g(min2) = g(min_{k != first argmin} a_k), equals the second smallest b (counted
// with multiplicity): if the minimum of b is attained once, it is attained at the unique argmin of a; if it is attained
// several times, min2_b == min1_b and every edge receives the same magnitude whichever way a_k == min1 falls.  Edges with
// a_k != min1 receive g(min1) = min1_b, and b_k != min1_b implies a_k != min1.  Hence the rule "magnitude = (b_k == min1_b ?
// min2_b : min1_b)" on the ROUNDED per-edge values is bit-identical to the reference, needs no argmin index, and the
// min / second-min / merge logic becomes binary32 FMNMX instead of binary64 compare+select chains.  The sign of b_k is the
// sign of v2c_k (v2c is never -0.0: it is a difference of which the minuend is never -0.0, and rounding keeps signs).
// beta < 0 is handled by the caller passing |beta| and folding the extra sign into `sgn_extra`.
template <int DC, bool REGULAR, int LPC>
__device__ __forceinline__ void ms_check_phase(int qb, int qe, int lane, const MsAddr &A, double prior, double beta, uint32_t sgn_extra)
{
    constexpr int SPL = (DC + LPC - 1) / LPC;       // slots per lane
    constexpr int CPP = 32 / LPC;                   // checks per pass
    constexpr bool EXACT = (LPC * SPL == DC);       // no lane owns a slot >= DC
    const int h = lane % LPC;                       // which slice of the row
    const int k0 = h * SPL;                         // first slot of the lane
    const float inf = __int_as_float(0x7f800000);
    for (int q0 = qb; q0 < qe; q0 += CPP) {
        const int q = q0 + lane / LPC;
        const bool act = q < qe;
        const uint32_t i = sld_u16(A.layer_chk + 2u * (uint32_t)(act ? q : qb));
        const uint32_t vt = A.var_tab + (uint32_t)k0 * A.m2 + 2u * i;    // &var_tab[k0*ms + i]
        const uint32_t cv = A.c2v + (uint32_t)k0 * A.m4 + 4u * i;        // &c2v[k0*ms + i]
        float bs[SPL];                              // signed b_k of the lane's own slots
        float m1 = inf, m2 = inf;                   // smallest / second smallest |b| (inf if none)
        uint32_t px = 0;                            // xor of the b_k bit patterns: bit 31 = parity of the negative signs
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            bs[s] = 0.0f;
            if (EXACT || k0 + s < DC) {
                const uint32_t joff = sld_u16(vt + (uint32_t)s * A.m2);       // byte offset of S_j, kPad past a short row
                if (REGULAR || joff != kPad) {
                    const double post = __dadd_rn(prior, (double)sld_f32(A.S + joff));                   // :173
                    const double v = __dsub_rn(post, (double)sld_f32(cv + (uint32_t)s * A.m4));          // :177
                    const float b = __double2float_rn(__dmul_rn(beta, v));                               // :167-168 (f64 product, f32 store)
                    bs[s] = b;
                    px ^= __float_as_uint(b);                                                             // :157-159
                    const float ab = fabsf(b);
                    m2 = fminf(m2, fmaxf(m1, ab));                                                        // :162-164
                    m1 = fminf(m1, ab);                                                                   // :160
                }
            }
        }
        // butterfly over the LPC lanes of the check
#pragma unroll
        for (int d = 1; d < LPC; d <<= 1) {
            const float o1 = __shfl_xor_sync(0xffffffffu, m1, d);
            const float o2 = __shfl_xor_sync(0xffffffffu, m2, d);
            px ^= __shfl_xor_sync(0xffffffffu, px, d);
            m2 = fminf(fmaxf(m1, o1), fminf(m2, o2));
            m1 = fminf(m1, o1);
        }
        if (act) {
            // inf -> 0: an empty / single-edge row (:165-166) or an overflowed binary32 message (:169)
            const float r1 = (m1 == inf) ? 0.0f : m1;
            const float r2 = (m2 == inf) ? 0.0f : m2;
            const uint32_t synbit = (sld_u32(A.syn + 4u * (i >> 5)) >> (i & 31u)) & 1u;
            const uint32_t P = (px ^ (synbit << 31) ^ sgn_extra) & 0x80000000u;               // sign product x syndrome sign (:151,:159)
            const uint32_t r1s = __float_as_uint(r1) | P, r2s = __float_as_uint(r2) | P;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                if (EXACT || k0 + s < DC) {
                    if (REGULAR || sld_u16(vt + (uint32_t)s * A.m2) != kPad) {
                        const uint32_t mag = (fabsf(bs[s]) == m1) ? r2s : r1s;
                        sst_u32(cv + (uint32_t)s * A.m4, mag ^ (__float_as_uint(bs[s]) & 0x80000000u));
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

// S_j = sequential binary32 sum of the c2v of variable j in ascending check order (decoders.py:172).  The first term
// is taken as is (0.0f + c only differs from c for c == -0.0f, and the sign of a zero sum is never observable).
template <int DV>
__device__ __forceinline__ float ms_colsum(uint32_t c2v, uint32_t vrow)
{
    constexpr int DVS = VnRow<DV>::DVS;
    uint32_t w[DVS / 2];
    if (DVS == 4) {
        const uint2 a = sld_v2(vrow);
        w[0] = a.x; w[1] = a.y;
    } else {
#pragma unroll
        for (int x = 0; x < DVS / 8; ++x) {
            const uint4 a = sld_v4(vrow + 16u * x);
            w[4 * x + 0] = a.x; w[4 * x + 1] = a.y; w[4 * x + 2] = a.z; w[4 * x + 3] = a.w;
        }
    }
    float term[DV];
#pragma unroll
    for (int x = 0; x < DV; ++x) {
        const uint32_t off = (x & 1) ? (w[x >> 1] >> 16) : (w[x >> 1] & 0xffffu);
        term[x] = sld_f32(c2v + off);
    }
    float s = term[0];
#pragma unroll
    for (int x = 1; x < DV; ++x) s = __fadd_rn(s, term[x]);
    return s;
}

// Cooperative parity update for the variables whose hard decision flipped (rare; one flipped variable per trip).
__device__ __forceinline__ void ms_apply_flips(uint32_t flips, uint32_t j, int lane, const MsAddr &A, uint32_t col_ptr, uint32_t col_chk, int &delta)
{
    while (flips) {
        const int src = __ffs(flips) - 1;
        flips &= flips - 1;
        const uint32_t jf = __shfl_sync(0xffffffffu, j, src);
        const uint32_t x0 = sld_u16(col_ptr + 2u * jf), x1 = sld_u16(col_ptr + 2u * jf + 2u);
        const uint32_t x = x0 + lane;
        if (x < x1) {                                                            // lane <-> check of the flipped variable
            const uint32_t ch = sld_u16(col_chk + 2u * x);
            const uint32_t bit = 1u << (ch & 31u);
            const uint32_t old = satom_xor(A.par + 4u * (ch >> 5), bit);
            delta += (old & bit) ? -1 : 1;
        }
    }
}

// One variable-node pass over TWO variables per lane (two independent load / add chains in flight): new sums, stores,
// flip detection.  Lists are padded to a multiple of 64 with the dummy variable n.
template <int DV>
__device__ __forceinline__ void ms_var_update2(uint32_t ja, uint32_t jb, int lane, const MsAddr &A, uint32_t vn_tab, uint32_t col_ptr,
                                               uint32_t col_chk, float Tf, int &delta)
{
    constexpr int DVS = VnRow<DV>::DVS;
    const uint32_t sa = A.S + 4u * ja, sb = A.S + 4u * jb;
    const float a_old = sld_f32(sa), b_old = sld_f32(sb);
    const float a = ms_colsum<DV>(A.c2v, vn_tab + (uint32_t)(2 * DVS) * ja);
    const float b = ms_colsum<DV>(A.c2v, vn_tab + (uint32_t)(2 * DVS) * jb);
    sst_f32(sa, a);
    sst_f32(sb, b);
    const uint32_t fa = __ballot_sync(0xffffffffu, (a < Tf) != (a_old < Tf));   // hard decision flipped (:173-174)
    const uint32_t fb = __ballot_sync(0xffffffffu, (b < Tf) != (b_old < Tf));
    if (fa | fb) {
        ms_apply_flips(fa, ja, lane, A, col_ptr, col_chk, delta);
        ms_apply_flips(fb, jb, lane, A, col_ptr, col_chk, delta);
    }
}

template <int DC, bool REGULAR, int DV>
__global__ void __launch_bounds__(kMsMaxWarps * 32, 1) ms_decode_kernel(Tables t, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io)
{
    extern __shared__ __align__(16) unsigned char smem[];
    {   // graph tables -> shared memory (once per CTA), 16 B per thread per trip
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < t.len / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(full, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform
    const MsSmemLayout lay = ms_layout(t);
    uint32_t tab;                                                         // shared-window address of the CTA tables
    // opaque to the optimiser on purpose: a plain __cvta_generic_to_shared gets rematerialised (S2UR + ULEA) in every loop
    asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(tab) : "l"(smem));
    const uint32_t wbase = tab + (uint32_t)((t.len * 2 + 15) & ~15) + (uint32_t)warp * (uint32_t)lay.bytes;
    MsAddr A;
    A.var_tab = tab + 2u * t.off_var;
    A.layer_chk = tab + 2u * t.off_layer_chk;
    A.c2v = wbase + lay.off_c2v;
    A.S = wbase + lay.off_S;
    A.par = wbase + lay.off_par;
    A.syn = wbase + lay.off_syn;
    A.m2 = 2u * t.ms;
    A.m4 = 4u * t.ms;
    const uint32_t col_ptr = tab + 2u * t.off_col_ptr, col_chk = tab + 2u * t.off_col_chk;
    const uint32_t layer_ptr = tab + 2u * t.off_layer_ptr, layer_lpc = tab + 2u * t.off_layer_lpc;
    const uint32_t lvar_ptr = tab + 2u * t.off_lvar_ptr, lvar_idx = tab + 2u * t.off_lvar_idx;
    const uint32_t vn_tab = tab + 2u * t.off_vn, rowpar = tab + 2u * t.off_rowpar;
    const int n = t.n;
    const float Tf = c.Tf;
    const bool init_bit = 0.0f < Tf;                    // decision of a variable whose sum is still 0 (only if L < 0)
    const int c2v_words = t.dc * t.ms + 4;

    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;

        // ---- initial state: c2v = 0 (decoders.py:150), S = 0, residual = syndrome (+ H.1 if the all-zero sums decide 1)
        for (int i = lane * 4; i < c2v_words; i += 128)
            asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" :: "r"(A.c2v + 4u * i), "f"(0.0f) : "memory");
        for (int i = lane; i <= n; i += 32) sst_f32(A.S + 4u * i, 0.0f);
        int unsat = 0;
        for (int i = lane; i < t.mw; i += 32) {
            const uint32_t w = io.syn[shot * t.mw + i];
            const uint32_t p0 = init_bit ? (w ^ sld_u32(rowpar + 4u * i)) : w;
            sst_u32(A.syn + 4u * i, w);
            sst_u32(A.par + 4u * i, p0);
            unsat += __popc(p0);
        }
        unsat = __reduce_add_sync(full, unsat);
        __syncwarp();

        bool converged = false;
        int it = 0;
        if (c.max_iter > 0) {
            // ---------------- first layer step: prior is the binary32-rounded L (decoders.py:148-149) and the variable
            // phase visits EVERY variable (the reference recomputes all posteriors; a variable outside layer 0 has
            // posterior L, which may be negative for p > 1/2)
            const int qb = sld_u16(layer_ptr), qe = sld_u16(layer_ptr + 2u);
            ms_check_phase<DC, REGULAR, 1>(qb, qe, lane, A, c.Lf, c.abeta, c.sgn);
            __syncwarp();
            int delta = 0;
            for (int q = lane; q < t.n_pad; q += 64)
                ms_var_update2<DV>((uint32_t)(q < n ? q : n), (uint32_t)(q + 32 < n ? q + 32 : n), lane, A, vn_tab, col_ptr, col_chk, Tf, delta);
            unsat += __reduce_add_sync(full, delta);
            __syncwarp();
            converged = unsat == 0;
        }
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = (it == 0 ? 1 : 0); l < t.nl; ++l) {
                // ---------------- check-node phase (decoders.py:156-169)
                const int qb = sld_u16(layer_ptr + 2u * l), qe = sld_u16(layer_ptr + 2u * l + 2u);
                const int lpc = sld_u16(layer_lpc + 2u * l);
                if (lpc == 1) ms_check_phase<DC, REGULAR, 1>(qb, qe, lane, A, c.L, c.abeta, c.sgn);
                else if (lpc == 2) ms_check_phase<DC, REGULAR, 2>(qb, qe, lane, A, c.L, c.abeta, c.sgn);
                else if (lpc == 4) ms_check_phase<DC, REGULAR, 4>(qb, qe, lane, A, c.L, c.abeta, c.sgn);
                else ms_check_phase<DC, REGULAR, 8>(qb, qe, lane, A, c.L, c.abeta, c.sgn);
                __syncwarp();
                // ---------------- variable-node phase (decoders.py:172-174) on the variables whose sums changed.  Every
                // lane runs the same number of trips (lists are padded to a multiple of 64 with the dummy variable n).
                const int vb = sld_u16(lvar_ptr + 2u * l), ve = sld_u16(lvar_ptr + 2u * l + 2u);
                int delta = 0;
                for (int q = vb + lane; q < ve; q += 64)
                    ms_var_update2<DV>(sld_u16(lvar_idx + 2u * q), sld_u16(lvar_idx + 2u * q + 64u), lane, A, vn_tab, col_ptr, col_chk, Tf, delta);
                unsat += __reduce_add_sync(full, delta);
                __syncwarp();
                // ---------------- H e == syndrome ?  (decoders.py:175-176)
                if (unsat == 0) { converged = true; break; }
            }
        }
        // iterations: it+1 of decoders.py:176 on a converging break (the outer ++it has already run; a convergence in the
        // very first step leaves it == 0), else max_iter (:182)
        const int iters = (converged && it == 0) ? 1 : it;
        // ---- outputs: e_j = (S_j < Tf)
        for (int w = 0; w < t.nw; ++w) {
            const int j = w * 32 + lane;
            const uint32_t bits = __ballot_sync(full, j < n && c.max_iter > 0 && sld_f32(A.S + 4u * (uint32_t)(j < n ? j : n)) < Tf);
            if (lane == 0) io.ehat[shot * t.nw + w] = bits;
        }
        if (lane == 0) {
            io.iters[shot] = iters;
            if (io.conv) io.conv[shot] = converged ? 1 : 0;
        }
        if (io.llr) {
            double *dst = io.llr + shot * (long long)n;
            for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)sld_f32(A.S + 4u * j));
        }
        if (!converged && io.fail_count) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(io.fail_count, 1);
            slot = __shfl_sync(full, slot, 0);
            if (slot < io.fail_cap) {
                if (lane == 0) io.fail_shot[slot] = (int)shot;
                double *dst = io.fail_llr + (long long)slot * n;
                for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)sld_f32(A.S + 4u * j));
            }
        }
        __syncwarp();
    }
}

}  // namespace qldpc
