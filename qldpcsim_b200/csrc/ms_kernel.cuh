// ms_kernel.cuh -- normalised min-sum syndrome decoder, flooding / layered / serial, one persistent launch.
//
// Semantics: decoders.py:110-182 of the reference (bit-level spec in SURVEY.md App. A.1, restated on the CPU
// in oracle/qldpc_oracle.c:ms_decode_one).  Mapping:
//   * one WARP -- or, for codes whose shot state leaves room for only ~10 shots per SM, a TEAM of two warps -- owns one
//     shot from its first layer step to its exit and then pulls the next shot from a global dispenser (shots converge
//     after 1..max_iter iterations, so there are no lock-step batches);
//   * the whole message state of the shot lives in shared memory: c2v as binary32 per edge, the binary32
//     column sums S_j, the residual syndrome H e + s as bit words;
//   * VARIABLE-MAJOR message layout.  Variables are renumbered by descending column weight (stable, so the
//     circulant blocks of a lifted code stay contiguous); the x-th check-to-variable message of variable j'
//     (ascending check order, the order of the reference's np.sum, decoders.py:172) lives at word
//     coff[x] + j' of the c2v array ("region x" holds the cnt[x] variables of degree > x).  The variable
//     phase therefore needs NO per-variable offset table -- its addresses are region base + 4*j' -- and the
//     bank of every access of a lane is j' mod 32: a group of lanes whose j' are distinct mod 32 is
//     conflict-free in all of its loads and stores (the plan builder forms the per-layer variable groups
//     that way).  The check phase reaches an edge's S_j' and c2v word through one packed 32-bit table entry
//     (two 16-bit byte offsets);
//   * v2c is never stored: v2c_e = fl64(fl64(prior + S_j) - c2v_e) is rebuilt from S_j and c2v_e, which is
//     exactly what decoders.py:173,:177 computes (prior = binary32-rounded L during the very first layer
//     step, decoders.py:148-149, L afterwards);
//   * check phase: LPC = 1, 2, 4 or 8 lanes share one check (chosen per layer so that the layer fills the
//     warp); min / second min are taken on the ROUNDED per-edge values in binary32 (see ms_check_phase); rows shorter
//     than the instantiated row weight are filled with padding edges whose posterior is +inf (no predicates);
//   * variable phase: lane <-> two / four variables adjacent to the layer per trip, re-summing ALL their c2v in
//     ascending check order in binary32 (decoders.py:172).  The plan lists a layer's LOW-DEGREE variables first (degree <=
//     DMIN, 62 % of the edges of the lifted-product codes: column weights 3 and 5) in fixed positions of every quad trip, where
//     they are summed with DMIN loads and no degree guards instead of DV loads and DV - DMIN selects (ms_colsum4).  Layers need not be column-disjoint
//     (simulator.py:230-234 hands the decoder the partition of the OTHER matrix), so the two phases are
//     separated by a warp barrier;
//   * hard decision: fl64(L + S_j) < 0  <=>  S_j < Tf with Tf = L negated and rounded UP to binary32 (the
//     double sum of two doubles is exact whenever it is tiny, so its sign is the sign of the real sum), which
//     needs one binary32 compare and no stored decision bits: the previous decision is S_j(old) < Tf;
//   * convergence is tested after every layer step (decoders.py:175-176) on an incrementally maintained count
//     of unsatisfied checks: a flipped decision toggles the parity bits of its checks (shared-memory atomics)
//     and adds +-1 per toggled bit.  Flips are NOT rare -- the shots that never converge oscillate through all
//     max_iter iterations and own a sixth of all steps on LP118_0 at p = 0.05: more than half of all quad trips
//     see one -- so all flipped variables of a quad trip are handled in one pass (ms_apply_flips_compact);
//   * the first layer step of a shot is an ordinary step whose prior is the binary32-rounded L: variables outside
//     layer 0 have no message yet, their sum stays 0 and their decision is the one the initial residual assumes;
//   * MERGED STEPS (SPEC instances).  Consecutive layers whose variable sets are pairwise disjoint -- the single-check layers of
//     the serial schedule inside one circulant block row, simulator.py:228-236 'S' -- read nothing that an earlier layer of the
//     run writes, so their check phases and variable phases are executed as ONE step (the plan builder forms the runs).  What
//     remains sequential is the reference's convergence test after every layer (decoders.py:175-176): the step must stop after
//     the first sub-layer whose updates satisfy all checks, with the later sub-layers' posteriors untouched.  The variable phase
//     of a merged step therefore keeps the new sums in registers, counts the flipped decisions F, and
//       - if unsat > dv_max * F no prefix of the run can reach zero unsatisfied checks: everything is committed at once;
//       - otherwise (a few steps per decode, near convergence) the flips are committed sub-layer by sub-layer in order, testing
//         the count after each, and on convergence only the sums of the sub-layers up to that one are stored.
//     Results are bit-identical to running the layers one by one; LP118_2 serial costs 15 steps per iteration instead of 450;
//   * control flow is kept warp-uniform everywhere (padded lists, selects instead of branches, cooperative flip
//     handling): a lane-divergent loop was measured to split warps into halves that never reconverged.
// Shared memory is addressed with explicit 32-bit shared-window addresses (ld.shared / st.shared) so that the
// hot loops carry one integer add per access and no generic-pointer arithmetic.
// No fused multiply-add may be formed in this file (compile with -fmad=false); every operation that the spec
// rounds individually uses an explicit _rn intrinsic anyway.
#pragma once
#include "common.cuh"

namespace qldpc {

// warps per CTA: 24 (768 threads, up to 85 registers per thread); the row-weight classes 4 and 8 are also instantiated for 32
// warps (64 registers, which they fit) and use that instance when the shot state is small enough for more than 24 shots per SM
constexpr int kMsWarps = 24, kMsWarpsBig = 32;
constexpr int kMsMaxDv = 16;

// Device view of the min-sum tables.  One uint16 blob per plan, copied to shared memory once per CTA; offsets are
// in uint16 units (32-bit tables start on even offsets).  j' = renumbered variable (descending column weight).
struct MsTables {
    int m, n, E;
    int dc;              // instantiated row weight (>= max row weight)
    int dv;              // max column weight
    int nl;              // layers
    int mw, nw;          // words(m), words(n)
    int ms;              // row stride of chk (>= m; ms = 4 mod 8 keeps the lane groups of a split check on disjoint banks)
    int n_pad;           // n rounded up to a multiple of 64 (no longer read by the kernels: the first layer step is an ordinary step)
    int c2v_words;       // words of the per-shot c2v array (regions start on multiples of 32 words -- bank 0 -- or, packed, of 16)
    int off_chk;         // u32 [dc*ms]   lo16 = 4*j' (byte offset of S_j'), hi16 = byte offset of the edge's c2v word.  Slots past the
                         //               end of a short row hold a PADDING EDGE: S entry n+1 (always +inf) and the scratch word S[n+2]
    int off_layer;       // u16 [nl][8]   16-byte record per step (a layer, or a merged run of layers): {qb, qe (range in layer_chk),
                         //               lanes per check (1, 2, 4 or 8), vb, ve (range in lvar, 32-bit entries, multiples of 32), 1 if
                         //               the second sub-group of the step's LAST pair-trip is empty (a single-variable trip is run
                         //               instead), number of sub-layers of a merged step (0: plain layer), edges of the step}
    int off_layer_chk;   // u16 [...]     check indices, layer by layer
    int off_lvar;        // u32 [...]     lo16 = 4*j'_a, hi16 = 4*j'_b: the two variables of (trip, lane); dummy = 4*n.  With
                         //               0 < DMIN < DV the FIRST entry of a quad trip (two consecutive entries) holds variables of degree <= DMIN only
    int off_lsub;        // u16 [...]     parallel to lvar: lo8 / hi8 = sub-layer (within its merged step) of variable a / b
    int off_col_chk;     // u16 [n+1][DV] checks of j' (flip handling), fixed stride DV = instantiated column weight, 0xFFFF past the
                         //               degree (variable n: all 0xFFFF)
    int off_rowpar;      // u32 [mw]      parity of the row weights as bit words
    int off_unperm;      // u16 [32*nw]   4*j' of original variable j (4*n past the end)
    int len;             // blob length in uint16 units (multiple of 8)
    int team;            // warps per shot (multi-warp teams get 16 bytes of flip-list scratch per warp behind the team box)
    int packed;          // 1: regions rounded to 16 words and shot states 16 bytes apart (ms_plan.h: packed), 0: 32 words / 128 bytes
    int cnt4[kMsMaxDv];  // 4 * number of variables of degree > x
    int coff4[kMsMaxDv]; // byte offset of region x in the c2v array
};

struct MsSmemLayout {
    // per-shot state, offsets in bytes from the warp's base; c2v and S are contiguous (one zero fill)
    int off_c2v;   // float [c2v_words]
    int off_S;     // float [n + 3] rounded up to 4 words: entry n = dummy variable of the padded lists, entry n+1 = +inf (the
                   // "posterior" of a padding edge: its |b| is +inf, so it never wins a minimum and its sign is +), entry n+2 =
                   // scratch c2v word of the padding edges
    int off_par;   // uint32 [mw]
    int off_syn;   // uint32 [mw]
    int off_team;  // 16 bytes: shot mailbox (int64; bytes 0..3 double as the sub-layer mask of a merged step) + unsatisfied-check
                   // count (int32) + flip count of a merged step (int32) of a multi-warp team
    int zero_words;
    int bytes;     // multiple of 128: S_j' and every c2v word of j' sit on bank j' mod 32
    int bytes16;   // the same rounded to 16 bytes only (ms_sub_kernel: its interleaved layout has no use for the 128-byte alignment)
};

__host__ __device__ inline MsSmemLayout ms_layout(const MsTables &t)
{
    MsSmemLayout l;
    int o = 0;
    l.off_c2v = o; o += 4 * t.c2v_words;
    l.off_S = o;   o += 4 * ((t.n + 3 + 3) & ~3);
    l.zero_words = o / 4;
    l.off_par = o; o += 4 * t.mw;
    l.off_syn = o; o += 4 * t.mw;
    o = (o + 7) & ~7;
    l.off_team = o; o += 16;
    if (t.team > 1) o += 16 * t.team;     // per-warp flip lists (a one-warp team uses its otherwise idle team box)
    l.bytes = (o + 127) & ~127;
    l.bytes16 = (o + 15) & ~15;
    return l;
}

__host__ __device__ inline int ms_table_bytes(const MsTables &t) { return (t.len * 2 + 127) & ~127; }

// ---- explicit shared-window accessors (addresses are 32-bit shared addresses)
__device__ __forceinline__ float sld_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t sld_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t sld_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sst_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sst_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sst_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sred_xor(uint32_t a, uint32_t v) { asm volatile("red.shared.xor.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t satom_xor(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.xor.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

struct MsAddr {          // shared-window byte addresses, warp-uniform
    uint32_t chk;        // u32 [dc*ms]   (CTA tables)
    uint32_t layer_chk;  // u16 [...]
    uint32_t col_chk;
    uint32_t c2v, S, par, syn;   // per-warp state
    uint32_t m4;         // 4*ms : byte stride of one slot row in chk
    uint32_t list;       // 16 bytes of per-warp scratch: the flipped variables of a quad trip
};

// Check-node phase of one layer with LPC lanes per check (decoders.py:156-169).
//
// The reference takes min1 / min2 of a_k = |v2c_k| in binary64 and stores c2v_k = +-fl32(fl64(beta * (a_k == min1 ? min2 : min1))).
// g(a) = fl32(fl64(beta * a)) is monotone non-decreasing for beta >= 0, so with b_k = g(a_k):  g(min1) = min_k b_k,  and the
// value handed to an edge with a_k == min1, g(min2) = g(min_{k != first argmin} a_k), equals the second smallest b (counted
// with multiplicity): if the minimum of b is attained once, it is attained at the unique argmin of a; if it is attained
// several times, min2_b == min1_b and every edge receives the same magnitude whichever way a_k == min1 falls.  Edges with
// a_k != min1 receive g(min1) = min1_b, and b_k != min1_b implies a_k != min1.  Hence the rule "magnitude = (b_k == min1_b ?
// min2_b : min1_b)" on the ROUNDED per-edge values is bit-identical to the reference, needs no argmin index (so the order of
// the edges within a check is free), and the min / second-min / merge logic is binary32 FMNMX instead of binary64
// compare+select chains.  The sign of b_k is the sign of v2c_k (v2c is never -0.0: it is a difference whose minuend is never
// -0.0, and rounding keeps signs).  beta < 0: the caller passes |beta| and folds the extra sign into `sgn_extra`.
template <int DC, int LPC>
__device__ __forceinline__ void ms_check_phase(int qb, int qe, int sub, int team_warps, int lane, const MsAddr &A, double prior, double beta,
                                               uint32_t sgn_extra)
{
    static_assert(DC % LPC == 0, "row-weight classes are multiples of the lane split");
    constexpr int SPL = DC / LPC;                   // slots per lane
    constexpr int CPP = 32 / LPC;                   // checks per pass
    const int h = lane % LPC;                       // which slice of the row
    const int k0 = h * SPL;                         // first slot of the lane
    const float inf = __int_as_float(0x7f800000);
    for (int q0 = qb + sub * CPP; q0 < qe; q0 += team_warps * CPP) {   // the passes of a layer go round-robin to the warps of the team
        const int q = q0 + lane / LPC;
        const bool act = q < qe;
        const uint32_t i = sld_u16(A.layer_chk + 2u * (uint32_t)(act ? q : qb));
        const uint32_t ct = A.chk + (uint32_t)k0 * A.m4 + 4u * i;        // &chk[k0*ms + i]
        float bs[SPL];                              // signed b_k of the lane's own slots
        uint32_t ca[SPL];                           // shared address of the slot's c2v word
        float m1 = inf, m2 = inf;                   // smallest / second smallest |b| (inf if none)
        uint32_t px = 0;                            // xor of the b_k bit patterns: bit 31 = parity of the negative signs
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const uint32_t e = sld_u32(ct + (uint32_t)s * A.m4);
            ca[s] = A.c2v + (e >> 16);
            const double post = __dadd_rn(prior, (double)sld_f32(A.S + (e & 0xffffu)));          // :173
            const double v = __dsub_rn(post, (double)sld_f32(ca[s]));                            // :177
            const float b = __double2float_rn(__dmul_rn(beta, v));                               // :167-168 (f64 product, f32 store)
            bs[s] = b;
            px ^= __float_as_uint(b);                                                             // :157-159
            const float ab = fabsf(b);
            m2 = fminf(m2, fmaxf(m1, ab));                                                        // :162-164
            m1 = fminf(m1, ab);                                                                   // :160
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
                const uint32_t mag = (fabsf(bs[s]) == m1) ? r2s : r1s;
                sst_u32(ca[s], mag ^ (__float_as_uint(bs[s]) & 0x80000000u));
            }
        }
    }
}

// S_j' = sequential binary32 sum of the c2v of variable j' in ascending check order (decoders.py:172): term x sits at
// region x + 4*j'.  The first DMIN regions hold every variable (plus a zero word for the dummy variable n).  The others
// are read unconditionally as well -- for j' >= cnt[x] the address falls into a later region or into S, always inside the
// warp's own state -- and the value is discarded by a select (cheaper than a predicated load, whose uniform-register
// operands the compiler guards with votes).  The first term is taken as is (0.0f + c only differs from c for c == -0.0f,
// and the sign of a zero sum is never observable).
template <int DV, int DMIN>
__device__ __forceinline__ float ms_colsum(uint32_t c2vj /* c2v + 4*j' */, uint32_t j4, const MsTables &t)
{
    float term[DV];
#pragma unroll
    for (int x = 0; x < DV; ++x) {
        term[x] = sld_f32(c2vj + (uint32_t)t.coff4[x]);
        if (x >= DMIN) term[x] = ((int)j4 < t.cnt4[x]) ? term[x] : 0.0f;
    }
    float s = term[0];
#pragma unroll
    for (int x = 1; x < DV; ++x) s = __fadd_rn(s, term[x]);
    return s;
}

// New sums of the four variables of a quad trip.  When only some regions hold every variable (0 < DMIN < DV) the plan fills the
// first two sub-groups of every quad with variables of degree <= DMIN (or dummies) only: they are summed with DMIN loads and
// no degree guards -- a static property of the list layout, no per-trip test (ms_plan.h: var_cost).
template <int DV, int DMIN, bool LOW>
__device__ __forceinline__ float ms_colsum_sel(uint32_t j4, const MsAddr &A, const MsTables &t)
{
    if constexpr (LOW && DMIN > 0 && DMIN < DV) return ms_colsum<DMIN, DMIN>(A.c2v + j4, j4, t);
    else return ms_colsum<DV, DMIN>(A.c2v + j4, j4, t);
}

template <int DV, int DMIN>
__device__ __forceinline__ void ms_colsum4(const uint32_t (&j4)[4], const MsAddr &A, const MsTables &t, float (&s)[4])
{
    s[0] = ms_colsum_sel<DV, DMIN, true>(j4[0], A, t);
    s[1] = ms_colsum_sel<DV, DMIN, true>(j4[1], A, t);
    s[2] = ms_colsum_sel<DV, DMIN, false>(j4[2], A, t);
    s[3] = ms_colsum_sel<DV, DMIN, false>(j4[3], A, t);
}

// Cooperative parity update for the variables whose hard decision flipped (rare; one flipped variable per trip).  The checks of
// variable j' sit in a fixed-stride table (DV entries per variable, 0xFFFF past its degree): lane x < DV toggles the x-th one.
template <int DV>
__device__ __forceinline__ void ms_apply_flips(uint32_t flips, uint32_t j4, int lane, const MsAddr &A, int &delta)
{
    while (flips) {
        const int src = __ffs(flips) - 1;
        flips &= flips - 1;
        const uint32_t jf4 = __shfl_sync(0xffffffffu, j4, src);
        if (lane < DV) {                                                         // lane <-> check of the flipped variable
            const uint32_t ch = sld_u16(A.col_chk + (jf4 >> 1) * (uint32_t)DV + 2u * (uint32_t)lane);
            if (ch != 0xffffu) {
                const uint32_t bit = 1u << (ch & 31u);
                const uint32_t old = satom_xor(A.par + 4u * (ch >> 5), bit);
                delta += (old & bit) ? -1 : 1;
            }
        }
    }
}

// One variable-node pass over TWO variables per lane (two independent load / add chains in flight): new sums, stores,
// flip detection.  ja4 / jb4 = 4*j'.
template <int DV, int DMIN>
__device__ __forceinline__ void ms_var_update2(uint32_t ja4, uint32_t jb4, int lane, const MsAddr &A, const MsTables &t, float Tf, int &delta)
{
    const uint32_t sa = A.S + ja4, sb = A.S + jb4;
    const float a_old = sld_f32(sa), b_old = sld_f32(sb);
    const float a = ms_colsum<DV, DMIN>(A.c2v + ja4, ja4, t);
    const float b = ms_colsum<DV, DMIN>(A.c2v + jb4, jb4, t);
    sst_f32(sa, a);
    sst_f32(sb, b);
    const uint32_t fa = __ballot_sync(0xffffffffu, (a < Tf) != (a_old < Tf));   // hard decision flipped (:173-174)
    const uint32_t fb = __ballot_sync(0xffffffffu, (b < Tf) != (b_old < Tf));
    if (fa | fb) {
        ms_apply_flips<DV>(fa, ja4, lane, A, delta);
        ms_apply_flips<DV>(fb, jb4, lane, A, delta);
    }
}

// Single variable per lane: layers that touch at most 32 variables (single-check layers of the serial schedule, bicycle).
template <int DV, int DMIN>
__device__ __forceinline__ void ms_var_update1(uint32_t ja4, int lane, const MsAddr &A, const MsTables &t, float Tf, int &delta)
{
    const uint32_t sa = A.S + ja4;
    const float a_old = sld_f32(sa);
    const float a = ms_colsum<DV, DMIN>(A.c2v + ja4, ja4, t);
    sst_f32(sa, a);
    const uint32_t fa = __ballot_sync(0xffffffffu, (a < Tf) != (a_old < Tf));   // hard decision flipped (:173-174)
    if (fa) ms_apply_flips<DV>(fa, ja4, lane, A, delta);
}

// All flipped variables of a quad trip in ONE pass (one-warp teams).  More than half of all quad trips see a flip (the shots that
// never converge oscillate through all max_iter iterations), 2.3 flipped variables on average, and the cooperative loop above
// spends a shuffle -> load -> atomic round trip and ~25 instructions on each.  Here every flipping lane drops its variable into a
// short list (slot = number of flips before it, from the ballots), and lane 5q + x (DV = 5) toggles the x-th check of the q-th
// listed variable: up to 32 / DV flips (at most 8) per pass, whatever their number; longer lists take the loop.  The list is
// per warp (A.list): a one-warp team uses its idle team box, the warps of a larger team 16 bytes each behind it.
template <int DV>
__device__ __forceinline__ bool ms_apply_flips_compact(const uint32_t (&f)[4], const uint32_t (&j4)[4], int lane, const MsAddr &A, int &delta)
{
    constexpr int CAP = (32 / DV) < 8 ? (32 / DV) : 8;
    const int c0 = __popc(f[0]), c1 = c0 + __popc(f[1]), c2 = c1 + __popc(f[2]), F = c2 + __popc(f[3]);
    if (F > CAP) return false;
    const uint32_t lt = (1u << lane) - 1u;
    if ((f[0] >> lane) & 1u) sst_u16(A.list + 2u * (uint32_t)__popc(f[0] & lt), j4[0]);
    if ((f[1] >> lane) & 1u) sst_u16(A.list + 2u * (uint32_t)(c0 + __popc(f[1] & lt)), j4[1]);
    if ((f[2] >> lane) & 1u) sst_u16(A.list + 2u * (uint32_t)(c1 + __popc(f[2] & lt)), j4[2]);
    if ((f[3] >> lane) & 1u) sst_u16(A.list + 2u * (uint32_t)(c2 + __popc(f[3] & lt)), j4[3]);
    __syncwarp();
    const int q = lane / DV, x = lane - q * DV;
    if (q < F) {
        const uint32_t jf4 = sld_u16(A.list + 2u * (uint32_t)q);
        const uint32_t ch = sld_u16(A.col_chk + (jf4 >> 1) * (uint32_t)DV + 2u * (uint32_t)x);
        if (ch != 0xffffu) {
            const uint32_t bit = 1u << (ch & 31u);
            const uint32_t old = satom_xor(A.par + 4u * (ch >> 5), bit);
            delta += (old & bit) ? -1 : 1;
        }
    }
    __syncwarp();                                                             // the list is rewritten by the next flipping trip
    return true;
}

// Same with FOUR variables per lane (two consecutive pair-trips of the layer's list at once): more independent chains in
// flight and half the loop overhead for the common layers whose variables fill four sub-groups.  The flips of the trip are
// handled in one pass (ms_apply_flips_compact) unless there are more than it takes.
template <int DV, int DMIN>
__device__ __forceinline__ void ms_var_update4(uint32_t e0, uint32_t e1, int lane, const MsAddr &A, const MsTables &t, float Tf, int &delta)
{
    const uint32_t j4[4] = {e0 & 0xffffu, e0 >> 16, e1 & 0xffffu, e1 >> 16};
    float s_old[4], s[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) s_old[v] = sld_f32(A.S + j4[v]);
    ms_colsum4<DV, DMIN>(j4, A, t, s);
    uint32_t f[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        sst_f32(A.S + j4[v], s[v]);
        f[v] = __ballot_sync(0xffffffffu, (s[v] < Tf) != (s_old[v] < Tf));      // hard decision flipped (:173-174)
    }
    if (f[0] | f[1] | f[2] | f[3]) {
        if (ms_apply_flips_compact<DV>(f, j4, lane, A, delta)) return;
#pragma unroll
        for (int v = 0; v < 4; ++v) ms_apply_flips<DV>(f[v], j4[v], lane, A, delta);
    }
}

// W: warps per shot ("team", see below), MAXW: warps per CTA the instance is compiled for (launch bound), DC: instantiated row weight (shorter rows are filled with
// padding edges), DV: instantiated column weight, DMIN: number of
// leading regions that hold every variable (0 = guard all), SPEC: the plan holds merged steps (see the header).
template <int DC, int DV, int DMIN, int MAXW, int W, bool SPEC>
__global__ void __launch_bounds__(MAXW * 32, 1) ms_decode_kernel(MsTables t, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io)
{
    extern __shared__ __align__(128) unsigned char smem[];
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
    // W > 1: a team of W warps shares one shot.  Codes whose shot state is large (LP118_2, Tanner: 19 KB) fit only ~10 shots per
    // SM, and 10 warps cannot hide the latency of their dependent shared-memory / FP64 chains; the team splits the passes of the
    // check phase and the trips of the variable phase and meets at a named barrier between the phases.  Every edge and every
    // variable is still computed by exactly one lane in the same arithmetic order, so the result does not depend on W.
    const int team = warp / W, sub = warp % W;
    constexpr int TT = 32 * W;
    const int tl = sub * 32 + lane;                                       // thread index within the team
    const uint32_t wbase = tab + (uint32_t)ms_table_bytes(t) + (uint32_t)team * (uint32_t)(t.packed ? lay.bytes16 : lay.bytes);
    auto team_sync = [&]() {
        if constexpr (W == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(TT) : "memory");
    };
    const uint32_t team_box = wbase + lay.off_team;                       // [0..7] shot mailbox, [8..11] unsatisfied-check count
    MsAddr A;
    A.chk = tab + 2u * t.off_chk;
    A.layer_chk = tab + 2u * t.off_layer_chk;
    A.col_chk = tab + 2u * t.off_col_chk;
    A.c2v = wbase + lay.off_c2v;
    A.S = wbase + lay.off_S;
    A.par = wbase + lay.off_par;
    A.syn = wbase + lay.off_syn;
    A.m4 = 4u * t.ms;
    A.list = W == 1 ? team_box : team_box + 16u + 16u * (uint32_t)sub;
    const uint32_t layer_rec = tab + 2u * t.off_layer, lvar = tab + 2u * t.off_lvar, lsub = tab + 2u * t.off_lsub;
    const uint32_t rowpar = tab + 2u * t.off_rowpar, unperm = tab + 2u * t.off_unperm;
    const int n = t.n;
    const uint32_t n4 = 4u * (uint32_t)n;
    const float Tf = c.Tf;
    const bool init_bit = 0.0f < Tf;                    // decision of a variable whose sum is still 0 (only if L < 0)

    unsigned long long edges_done = 0;                  // check-to-variable messages computed by this team (executed-work counter)
    for (;;) {
        long long shot = 0;
        if (tl == 0) {
            shot = (long long)atomicAdd(io.work_counter, 1ull);
            if constexpr (W > 1) asm volatile("st.shared.u64 [%0], %1;" :: "r"(team_box), "l"(shot) : "memory");
        }
        if constexpr (W == 1) shot = __shfl_sync(full, shot, 0);
        else {
            team_sync();
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(shot) : "r"(team_box) : "memory");
        }
        if (shot >= io.shots) break;

        // ---- initial state: c2v = 0 (decoders.py:150), S = 0, residual = syndrome (+ H.1 if the all-zero sums decide 1)
        for (int i = tl * 4; i < lay.zero_words; i += 4 * TT)
            asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" :: "r"(A.c2v + 4u * i), "f"(0.0f) : "memory");
        team_sync();
        if (tl == 0) {
            sst_u32(A.S + n4 + 4u, 0x7f800000u);                         // S[n+1] = +inf: the padding edges
            if constexpr (SPEC && W > 1) { sst_u32(team_box, 0u); sst_u32(team_box + 12u, 0u); }   // exchange words of the merged steps
        }
        int unsat = 0;
        if (sub == 0) {
            for (int i = lane; i < t.mw; i += 32) {
                const uint32_t w = io.syn[shot * t.mw + i];
                const uint32_t p0 = init_bit ? (w ^ sld_u32(rowpar + 4u * i)) : w;
                sst_u32(A.syn + 4u * i, w);
                sst_u32(A.par + 4u * i, p0);
                unsat += __popc(p0);
            }
            unsat = __reduce_add_sync(full, unsat);
            if constexpr (W > 1) { if (lane == 0) sst_u32(team_box + 8u, (uint32_t)unsat); }
        }
        team_sync();
        // unsat: a register in every lane for W == 1; for a team the shared word is the truth and `unsat` its copy after a barrier
        auto settle = [&](int delta) {
            delta = __reduce_add_sync(full, delta);
            if constexpr (W == 1) { unsat += delta; __syncwarp(); }
            else {
                if (lane == 0 && delta != 0) asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(team_box + 8u), "r"(delta) : "memory");
                team_sync();
                unsat = (int)sld_u32(team_box + 8u);      // read by every warp before it can reach the next barrier; next written after it
            }
        };

        bool converged = false;
        int it = 0;
        uint32_t shot_edges = 0;
        {
            // The first layer step of a shot is an ordinary step whose prior is the binary32-rounded L (decoders.py:148-149).  The
            // reference recomputes the posterior of EVERY variable after it; a variable outside layer 0 has no message yet, its sum
            // stays 0 and its decision (0 < Tf, i.e. L < 0: p > 1/2) is the one the initial residual already assumes, so only the
            // variables of layer 0 need a visit -- the sweep over all n variables this kernel used to make cost 7 % of its
            // instructions.
        }
        for (; it < c.max_iter && !converged; ++it) {
            for (int l = 0; l < t.nl; ++l) {
                // ---------------- check-node phase (decoders.py:156-169)
                const double prior = (it == 0 && l == 0) ? c.Lf : c.L;
                uint32_t r0, r1, r2, r3;                                       // the layer's 16-byte record
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(layer_rec + 16u * l));
                const int qb = r0 & 0xffffu, qe = r0 >> 16, lpc = r1 & 0xffffu;
                shot_edges += r3 >> 16;
                if (lpc == 1) ms_check_phase<DC, 1>(qb, qe, sub, W, lane, A, prior, c.abeta, c.sgn);
                else if (lpc == 2) ms_check_phase<DC, 2>(qb, qe, sub, W, lane, A, prior, c.abeta, c.sgn);
                else if (lpc == 4) ms_check_phase<DC, 4>(qb, qe, sub, W, lane, A, prior, c.abeta, c.sgn);
                else if constexpr (DC % 8 == 0) ms_check_phase<DC, 8>(qb, qe, sub, W, lane, A, prior, c.abeta, c.sgn);   // DC = 4: lpc <= 4
                team_sync();
                // ---------------- variable-node phase (decoders.py:172-174) on the variables whose sums changed.  Every
                // lane runs the same number of trips (lists are padded to whole trips with the dummy variable n).
                const int vb = r1 >> 16, ve = r2 & 0xffffu;
                const int P = (ve - vb) >> 5;                                  // pair-trips of the step
                if (SPEC && (r3 & 0xffffu)) {
                    // ---------------- merged step: at most one quad trip per warp (the plan builder caps the run at 128 W
                    // variables); sums stay in registers until it is known how far the run may be committed
                    const int p0 = 2 * sub;
                    const bool h0 = p0 < P, h1 = p0 + 1 < P;                   // warp-uniform
                    const uint32_t q0 = (uint32_t)(vb + 32 * p0 + lane);
                    const uint32_t dummy = n4 | (n4 << 16);
                    const uint32_t e0 = h0 ? sld_u32(lvar + 4u * q0) : dummy, e1 = h1 ? sld_u32(lvar + 4u * q0 + 128u) : dummy;
                    const uint32_t j4[4] = {e0 & 0xffffu, e0 >> 16, e1 & 0xffffu, e1 >> 16};
                    float s_old[4], s_new[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) s_old[v] = sld_f32(A.S + j4[v]);
                    if (h1) ms_colsum4<DV, DMIN>(j4, A, t, s_new);             // a whole quad: [low, low, any, any]
                    else {                                                     // a lone pair-trip is generic (the other two are dummies)
                        s_new[0] = ms_colsum<DV, DMIN>(A.c2v + j4[0], j4[0], t);
                        s_new[1] = ms_colsum<DV, DMIN>(A.c2v + j4[1], j4[1], t);
                        s_new[2] = s_old[2]; s_new[3] = s_old[3];
                    }
                    uint32_t f[4];
                    int F = 0;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        f[v] = __ballot_sync(full, (s_new[v] < Tf) != (s_old[v] < Tf));      // hard decision flipped (:173-174)
                        F += __popc(f[v]);
                    }
                    uint32_t ks0 = 0, ks1 = 0, kmask = 0;                      // sub-layers of my variables; sub-layers that hold flips
                    if (F) {
                        ks0 = h0 ? sld_u16(lsub + 2u * q0) : 0u;
                        ks1 = h1 ? sld_u16(lsub + 2u * q0 + 64u) : 0u;
                        const uint32_t kk[4] = {ks0 & 0xffu, ks0 >> 8, ks1 & 0xffu, ks1 >> 8};
                        uint32_t mine = 0;
#pragma unroll
                        for (int v = 0; v < 4; ++v) mine |= ((f[v] >> lane) & 1u) << kk[v];
                        kmask = __reduce_or_sync(full, mine);
                    }
                    int Ft = F;
                    if constexpr (W > 1) {
                        if (lane == 0 && F) {
                            asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(team_box + 12u), "r"(F) : "memory");
                            asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(team_box), "r"(kmask) : "memory");
                        }
                        team_sync();
                        Ft = (int)sld_u32(team_box + 12u);
                        kmask = sld_u32(team_box);
                    }
                    if (unsat > t.dv * Ft) {
                        // no prefix of the run can satisfy every check (a flip toggles at most dv of them): commit everything
#pragma unroll
                        for (int v = 0; v < 4; ++v) sst_f32(A.S + j4[v], s_new[v]);
                        int delta = 0;
                        if (F && !ms_apply_flips_compact<DV>(f, j4, lane, A, delta)) {
#pragma unroll
                            for (int v = 0; v < 4; ++v) ms_apply_flips<DV>(f[v], j4[v], lane, A, delta);
                        }
                        settle(delta);
                    } else {
                        // commit sub-layer by sub-layer, testing after each one (decoders.py:175-176)
                        if (!F) {
                            ks0 = h0 ? sld_u16(lsub + 2u * q0) : 0u;
                            ks1 = h1 ? sld_u16(lsub + 2u * q0 + 64u) : 0u;
                        }
                        const uint32_t kk[4] = {ks0 & 0xffu, ks0 >> 8, ks1 & 0xffu, ks1 >> 8};
                        uint32_t kconv = 64u;
                        while (kmask) {
                            const uint32_t k = (uint32_t)__ffs(kmask) - 1u;
                            kmask &= kmask - 1u;
                            int delta = 0;
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const uint32_t mk = __ballot_sync(full, ((f[v] >> lane) & 1u) && kk[v] == k);
                                ms_apply_flips<DV>(mk, j4[v], lane, A, delta);
                            }
                            settle(delta);
                            if constexpr (W > 1) team_sync();              // every warp has read the count before the next round adds to it
                            if (unsat == 0) { kconv = k; break; }
                        }
#pragma unroll
                        for (int v = 0; v < 4; ++v) if (kk[v] <= kconv) sst_f32(A.S + j4[v], s_new[v]);
                        team_sync();
                    }
                    if constexpr (W > 1) {
                        // the exchange words are next touched after the barrier that follows the next check phase
                        if (tl == 0 && Ft) { sst_u32(team_box, 0u); sst_u32(team_box + 12u, 0u); }
                    }
                    if (unsat == 0) { converged = true; break; }
                    continue;
                }
                int delta = 0;
                for (int p = 2 * sub; p + 1 < P; p += 2 * W) {
                    const int q = vb + 32 * p + lane;
                    ms_var_update4<DV, DMIN>(sld_u32(lvar + 4u * q), sld_u32(lvar + 4u * q + 128u), lane, A, t, Tf, delta);
                }
                if ((P & 1) && ((P >> 1) % W) == sub) {                        // odd pair-trip at the end
                    const uint32_t e = sld_u32(lvar + 4u * (uint32_t)(ve - 32 + lane));
                    if (r2 >> 16) ms_var_update1<DV, DMIN>(e & 0xffffu, lane, A, t, Tf, delta);      // it holds one sub-group only
                    else ms_var_update2<DV, DMIN>(e & 0xffffu, e >> 16, lane, A, t, Tf, delta);
                }
                settle(delta);
                // ---------------- H e == syndrome ?  (decoders.py:175-176)
                if (unsat == 0) { converged = true; break; }
            }
        }
        // iterations: it+1 of decoders.py:176 on a converging break (the outer ++it has already run; a convergence in the
        // very first step leaves it == 0), else max_iter (:182)
        const int iters = (converged && it == 0) ? 1 : it;
        edges_done += shot_edges;
        // ---- outputs: e_j = (S_j' < Tf)
        // (the k-th word a warp forms stays in lane k mod 32; one coalesced store per 32 words instead of one store, with its
        // 64-bit address arithmetic, per word: the loop is 4 % of the kernel's instructions otherwise)
        {
            uint32_t *erow = io.ehat + shot * t.nw;
            uint32_t mine = 0;
            int k = 0;
            for (int w = sub; w < t.nw; w += W, ++k) {
                const int j = w * 32 + lane;
                const uint32_t bits = __ballot_sync(full, j < n && c.max_iter > 0 && sld_f32(A.S + sld_u16(unperm + 2u * j)) < Tf);
                if (lane == (k & 31)) mine = bits;
                if ((k & 31) == 31) {                                       // rows of more than 32 words per warp
                    erow[sub + (k - 31 + lane) * W] = mine;
                    mine = 0;
                }
            }
            const int rest = k & 31, k0 = k - rest;
            if (lane < rest) erow[sub + (k0 + lane) * W] = mine;
        }
        if (tl == 0) {
            io.iters[shot] = iters;
            if (io.conv) io.conv[shot] = converged ? 1 : 0;
        }
        if (io.llr) {
            double *dst = io.llr + shot * (long long)n;
            for (int j = tl; j < n; j += TT) dst[j] = __dadd_rn(c.L, (double)sld_f32(A.S + sld_u16(unperm + 2u * j)));
        }
        if (!converged && io.fail_count) {
            int slot = 0;
            if (tl == 0) {
                slot = atomicAdd(io.fail_count, 1);
                if (slot < io.fail_cap) io.fail_shot[slot] = (int)shot;
                if constexpr (W > 1) sst_u32(team_box, (uint32_t)slot);       // the mailbox is free: every warp has read the shot index
            }
            if constexpr (W == 1) slot = __shfl_sync(full, slot, 0);
            else { team_sync(); slot = (int)sld_u32(team_box); }
            if (slot < io.fail_cap) {
                double *dst = io.fail_llr + (long long)slot * n;
                for (int j = tl; j < n; j += TT) dst[j] = __dadd_rn(c.L, (double)sld_f32(A.S + sld_u16(unperm + 2u * j)));
            }
        }
        team_sync();
    }
    if (io.work_done && tl == 0 && edges_done) atomicAdd(io.work_done, edges_done);
}

}  // namespace qldpc
