// ms_sub_kernel.cuh -- normalised min-sum for schedules whose every layer is ONE check that shares variables with its
// neighbours (bicycle code: 73 single-check layers of 18 edges, any two checks overlap, so nothing can be merged into larger
// steps): EIGHT LANES PER SHOT, FOUR SHOTS PER WARP.
//
// Semantics and arithmetic are those of ms_kernel.cuh (decoders.py:110-182; same tables, same
// variable-major message layout, same rounded-minimum rule, same hard-decision threshold).  What changes is the mapping:
//   * a warp-per-shot step on such a schedule keeps 8 of 32 lanes busy in the check phase and 18 in the variable phase and
//     pays the fixed cost of a step for 18 edges.  Here the four 8-lane groups of a warp decode four different shots with one
//     instruction stream: check phase = DCS/8 slots per lane and a 3-round butterfly inside the group, variable phase = DCS/8
//     trips of 8 variables;
//   * the groups are NOT in lock-step: each has its own shot, layer index and iteration count (per-lane registers, equal
//     within a group), so a group whose shot finishes takes the next shot at once and restarts at layer 0 while the others
//     continue -- every table access is per lane anyway, and the control flow stays warp-uniform because all layers have the
//     same shape;
//   * the four shots of a warp are INTERLEAVED word by word in shared memory (word w of shot g at 16 w + 4 g): lanes of
//     different groups can never collide on a bank, whatever layers they are at; inside a group the eight accesses of an
//     instruction fall on banks 4 (w mod 8) + g;
//   * finishing / refilling is done by the whole warp for one group at a time (32-lane ballots for the estimate words,
//     32-lane zero fill);
//   * flipped hard decisions are NOT rare here: on the bicycle code at p = 0.03 half of all steps belong to the 5 % of shots
//     that never converge and keep oscillating, and 3 of 4 warp steps see at least one flip.  Every lane therefore toggles the
//     residual parities of its own flipped variables in parallel -- with at most four parity words per shot (MW > 0) through
//     one precomputed parity mask per variable, XOR-ed into the shot's parity words by fire-and-forget shared-memory
//     reductions -- and the unsatisfied-check count is then recounted by population count (no serial loop over the flipped
//     variables, no atomics that return a value);
//   * a step is one dependent chain (a warp has nothing else to do and an SM holds only 8 such warps), so it is kept SHORT:
//     lane h owns cell (s, h) of the check in slot step s of the check phase AND the same edge's variable in trip s of the
//     variable phase (ms_plan.h: sub8_deal), hence (i) one record per (layer, lane) -- the SPL packed table entries of the lane's
//     edges -- is all a step needs, and it is fetched one step ahead; (ii) the message a lane writes in the check phase is
//     read back only by itself, so no warp barrier separates the phases; (iii) the posterior read in the check phase is the
//     "old" sum of the variable phase.  All loads of a phase are issued before the first dependent operation.
// Results are bit-identical to the warp-per-shot kernel and to the reference.
#pragma once
#include "ms_kernel.cuh"

namespace qldpc {

constexpr int kMsSubWarps = 16;          // launch bound: 512 threads, 128 registers

// extra table of the sub-warp kernel inside the same blob: off_srec u32 [nl][8][SPL] = the packed entries of the cells (s, h),
// s < SPL, of the check of layer l, PRE-SCALED for the interleaved layout: lo16 = 16*j' (byte offset of S_j' from the shot's S
// array), hi16 = 4 * (byte offset of the edge's c2v word in the single-shot c2v array); padding cell past the row: S entry n+1
// (+inf) and the scratch word S[n+2].  The check of layer l is layer_chk[l].
struct MsSubTables {
    int off_srec;
    int off_colmask;     // u32 [n][MW] (MW = mw <= 4 instances only): bit i of the mask of j' = check i is adjacent to j'
    int c16[kMsMaxDv];   // 16 * word offset of region x: byte offset of the region in the interleaved c2v array
};

template <int SPL>
__device__ __forceinline__ void sub_load_rec(uint32_t a, uint32_t (&e)[SPL])
{
#pragma unroll
    for (int s = 0; s < SPL; ++s) e[s] = sld_u32(a + 4u * (uint32_t)s);
}

template <int DCS, int DV, int DMIN, int MW>
__global__ void __launch_bounds__(kMsSubWarps * 32, 1) ms_sub_kernel(MsTables t, MsSubTables ts, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io)
{
    static_assert(DCS % 8 == 0 && DCS <= 32, "eight lanes per check");
    constexpr int SPL = DCS / 8;          // slots per lane in the check phase = trips of 8 variables in the variable phase
    extern __shared__ __align__(128) unsigned char smem[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < t.len / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(full, (int)(threadIdx.x >> 5), 0);
    const int grp = lane >> 3, h = lane & 7;
    const MsSmemLayout lay = ms_layout(t);
    uint32_t tab;
    asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(tab) : "l"(smem));
    const uint32_t wbase = tab + (uint32_t)ms_table_bytes(t) + (uint32_t)warp * 4u * (uint32_t)lay.bytes16;   // four interleaved shots
    const uint32_t sbase = wbase + 4u * (uint32_t)grp;                                                      // my group's shot
    // byte offset `off` of the single-shot layout -> address in the interleaved layout
    auto at = [&](uint32_t base, uint32_t off) { return base + (off << 2); };
    const uint32_t layer_chk = tab + 2u * t.off_layer_chk, col_chk = tab + 2u * t.off_col_chk;
    const uint32_t rowpar = tab + 2u * t.off_rowpar, unperm = tab + 2u * t.off_unperm;
    const uint32_t srec = tab + 2u * ts.off_srec + (uint32_t)h * (4u * SPL);     // my lane's record of layer 0
    const int n = t.n;
    const uint32_t n4 = 4u * (uint32_t)n, n16 = 16u * (uint32_t)n;
    const uint32_t oC = (uint32_t)lay.off_c2v, oS = (uint32_t)lay.off_S, oP = (uint32_t)lay.off_par, oY = (uint32_t)lay.off_syn;
    const uint32_t cS = sbase + 4u * oS, cC = sbase + 4u * oC, cP = sbase + 4u * oP;   // my shot's S / c2v / parity arrays (interleaved addresses)
    const uint32_t colmask = tab + 2u * ts.off_colmask;
    float Tf = c.Tf;
    asm volatile("" : "+f"(Tf));              // stays in a register (the compiler would re-load it from the constant bank per use)
    const bool init_bit = 0.0f < Tf;

    const float inf = __int_as_float(0x7f800000);

    // per-group state (equal in the 8 lanes of a group)
    long long shot = -1;
    int l = 0, it = 0, unsat = 0, iters_out = 0;
    bool active = false, exhausted = false, fin = false, conv_out = false;
    uint32_t e[SPL], i = 0;                   // record and check of layer l (fetched one step ahead)
#pragma unroll
    for (int s = 0; s < SPL; ++s) e[s] = 0;

    for (;;) {
        // ---------------- groups without a running shot: write the finished shot's results, take the next shot (whole warp, one
        // group at a time)
        const uint32_t needmask = __ballot_sync(full, !active && !exhausted);
        if (needmask) {
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                if (!((needmask >> (8 * g)) & 1u)) continue;                      // warp-uniform
                const uint32_t gb = wbase + 4u * (uint32_t)g;
                const bool g_fin = __shfl_sync(full, fin ? 1 : 0, 8 * g) != 0;
                if (g_fin) {
                    const long long sh = __shfl_sync(full, shot, 8 * g);
                    const int g_it = __shfl_sync(full, iters_out, 8 * g);
                    const bool g_cv = __shfl_sync(full, conv_out ? 1 : 0, 8 * g) != 0;
                    for (int w = 0; w < t.nw; ++w) {
                        const int j = w * 32 + lane;
                        const uint32_t bits = __ballot_sync(full, j < n && c.max_iter > 0 && sld_f32(at(gb, oS + sld_u16(unperm + 2u * j))) < Tf);
                        if (lane == 0) io.ehat[sh * t.nw + w] = bits;
                    }
                    if (lane == 0) {
                        io.iters[sh] = g_it;
                        if (io.conv) io.conv[sh] = g_cv ? 1 : 0;
                    }
                    if (io.llr) {
                        double *dst = io.llr + sh * (long long)n;
                        for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)sld_f32(at(gb, oS + sld_u16(unperm + 2u * j))));
                    }
                    if (!g_cv && io.fail_count) {
                        int slot = 0;
                        if (lane == 0) {
                            slot = atomicAdd(io.fail_count, 1);
                            if (slot < io.fail_cap) io.fail_shot[slot] = (int)sh;
                        }
                        slot = __shfl_sync(full, slot, 0);
                        if (slot < io.fail_cap) {
                            double *dst = io.fail_llr + (long long)slot * n;
                            for (int j = lane; j < n; j += 32) dst[j] = __dadd_rn(c.L, (double)sld_f32(at(gb, oS + sld_u16(unperm + 2u * j))));
                        }
                    }
                }
                long long ns = 0;
                if (lane == 0) ns = (long long)atomicAdd(io.work_counter, 1ull);
                ns = __shfl_sync(full, ns, 0);
                const bool got = ns < io.shots;
                int u = 0;
                if (got) {
                    // initial state: c2v = 0 (decoders.py:150), S = 0, residual = syndrome (+ H.1 if the all-zero sums decide 1)
                    for (int k = lane; k < lay.zero_words; k += 32) sst_f32(at(gb, 4u * (uint32_t)k), 0.0f);
                    __syncwarp();
                    if (lane == 0) sst_u32(at(gb, oS + n4 + 4u), 0x7f800000u);        // S[n+1] = +inf: the padding edges
                    for (int k = lane; k < t.mw; k += 32) {
                        const uint32_t w = io.syn[ns * t.mw + k];
                        const uint32_t p0 = init_bit ? (w ^ sld_u32(rowpar + 4u * k)) : w;
                        sst_u32(at(gb, oY + 4u * k), w);
                        sst_u32(at(gb, oP + 4u * k), p0);
                        u += __popc(p0);
                    }
                    u = __reduce_add_sync(full, u);
                }
                if (grp == g) {
                    fin = false;
                    if (got) {
                        shot = ns; l = 0; it = 0; unsat = u;
                        sub_load_rec<SPL>(srec, e);
                        i = sld_u16(layer_chk);
                        if (c.max_iter > 0) active = true;
                        else { fin = true; iters_out = 0; conv_out = false; }          // decoders.py:153 with max_iter = 0: no step at all
                    } else exhausted = true;
                }
                __syncwarp();
            }
            continue;                                                              // re-evaluate (a max_iter = 0 shot finishes at once)
        }
        if (!__any_sync(full, active)) break;

        // ---------------- one layer step of every group (its own shot, its own layer).  Groups without a shot run along on stale
        // state with every store / flip disabled, so the warp stays convergent.
        const double prior = (it == 0 && l == 0) ? c.Lf : c.L;                      // binary32-rounded prior in the very first step (:148-149)
        int l_next = l + 1;
        if (l_next == t.nl) l_next = 0;
        // all loads of the check phase, then the record of the next step
        uint32_t sa[SPL], ca[SPL];
        float s_old[SPL], cv[SPL];
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            sa[s] = cS + (e[s] & 0xffffu);
            ca[s] = cC + (e[s] >> 16);
            s_old[s] = sld_f32(sa[s]);
            cv[s] = sld_f32(ca[s]);
        }
        const uint32_t synw = sld_u32(at(sbase, oY + 4u * (i >> 5)));
        uint32_t e_next[SPL];
        sub_load_rec<SPL>(srec + (uint32_t)l_next * (32u * SPL), e_next);
        const uint32_t i_next = sld_u16(layer_chk + 2u * (uint32_t)l_next);
        {   // ---- check phase (decoders.py:156-169): lane h holds cells (s, h), s < SPL, of the check
            float bs[SPL];
            float m1 = inf, m2 = inf;
            uint32_t px = 0;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const double post = __dadd_rn(prior, (double)s_old[s]);                                    // :173
                const double v = __dsub_rn(post, (double)cv[s]);                                           // :177
                const float b = __double2float_rn(__dmul_rn(c.abeta, v));                                  // :167-168
                bs[s] = b;
                px ^= __float_as_uint(b);
                const float ab = fabsf(b);
                m2 = fminf(m2, fmaxf(m1, ab));
                m1 = fminf(m1, ab);
            }
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const float o1 = __shfl_xor_sync(full, m1, d);
                const float o2 = __shfl_xor_sync(full, m2, d);
                px ^= __shfl_xor_sync(full, px, d);
                m2 = fminf(fmaxf(m1, o1), fminf(m2, o2));
                m1 = fminf(m1, o1);
            }
            const float r1 = (m1 == inf) ? 0.0f : m1;                               // inf -> 0 (:165-166, :169)
            const float r2 = (m2 == inf) ? 0.0f : m2;
            const uint32_t synbit = (synw >> (i & 31u)) & 1u;
            const uint32_t P = (px ^ (synbit << 31) ^ c.sgn) & 0x80000000u;
            const uint32_t r1s = __float_as_uint(r1) | P, r2s = __float_as_uint(r2) | P;
            if (active) {
#pragma unroll
                for (int s = 0; s < SPL; ++s) {
                    const uint32_t mag = (fabsf(bs[s]) == m1) ? r2s : r1s;
                    sst_u32(ca[s], mag ^ (__float_as_uint(bs[s]) & 0x80000000u));
                }
            }
        }
        // ---- variable phase (decoders.py:172-174): trip s = the variables of my cells (s, h); the only message of theirs that
        // changed is the one I stored myself
        float term[SPL][DV];
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            uint32_t cj = cC + (e[s] & 0xffffu);
            asm volatile("" : "+r"(cj));      // keep (per-lane base) + (uniform offset): one address register per variable
#pragma unroll
            for (int x = 0; x < DV; ++x) term[s][x] = sld_f32(cj + (uint32_t)ts.c16[x]);
        }
        bool fl[SPL], anyf = false;
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const uint32_t j16 = e[s] & 0xffffu;
            const bool real = j16 < n16;                                            // not a padding cell
#pragma unroll
            for (int x = DMIN; x < DV; ++x) term[s][x] = ((int)j16 < 4 * t.cnt4[x]) ? term[s][x] : 0.0f;
            float sum = term[s][0];
#pragma unroll
            for (int x = 1; x < DV; ++x) sum = __fadd_rn(sum, term[s][x]);
            if (active && real) sst_f32(sa[s], sum);
            fl[s] = active && real && ((sum < Tf) != (s_old[s] < Tf));             // hard decision flipped (:173-174)
            anyf = anyf || fl[s];
        }
        if (__any_sync(full, anyf)) {
            // every lane toggles the residual parities of its own flipped variables; the count of unsatisfied checks is recounted
            if constexpr (MW > 0) {
                uint32_t mk[MW];
#pragma unroll
                for (int w = 0; w < MW; ++w) mk[w] = 0;
#pragma unroll
                for (int s = 0; s < SPL; ++s) {
                    const uint32_t a = colmask + ((e[s] & 0xffffu) >> 4) * (4u * MW);
#pragma unroll
                    for (int w = 0; w < MW; ++w) { const uint32_t v = fl[s] ? sld_u32(a + 4u * w) : 0u; mk[w] ^= v; }
                }
#pragma unroll
                for (int w = 0; w < MW; ++w) if (mk[w]) sred_xor(cP + 16u * w, mk[w]);
                __syncwarp();
                int u = 0;
#pragma unroll
                for (int w = 0; w < MW; ++w) u += __popc(sld_u32(cP + 16u * w));
                unsat = u;
            } else {
#pragma unroll
                for (int s = 0; s < SPL; ++s) {
                    if (fl[s]) {
                        const uint32_t a = col_chk + ((e[s] & 0xffffu) >> 3) * (uint32_t)DV;
#pragma unroll 4
                        for (int x = 0; x < DV; ++x) {
                            const uint32_t ch = sld_u16(a + 2u * (uint32_t)x);
                            if (ch != 0xffffu) sred_xor(cP + 16u * (ch >> 5), 1u << (ch & 31u));
                        }
                    }
                }
                __syncwarp();
                int u = 0;
                for (int w = h; w < t.mw; w += 8) u += __popc(sld_u32(cP + 16u * (uint32_t)w));
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) u += __shfl_xor_sync(full, u, d);
                unsat = u;
            }
        }
        __syncwarp();
        // ---- H e == syndrome ?  (decoders.py:175-176); next layer / iteration
        {   // (groups without a shot advance along as well: their layer state is rewritten when they take a shot)
            const bool conv_now = unsat == 0;
            const int it_next = l_next == 0 ? it + 1 : it;
            if (active && (conv_now || it_next >= c.max_iter)) {
                active = false; fin = true;
                conv_out = conv_now;
                iters_out = conv_now ? it + 1 : c.max_iter;                         // :176 / :182
            }
            l = l_next; it = it_next;
#pragma unroll
            for (int s = 0; s < SPL; ++s) e[s] = e_next[s];
            i = i_next;
        }
    }
}

}  // namespace qldpc
