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
//     32-lane zero fill), flips of hard decisions are handled by the whole warp for one flipped variable at a time.
// Results are bit-identical to the warp-per-shot kernel and to the reference.
#pragma once
#include "ms_kernel.cuh"

namespace qldpc {

constexpr int kMsSubWarps = 16;          // launch bound: 512 threads, 128 registers

// extra tables of the sub-warp kernel inside the same blob: off_svar u16 [nl][DCS] = 4*j' of the variables of the layer's check
// in (trip, lane-in-group) order, dummy 4*n; the check of layer l is layer_chk[l]
struct MsSubTables {
    int off_svar;
};

template <int DCS, int DV, int DMIN>
__global__ void __launch_bounds__(kMsSubWarps * 32, 1) ms_sub_kernel(MsTables t, MsSubTables ts, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io)
{
    static_assert(DCS % 8 == 0, "eight lanes per check");
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
    const uint32_t chk = tab + 2u * t.off_chk, layer_chk = tab + 2u * t.off_layer_chk, col_chk = tab + 2u * t.off_col_chk;
    const uint32_t rowpar = tab + 2u * t.off_rowpar, unperm = tab + 2u * t.off_unperm, svar = tab + 2u * ts.off_svar;
    const uint32_t m4 = 4u * t.ms;
    const int n = t.n;
    const uint32_t n4 = 4u * (uint32_t)n;
    const uint32_t oC = (uint32_t)lay.off_c2v, oS = (uint32_t)lay.off_S, oP = (uint32_t)lay.off_par, oY = (uint32_t)lay.off_syn;
    const float Tf = c.Tf;
    const bool init_bit = 0.0f < Tf;
    const float inf = __int_as_float(0x7f800000);

    // per-group state (equal in the 8 lanes of a group)
    long long shot = -1;
    int l = 0, it = 0, unsat = 0, iters_out = 0;
    bool active = false, exhausted = false, fin = false, conv_out = false;

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
                    for (int i = lane; i < lay.zero_words; i += 32) sst_f32(at(gb, 4u * (uint32_t)i), 0.0f);
                    __syncwarp();
                    if (lane == 0) sst_u32(at(gb, oS + n4 + 4u), 0x7f800000u);        // S[n+1] = +inf: the padding edges
                    for (int i = lane; i < t.mw; i += 32) {
                        const uint32_t w = io.syn[ns * t.mw + i];
                        const uint32_t p0 = init_bit ? (w ^ sld_u32(rowpar + 4u * i)) : w;
                        sst_u32(at(gb, oY + 4u * i), w);
                        sst_u32(at(gb, oP + 4u * i), p0);
                        u += __popc(p0);
                    }
                    u = __reduce_add_sync(full, u);
                }
                if (grp == g) {
                    fin = false;
                    if (got) {
                        shot = ns; l = 0; it = 0; unsat = u;
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
        const uint32_t i = sld_u16(layer_chk + 2u * (uint32_t)l);
        const double prior = (it == 0 && l == 0) ? c.Lf : c.L;                      // binary32-rounded prior in the very first step (:148-149)
        {   // ---- check phase (decoders.py:156-169): lane h holds slots h*SPL .. h*SPL+SPL-1 of the check
            const uint32_t ct = chk + (uint32_t)(h * SPL) * m4 + 4u * i;
            float bs[SPL];
            uint32_t ca[SPL];
            float m1 = inf, m2 = inf;
            uint32_t px = 0;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const uint32_t e = sld_u32(ct + (uint32_t)s * m4);
                ca[s] = at(sbase, oC + (e >> 16));
                const double post = __dadd_rn(prior, (double)sld_f32(at(sbase, oS + (e & 0xffffu))));      // :173
                const double v = __dsub_rn(post, (double)sld_f32(ca[s]));                                  // :177
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
            const uint32_t synbit = (sld_u32(at(sbase, oY + 4u * (i >> 5))) >> (i & 31u)) & 1u;
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
        __syncwarp();
        // ---- variable phase (decoders.py:172-174): SPL trips of 8 variables per group
        int delta = 0;
#pragma unroll
        for (int tr = 0; tr < SPL; ++tr) {
            const uint32_t j4 = sld_u16(svar + 2u * (uint32_t)((l * SPL + tr) * 8 + h));
            const uint32_t sa = at(sbase, oS + j4);
            const float s_old = sld_f32(sa);
            float term[DV];
#pragma unroll
            for (int x = 0; x < DV; ++x) {
                term[x] = sld_f32(at(sbase, oC + (uint32_t)t.coff4[x] + j4));
                if (x >= DMIN) term[x] = ((int)j4 < t.cnt4[x]) ? term[x] : 0.0f;
            }
            float s = term[0];
#pragma unroll
            for (int x = 1; x < DV; ++x) s = __fadd_rn(s, term[x]);
            if (active) sst_f32(sa, s);
            uint32_t flips = __ballot_sync(full, active && ((s < Tf) != (s_old < Tf)));       // hard decision flipped (:173-174)
            while (flips) {                                                        // rare; the whole warp toggles the checks of one flipped variable
                const int src = __ffs(flips) - 1;
                flips &= flips - 1;
                const uint32_t jf4 = __shfl_sync(full, j4, src);
                int d1 = 0;
                if (lane < DV) {
                    const uint32_t ch = sld_u16(col_chk + (jf4 >> 1) * (uint32_t)DV + 2u * (uint32_t)lane);
                    if (ch != 0xffffu) {
                        const uint32_t bit = 1u << (ch & 31u);
                        const uint32_t old = satom_xor(at(wbase + 4u * (uint32_t)(src >> 3), oP + 4u * (ch >> 5)), bit);
                        d1 = (old & bit) ? -1 : 1;
                    }
                }
                d1 = __reduce_add_sync(full, d1);
                if (grp == (src >> 3)) delta += d1;
            }
        }
        unsat += delta;
        __syncwarp();
        // ---- H e == syndrome ?  (decoders.py:175-176); next layer / iteration
        if (active) {
            const bool conv_now = unsat == 0;
            int it_next = it, l_next = l + 1;
            if (l_next == t.nl) { l_next = 0; it_next = it + 1; }
            if (conv_now || it_next >= c.max_iter) {
                active = false; fin = true;
                conv_out = conv_now;
                iters_out = conv_now ? it + 1 : c.max_iter;                         // :176 / :182
            }
            l = l_next; it = it_next;
        }
    }
}

}  // namespace qldpc
