// ms_lane_kernel.cuh -- min-sum decoder, LANE-PER-SHOT variant for schedules with tiny layers (serial: one check
// per layer step; decoders.py:110-182 with the `layers` of simulator.py:228-236, 'S').
//
// With one check per layer step a shot exposes only `row weight` (8) independent edges per step, so the warp-per-shot
// kernel (ms_kernel.cuh) leaves 24 of 32 lanes idle and pays its fixed per-step cost for 8 edges.  Here every lane owns
// a whole shot and the warp walks the layer list in lock-step (graph indices are warp-uniform, no shuffles, no
// atomics, no bank conflicts); parallelism comes from shots only.  The message state does not fit on chip at 32 shots
// per warp, so it is streamed through L2/HBM in a shot-minor layout (element x of the 32 shots of a warp is one 128-byte
// line): c2v[e][lane], S[j][lane], residual-syndrome and decision bit words [w][lane].  Every access of the warp is one
// fully coalesced line; the traffic model is 40 B per edge-iteration for row weight 8 / column weight 5.
//
// Arithmetic is the same bit-exact specification as ms_kernel.cuh.  Differences in bookkeeping:
//   * nothing is zeroed when a lane takes a new shot: during the shot's first iteration a c2v entry counts as 0 until
//     the layer that first processes its check has run, and S_j counts as 0 until the layer that first touches j has
//     run (tables fl_chk / fl_var);
//   * the reference's full posterior sweep after the very first layer step is a no-op for untouched variables when the
//     prior is non-negative (S_j = 0 decides 0), so it is skipped; plans with a negative prior use the warp kernel;
//   * a lane whose shot finishes mid-iteration idles until the next iteration boundary, where finished lanes fetch new
//     shots with one aggregated atomic;
//   * a lock-step layer step costs ~10 us of dependent L2/HBM round trips, so a shot that does not converge would hold
//     its lane for max_iter * layers steps (225 ms for LP118_2, 50 iterations) and, at the tail of a batch, keep a whole
//     SM waiting: shots still running after `defer_iters` iterations are handed to the warp-per-shot kernel instead.
#pragma once
#include "common.cuh"

namespace qldpc {

struct LaneTables {          // offsets into a uint16 blob (copied to shared memory per CTA)
    int m, n, dc, dv, nl, mw, nw;
    int off_chk_var;         // [m*dc]  variable of (check i, slot k), check-major; kPad past a short row
    int off_var_ptr;         // [n+1]
    int off_var_edge;        // [E]     edge id (i*dc+k) of the edges of variable j, ascending check
    int off_var_chk;         // [E]     check of those edges
    int off_var_fl;          // [E]     fl_chk of that check
    int off_layer_ptr, off_layer_chk, off_lvar_ptr, off_lvar_idx;
    int off_fl_chk;          // [m]     first layer that processes check i (0xFFFF: never)
    int off_fl_var;          // [n]     first layer that touches variable j (0xFFFF: never)
    int len;
};

struct LaneScratch {
    float *c2v;              // [warps][m*dc][32]
    float *S;                // [warps][n][32]
    uint32_t *par;           // [warps][2*mw][32]  residual syndrome words, then the syndrome words themselves
    uint32_t *eb;            // [warps][nw][32]
};

template <int DC, int DV>
__global__ void __launch_bounds__(256, 2) ms_lane_kernel(LaneTables t, const uint16_t *__restrict__ blob, MsConst c, DecodeIO io,
                                                        LaneScratch sc)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < t.len / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint16_t *chk_var = tab + t.off_chk_var;
    const uint16_t *var_ptr = tab + t.off_var_ptr;
    const uint16_t *var_edge = tab + t.off_var_edge;
    const uint16_t *var_chk = tab + t.off_var_chk;
    const uint16_t *var_fl = tab + t.off_var_fl;
    const uint16_t *layer_ptr = tab + t.off_layer_ptr;
    const uint16_t *layer_chk = tab + t.off_layer_chk;
    const uint16_t *lvar_ptr = tab + t.off_lvar_ptr;
    const uint16_t *lvar_idx = tab + t.off_lvar_idx;
    const uint16_t *fl_chk = tab + t.off_fl_chk;
    const uint16_t *fl_var = tab + t.off_fl_var;

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float *c2v = sc.c2v + gw * (long long)t.m * t.dc * 32 + lane;
    float *S = sc.S + gw * (long long)t.n * 32 + lane;
    uint32_t *par = sc.par + gw * (long long)t.mw * 64 + lane;
    const uint32_t *synw = par + t.mw * 32;
    uint32_t *eb = sc.eb + gw * (long long)t.nw * 32 + lane;
    const int n = t.n, dc = t.dc;
    const float Tf = c.Tf;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);

    long long shot = -1;
    int it = 0, unsat = 0;
    bool done = true, idle = false;

    for (;;) {
        // ---------------- iteration boundary: finished lanes take new shots (one aggregated atomic per warp)
        const bool need = done && !idle;
        const uint32_t nm = __ballot_sync(full, need);
        if (nm) {
            long long base = 0;
            if (lane == 0) base = (long long)atomicAdd(io.work_counter, (unsigned long long)__popc(nm));
            base = __shfl_sync(full, base, 0);
            if (need) {
                const long long s = base + __popc(nm & ((1u << lane) - 1u));
                if (s < io.shots) { shot = s; it = 0; done = false; unsat = 0; }
                else idle = true;
            }
            const bool fresh = need && !idle;
            for (int w = 0; w < t.mw; ++w) {
                if (fresh) { const uint32_t v = io.syn[shot * t.mw + w]; par[w * 32] = v; par[(t.mw + w) * 32] = v; unsat += __popc(v); }
            }
            for (int w = 0; w < t.nw; ++w) if (fresh) eb[w * 32] = 0u;
        }
        if (!__any_sync(full, !done)) break;

        for (int l = 0; l < t.nl; ++l) {
            const bool active = !done;
            const bool first_it = it == 0;
            const int thr_lt = first_it ? l : 0xFFFE;          // entries first written at a layer <  thr_lt+... (see uses)
            const double prior = (first_it && l == 0) ? c.Lf : c.L;     // binary32-rounded prior in the very first step (:148-149)
            // ---------------- check-node phase (decoders.py:156-169), checks of the layer one after the other
            const int qb = layer_ptr[l], qe = layer_ptr[l + 1];
            for (int q = qb; q < qe; ++q) {
                const int i = layer_chk[q];
                const bool own_valid = fl_chk[i] < thr_lt;      // processed in an earlier layer of this shot (or it > 0)
                float sv[DC], cv[DC];
                uint32_t jv[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    jv[k] = (k < dc) ? chk_var[i * dc + k] : (uint32_t)kPad;
                    const bool e = jv[k] != kPad;
                    const uint32_t j = e ? jv[k] : 0u;
                    sv[k] = (e && fl_var[j] < thr_lt) ? S[j * 32] : 0.0f;
                    cv[k] = (e && own_valid) ? c2v[(i * dc + (k < dc ? k : 0)) * 32] : 0.0f;
                }
                double m1 = inf, m2 = inf;
                int k1 = 0;
                uint32_t sb = 0;
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    if (jv[k] != kPad) {                         // warp-uniform
                        const double v = __dsub_rn(__dadd_rn(prior, (double)sv[k]), (double)cv[k]);   // :173, :177
                        const double av = fabs(v);
                        sb |= (v < 0.0 ? 1u : 0u) << k;                                                // :157-158
                        const bool lt1 = av < m1, lt2 = av < m2;
                        m2 = lt1 ? m1 : (lt2 ? av : m2);                                               // :162-164
                        m1 = lt1 ? av : m1;                                                            // :161
                        k1 = lt1 ? k : k1;
                    }
                }
                m1 = (m1 == inf) ? 0.0 : m1;                                                           // :165
                m2 = (m2 == inf) ? 0.0 : m2;                                                           // :166
                float r1 = __double2float_rn(__dmul_rn(c.beta, m1));                                   // :167
                float r2 = __double2float_rn(__dmul_rn(c.beta, m2));                                   // :168
                r1 = (r1 == __int_as_float(0x7f800000)) ? 0.0f : r1;                                   // :169
                r2 = (r2 == __int_as_float(0x7f800000)) ? 0.0f : r2;
                const uint32_t synbit = (synw[(i >> 5) * 32] >> (i & 31)) & 1u;
                const uint32_t P = (__popc(sb) & 1u) ^ synbit;                                         // :151, :159
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    if (jv[k] != kPad && active) {
                        const float mag = (k == k1) ? r2 : r1;
                        c2v[(i * dc + k) * 32] = __uint_as_float(__float_as_uint(mag) ^ ((((sb >> k) & 1u) ^ P) << 31));
                    }
                }
            }
            // ---------------- variable-node phase (decoders.py:172-174) on the variables adjacent to the layer
            const int thr_le = first_it ? l : 0xFFFE;          // c2v written at a layer <= l of this shot (or it > 0)
            const int vb = lvar_ptr[l], ve = lvar_ptr[l + 1];
            constexpr int VC = 8;                              // variables per batch: all their loads are issued before any use
            for (int q0 = vb; q0 < ve; q0 += VC) {
                float term[VC][DV], s_old[VC];
                int jj[VC], tt[VC], dd[VC];
#pragma unroll
                for (int v = 0; v < VC; ++v) {
                    const bool in = q0 + v < ve;
                    const int j = lvar_idx[in ? q0 + v : vb];
                    const int t0 = var_ptr[j], deg = in ? var_ptr[j + 1] - t0 : 0;
                    jj[v] = j; tt[v] = t0; dd[v] = deg;
#pragma unroll
                    for (int x = 0; x < DV; ++x) {
                        const bool e = x < deg;
                        const int xe = t0 + (e ? x : 0);
                        term[v][x] = (e && var_fl[xe] <= thr_le) ? c2v[(int)var_edge[xe] * 32] : 0.0f;
                    }
                    s_old[v] = (in && fl_var[j] < thr_lt) ? S[j * 32] : 0.0f;
                }
#pragma unroll
                for (int v = 0; v < VC; ++v) {
                    if (q0 + v < ve) {                          // warp-uniform
                        const int j = jj[v], deg = dd[v];
                        float s = term[v][0];
#pragma unroll
                        for (int x = 1; x < DV; ++x) s = (x < deg) ? __fadd_rn(s, term[v][x]) : s;    // :172
                        if (active) S[j * 32] = s;
                        const bool flip = active && ((s < Tf) != (s_old[v] < Tf));                     // :173-174
                        if (__any_sync(full, flip)) {
                            const uint32_t jb = 1u << (j & 31);
                            if (flip) eb[(j >> 5) * 32] ^= jb;
                            for (int x = 0; x < deg; ++x) {
                                const int ch = var_chk[tt[v] + x];
                                const uint32_t bit = 1u << (ch & 31);
                                if (flip) {
                                    const uint32_t old = par[(ch >> 5) * 32];
                                    par[(ch >> 5) * 32] = old ^ bit;
                                    unsat += (old & bit) ? -1 : 1;
                                }
                            }
                        }
                    }
                }
            }
            // ---------------- H e == syndrome ?  (decoders.py:175-176)
            const bool conv_now = active && unsat == 0;
            if (__any_sync(full, conv_now)) {
                if (conv_now) {
                    io.iters[shot] = it + 1;
                    if (io.conv) io.conv[shot] = 1;
                    done = true;
                }
                for (int w = 0; w < t.nw; ++w) if (conv_now) io.ehat[shot * t.nw + w] = eb[w * 32];
                if (io.llr) {
                    for (int j = 0; j < n; ++j)
                        if (conv_now) io.llr[shot * (long long)n + j] = __dadd_rn(c.L, (double)((fl_var[j] <= thr_le) ? S[j * 32] : 0.0f));
                }
            }
        }
        // ---------------- end of the iteration
        if (!done) ++it;
        // hand-over: a shot that needs more than defer_iters iterations would pin its lane (and, at the tail of the batch, its
        // whole warp) for max_iter slow lock-step iterations; it is abandoned here and decoded from scratch by the
        // warp-per-shot kernel, which is bit-identical
        const bool defer_now = !done && io.defer_list && it >= io.defer_iters && it < c.max_iter;
        if (__any_sync(full, defer_now)) {
            if (defer_now) { io.defer_list[atomicAdd(io.defer_count, 1)] = (int)shot; done = true; }
        }
        const bool fail_now = !done && it >= c.max_iter;                                               // :182
        if (__any_sync(full, fail_now)) {
            int slot = -1;
            if (fail_now) {
                io.iters[shot] = c.max_iter;
                if (io.conv) io.conv[shot] = 0;
                if (io.fail_count) {
                    slot = atomicAdd(io.fail_count, 1);
                    if (slot < io.fail_cap) io.fail_shot[slot] = (int)shot; else slot = -1;
                }
                done = true;
            }
            for (int w = 0; w < t.nw; ++w) if (fail_now) io.ehat[shot * t.nw + w] = eb[w * 32];
            if (io.llr || io.fail_count) {
                for (int j = 0; j < n; ++j) {
                    if (fail_now) {
                        const double v = __dadd_rn(c.L, (double)((fl_var[j] != 0xFFFF) ? S[j * 32] : 0.0f));
                        if (io.llr) io.llr[shot * (long long)n + j] = v;
                        if (slot >= 0) io.fail_llr[(long long)slot * n + j] = v;
                    }
                }
            }
        }
    }
}

}  // namespace qldpc
