// hard_kernels.cuh -- integer decoders (bit flipping, naive greedy) and the outcome classifier.
// One warp per shot, persistent CTAs, shot state in shared memory, CSR/CSC tables read through L1.
#pragma once
#include "common.cuh"

namespace qldpc {

struct GraphDev {   // int32 CSR / CSC in global memory
    int m, n, mw, nw;
    const int32_t *row_ptr, *col_idx;   // CSR: variables of check i, ascending
    const int32_t *col_ptr, *row_idx;   // CSC: checks of variable j, ascending
};

__device__ __forceinline__ uint32_t get_bit(const uint32_t *w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }

// ---------------------------------------------------------------------------------------------------------
// Parallel bit flipping -- decoders.py:74-102.
//   nuc_j = #unsatisfied checks on j (:95); flip where nuc_j > w_j/2 (:96); the residual is
//   (ANY flipped neighbour) xor syndrome -- an OR, not a parity (:97-98); stop when it is all-zero (:99-100).
// Shared memory per warp: e bits [nw] | r bits [mw] | syndrome bits [mw].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bf_decode_kernel(GraphDev g, int max_iter, DecodeIO io)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = g.nw + 2 * g.mw;
    uint32_t *eb = reinterpret_cast<uint32_t *>(smem) + (size_t)warp * per;
    uint32_t *rb = eb + g.nw;
    uint32_t *sb = rb + g.mw;
    const unsigned full = 0xffffffffu;
    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;
        for (int i = lane; i < g.nw; i += 32) eb[i] = 0u;
        for (int i = lane; i < g.mw; i += 32) { uint32_t w = io.syn[shot * g.mw + i]; rb[i] = w; sb[i] = w; }   // r = syndrome (:90)
        __syncwarp();
        int iters = max_iter;
        bool converged = false;
        for (int it = 0; it < max_iter; ++it) {
            for (int w = 0; w < g.nw; ++w) {                    // one word of variables per trip
                const int j = w * 32 + lane;
                bool flip = false;
                if (j < g.n) {
                    const int t0 = g.col_ptr[j], t1 = g.col_ptr[j + 1];
                    int nuc = 0;
                    for (int x = t0; x < t1; ++x) nuc += get_bit(rb, g.row_idx[x]);       // :95
                    flip = 2 * nuc > (t1 - t0);                                            // nuc > nChecks/2. (:96)
                }
                const uint32_t fw = __ballot_sync(full, flip);
                if (lane == 0) eb[w] ^= fw;
            }
            __syncwarp();
            uint32_t any = 0;
            for (int w = 0; w < g.mw; ++w) {
                const int i = w * 32 + lane;
                bool r = false;
                if (i < g.m) {
                    uint32_t cnt = 0;
                    for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) cnt |= get_bit(eb, g.col_idx[x]);   // s_hat != 0 (:97)
                    r = (cnt ^ get_bit(sb, i)) != 0;                                                             // :98
                }
                const uint32_t rw = __ballot_sync(full, r);
                if (lane == 0) rb[w] = rw;
                any |= rw;
            }
            __syncwarp();
            if (any == 0) { iters = it + 1; converged = true; break; }                     // :99-100
        }
        for (int w = lane; w < g.nw; w += 32) io.ehat[shot * g.nw + w] = eb[w];
        if (lane == 0) { io.iters[shot] = iters; if (io.conv) io.conv[shot] = converged ? 1 : 0; }
        __syncwarp();
    }
}

// Sparse formulation of the same decoder (used whenever words(m) <= 32).  Both matrix-vector products of an iteration
// have sparse inputs -- r (the unsatisfied checks) and e (the flipped variables) -- so instead of visiting every edge twice
//   * the unsatisfied checks are compacted into a list (ballot + popc), one lane per listed check adds 1 to the counters of
//     its variables (shared-memory atomics),
//   * one dense pass over the n counters decides the flips and clears the counters,
//   * "any flipped neighbour" is the OR of the bit-packed columns of H selected by the set bits of e (every lane walks its
//     own word of e, one warp OR-reduction per residual word).
// Same integers as the dense kernel above (kept for matrices with more than 1024 checks); ~3.5x fewer instructions.
// Shared memory per warp: e bits [nw] | r bits [mw] | syndrome bits [mw] | counters uint32 [32*nw] | list uint16 [m rounded to 2].
// Per CTA (after the warps): column weights uint16 [n] (a column has at most m <= 1024 entries here).
template <int VH>
__global__ void __launch_bounds__(256) bf_sparse_kernel(GraphDev g, const uint32_t *__restrict__ hcol, int max_iter, DecodeIO io)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int per = g.nw + 2 * g.mw + 32 * g.nw + (g.m + 1) / 2;          // words per warp
    uint32_t *eb = reinterpret_cast<uint32_t *>(smem) + (size_t)warp * per;
    uint32_t *rb = eb + g.nw;
    uint32_t *sb = rb + g.mw;
    uint32_t *nuc = sb + g.mw;
    uint16_t *list = reinterpret_cast<uint16_t *>(nuc + 32 * g.nw);
    uint16_t *colw = reinterpret_cast<uint16_t *>(reinterpret_cast<uint32_t *>(smem) + (size_t)nwarps * per);
    for (int j = threadIdx.x; j < g.n; j += blockDim.x) colw[j] = (uint16_t)(g.col_ptr[j + 1] - g.col_ptr[j]);
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint4 *cols = reinterpret_cast<const uint4 *>(hcol);
    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;
        for (int i = lane; i < g.nw; i += 32) eb[i] = 0u;
        for (int i = lane; i < 32 * g.nw; i += 32) nuc[i] = 0u;
        for (int i = lane; i < g.mw; i += 32) { uint32_t w = io.syn[shot * g.mw + i]; rb[i] = w; sb[i] = w; }   // r = syndrome (:90)
        __syncwarp();
        int iters = max_iter;
        bool converged = false;
        for (int it = 0; it < max_iter; ++it) {
            // ---- unsatisfied checks -> list
            int cnt = 0;
            for (int t = 0; t < g.mw; ++t) {
                const uint32_t w = rb[t];                                    // warp-uniform
                if ((w >> lane) & 1u) list[cnt + __popc(w & lt_mask)] = (uint16_t)(t * 32 + lane);
                cnt += __popc(w);
            }
            __syncwarp();
            // ---- nuc_j = #unsatisfied checks on j (:95)
            for (int q = lane; q < cnt; q += 32) {
                const int i = list[q];
                for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) atomicAdd(&nuc[g.col_idx[x]], 1u);
            }
            __syncwarp();
            // ---- flip where nuc_j > w_j / 2 (:96); clear the counters for the next iteration
            for (int w = 0; w < g.nw; ++w) {
                const int j = w * 32 + lane;
                const uint32_t c = nuc[j];
                nuc[j] = 0u;
                const bool flip = (j < g.n) && (2u * c > (uint32_t)colw[j < g.n ? j : 0]);
                const uint32_t fw = __ballot_sync(full, flip);
                if (lane == 0) eb[w] ^= fw;
            }
            __syncwarp();
            // ---- r = (any flipped neighbour) xor syndrome (:97-98): OR of the columns selected by e
            uint32_t acc[4 * VH];
#pragma unroll
            for (int k = 0; k < 4 * VH; ++k) acc[k] = 0u;
            for (int w0 = 0; w0 < g.nw; w0 += 32) {
                uint32_t word = (w0 + lane < g.nw) ? eb[w0 + lane] : 0u;
                while (__any_sync(full, word != 0u)) {
                    if (word) {
                        const int j = (w0 + lane) * 32 + __ffs(word) - 1;
                        word &= word - 1;
                        const uint4 *c = cols + (size_t)j * (kColStride / 4);
#pragma unroll
                        for (int v = 0; v < VH; ++v) {
                            const uint4 q = __ldg(c + v);
                            acc[4 * v + 0] |= q.x; acc[4 * v + 1] |= q.y; acc[4 * v + 2] |= q.z; acc[4 * v + 3] |= q.w;
                        }
                    }
                }
            }
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 4 * VH; ++k) {
                if (k < g.mw) {                                              // warp-uniform
                    const uint32_t rw = __reduce_or_sync(full, acc[k]) ^ sb[k];
                    if (lane == 0) rb[k] = rw;
                    any |= rw;
                }
            }
            __syncwarp();
            if (any == 0) { iters = it + 1; converged = true; break; }                     // :99-100
        }
        for (int w = lane; w < g.nw; w += 32) io.ehat[shot * g.nw + w] = eb[w];
        if (lane == 0) { io.iters[shot] = iters; if (io.conv) io.conv[shot] = converged ? 1 : 0; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Naive greedy -- decoders.py:27-66.
//   Repeat (at most 2n times, :47): score_v = #failing checks on v (:52-56); stop if the residual is zero (:49)
//   or the best score is 0 (:57-58); flip the LOWEST-index variable of maximal score (np.argmax, :59) and
//   toggle its checks (:61-64).  The scores are maintained incrementally (a toggled check adds +-1 to its
//   variables), which yields the same integers as the reference's recomputation.
// Shared memory per warp: est bits [nw] | residual bits [mw] | score int32 [n].
// ---------------------------------------------------------------------------------------------------------
// TAB16: CSR / CSC copied to shared memory as uint16 once per CTA (every index < 65536); otherwise read through L1.
template <bool TAB16>
__global__ void __launch_bounds__(256) ng_decode_kernel(GraphDev g, DecodeIO io)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int per = g.nw + g.mw + g.n;
    uint32_t *eb = reinterpret_cast<uint32_t *>(smem) + (size_t)warp * per;
    uint32_t *rb = eb + g.nw;
    int *score = reinterpret_cast<int *>(rb + g.mw);
    // per-CTA tables behind the per-warp state: row_ptr [m+1] | col_ptr [n+1] | col_idx [E] | row_idx [E]
    const int E = g.row_ptr[g.m];
    uint16_t *t_row_ptr = reinterpret_cast<uint16_t *>(reinterpret_cast<uint32_t *>(smem) + (size_t)nwarps * per);
    uint16_t *t_col_ptr = t_row_ptr + (g.m + 1);
    uint16_t *t_col_idx = t_col_ptr + (g.n + 1);
    uint16_t *t_row_idx = t_col_idx + E;
    if (TAB16) {
        for (int i = threadIdx.x; i <= g.m; i += blockDim.x) t_row_ptr[i] = (uint16_t)g.row_ptr[i];
        for (int i = threadIdx.x; i <= g.n; i += blockDim.x) t_col_ptr[i] = (uint16_t)g.col_ptr[i];
        for (int i = threadIdx.x; i < E; i += blockDim.x) { t_col_idx[i] = (uint16_t)g.col_idx[i]; t_row_idx[i] = (uint16_t)g.row_idx[i]; }
        __syncthreads();
    }
    auto row_ptr = [&](int i) -> int { return TAB16 ? (int)t_row_ptr[i] : g.row_ptr[i]; };
    auto col_ptr = [&](int j) -> int { return TAB16 ? (int)t_col_ptr[j] : g.col_ptr[j]; };
    auto col_idx = [&](int x) -> int { return TAB16 ? (int)t_col_idx[x] : g.col_idx[x]; };
    auto row_idx = [&](int x) -> int { return TAB16 ? (int)t_row_idx[x] : g.row_idx[x]; };
    const unsigned full = 0xffffffffu;
    for (;;) {
        long long shot = 0;
        if (lane == 0) shot = (long long)atomicAdd(io.work_counter, 1ull);
        shot = __shfl_sync(full, shot, 0);
        if (shot >= io.shots) break;
        for (int i = lane; i < g.nw; i += 32) eb[i] = 0u;
        int rsum = 0;
        for (int i = lane; i < g.mw; i += 32) { uint32_t w = io.syn[shot * g.mw + i]; rb[i] = w; rsum += __popc(w); }
        rsum = __reduce_add_sync(full, rsum);
        __syncwarp();
        for (int j = lane; j < g.n; j += 32) {
            int sc = 0;
            for (int x = col_ptr(j); x < col_ptr(j + 1); ++x) sc += get_bit(rb, row_idx(x));
            score[j] = sc;
        }
        __syncwarp();
        int steps = 0;
        const int max_steps = 2 * g.n;                                   // :47
        while (rsum > 0 && steps < max_steps) {                          // :49
            ++steps;
            int best = 0, arg = 0x7fffffff;
            for (int j = lane; j < g.n; j += 32) {                       // ascending j per lane: strict > keeps the first
                const int sc = score[j];
                if (sc > best) { best = sc; arg = j; }
            }
            const int wbest = __reduce_max_sync(full, best);
            if (wbest == 0) break;                                       // :57-58
            const int v = __reduce_min_sync(full, best == wbest ? arg : 0x7fffffff);   // first maximum (:59)
            if (lane == 0) eb[v >> 5] ^= 1u << (v & 31);                 // :61
            const int t0 = col_ptr(v), t1 = col_ptr(v + 1);
            for (int x = t0; x < t1; ++x) {                              // :63-64, one check at a time
                const int ch = row_idx(x);
                const int was = get_bit(rb, ch);
                const int delta = was ? -1 : 1;
                __syncwarp();
                if (lane == 0) rb[ch >> 5] ^= 1u << (ch & 31);
                for (int y = row_ptr(ch) + lane; y < row_ptr(ch + 1); y += 32) score[col_idx(y)] += delta;
                rsum += delta;
                __syncwarp();
            }
        }
        for (int w = lane; w < g.nw; w += 32) io.ehat[shot * g.nw + w] = eb[w];
        if (lane == 0) { io.iters[shot] = steps; if (io.conv) io.conv[shot] = (rsum == 0) ? 1 : 0; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Outcome classification + counter reduction -- simulator.py:291-303.
//   exact : errX == eX and errZ == eZ (:294-295)
//   degen : not exact and Hz (errX^eX) == 0 and Hx (errZ^eZ) == 0 as INTEGER products, i.e. the difference
//           touches no check at all (:296-298, no mod 2) <=> diff & colmask == 0, colmask = OR of the rows
//   failX : sy_z != Hz eX mod 2 (:300-301),  failZ : sy_x != Hx eZ mod 2 (:302-303)
// One warp per shot (grid-stride); counters accumulated per CTA then atomically into int64[QLDPC_NUM_COUNTERS].
// ---------------------------------------------------------------------------------------------------------
struct ClassifyArgs {
    GraphDev gz, gx;                       // gz = Hz (X-error decode), gx = Hx (Z-error decode)
    const uint32_t *colmask_z, *colmask_x; // [nw] OR of the rows of Hz / Hx
    const uint32_t *hcol_z, *hcol_x;       // [n][kColStride] bit-packed columns of Hz / Hx, zero padded
    const uint32_t *lcol_z, *lcol_x;       // [n][kColStride] bit-packed columns of the logical-Z / logical-X bases, or null
    const uint32_t *errx, *errz, *ehx, *ehz, *synz, *synx;
    const int32_t *itx, *itz;
    long long shots;
    unsigned long long *counters;
};

// M d (mod 2) for a sparse bit-packed vector d = x ^ y (y may be null), as the XOR of the bit-packed COLUMNS of M selected
// by the set bits of d (the estimates and residuals are sparse: ~2pn/3 bits).  Every lane walks the set bits of its own
// word(s) of d and XORs whole columns (V4 128-bit loads each) into private registers -- no cross-lane traffic per bit --
// and the 4*V4 result words are combined at the end with one warp XOR-reduction (REDUX) each, XOR-ed with `init` (the
// syndrome words, or null) and tested for zero.  cols: [n][kColStride] words (one 128-byte line per column, zero padded), of
// which the first 4*V4 are read.
template <int V4>
__device__ __forceinline__ void xor_columns(const uint4 *__restrict__ cols, int nw, const uint32_t *x, const uint32_t *y, int lane,
                                            uint32_t acc[4 * V4])
{
#pragma unroll
    for (int k = 0; k < 4 * V4; ++k) acc[k] = 0u;
    for (int w0 = 0; w0 < nw; w0 += 32) {
        uint32_t word = 0u;
        if (w0 + lane < nw) word = y ? (x[w0 + lane] ^ y[w0 + lane]) : x[w0 + lane];
        while (__any_sync(0xffffffffu, word != 0u)) {
            if (word) {
                const int j = (w0 + lane) * 32 + __ffs(word) - 1;
                word &= word - 1;
                const uint4 *c = cols + (size_t)j * (kColStride / 4);
#pragma unroll
                for (int v = 0; v < V4; ++v) {
                    const uint4 q = __ldg(c + v);
                    acc[4 * v + 0] ^= q.x; acc[4 * v + 1] ^= q.y; acc[4 * v + 2] ^= q.z; acc[4 * v + 3] ^= q.w;
                }
            }
        }
    }
}

template <int V4>
__device__ __forceinline__ bool xor_columns_nonzero(const uint4 *__restrict__ cols, int nw, const uint32_t *x, const uint32_t *y,
                                                    const uint32_t *init, int init_words, int lane)
{
    uint32_t acc[4 * V4];
    xor_columns<V4>(cols, nw, x, y, lane, acc);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 4 * V4; ++k) {
        uint32_t r = __reduce_xor_sync(0xffffffffu, acc[k]);
        if (init && k < init_words) r ^= init[k];            // warp-uniform (broadcast) load
        bad |= r != 0u;
    }
    return bad;
}

// M e + s != 0 (mod 2) row by row through the CSR tables: matrices with more than 1024 checks, whose bit-packed columns do not
// fit the 128-byte column lines (e, s: global memory, read through L1)
__device__ __forceinline__ bool rows_parity_nonzero(const GraphDev &g, const uint32_t *e, const uint32_t *syn, int lane)
{
    bool bad = false;
    for (int i = lane; i < g.m; i += 32) {
        uint32_t par = get_bit(syn, i);
        for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; ++x) par ^= get_bit(e, g.col_idx[x]);
        bad |= par != 0u;
    }
    return __any_sync(0xffffffffu, bad);
}

// VH: 128-bit loads per column of H (0: row-wise parities, more than 1024 checks); VL: per column of the logical bases
template <int VH, int VL>
__global__ void __launch_bounds__(256) classify_kernel(ClassifyArgs a)
{
    __shared__ unsigned long long cta_cnt[QLDPC_NUM_COUNTERS];
    if (threadIdx.x < QLDPC_NUM_COUNTERS) cta_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nw = a.gz.nw;
    unsigned long long c_fx = 0, c_fz = 0, c_ex = 0, c_dg = 0, c_ix = 0, c_iz = 0, c_sh = 0, c_td = 0, c_lg = 0, c_fa = 0;
    const bool have_logicals = a.lcol_z != nullptr && a.lcol_x != nullptr;
    for (long long s = warp_global; s < a.shots; s += nwarps) {
        const uint32_t *ex = a.errx + s * nw, *ez = a.errz + s * nw, *hx = a.ehx + s * nw, *hz = a.ehz + s * nw;
        bool diff = false, touch = false;
        for (int w = lane; w < nw; w += 32) {
            const uint32_t dx = ex[w] ^ hx[w], dz = ez[w] ^ hz[w];
            diff |= (dx | dz) != 0;
            touch |= ((dx & a.colmask_z[w]) | (dz & a.colmask_x[w])) != 0;
        }
        const bool exact = !__any_sync(0xffffffffu, diff);
        const bool degen = !exact && !__any_sync(0xffffffffu, touch);
        bool fx, fz;
        if constexpr (VH > 0) {
            fx = xor_columns_nonzero<VH>(reinterpret_cast<const uint4 *>(a.hcol_z), nw, hx, nullptr, a.synz + s * a.gz.mw, a.gz.mw, lane);
            fz = xor_columns_nonzero<VH>(reinterpret_cast<const uint4 *>(a.hcol_x), nw, hz, nullptr, a.synx + s * a.gx.mw, a.gx.mw, lane);
        } else {
            fx = rows_parity_nonzero(a.gz, hx, a.synz + s * a.gz.mw, lane);
            fz = rows_parity_nonzero(a.gx, hz, a.synx + s * a.gx.mw, lane);
        }
        // README classes (README.md:15-22): the residual of a shot whose two syndromes are reproduced lies in the normaliser;
        // it is a stabiliser iff its X part commutes with every logical Z and its Z part with every logical X
        bool logical = false;
        if (have_logicals && !exact && !fx && !fz)                                   // warp-uniform
            logical = xor_columns_nonzero<VL>(reinterpret_cast<const uint4 *>(a.lcol_z), nw, ex, hx, nullptr, 0, lane) |
                      xor_columns_nonzero<VL>(reinterpret_cast<const uint4 *>(a.lcol_x), nw, ez, hz, nullptr, 0, lane);
        if (lane == 0) {
            const bool fail_any = fx | fz;
            c_fa += fail_any;
            if (have_logicals && !exact && !fail_any) { c_lg += logical; c_td += !logical; }
            c_fx += fx; c_fz += fz; c_ex += exact; c_dg += degen;
            c_ix += (unsigned long long)a.itx[s]; c_iz += (unsigned long long)a.itz[s]; c_sh += 1;
        }
    }
    if (lane == 0) {
        atomicAdd(&cta_cnt[QLDPC_CNT_FAIL_X], c_fx);
        atomicAdd(&cta_cnt[QLDPC_CNT_FAIL_Z], c_fz);
        atomicAdd(&cta_cnt[QLDPC_CNT_EXACT], c_ex);
        atomicAdd(&cta_cnt[QLDPC_CNT_DEGEN], c_dg);
        atomicAdd(&cta_cnt[QLDPC_CNT_ITERS_X], c_ix);
        atomicAdd(&cta_cnt[QLDPC_CNT_ITERS_Z], c_iz);
        atomicAdd(&cta_cnt[QLDPC_CNT_SHOTS], c_sh);
        atomicAdd(&cta_cnt[QLDPC_CNT_TRUE_DEGEN], c_td);
        atomicAdd(&cta_cnt[QLDPC_CNT_LOGICAL], c_lg);
        atomicAdd(&cta_cnt[QLDPC_CNT_FAIL_ANY], c_fa);
    }
    __syncthreads();
    if (threadIdx.x < QLDPC_NUM_COUNTERS && cta_cnt[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], cta_cnt[threadIdx.x]);
}

}  // namespace qldpc
