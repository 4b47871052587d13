// osd_kernel.cuh -- ordered-statistics post-processing: bit-packed GF(2) Gauss-Jordan, one CTA per shot.
//
// Semantics: decoders.py:299-370 with gf2math.rank / gf2math.REF (gf2math.py:91-187); SURVEY.md App. A.5, B-8,
// B-9; CPU restatement oracle/qldpc_oracle.c:orc_osd.
//   1. reliability rel_j = max(P, 1-P), P = 1/(1+exp(clip(LLR, +-100)))               (decoders.py:320-324)
//   2. column order: ascending rel; the library's own order is the STABLE one (ties by index) computed by
//      rank counting; a caller-supplied `perm` replaces it (NumPy's argsort at :325 is unstable)
//   3. walk the columns in that order keeping those that raise the rank (= pivot columns of a Gauss-Jordan
//      elimination in that column order) until rank(H) are kept                         (decoders.py:329-342)
//   4. information bits keep the decoder's hard decision -- for order 1 the first information bit is flipped,
//      for order 0 and >= 2 none (the reference's order loop aliases its buffers, App. B-8) -- and the basis
//      bits are the unique solution of H_J e_J = s + H_I e_I                           (decoders.py:347-368)
// TWO KERNELS.  osd_reg_kernel<W> (below, the one every code of the reference's library uses: m <= 512 checks, n <= 1152
// columns) keeps one ROW PER THREAD IN REGISTERS, in permuted column order; osd_kernel<TRIPS> is the shared-memory formulation
// for larger matrices.  Implementation of osd_kernel: the matrix is NOT permuted; rows stay bit-packed in original column numbering with one
// extra word for the right-hand side, initialised to the residual syndrome s + H e.  Solving for the basis
// FLIPS d_J (H_J d_J = residual) is equivalent and needs no knowledge of I before the elimination.
//   * FORWARD elimination only (the pivot column is cleared from the rows that are not pivot rows yet), then a
//     back-substitution by one warp: half the row operations of Gauss-Jordan, same (unique) solution;
//   * warp per row, lane per 32-bit word: a row is touched only if it holds the pivot bit (one broadcast load
//     decides for the warp), the pivot row is held in registers, unused rows are kept as a compacted index list
//     that shrinks with every pivot;
//   * the same pass ORs the updated rows into `live` words: a column whose live bit is clear has no pivot among
//     the unused rows, so dependent columns are skipped 32 at a time by a ballot instead of costing a pivot search
//     and two barriers each;
//   * two CTA barriers per pivot (pivot row known / rows updated); the scratch words they protect are double
//     buffered by iteration parity.
#pragma once
#include "common.cuh"

namespace qldpc {

struct OsdArgs {
    int m, n, mw, nw;
    const int32_t *row_ptr, *col_idx;   // CSR of H (osd_reg_kernel builds its permuted rows from it)
    const uint32_t *hbits;     // [m][nw]
    uint32_t *ehat;            // [*][nw] in/out
    const uint32_t *syn;       // [*][mw]
    const double *llr;         // [count][n] (compact, indexed by list position)
    const int32_t *perm;       // [count][n] or null
    const int *shot_ids;       // list position -> shot row in ehat/syn; null = identity
    const int *count_dev;      // device-side count (null: use `count`)
    int count;
    int order;
    int rank_h;                // GF(2) rank of H (gf2math.rank(Hp), decoders.py:330): the walk stops once reached
};

__device__ __forceinline__ double osd_reliability(double llr)
{
    const double sg = (llr > 0.0) ? 1.0 : ((llr < 0.0) ? -1.0 : 0.0);
    const double sat = (fabs(llr) < 100.0) ? llr : 100.0 * sg;       // decoders.py:320-322
    const double P = 1.0 / (1.0 + exp(sat));                          // :323
    return (P > 0.5) ? P : 1.0 - P;                                   // :324
}

constexpr int kOsdThreads = 256;
constexpr int kOsdWarps = kOsdThreads / 32;
constexpr int kOsdMaxTrips = 2;           // words per lane: rows of up to 64 words (n <= 2016)

// dynamic shared memory: A [m][nw+1] uint32 | rel [n] double | perm [n] int | pivrow [n] int | rows [m] int |
// pcol [m] int | prow [m] int | live [2][rw] | scalars   -- sized by osd_smem_bytes()
inline size_t osd_smem_bytes(int m, int n, int nw)
{
    size_t b = (size_t)m * (nw + 1) * 4;
    b = (b + 7) & ~size_t(7);
    b += (size_t)n * 8;      // rel
    b += (size_t)n * 4;      // perm
    b += (size_t)n * 4;      // pivot row of column (or -1)
    b += (size_t)m * 4 * 3;  // unused-row list, pivot columns, pivot rows (in pivot order)
    b += (size_t)(nw + 1) * 4 * 2;   // live words, double buffered
    return b + 64;
}

// TRIPS = ceil((nw + 1) / 32): words of an augmented row per lane
template <int TRIPS>
__global__ void __launch_bounds__(kOsdThreads) osd_kernel(OsdArgs a)
{
    constexpr int kOsdTrips = TRIPS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int m = a.m, n = a.n, nw = a.nw, rw = a.nw + 1;   // rw: words per augmented row
    uint32_t *A = reinterpret_cast<uint32_t *>(smem);
    size_t off = ((size_t)m * rw * 4 + 7) & ~size_t(7);
    double *rel = reinterpret_cast<double *>(smem + off); off += (size_t)n * 8;
    int *perm = reinterpret_cast<int *>(smem + off); off += (size_t)n * 4;
    int *pivrow = reinterpret_cast<int *>(smem + off); off += (size_t)n * 4;
    int *rows = reinterpret_cast<int *>(smem + off); off += (size_t)m * 4;
    int *pcol = reinterpret_cast<int *>(smem + off); off += (size_t)m * 4;
    int *prowl = reinterpret_cast<int *>(smem + off); off += (size_t)m * 4;
    uint32_t *live = reinterpret_cast<uint32_t *>(smem + off); off += (size_t)rw * 4 * 2;
    int *sh = reinterpret_cast<int *>(smem + off);          // sh[0..1]: pivot candidate (list position), double buffered
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const int count = a.count_dev ? min(*a.count_dev, a.count) : a.count;
    const int INF = 0x7fffffff;

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const long long shot = a.shot_ids ? a.shot_ids[item] : item;
        uint32_t *e = a.ehat + shot * nw;
        const uint32_t *s = a.syn + shot * a.mw;
        const double *llr = a.llr + (long long)item * n;
        // ---- load H, residual right-hand side, reliabilities
        for (int x = tid; x < m * nw; x += kOsdThreads) {
            const int i = x / nw, w = x - i * nw;
            A[i * rw + w] = a.hbits[x];
        }
        for (int j = tid; j < n; j += kOsdThreads) { rel[j] = osd_reliability(llr[j]); pivrow[j] = -1; }
        for (int i = tid; i < m; i += kOsdThreads) rows[i] = i * rw;          // unused-row list: word offset of the row
        for (int w = tid; w < 2 * rw; w += kOsdThreads) live[w] = 0u;
        if (tid == 0) { sh[0] = INF; sh[1] = INF; }
        __syncthreads();
        for (int i = tid; i < m; i += kOsdThreads) {
            uint32_t par = 0;
            for (int w = 0; w < nw; ++w) par ^= A[i * rw + w] & e[w];
            A[i * rw + nw] = (__popc(par) & 1u) ^ ((s[i >> 5] >> (i & 31)) & 1u);    // residual s + H e
        }
        // ---- column order
        if (a.perm) {
            const int32_t *pp = a.perm + (long long)item * n;
            for (int j = tid; j < n; j += kOsdThreads) perm[j] = pp[j];
        } else {
            for (int j = tid; j < n; j += kOsdThreads) {       // stable rank by counting
                const double rj = rel[j];
                int rank = 0;
                for (int i = 0; i < n; ++i) { const double ri = rel[i]; rank += (ri < rj) || (ri == rj && i < j); }
                perm[rank] = j;
            }
        }
        // ---- live words of the untouched matrix (buffer 0)
        {
            uint32_t acc[kOsdTrips] = {};
            for (int i = warp; i < m; i += kOsdWarps)
#pragma unroll
                for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; if (w < nw) acc[t] |= A[i * rw + w]; }
#pragma unroll
            for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; if (w < nw && acc[t]) atomicOr(&live[w], acc[t]); }
        }
        __syncthreads();
        // ---- forward elimination in the given column order
        int rank = 0, nun = m, cpos = 0, it = 0;
        int first_info = -1;                                     // first non-pivot column met (order-1 flip)
        while (rank < a.rank_h && cpos < n) {
            const uint32_t *lv = live + (it & 1) * rw;
            uint32_t *lv_next = live + ((it + 1) & 1) * rw;
            // next column, in order, that still has a 1 in an unused row (every warp computes the same result)
            int found = -1;
            while (cpos < n) {
                const int c = cpos + lane;
                bool bit = false;
                if (c < n) { const int col = perm[c]; bit = (lv[col >> 5] >> (col & 31)) & 1u; }
                const uint32_t valid = (n - cpos >= 32) ? full : ((1u << (n - cpos)) - 1u);
                const uint32_t bal = __ballot_sync(full, bit);
                const uint32_t clear_before = ~bal & valid & (bal ? ((1u << (__ffs(bal) - 1)) - 1u) : full);
                if (first_info < 0 && clear_before) first_info = perm[cpos + __ffs(clear_before) - 1];
                if (bal) { found = cpos + __ffs(bal) - 1; break; }
                cpos += 32;
            }
            if (found < 0) break;
            const int col = perm[found];
            const int cw = col >> 5;
            const uint32_t cb = 1u << (col & 31);
            // pivot: any unused row with a 1 in this column (the solution does not depend on the choice); lowest list position
            for (int x0 = 0; x0 < nun; x0 += kOsdThreads) {
                const int x = x0 + tid;
                const bool hit = (x < nun) && (A[rows[x < nun ? x : 0] + cw] & cb);
                const uint32_t bal = __ballot_sync(full, hit);
                if (bal && lane == 0) atomicMin(&sh[it & 1], x0 + (tid & ~31) + (__ffs(bal) - 1));
            }
            if (tid < rw) lv_next[tid] = 0u;                     // rw <= 64 < kOsdThreads; nobody reads this buffer any more
            __syncthreads();
            const int pidx = sh[it & 1];
            if (pidx == INF) break;                              // cannot happen: a set live bit means an unused row holds the column
            const int prow = rows[pidx];                         // word offset of the pivot row
            uint32_t pw[kOsdTrips], acc[kOsdTrips] = {};
#pragma unroll
            for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; pw[t] = (w < rw) ? A[prow + w] : 0u; }
            const int cl = cw & 31, ct = cw >> 5;                // lane / trip that holds the pivot word of a row
            // two rows per trip of the loop (independent loads in flight); a row is updated only if it holds the pivot bit,
            // which its own pivot word tells (one shuffle, no extra load)
            for (int x = warp; x < nun; x += 2 * kOsdWarps) {
                const int x2 = x + kOsdWarps;
                const bool ok1 = x != pidx, ok2 = x2 < nun && x2 != pidx;          // warp-uniform
                const int i1 = rows[x], i2 = rows[x2 < nun ? x2 : x];
                uint32_t v1[kOsdTrips], v2[kOsdTrips];
#pragma unroll
                for (int t = 0; t < kOsdTrips; ++t) {
                    const int w = lane + 32 * t;
                    const bool in = (t + 1 < kOsdTrips) || (w < rw);
                    v1[t] = in ? A[i1 + w] : 0u;
                    v2[t] = in ? A[i2 + w] : 0u;
                }
                uint32_t k1 = v1[0], k2 = v2[0];
#pragma unroll
                for (int t = 1; t < kOsdTrips; ++t) { k1 = (ct == t) ? v1[t] : k1; k2 = (ct == t) ? v2[t] : k2; }
                const bool has1 = ok1 && (__shfl_sync(full, k1, cl) & cb);
                const bool has2 = ok2 && (__shfl_sync(full, k2, cl) & cb);
#pragma unroll
                for (int t = 0; t < kOsdTrips; ++t) {
                    const int w = lane + 32 * t;
                    const bool in = (t + 1 < kOsdTrips) || (w < rw);
                    if (has1) { v1[t] ^= pw[t]; if (in) A[i1 + w] = v1[t]; }
                    if (has2) { v2[t] ^= pw[t]; if (in) A[i2 + w] = v2[t]; }
                    const bool data = (t + 1 < kOsdTrips) ? true : (w < nw);       // the right-hand-side word is not a column
                    if (data) acc[t] |= (ok1 ? v1[t] : 0u) | (ok2 ? v2[t] : 0u);
                }
            }
#pragma unroll
            for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; if (w < nw && acc[t]) atomicOr(&lv_next[w], acc[t]); }
            if (tid == 0) {
                rows[pidx] = rows[nun - 1];                       // entry pidx is skipped by every reader of this pass
                pcol[rank] = col; prowl[rank] = prow; pivrow[col] = prow / rw;
                sh[(it + 1) & 1] = INF;
            }
            __syncthreads();
            ++rank; --nun; ++it; cpos = found + 1;
        }
        // remaining columns (the loop stops once rank(H) columns are kept) are information columns
        if (first_info < 0) {
            for (int c = 0; c < n; ++c) if (pivrow[perm[c]] < 0) { first_info = perm[c]; break; }
        }
        // ---- order 1: flip the first information bit (App. B-8) and move its column to the right-hand side
        if (a.order == 1 && first_info >= 0) {
            const int cw = first_info >> 5;
            const uint32_t cb = 1u << (first_info & 31);
            for (int i = tid; i < m; i += kOsdThreads) if (A[i * rw + cw] & cb) A[i * rw + nw] ^= 1u;
            if (tid == 0) e[cw] ^= cb;
        }
        __syncthreads();
        // ---- back-substitution (warp 0): pivot r's row holds 1s only in its own column, in LATER pivot columns and in
        // information columns (whose flips are 0): d_col = rhs + <row, d>
        if (warp == 0) {
            uint32_t d[kOsdTrips] = {};
            for (int r = rank - 1; r >= 0; --r) {
                const int pr = prowl[r], col = pcol[r];
                uint32_t par = 0;
#pragma unroll
                for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; if (w < nw) par ^= A[pr + w] & d[t]; }
                const uint32_t odd = __popc(__ballot_sync(full, __popc(par) & 1u)) & 1u;
                const uint32_t bit = (A[pr + nw] & 1u) ^ odd;
                const int cw = col >> 5;
#pragma unroll
                for (int t = 0; t < kOsdTrips; ++t) if (bit && lane == (cw & 31) && (cw >> 5) == t) d[t] |= 1u << (col & 31);
            }
#pragma unroll
            for (int t = 0; t < kOsdTrips; ++t) { const int w = lane + 32 * t; if (w < nw) e[w] ^= d[t]; }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------
// osd_reg_kernel<W>: Gauss-Jordan with the matrix in REGISTERS.
//   * thread i owns row i of H, permuted into reliability order (bit c' of the row = H[i][perm[c']]), as W 32-bit registers
//     plus the right-hand-side bit; the rows are built from the CSR lists (8 entries per row on the lifted-product codes)
//     through the inverse permutation, not by gathering n bits;
//   * column order: bitonic sort of (reliability bits, index) in shared memory -- the stable order of the library (a
//     caller-supplied permutation replaces it);
//   * the column walk is unrolled over the W words of a row, so every register index is static.  Inside word k: the OR over the
//     unused rows of (word k & columns not yet examined) -- a warp REDUX, one shared word per warp, barrier -- tells every
//     thread the next pivot column (lowest set bit; the clear bits below it are information columns) or that the word is
//     exhausted; the pivot row is the first unused row holding that bit (first such warp, first such lane); it publishes its
//     words k..W-1 and its right-hand side, barrier, and EVERY other row holding the bit -- used or not: Gauss-Jordan, so no
//     back-substitution -- XORs them in.  Words below k are never needed again (only pivot columns and the right-hand side
//     enter the solution), so the work per pivot shrinks as the walk advances;
//   * two CTA barriers and ~2 (W - k) + 25 instructions per thread and pivot, against ~40 instructions per ROW and pivot in
//     the shared-memory formulation: 10 x fewer instructions per solve, 10 x shorter latency per solve;
//   * solution: pivot row p gives flip(perm[col_p]) = rhs_p (+ bit of the first information column for order 1, App. B-8).
// Same unique solution as any elimination given the column order (bit-exact with the oracle and the reference goldens).
// ---------------------------------------------------------------------------------------------------------
constexpr int kOsdRegMaxThreads = 512;

inline size_t osd_reg_smem_bytes(int n, int nw, int W)
{
    int np = 1;
    while (np < n) np <<= 1;
    size_t b = (size_t)np * 8;            // sort keys
    b += (size_t)np * 4;                  // sorted indices = perm
    b += (((size_t)n * 2 + 15) & ~size_t(15));   // inverse permutation (u16)
    b += (size_t)(W + 4) * 4;             // published pivot row + right-hand side
    b += 32 * 4;                          // per-warp OR words
    b += (size_t)nw * 4;                  // the shot's estimate words
    return b + 64;
}

template <int W>
struct OsdWalk {
    uint32_t row[W];      // my row, permuted column order
    uint32_t rhs;         // right-hand-side bit
    bool used;            // my row is a pivot row (or lies beyond m)
    int mycol;            // permuted column my row is the pivot of, or -1
    int rank, first_info, last_piv;
    uint32_t fbit;        // my row's bit in the first information column
};

// Word K of the column walk (see osd_reg_kernel); recursion on K keeps every register index static.
template <int W, int K>
__device__ __forceinline__ void osd_walk(OsdWalk<W> &st, int rank_h, int n, int lane, int warp, int nwarps, uint32_t *P, uint32_t *wor)
{
    static_assert(W % 4 == 0, "rows are published and read back as 128-bit vectors");
    const unsigned full = 0xffffffffu;
    constexpr int K4 = K & ~3;                                                // first word of the vector that holds word K
    if (st.rank < rank_h && 32 * K < n) {
        uint32_t cmask = (n - 32 * K >= 32) ? full : ((1u << (n - 32 * K)) - 1u);      // columns of this word not examined yet
        while (st.rank < rank_h) {
            const uint32_t wo = __reduce_or_sync(full, st.used ? 0u : (st.row[K] & cmask));
            if (lane == 0) wor[warp] = wo;
            __syncthreads();
            const uint32_t wv = lane < nwarps ? wor[lane] : 0u;              // lane x <-> warp x (at most 16 warps)
            const uint32_t live = __reduce_or_sync(full, wv);
            const uint32_t below = live ? (cmask & ((live & (0u - live)) - 1u)) : cmask;   // information columns passed over
            if (st.first_info < 0 && below) {
                const int c0 = __ffs(below) - 1;
                st.first_info = 32 * K + c0;
                st.fbit = (st.row[K] >> c0) & 1u;
            }
            if (!live) { __syncthreads(); break; }                          // word exhausted (wor is rewritten after this barrier)
            const int c = __ffs(live) - 1;
            const uint32_t bit = 1u << c;
            const int pw = __ffs(__ballot_sync(full, (wv & bit) != 0u)) - 1;   // first warp with an unused row holding the column
            const bool has = (st.row[K] & bit) != 0u;
            const uint32_t cand = __ballot_sync(full, has && !st.used);
            const bool is_piv = warp == pw && has && !st.used && lane == __ffs(cand) - 1;
            if (is_piv) {
#pragma unroll
                for (int w = K4; w < W; w += 4) *reinterpret_cast<uint4 *>(P + w) = make_uint4(st.row[w], st.row[w + 1], st.row[w + 2], st.row[w + 3]);
                P[W] = st.rhs;
                st.used = true;
                st.mycol = 32 * K + c;
            }
            __syncthreads();
            if (has && !is_piv) {                                            // words K4 .. K-1 are dead: XOR-ing them as well is harmless
#pragma unroll
                for (int w = K4; w < W; w += 4) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(P + w);
                    st.row[w] ^= q.x; st.row[w + 1] ^= q.y; st.row[w + 2] ^= q.z; st.row[w + 3] ^= q.w;
                }
                st.rhs ^= P[W];
            }
            cmask &= ~(bit | (bit - 1u));
            st.last_piv = 32 * K + c;
            ++st.rank;
        }
    }
    if constexpr (K + 1 < W) osd_walk<W, K + 1>(st, rank_h, n, lane, warp, nwarps, P, wor);
}

template <int W, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) osd_reg_kernel(OsdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int m = a.m, n = a.n, nw = a.nw;
    int np = 1;
    while (np < n) np <<= 1;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(smem);
    int *perm = reinterpret_cast<int *>(smem + (size_t)np * 8);
    uint16_t *pos = reinterpret_cast<uint16_t *>(smem + (size_t)np * 12);
    uint32_t *P = reinterpret_cast<uint32_t *>(smem + (size_t)np * 12 + (((size_t)n * 2 + 15) & ~size_t(15)));   // 16-byte aligned
    uint32_t *wor = P + (W + 4);
    uint32_t *eb = wor + 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nwarps = T >> 5;
    const unsigned full = 0xffffffffu;
    const int count = a.count_dev ? min(*a.count_dev, a.count) : a.count;
    const bool valid = tid < m;

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const long long shot = a.shot_ids ? a.shot_ids[item] : item;
        uint32_t *e = a.ehat + shot * nw;
        const uint32_t *s = a.syn + shot * a.mw;
        // ---- column order
        if (a.perm) {
            const int32_t *pp = a.perm + (long long)item * n;
            for (int j = tid; j < n; j += T) perm[j] = pp[j];
        } else {
            const double *llr = a.llr + (long long)item * n;
            for (int j = tid; j < np; j += T) {
                key[j] = j < n ? (unsigned long long)__double_as_longlong(osd_reliability(llr[j])) : ~0ull;   // rel in [0.5, 1]: bit order = value order
                perm[j] = j;
            }
            __syncthreads();
            for (int k2 = 2; k2 <= np; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (np >> 1); t += T) {
                        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                        const unsigned long long ka = key[lo], kb = key[hi];
                        const int ia = perm[lo], ib = perm[hi];
                        const bool gt = ka > kb || (ka == kb && ia > ib);          // ties by index: the stable order
                        if (gt == ((lo & k2) == 0)) { key[lo] = kb; key[hi] = ka; perm[lo] = ib; perm[hi] = ia; }
                    }
                    __syncthreads();
                }
        }
        for (int w = tid; w < nw; w += T) eb[w] = e[w];
        __syncthreads();
        for (int c = tid; c < n; c += T) pos[perm[c]] = (uint16_t)c;
        __syncthreads();
        // ---- my row in permuted column order, residual right-hand side s + H e
        uint32_t row[W];
#pragma unroll
        for (int w = 0; w < W; ++w) row[w] = 0u;
        uint32_t rhs = 0;
        if (valid) {
            rhs = (s[tid >> 5] >> (tid & 31)) & 1u;
            for (int x = a.row_ptr[tid]; x < a.row_ptr[tid + 1]; ++x) {
                const int j = a.col_idx[x];
                rhs ^= (eb[j >> 5] >> (j & 31)) & 1u;
                const uint32_t c = pos[j], cw = c >> 5, bit = 1u << (c & 31u);
#pragma unroll
                for (int w = 0; w < W; ++w) row[w] |= (cw == (uint32_t)w) ? bit : 0u;
            }
        }
        // ---- Gauss-Jordan over the columns in order (osd_walk: one instantiation per word of a row, static register indices)
        OsdWalk<W> st;
#pragma unroll
        for (int w = 0; w < W; ++w) st.row[w] = row[w];
        st.rhs = rhs; st.used = !valid; st.mycol = -1; st.rank = 0; st.first_info = -1; st.last_piv = -1; st.fbit = 0;
        osd_walk<W, 0>(st, a.rank_h, n, lane, warp, nwarps, P, wor);
#pragma unroll
        for (int w = 0; w < W; ++w) row[w] = st.row[w];
        rhs = st.rhs;
        const int mycol = st.mycol, first_info = st.first_info, last_piv = st.last_piv;
        uint32_t fbit = st.fbit;
        // ---- order 1: the first information column is flipped (App. B-8)
        if (a.order == 1) {
            int f = first_info;
            if (f < 0 && last_piv + 1 < n) {                                  // no column was passed over: the one after the last pivot
                f = last_piv + 1;
                uint32_t wsel = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) wsel = ((f >> 5) == w) ? row[w] : wsel;
                fbit = (wsel >> (f & 31)) & 1u;
            }
            if (f >= 0) {
                if (mycol >= 0) rhs ^= fbit;
                if (tid == 0) atomicXor(&eb[perm[f] >> 5], 1u << (perm[f] & 31));
            }
        }
        // ---- solution: the flip of a pivot column is the right-hand side of its row
        if (mycol >= 0 && (rhs & 1u)) { const int col = perm[mycol]; atomicXor(&eb[col >> 5], 1u << (col & 31)); }
        __syncthreads();
        for (int w = tid; w < nw; w += T) e[w] = eb[w];
        __syncthreads();
    }
}

inline int osd_launch(const OsdArgs &a, int sm_count, cudaStream_t st)
{
    static const int force_old = [] { const char *ev = getenv("QLDPC_OSD_KERNEL"); return ev && ev[0] == 's' ? 1 : 0; }();   // 's': shared-memory kernel (tests)
    const int threads = ((a.m + 31) / 32) * 32;
    if (!force_old && threads <= kOsdRegMaxThreads && a.nw <= 36 && a.row_ptr) {
        // instances: (words per row, launch bound): small rows run four CTAs of 256 threads per SM, the 36-word rows of LP118_2 /
        // Tanner (450 / 465 checks) one CTA of up to 512
        const int W = (a.nw <= 8 && threads <= 256) ? 8 : ((a.nw <= 20 && threads <= 256) ? 20 : 36);
        void (*fn)(OsdArgs) = W == 8 ? osd_reg_kernel<8, 256, 6> : (W == 20 ? osd_reg_kernel<20, 256, 4> : osd_reg_kernel<36, 512, 1>);
        const size_t smem = osd_reg_smem_bytes(a.n, a.nw, W);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
        if (e != cudaSuccess) return (int)e;
        const int grid = std::max(1, std::min(a.count, sm_count * std::max(1, per_sm)));
        fn<<<grid, threads, smem, st>>>(a);
        return (int)cudaGetLastError();
    }
    const size_t smem = osd_smem_bytes(a.m, a.n, a.nw);
    if (smem > (size_t)kMaxSmemPerCta || a.nw + 1 > 32 * kOsdMaxTrips) return (int)cudaErrorInvalidValue;
    void (*fn)(OsdArgs) = (a.nw + 1 <= 32) ? osd_kernel<1> : osd_kernel<2>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta);
    if (e != cudaSuccess) return (int)e;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)kMaxSmemPerCta / smem));
    const int grid = std::max(1, std::min(a.count, sm_count * per_sm));
    fn<<<grid, kOsdThreads, smem, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace qldpc
