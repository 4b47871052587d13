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
// Implementation: the matrix is NOT permuted; rows stay bit-packed in original column numbering with one
// extra word for the right-hand side, initialised to the residual syndrome s + H e.  Solving for the basis
// FLIPS d_J (H_J d_J = residual) is equivalent and needs no knowledge of I before the elimination.  Pivot
// search uses warp ballots over the candidate rows; the row updates run over (row, word) pairs.
#pragma once
#include "common.cuh"

namespace qldpc {

struct OsdArgs {
    int m, n, mw, nw;
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

// dynamic shared memory: A [m][nw+1] uint32 | rel [n] double | perm [n] int | used [m] uint8 (as uint32 words)
// | pivrow [n] int16-ish (int)   -- sized by osd_smem_bytes()
inline size_t osd_smem_bytes(int m, int n, int nw)
{
    size_t b = (size_t)m * (nw + 1) * 4;
    b = (b + 7) & ~size_t(7);
    b += (size_t)n * 8;      // rel
    b += (size_t)n * 4;      // perm
    b += (size_t)n * 4;      // pivot row of column (or -1)
    b += (size_t)m * 4;      // row used flag
    return b + 64;
}

__global__ void __launch_bounds__(kOsdThreads) osd_kernel(OsdArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int m = a.m, n = a.n, nw = a.nw, rw = a.nw + 1;   // rw: words per augmented row
    uint32_t *A = reinterpret_cast<uint32_t *>(smem);
    size_t off = ((size_t)m * rw * 4 + 7) & ~size_t(7);
    double *rel = reinterpret_cast<double *>(smem + off); off += (size_t)n * 8;
    int *perm = reinterpret_cast<int *>(smem + off); off += (size_t)n * 4;
    int *pivrow = reinterpret_cast<int *>(smem + off); off += (size_t)n * 4;
    int *used = reinterpret_cast<int *>(smem + off); off += (size_t)m * 4;
    int *sh = reinterpret_cast<int *>(smem + off);          // sh[0]: pivot row candidate, sh[1]: rank so far
    const int tid = threadIdx.x, lane = tid & 31;
    const int count = a.count_dev ? min(*a.count_dev, a.count) : a.count;

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
        for (int i = tid; i < m; i += kOsdThreads) used[i] = 0;
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
        if (tid == 0) { sh[1] = 0; }
        __syncthreads();
        // ---- Gauss-Jordan in the given column order
        int rank = 0;
        int first_info = -1;                                     // first non-pivot column met (order-1 flip)
        for (int cpos = 0; cpos < n && rank < a.rank_h; ++cpos) {
            const int col = perm[cpos];
            const int cw = col >> 5;
            const uint32_t cb = 1u << (col & 31);
            if (tid == 0) sh[0] = 0x7fffffff;
            __syncthreads();
            // pivot: lowest-index unused row with a 1 in this column (any choice gives the same solution)
            for (int i0 = 0; i0 < m; i0 += kOsdThreads) {
                const int i = i0 + tid;
                const bool hit = (i < m) && !used[i] && (A[i * rw + cw] & cb);
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (bal && lane == 0) atomicMin(&sh[0], i0 + (tid & ~31) + (__ffs(bal) - 1));
            }
            __syncthreads();
            const int prow = sh[0];
            if (prow == 0x7fffffff) {                            // dependent column -> information set
                if (first_info < 0) first_info = col;
                __syncthreads();
                continue;
            }
            // eliminate the column from every other row: (row, word) pairs, pivot row read-only
            for (int x = tid; x < m * rw; x += kOsdThreads) {
                const int i = x / rw, w = x - i * rw;
                if (i != prow && (A[i * rw + cw] & cb)) {
                    // the word holding the pivot bit is cleared last by the thread that owns it, so every
                    // thread of this row still sees the bit set: defer that word
                    if (w != cw) A[x] ^= A[prow * rw + w];
                }
            }
            __syncthreads();
            for (int i = tid; i < m; i += kOsdThreads)
                if (i != prow && (A[i * rw + cw] & cb)) A[i * rw + cw] ^= A[prow * rw + cw];
            if (tid == 0) { used[prow] = 1; pivrow[col] = prow; }
            ++rank;
            __syncthreads();
        }
        // remaining columns (the loop stops once rank(H) columns are kept) are information columns
        if (first_info < 0) {
            for (int cpos = 0; cpos < n; ++cpos) if (pivrow[perm[cpos]] < 0) { first_info = perm[cpos]; break; }
        }
        __syncthreads();
        // ---- order 1: flip the first information bit (App. B-8) and add its reduced column to the rhs
        if (a.order == 1 && first_info >= 0) {
            const int cw = first_info >> 5;
            const uint32_t cb = 1u << (first_info & 31);
            for (int i = tid; i < m; i += kOsdThreads) if (A[i * rw + cw] & cb) A[i * rw + nw] ^= 1u;
            if (tid == 0) e[cw] ^= cb;
            __syncthreads();
        }
        // ---- basis flips: d_col = rhs[pivot row of col]
        for (int w = tid; w < nw; w += kOsdThreads) {
            uint32_t flip = 0;
            for (int b = 0; b < 32; ++b) {
                const int j = w * 32 + b;
                if (j < n) { const int pr = pivrow[j]; if (pr >= 0 && (A[pr * rw + nw] & 1u)) flip |= 1u << b; }
            }
            e[w] ^= flip;
        }
        __syncthreads();
    }
}

inline int osd_launch(const OsdArgs &a, int sm_count, cudaStream_t st)
{
    const size_t smem = osd_smem_bytes(a.m, a.n, a.nw);
    if (smem > (size_t)kMaxSmemPerCta) return (int)cudaErrorInvalidValue;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(osd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)kMaxSmemPerCta / smem));
    const int grid = std::max(1, std::min(a.count, sm_count * per_sm));
    osd_kernel<<<grid, kOsdThreads, smem, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace qldpc
