// common.cuh -- shared definitions of the qldpc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/qldpc_b200.h"

namespace qldpc {

constexpr int kWarp = 32;
constexpr uint16_t kPad = 0xFFFF;          // padding entry of the slot-major variable table
constexpr int kMaxSmemPerCta = 227 * 1024; // B200: 227 KB opt-in dynamic shared memory per CTA

// Device view of the graph tables of the SUM-PRODUCT kernel (the min-sum kernel has its own, MsTables in ms_kernel.cuh;
// m, n, E, mw, nw, nl, dv are filled for every plan).  All tables live in one uint16 blob that every CTA copies into shared
// memory once; offsets are in uint16 units.  Edge storage is SLOT-MAJOR: the k-th edge (ascending variable)
// of check i sits at position k*m + i, so that a warp whose lanes hold consecutive checks touches
// consecutive shared-memory words (no bank conflicts), and -- for circulant-lifted codes -- consecutive
// variables as well.
struct Tables {
    int m, n, E;
    int dc, dv;          // max row / column weight
    int nl;              // number of layers
    int mw, nw;          // words(m), words(n)
    int off_var;         // [dc*m]   BYTE offset (4*j) of the variable of (slot k, check i); kPad past the end of a short row
    int off_col_ptr;     // [n+1]
    int off_col_pos;     // [E]      slot-major positions of the edges of variable j, ascending check
    int off_col_chk;     // [E]      check of those edges
    int off_layer_ptr;   // [nl+1]
    int off_layer_chk;   // [layer_ptr[nl]]
    int off_lvar_ptr;    // [nl+1]
    int off_lvar_idx;    // [lvar_ptr[nl]]  sorted distinct variables adjacent to the checks of layer l
    int off_colf;        // [n+1][dv] the same positions with a fixed stride of dv entries per variable (columns of fewer than 8
                         //          checks: sum-product column sums without pointer loads); past the degree -- and for the dummy
                         //          variable n -- the always-zero word dc*ms of the c2v array
    int off_rowpar;      // [2*mw]   parity of the row weights as bit words (low half, high half)
    int n_pad;           // n rounded up to a multiple of 64 (first-step variable sweep, two variables per lane and trip)
    int ms;              // slot stride of the slot-major edge arrays (>= m; padded so that the lane groups of the
                         // min-sum check phase fall on disjoint shared-memory banks)
    int len;             // blob length in uint16 units (padded to a multiple of 8)
};

struct DecodeIO {
    const uint32_t *syn;     // [shots][mw]
    uint32_t *ehat;          // [shots][nw]
    int32_t *iters;          // [shots]
    uint8_t *conv;           // [shots] or null
    double *llr;             // [shots][n] or null
    long long shots;
    unsigned long long *work_counter;   // global shot dispenser (persistent CTAs pull the next shot)
    // compacted list of unconverged shots (for OSD); null when unused
    int *fail_count;
    int *fail_shot;          // [fail_cap]
    double *fail_llr;        // [fail_cap][n]
    int fail_cap;
    unsigned long long *work_done;      // min-sum: += check-to-variable messages actually computed (one add per warp at kernel exit)
};

struct MsConst {
    double L;        // prior LLR, binary64 (decoders.py:147)
    double Lf;       // (double)(float)L : the prior as first stored into the binary32 v2c array (decoders.py:148-149)
    double beta;     // normalisation (decoders.py:115)
    double abeta;    // |beta|
    uint32_t sgn;    // 0x80000000 if beta < 0 (extra sign of every check-to-variable message), else 0
    float Tf;        // -L rounded UP to binary32: fl64(L + S) < 0  <=>  S < Tf for binary32 S
    int max_iter;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// bit-packed column arrays (columns of H, of the logical bases): one 128-byte line per column, of which the classifier reads
// the first col_words(words) words with 1, 2, 4 or 8 128-bit loads
constexpr int kColStride = 32;
__host__ __device__ inline int col_words(int words) { return words <= 4 ? 4 : (words <= 8 ? 8 : (words <= 16 ? 16 : 32)); }

}  // namespace qldpc

// Host-side plan object (opaque to C callers).
struct qldpc_plan {
    int device = 0;
    qldpc_opts opts{};
    qldpc::Tables tab{};
    std::vector<uint16_t> h_blob;
    // host copies of the graph (CSR / CSC) for the kernels that want int32 tables in global memory
    std::vector<int32_t> row_ptr, col_idx, col_ptr, row_idx, layer_ptr, layer_chk;
    uint16_t *d_blob = nullptr;
    int32_t *d_row_ptr = nullptr, *d_col_idx = nullptr, *d_col_ptr = nullptr, *d_row_idx = nullptr;
    uint32_t *d_hbits = nullptr;   // [m][nw] bit-packed rows of H (OSD, sampler, classification)
    uint32_t *d_hcol = nullptr;    // [n][kColStride] bit-packed columns of H (syndrome of a sparse vector), zero padded
    uint32_t *d_lcol = nullptr;    // [n][kColStride] bit-packed columns of the attached logical-operator basis (or null)
    int logical_k = 0, lkw = 0;
    unsigned long long *d_work = nullptr;
    int *d_fail_count = nullptr;
    unsigned long long *d_work_done = nullptr;   // executed edge updates of the min-sum kernel since creation / the last reset
    int sm_count = 0;
    int rank_h = 0;                // GF(2) rank of H
    int row_w = 0;                 // true max row weight (tab.dc may be padded to the instantiated kernel shape)
    int grid = 0, threads = 0, shots_per_cta = 0;
    size_t smem_bytes = 0;
    size_t state_bytes = 0;        // per-shot shared-memory state
    long long plan_wavefronts = 0, plan_wavefronts_ideal = 0;   // min-sum layout planner: modelled / conflict-free wavefronts per iteration
    // scratch for qldpc_decode_host / OSD
    void *scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t events[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t pinned_bytes[4] = {0, 0, 0, 0};
};
