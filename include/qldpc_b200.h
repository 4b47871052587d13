/*
 * qldpc_b200.h -- C ABI of the B200-native batched quantum-LDPC decoder library (libqldpc_b200.so).
 *
 * The reference (albertogp71/qLDPCsim v0.2.2) is pure Python and has no FFI; its boundary for this path is
 * "module decoders, five free functions, called once per shot from simulator.py:270-284".  This header is
 * the batched, device-side replacement of exactly that boundary: one plan per (parity-check matrix, decoder
 * configuration), one call per batch of shots.  Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative QLDPC_E* code and never throws;
 *     qldpc_last_error() gives the message of the last failure on the calling thread.
 *   - "dev" pointers are CUDA device pointers on the plan's device, owned by the caller (e.g. PyTorch
 *     tensors); work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = default stream) and
 *     the call does not synchronise unless stated.  "host" pointers are ordinary host memory.
 *   - bit-packed rows: row r of a 0/1 matrix with `cols` columns is qldpc_words(cols) little-endian
 *     uint32 words, bit j of word w = column 32*w + j, padding bits zero.
 *   - a plan is immutable after creation except for its internal work counters, so calls on one plan must
 *     be issued from one thread/stream at a time.
 */
#ifndef QLDPC_B200_H
#define QLDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLDPC_ABI_VERSION 2

enum { QLDPC_OK = 0, QLDPC_EINVAL = -1, QLDPC_ECUDA = -2, QLDPC_ETOOBIG = -3, QLDPC_ENOMEM = -4 };

/* decType of simulator.py:270-284 */
enum { QLDPC_NG = 0, QLDPC_BF = 1, QLDPC_MS = 2, QLDPC_BP = 3 };

/* Outcome counters of simulator.py:238-243 / :308-315, in this order. */
enum {
    QLDPC_CNT_FAIL_X = 0,   /* DecFailures_X      simulator.py:300-301 */
    QLDPC_CNT_FAIL_Z = 1,   /* DecFailures_Z      simulator.py:302-303 */
    QLDPC_CNT_EXACT  = 2,   /* decSuccessExact    simulator.py:294-295 */
    QLDPC_CNT_DEGEN  = 3,   /* decSuccessDegen    simulator.py:296-298 (integer products, no mod 2) */
    QLDPC_CNT_ITERS_X = 4,  /* sum of nIterX      simulator.py:291 */
    QLDPC_CNT_ITERS_Z = 5,  /* sum of nIterZ      simulator.py:292 */
    QLDPC_CNT_SHOTS  = 6,
    /* Extension: the four outcome classes of README.md:15-22, which the reference's own
     * counters do not implement (its "degenerate" test is vacuous, simulator.py:296-298).  Filled only when both
     * plans carry logical operators (qldpc_plan_set_logicals); shots = exact + true_degen + logical + fail_any. */
    QLDPC_CNT_TRUE_DEGEN = 7, /* both syndromes reproduced, residual is a stabiliser, not an exact match   */
    QLDPC_CNT_LOGICAL = 8,    /* both syndromes reproduced, residual acts on the logical qubits            */
    QLDPC_CNT_FAIL_ANY = 9,   /* at least one of the two syndromes not reproduced ("decoder failure")      */
    QLDPC_NUM_COUNTERS = 10
};

/* Tanner graph + check schedule, host memory, as produced by the PCM compiler (qldpcsim_b200/pcm.py).
 * Replaces the dense `H` and `layers` arguments of decoders.py:27, :74, :110-117, :189-195. */
typedef struct {
    int32_t m, n;               /* H is m x n                                                           */
    int32_t nnz;                /* number of ones                                                       */
    const int32_t *row_ptr;     /* CSR, m+1 entries                                                     */
    const int32_t *col_idx;     /* CSR, nnz entries, ascending within a row (== np.where(H) order)      */
    int32_t n_layers;           /* `layers` of MS/BP (decoders.py:114, :193); ignored by NG/BF          */
    const int32_t *layer_ptr;   /* n_layers+1 entries                                                   */
    const int32_t *layer_chk;   /* check indices, layer by layer, in list order                         */
} qldpc_graph;

/* Decoder configuration: the keyword arguments of the reference decoders. */
typedef struct {
    int32_t dec_type;           /* QLDPC_NG / BF / MS / BP                                              */
    int32_t max_iter;           /* max_iter (decoders.py:74, :113, :192); NG ignores it (2n steps, :47) */
    double  prior_llr;          /* log((1-p)/max(p,eps)) as float64 (decoders.py:147, :232)             */
    double  beta;               /* MS normalisation (decoders.py:115), 0.75 in the driver               */
    double  eps;                /* BP clamp shift (decoders.py:195, :257-258)                           */
    int32_t osd_order;          /* OSDorder (decoders.py:116, :194); < 0 disables                       */
    int32_t reserved;           /* min-sum A/B switches: bit 0 never merge runs of disjoint layers, bit 1 no 8-lane kernel */
} qldpc_opts;

typedef struct qldpc_plan qldpc_plan;

int         qldpc_abi_version(void);
const char *qldpc_last_error(void);
int         qldpc_words(int32_t nbits);             /* ceil(nbits/32) */

/* Builds the device-resident plan (CSC, the decoder's shared-memory tables -- for min-sum the variable-major message
 * layout chosen by the layout planner, for sum-product slot-major edge tables --, per-layer variable lists, bit-packed
 * rows and columns of H, launch geometry).
 * `device` is a CUDA ordinal. */
int qldpc_plan_create(const qldpc_graph *graph, const qldpc_opts *opts, int device, qldpc_plan **out);
int qldpc_plan_destroy(qldpc_plan *plan);

/* Plan introspection: what = 0 m, 1 n, 2 nnz, 3 n_layers, 4 CTAs launched, 5 threads per CTA,
 * 6 dynamic shared memory bytes per CTA, 7 shots resident per CTA, 8 max row weight, 9 max column weight,
 * 10 GF(2) rank of H (gf2math.rank, gf2math.py:91-135), 11 reserved (0),
 * 12 / 13 shared-memory wavefronts per iteration modelled by the min-sum layout planner / its conflict-free bound,
 * 14 bytes of shared-memory state per shot, 15 number of attached logical operators, 16 steps per iteration of the min-sum
 * kernel (< n_layers when runs of layers with disjoint variable sets were merged into one step), 17 warps per shot. */
int64_t qldpc_plan_info(const qldpc_plan *plan, int what);

/* Executed-work counter of a min-sum plan: check-to-variable messages actually computed by the kernel since plan creation or the
 * last reset (a decode that converges in the middle of an iteration, decoders.py:175-176, executes fewer than E x iterations;
 * bench.py states the roofline fraction of executed work beside the one of SURVEY's whole-iteration accounting).  Synchronises
 * the device.  Returns -1 on error, 0 for other decoders. */
int64_t qldpc_plan_work(qldpc_plan *plan, int reset);

/* Decode `shots` syndromes.  Replaces the per-shot calls NG_decoder / BF_decoder / MS_decoder / BP_decoder
 * (decoders.py:27, :74, :110, :189 as driven by simulator.py:270-284).
 *   syn_bits  dev  [shots][qldpc_words(m)]   in   syndromes, bit-packed
 *   ehat_bits dev  [shots][qldpc_words(n)]   out  error estimates, bit-packed
 *   iters     dev  [shots] int32             out  second return value of the reference decoder
 *   converged dev  [shots] uint8 or NULL     out  1 if the decoder returned from its early-exit branch
 *                                                 (decoders.py:99-100, :175-176, :283-285; for NG: residual
 *                                                 reached zero, decoders.py:49)
 *   llr_out   dev  [shots][n] float64 or NULL out posterior LLRs of the last layer step (MS/BP), the
 *                                                 vector the reference hands to OSDdec (decoders.py:179-180)
 * OSD (opts.osd_order >= 0) is applied by this call to the shots that did not converge. */
int qldpc_decode(qldpc_plan *plan, const uint32_t *syn_bits, int64_t shots, uint32_t *ehat_bits,
                 int32_t *iters, uint8_t *converged, double *llr_out, void *stream);

/* Same, with HOST buffers: chunks the batch, overlaps host->device copies, kernels and device->host copies
 * on internal streams, and returns when the outputs are in host memory.  llr_out may be NULL. */
int qldpc_decode_host(qldpc_plan *plan, const uint32_t *syn_bits, int64_t shots, uint32_t *ehat_bits,
                      int32_t *iters, uint8_t *converged, double *llr_out);

/* Ordered-statistics post-processing of a batch (decoders.py:299-370 with gf2math.py:91-187).
 *   ehat_bits dev [shots][words(n)]  in/out (mutated in place like the reference, decoders.py:368)
 *   syn_bits  dev [shots][words(m)]  in
 *   llr       dev [shots][n] float64 in     posteriorLLRs
 *   perm      dev [shots][n] int32 or NULL  column order to use instead of the library's own stable sort by
 *                                           (reliability, index) -- see DESIGN.md on NumPy's unstable argsort */
int qldpc_osd(qldpc_plan *plan, uint32_t *ehat_bits, const uint32_t *syn_bits, const double *llr,
              const int32_t *perm, int64_t shots, int32_t order, void *stream);

/* Per-shot outcome classification and counter reduction (simulator.py:291-303).
 *   plan_x : plan built from Hz (it produced the X-error estimate), plan_z : plan built from Hx.
 *   err*_bits / ehat*_bits dev [shots][words(n)], syn_z_bits dev [shots][words(m_z)], syn_x_bits [shots][words(m_x)]
 *   counters dev int64[QLDPC_NUM_COUNTERS]: ACCUMULATED into (zero it first). */
int qldpc_classify(const qldpc_plan *plan_x, const qldpc_plan *plan_z,
                   const uint32_t *errx_bits, const uint32_t *errz_bits,
                   const uint32_t *ehatx_bits, const uint32_t *ehatz_bits,
                   const uint32_t *syn_z_bits, const uint32_t *syn_x_bits,
                   const int32_t *iters_x, const int32_t *iters_z,
                   int64_t shots, int64_t *counters, void *stream);

/* Attach a basis of logical operators to a plan (extension; no counterpart in the reference,
 * whose README.md:15-22 defines the classes but whose simulator.py:296-298 cannot tell them apart).
 *   logical_rows host [k][words(n)] bit-packed.  For the plan built from Hz (X-error decode) pass a basis of the logical
 *   Z operators -- ker(Hx) modulo rowspace(Hz); for the plan built from Hx pass the logical X operators.  With both
 *   attached, qldpc_classify also fills QLDPC_CNT_TRUE_DEGEN / _LOGICAL (it always fills _FAIL_ANY).  k = 0 detaches. */
int qldpc_plan_set_logicals(qldpc_plan *plan, const uint32_t *logical_rows, int32_t k);

/* The shot loop of simulator.py:244-304 on HOST buffers (pinned or pageable): the bit-packed measurement record -- its four
 * column groups [sy_z | sy_x | errX | errZ] (simulator.py:249-252) as four arrays -- in, the counters of qldpc_classify out
 * (host int64[QLDPC_NUM_COUNTERS], overwritten).  plan_x decodes the X errors (built from Hz), plan_z the Z errors (from Hx).
 * Host-to-device copies of one chunk overlap the decodes and the classification of the previous one; blocks until done. */
int qldpc_simulate_host(qldpc_plan *plan_x, qldpc_plan *plan_z, const uint32_t *syn_z_bits, const uint32_t *syn_x_bits,
                        const uint32_t *errx_bits, const uint32_t *errz_bits, int64_t shots, int64_t *counters);

/* On-device depolarizing sampler + syndrome generator: the stand-in for Stim's sampler
 * (simulator.py:43-160, :196-197).  Qubit q of global shot s draws u from a counter-based generator keyed by
 * (seed, s, q), so the batch is independent of how shots are sharded over GPUs.
 *   X: u < p/3, Y: p/3 <= u < 2p/3, Z: 2p/3 <= u < p; errX = X|Y, errZ = Z|Y;
 *   syn_z = Hz errX, syn_x = Hx errZ (mod 2).  All outputs bit-packed, dev. */
int qldpc_sample(const qldpc_plan *plan_x, const qldpc_plan *plan_z, double p, uint64_t seed,
                 int64_t first_shot, int64_t shots,
                 uint32_t *errx_bits, uint32_t *errz_bits, uint32_t *syn_z_bits, uint32_t *syn_x_bits,
                 void *stream);

/* Number of kernels this library has launched since load (for bench accounting). */
int64_t qldpc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_B200_H */
