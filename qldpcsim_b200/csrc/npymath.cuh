// npymath.cuh -- the two float64 routines NumPy 2.3.5 evaluates for np.tanh and np.arctanh (x86-64, AVX512_SKX dispatch), as
// device functions.  The reference's sum-product decoder (decoders.py:254, :256, :259) depends on the last bit of both, and
// neither is libm's nor correctly rounded (np.tanh differs from glibc's / CUDA's on a quarter of all arguments), so the kernel
// evaluates the same algorithms with the same constants:
//   npym_tanh    : NumPy's own SIMD kernel (numpy/_core/src/umath/loops_hyperbolic.dispatch.c.src, simd_tanh_f64) -- 16
//                  intervals chosen by exponent and leading mantissa bit, degree-16 polynomial in |x| - b by Horner with FMAs;
//   npym_arctanh : Intel SVML __svml_atanh8_ha as vendored by NumPy (umath/svml/linux/avx512/svml_z0_atanh_d_ha.s) -- two
//                  logarithms reduced by reciprocals rounded to 1+4 bits, 16-entry table, degree-10 log1p polynomial.
// Every operation is an individually rounded IEEE binary64 add / multiply / FMA, so the device results equal NumPy's bit for
// bit (tests: the kernel against 4800 decodes of the unmodified reference; the CPU twin of these functions against NumPy
// itself in tests/test_npymath.py).  Tables (npymath_tables.inc, data only) are staged in shared memory: the index varies
// per lane, and [coefficient][interval] rows of 16 doubles are conflict-free for distinct intervals.  Both functions are
// branch-free on the main path: ~20 + ~45 FP64 instructions against ~250 for CUDA's tanh + atanh.
#pragma once
#include <stdint.h>

namespace qldpc {

#define NPYM_TABLE static __device__ const
#include "npymath_tables.inc"
#undef NPYM_TABLE

struct NpymTables {              // shared-memory copy, 2880 bytes
    double tanh_lut[18 * 16];    // [b, c0 .. c16][interval]
    double atanh_T[16], atanh_U[16];
    uint32_t rcp_bucket[64];     // per 1/64 of the mantissa range: thresholds below it (low 8 bits) | the one inside it << 8
    uint32_t pad[16];
};

// all threads of the CTA; followed by a __syncthreads() of the caller
__device__ __forceinline__ void npym_stage_tables(NpymTables *s)
{
    for (int i = threadIdx.x; i < 18 * 16; i += blockDim.x) s->tanh_lut[i] = __longlong_as_double((long long)NPYM_TANH_LUT[i]);
    for (int i = threadIdx.x; i < 16; i += blockDim.x) {
        s->atanh_T[i] = __longlong_as_double((long long)NPYM_ATANH_T[i]);
        s->atanh_U[i] = __longlong_as_double((long long)NPYM_ATANH_U[i]);
    }
    for (int b = threadIdx.x; b < 64; b += blockDim.x) {
        uint32_t below = 0, inside = 0x10000u;                       // buckets are narrower than the gaps between thresholds
        for (int i = 0; i < 16; ++i) {
            const uint32_t th = NPYM_RCP14_R5_THR[i];
            if (th < ((uint32_t)b << 10)) ++below;
            else if (th < ((uint32_t)(b + 1) << 10)) inside = th;
        }
        s->rcp_bucket[b] = below | (inside << 8);
    }
}

__device__ __forceinline__ double npym_tanh(double x, const NpymTables &s)
{
    const int hi = __double2hiint(x);
    const int nd = hi & 0x7ff80000;                                  // exponent and leading mantissa bit
    const int idx = min(max(nd - 0x3fc00000, 0), 0x780000) >> 19;
    const double *c = s.tanh_lut + idx;
    const double y = __dsub_rn(fabs(x), c[0]);
    double r = c[17 * 16];
#pragma unroll
    for (int k = 16; k >= 1; --k) r = __fma_rn(r, y, c[k * 16]);
    r = (nd > 0x7fe00000) ? 1.0 : r;                                 // beyond the last interval (and infinities)
    r = __hiloint2double(__double2hiint(r) | (hi & 0x80000000), __double2loint(r));
    return (x != x) ? __longlong_as_double(0x7ff8000000000000ll) : r;
}

// VRCP14PD + the routine's rounding to 1+4 significant bits: bits of the rounded reciprocal of a positive normal operand
__device__ __forceinline__ int npym_rcp14_r5_hi(double x, const NpymTables &s)
{
    const int hi = __double2hiint(x);
    const uint32_t m16 = ((uint32_t)hi >> 4) & 0xffffu;              // 16 leading mantissa bits
    const uint32_t bk = s.rcp_bucket[m16 >> 10];
    const int k = (int)(bk & 0xffu) + (m16 >= (bk >> 8) ? 1 : 0);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    return 0x3ff00000 - (k << 16) - (e << 20);                       // high word; the low word is zero
}

__device__ __forceinline__ double npym_arctanh(double x, const NpymTables &s)
{
    const double ax = fabs(x);
    if (!(ax < 1.0)) return atanh(x);                                // SVML's scalar fall-back: +-inf at +-1, NaN beyond
    const double P = __dadd_rn(ax, 1.0), Q = __dsub_rn(1.0, ax);
    const double Pl = __dsub_rn(ax, __dsub_rn(P, 1.0)), Ql = __dadd_rn(ax, __dsub_rn(Q, 1.0));
    const int rph = npym_rcp14_r5_hi(P, s), rqh = npym_rcp14_r5_hi(Q, s);
    const double Rp = __hiloint2double(rph, 0), Rq = __hiloint2double(rqh, 0);
    double Ep = __fma_rn(Rp, P, -1.0); Ep = __fma_rn(Pl, Rp, Ep);
    double Eq = __fma_rn(Q, Rq, -1.0); Eq = __fma_rn(-Ql, Rq, Eq);
    const double de = (double)(((rqh >> 20) & 0x7ff) - ((rph >> 20) & 0x7ff));
    const int ip = (rph >> 16) & 15, iq = (rqh >> 16) & 15;
    const double dU = __dsub_rn(s.atanh_U[iq], s.atanh_U[ip]);
    const double dT = __dsub_rn(s.atanh_T[iq], s.atanh_T[ip]);
#define NPYM_D(c) __longlong_as_double((long long)(c))
    double pp = __fma_rn(NPYM_D(NPYM_ATANH_C8), Ep, NPYM_D(NPYM_ATANH_C7)), pq = __fma_rn(NPYM_D(NPYM_ATANH_C8), Eq, NPYM_D(NPYM_ATANH_C7));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C6)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C6));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C5)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C5));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C4)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C4));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C3)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C3));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C2)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C2));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C1)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C1));
    pp = __fma_rn(pp, Ep, NPYM_D(NPYM_ATANH_C0)); pq = __fma_rn(pq, Eq, NPYM_D(NPYM_ATANH_C0));
    const double H = __fma_rn(NPYM_D(NPYM_ATANH_L2H), de, dT), Lo = __fma_rn(NPYM_D(NPYM_ATANH_L2L), de, dU);
#undef NPYM_D
    const double A = __dadd_rn(Ep, H), B = __dsub_rn(A, Eq);
    const double errA = __dadd_rn(Ep, __dsub_rn(H, A)), errB = __dadd_rn(Eq, __dsub_rn(B, A));
    pp = __fma_rn(__dmul_rn(Ep, Ep), pp, Lo);
    pq = __fma_rn(-__dmul_rn(Eq, Eq), pq, errA);
    const double r = __dadd_rn(B, __dsub_rn(__dadd_rn(pp, pq), errB));
    return __dmul_rn(r, __hiloint2double((__double2hiint(x) & 0x80000000) | 0x3fe00000, 0));
}

}  // namespace qldpc
