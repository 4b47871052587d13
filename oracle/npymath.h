/* npymath.h -- CPU restatement of the two float64 routines NumPy 2.3.5 runs for np.tanh and np.arctanh on an AVX-512 x86-64
 * host (TEST INFRASTRUCTURE: part of the oracle).  The reference's sum-product decoder calls them at decoders.py:254, :256
 * and :259; its results depend on their last bit, and they are neither libm's nor correctly rounded (np.tanh differs from
 * glibc's on 26 % of arguments, np.arctanh on 4 %), so a bit-exact oracle has to evaluate the same algorithms:
 *   npym_tanh    -- NumPy's own SIMD kernel simd_tanh_f64 (numpy/_core/src/umath/loops_hyperbolic.dispatch.c.src),
 *   npym_arctanh -- Intel SVML __svml_atanh8_ha, vendored by NumPy (umath/svml/linux/avx512/svml_z0_atanh_d_ha.s).
 * NumPy is a third-party dependency of the reference (pyproject.toml:20, numpy>=2.3.5; 2.3.5 is what the build container and the
 * GPU box hold), absent from /root/reference; the algorithms are restated from the published sources, the tables are data
 * (npymath_tables.inc, regenerated and re-verified by oracle/make_npymath_tables.py).
 * Pinned: tests/test_npymath.py checks both functions bit for bit against tests/golden/numpy_tanh_arctanh.npz (values produced
 * by NumPy itself in the build container) and, where the running NumPy dispatches to AVX512_SKX, against NumPy live. */
#ifndef ORC_NPYMATH_H
#define ORC_NPYMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NPYM_TABLE static const
#include "npymath_tables.inc"

static inline double npym_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t npym_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* simd_tanh_f64: interval index from (exponent, leading mantissa bit) of |x| relative to 2^-3, clamped to 0..15; Horner with
 * FMAs in y = |x| - b; 1.0 beyond the last interval's range; sign restored; NaN -> quiet NaN. */
static inline double npym_tanh(double x)
{
    const uint64_t bits = npym_d2u(x);
    if (x != x) return npym_u2d(0x7ff8000000000000ull);
    const uint64_t nd = bits & 0x7ff8000000000000ull;
    int32_t hi = (int32_t)(nd >> 32) - (int32_t)0x3fc00000;
    if (hi < 0) hi = 0;
    if (hi > 0x780000) hi = 0x780000;
    const int idx = hi >> 19;
    const double y = npym_u2d(bits & 0x7fffffffffffffffull) - npym_u2d(NPYM_TANH_LUT[idx]);
    double r = npym_u2d(NPYM_TANH_LUT[17 * 16 + idx]);
    for (int k = 16; k >= 1; --k) r = fma(r, y, npym_u2d(NPYM_TANH_LUT[k * 16 + idx]));
    if (nd > 0x7fe0000000000000ull) r = 1.0;
    return npym_u2d(npym_d2u(r) | (bits & 0x8000000000000000ull));
}

/* VRCP14PD followed by the routine's rounding to 1+4 significant bits, for a positive normal operand */
static inline uint64_t npym_rcp14_r5(double x)
{
    const uint64_t b = npym_d2u(x);
    const uint32_t m16 = (uint32_t)((b >> 36) & 0xffff);
    int k = 0;
    for (int i = 0; i < 16; i++) k += (m16 >= NPYM_RCP14_R5_THR[i]);
    const int64_t e = (int64_t)((b >> 52) & 0x7ff) - 1023;
    return 0x3ff0000000000000ull - ((uint64_t)k << 48) - ((uint64_t)e << 52);
}

/* __svml_atanh8_ha, main path (|x| < 1); other arguments take SVML's scalar fall-back, for which libm's values are used
 * (atanh(+-1) = +-inf, NaN beyond) */
static inline double npym_arctanh(double x)
{
    const double ax = fabs(x);
    if (!(ax < 1.0)) return atanh(x);
    const double P = ax + 1.0, Q = 1.0 - ax;                 /* 1 + |x|, 1 - |x| and the parts lost to rounding */
    const double Pl = ax - (P - 1.0), Ql = ax + (Q - 1.0);
    const uint64_t rp = npym_rcp14_r5(P), rq = npym_rcp14_r5(Q);
    const double Rp = npym_u2d(rp), Rq = npym_u2d(rq);
    double Ep = fma(Rp, P, -1.0); Ep = fma(Pl, Rp, Ep);      /* reduced arguments: Rp (1+|x|) - 1, Rq (1-|x|) - 1 */
    double Eq = fma(Q, Rq, -1.0); Eq = fma(-Ql, Rq, Eq);
    const double de = (double)((int)((rq >> 52) & 0x7ff) - 1023) - (double)((int)((rp >> 52) & 0x7ff) - 1023);
    const int ip = (int)(rp >> 48) & 15, iq = (int)(rq >> 48) & 15;
    const double dU = npym_u2d(NPYM_ATANH_U[iq]) - npym_u2d(NPYM_ATANH_U[ip]);
    const double dT = npym_u2d(NPYM_ATANH_T[iq]) - npym_u2d(NPYM_ATANH_T[ip]);
    double pp = fma(npym_u2d(NPYM_ATANH_C8), Ep, npym_u2d(NPYM_ATANH_C7)), pq = fma(npym_u2d(NPYM_ATANH_C8), Eq, npym_u2d(NPYM_ATANH_C7));
    static const uint64_t c[7] = {NPYM_ATANH_C6, NPYM_ATANH_C5, NPYM_ATANH_C4, NPYM_ATANH_C3, NPYM_ATANH_C2, NPYM_ATANH_C1, NPYM_ATANH_C0};
    for (int k = 0; k < 7; k++) { pp = fma(pp, Ep, npym_u2d(c[k])); pq = fma(pq, Eq, npym_u2d(c[k])); }
    const double H = fma(npym_u2d(NPYM_ATANH_L2H), de, dT), Lo = fma(npym_u2d(NPYM_ATANH_L2L), de, dU);
    const double A = Ep + H, B = A - Eq;
    const double errA = Ep + (H - A), errB = Eq + (B - A);
    pp = fma(Ep * Ep, pp, Lo);
    pq = fma(-(Eq * Eq), pq, errA);
    const double r = B + ((pp + pq) - errB);
    return r * npym_u2d((npym_d2u(x) & 0x8000000000000000ull) | 0x3fe0000000000000ull);
}
#endif
