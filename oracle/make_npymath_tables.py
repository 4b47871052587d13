#!/usr/bin/env python
"""Regenerates npymath_tables.inc (oracle/ and qldpcsim_b200/csrc/ hold identical copies) from the NumPy installed here.

    python oracle/make_npymath_tables.py [--check]

What it reads (data only; the evaluating code in npymath.h / npymath.cuh is written from the published algorithms):
  * np.tanh, float64: the 18 x 16 look-up table of simd_tanh_f64 (numpy/_core/src/umath/loops_hyperbolic.dispatch.c.src),
    found in .rodata of numpy/_core/_multiarray_umath*.so by its first row, the interval centres 0, 7/32, 5/16, 7/16, ...;
  * np.arctanh, float64: the data block of Intel SVML's __svml_atanh8_ha (symbol __svml_datanh_ha_data_internal_avx512);
  * VRCP14PD + the routine's rounding ((bits + 2^47) & ~(2^48 - 1)) as a step function of the operand's mantissa: the 16
    thresholds are found by bisection WITH THE INSTRUCTION ITSELF (a small C program, needs an AVX-512 host and gcc), then
    checked on 2.6e8 operands (random and around every threshold) and for independence of the exponent.
Afterwards the restated functions are compared bit for bit with np.tanh / np.arctanh on 2e6 arguments (needs NumPy's
AVX512_SKX dispatch, which is what produced the reference goldens).  NumPy 2.3.5 is pinned by the image; with another version
the script fails loudly rather than writing tables that do not reproduce it.
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = [os.path.join(HERE, "npymath_tables.inc"), os.path.join(ROOT, "qldpcsim_b200", "csrc", "npymath_tables.inc")]

RCP_C = r"""
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t R(uint64_t b){ double x; memcpy(&x,&b,8); __m128d v=_mm_set_sd(x); double y=_mm_cvtsd_f64(_mm_rcp14_sd(v,v)); uint64_t u; memcpy(&u,&y,8); return u; }
static inline uint64_t G(uint64_t M){ return (R(0x3ff0000000000000ull|M)+0x0000800000000000ull)&0xffff000000000000ull; }
int main(void){
    const uint64_t top=(1ull<<52)-1; uint64_t lo[128],hi[128],thr[64],val[64]; int sp=1,nt=0; lo[0]=0; hi[0]=top;
    while(sp){ uint64_t a=lo[--sp],b=hi[sp]; if(G(a)==G(b)) continue; if(b-a==1){thr[nt]=b;val[nt++]=G(b);continue;}
        uint64_t m=a+(b-a)/2; lo[sp]=a;hi[sp++]=m; lo[sp]=m;hi[sp++]=b; }
    for(int i=0;i<nt;i++)for(int j=i+1;j<nt;j++) if(thr[j]<thr[i]){uint64_t t=thr[i];thr[i]=thr[j];thr[j]=t;t=val[i];val[i]=val[j];val[j]=t;}
    uint64_t bad=0; srand48(7);
    for(long t=0;t<200000000;t++){ uint64_t M=((uint64_t)lrand48()<<31^(uint64_t)lrand48()<<10^(uint64_t)lrand48())&top,e=G(0); for(int i=0;i<nt;i++) if(M>=thr[i]) e=val[i]; bad+=G(M)!=e; }
    for(int i=0;i<nt;i++) for(long d=-2000000;d<=2000000;d++){ uint64_t M=thr[i]+d; if(M>top) continue; uint64_t e=G(0); for(int k=0;k<nt;k++) if(M>=thr[k]) e=val[k]; bad+=G(M)!=e; }
    for(long t=0;t<20000000;t++){ uint64_t M=((uint64_t)lrand48()<<31^(uint64_t)lrand48()<<10^(uint64_t)lrand48())&top; int e=(int)(lrand48()%60)-50; bad+=R(((uint64_t)(1023+e)<<52)|M)!=R(0x3ff0000000000000ull|M)-((uint64_t)e<<52); }
    printf("%d %lu %016lx\n",nt,bad,G(0));
    for(int i=0;i<nt;i++) printf("%013lx %016lx\n",thr[i],val[i]);
    return 0; }
"""


TANH_C = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
static double u2d(uint64_t u){double d;memcpy(&d,&u,8);return d;}
void eval(const uint64_t *lut, const double *x, double *y, long n){
    for(long i=0;i<n;i++){ uint64_t b; memcpy(&b,&x[i],8); int32_t hi=(int32_t)((b&0x7ff8000000000000ull)>>32)-0x3fc00000;
        if(hi<0)hi=0; if(hi>0x780000)hi=0x780000; int k=hi>>19; double v=fabs(x[i])-u2d(lut[k]),r=u2d(lut[17*16+k]);
        for(int c=16;c>=1;--c) r=fma(r,v,u2d(lut[c*16+k])); y[i]=copysign(r,x[i]); } }
"""


def tanh_table_reproduces_numpy(lut):
    """Evaluates a candidate [18][16] table with the Horner/FMA scheme and compares with np.tanh on |x| < 20."""
    import ctypes
    with tempfile.TemporaryDirectory() as td:
        src, lib = os.path.join(td, "t.c"), os.path.join(td, "t.so")
        open(src, "w").write(TANH_C)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", lib, src, "-lm"], check=True)
        L = ctypes.CDLL(lib)
        x = np.random.default_rng(3).uniform(-20, 20, 200000)
        y = np.empty_like(x)
        L.eval(lut.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    return bool((y.view(np.uint64) == np.tanh(x).view(np.uint64)).all())


def find_tables():
    so = np._core._multiarray_umath.__file__
    data = open(so, "rb").read()
    centres = np.array([0, 0.21875, 0.3125, 0.4375, 0.625, 0.875, 1.25, 1.75, 2.5, 3.5, 5, 7, 10, 14, 20, 0], np.float64).tobytes()
    hits, pos = [], data.find(centres)
    while pos >= 0:                                   # SVML's own tanh tables (unused by NumPy for float64) start with the same row
        hits.append(np.frombuffer(data[pos:pos + 18 * 16 * 8], dtype=np.uint64).copy())
        pos = data.find(centres, pos + 1)
    good = [h for h in hits if tanh_table_reproduces_numpy(h)]
    assert good and all((g == good[0]).all() for g in good), "tanh look-up table not found (or ambiguous) in " + so
    lut = good[0]
    nm = subprocess.run(["nm", so], capture_output=True, text=True, check=True).stdout
    addr = next(int(l.split()[0], 16) for l in nm.splitlines() if l.endswith(" __svml_datanh_ha_data_internal_avx512"))
    # .rodata is mapped at its file offset in this binary (readelf -S: Addr == Off); verified by the constants below
    d = np.frombuffer(data[addr:addr + 0x500], dtype=np.uint64).reshape(-1, 8)
    assert d[0x100 // 64][0] == 0x3ff0000000000000 and d[0x180 // 64][0] == 0x0000800000000000 and d[0x1c0 // 64][0] == 0xffff000000000000
    return lut, d


def rcp_thresholds():
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "rcp.c"), os.path.join(td, "rcp")
        open(src, "w").write(RCP_C)
        subprocess.run(["gcc", "-O2", "-mavx512f", "-mavx512vl", "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    nt, bad, g0 = int(out[0]), int(out[1]), int(out[2], 16)
    assert nt == 16 and bad == 0 and g0 == 0x3ff0000000000000, out[:3]
    thr = [int(x, 16) for x in out[3::2]]
    val = [int(x, 16) for x in out[4::2]]
    assert all(t & ((1 << 36) - 1) == 0 for t in thr), "thresholds depend on more than 16 mantissa bits"
    assert val == [0x3ff0000000000000 - ((k + 1) << 48) for k in range(16)]
    return [t >> 36 for t in thr]


def render(lut, d, thr):
    def rows(v, per=4):
        return "\n".join("    " + ", ".join("0x%016xull" % int(x) for x in v[i:i + per]) + "," for i in range(0, len(v), per))
    head = open(OUT[0]).read().split("NPYM_TABLE", 1)[0] if os.path.exists(OUT[0]) else "/* NumPy math tables */\n"
    s = head
    s += "NPYM_TABLE unsigned long long NPYM_TANH_LUT[18 * 16] = {   /* [b, c0 .. c16][interval] */\n" + rows(lut) + "\n};\n"
    s += "NPYM_TABLE unsigned long long NPYM_ATANH_T[16] = {\n" + rows(np.concatenate([d[0], d[1]])) + "\n};\n"
    s += "NPYM_TABLE unsigned long long NPYM_ATANH_U[16] = {\n" + rows(np.concatenate([d[2], d[3]])) + "\n};\n"
    s += "NPYM_TABLE unsigned short NPYM_RCP14_R5_THR[16] = {" + ", ".join("0x%04x" % x for x in thr) + "};\n"
    for off, nm in ((0x200, "C8"), (0x240, "C7"), (0x280, "C6"), (0x2c0, "C5"), (0x300, "C4"), (0x340, "C3"), (0x380, "C2"),
                    (0x3c0, "C1"), (0x400, "C0"), (0x440, "L2H"), (0x480, "L2L")):
        s += "#define NPYM_ATANH_%s 0x%016xull\n" % (nm, int(d[off // 64][0]))
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="do not write; fail if the committed tables differ")
    a = ap.parse_args()
    assert np._core._multiarray_umath.__cpu_features__.get("AVX512_SKX"), "needs a host on which NumPy dispatches to AVX512_SKX"
    lut, d = find_tables()
    text = render(lut, d, rcp_thresholds())
    if a.check:
        assert all(open(p).read() == text for p in OUT), "committed tables differ from the installed NumPy's"
    else:
        for p in OUT:
            open(p, "w").write(text)
    sys.path.insert(0, ROOT)
    from oracle import oracle
    oracle.build(force=not a.check)
    rng = np.random.default_rng(0)
    xt = np.concatenate([rng.uniform(-30, 30, 1000000), np.sign(rng.uniform(-1, 1, 1000000)) * 10.0 ** rng.uniform(-300, 2, 1000000)])
    xa = np.concatenate([rng.uniform(-1, 1, 1000000), np.sign(rng.uniform(-1, 1, 1000000)) * (1 - 10.0 ** rng.uniform(-16, 0, 1000000))])
    assert (oracle.npy_tanh(xt).view(np.uint64) == np.tanh(xt).view(np.uint64)).all()
    assert (oracle.npy_arctanh(xa).view(np.uint64) == np.arctanh(xa).view(np.uint64)).all()
    print("tables", "match" if a.check else "written", "and reproduce NumPy", np.__version__, "bit for bit on 4e6 arguments")


if __name__ == "__main__":
    main()
