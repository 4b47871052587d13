#!/usr/bin/env python
"""
Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(albertogp71/qLDPCsim, /root/reference) in a container where it is reachable.

    python tests/golden/make_golden.py [--only NAME_SUBSTR] [--jobs 8]

For every case the per-shot decoder calls are made exactly as simulator.py:270-282 makes them (prior p/3,
layersX -- built from Hx -- used with Hz and vice versa, BF/NG without iteration arguments) on the record of
the deterministic sampler of SURVEY.md section 8d (seed 1234).  Stored per case (bit-packed, see
qldpcsim_b200/bitpack.py): the input record, the reference's eX/eZ estimates, iteration counts and -- for
`counters=True` cases -- the dict returned by the reference's own simulate_p run behind an inert stim.
OSD cases also store the float64 posterior handed to OSDdec and the permutation np.argsort produced for it
(decoders.py:325), because that argsort is unstable and ties are frequent.

The reference cannot travel to the GPU box; these fixtures can.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from qldpcsim_b200 import bitpack, sampler  # noqa: E402

SEED = 1234
OUT = os.path.dirname(os.path.abspath(__file__))

# name, code, decType, schedule, p, shots, decIterations, OSDorder, counters
CASES = [
    # BASELINE.json configs[0]: Steane, MS flooding, 50 it, four p, 1000 shots
    ("steane_MS_F_p01", "steane", "MS", "F", 0.01, 1000, 50, -1, True),
    ("steane_MS_F_p02", "steane", "MS", "F", 0.02, 1000, 50, -1, True),
    ("steane_MS_F_p05", "steane", "MS", "F", 0.05, 1000, 50, -1, True),
    ("steane_MS_F_p10", "steane", "MS", "F", 0.10, 1000, 50, -1, True),
    ("steane_MS_L_p10", "steane", "MS", "L", 0.10, 500, 50, -1, False),
    ("steane_BP_F_p05", "steane", "BP", "F", 0.05, 1000, 50, -1, True),
    ("steane_BP_S_p10", "steane", "BP", "S", 0.10, 300, 50, -1, False),
    ("steane_BF_p05", "steane", "BF", "F", 0.05, 1000, 50, -1, True),
    ("steane_NG_p05", "steane", "NG", "F", 0.05, 1000, 50, -1, True),
    ("shor_NG_p05", "shor", "NG", "F", 0.05, 200, 50, -1, True),
    ("shor_BF_p05", "shor", "BF", "F", 0.05, 200, 50, -1, True),
    ("LP04_0_MS_L_p02", "LP04_0", "MS", "L", 0.02, 200, 50, -1, True),
    ("LP04_0_MS_L_p05", "LP04_0", "MS", "L", 0.05, 200, 50, -1, True),
    ("LP04_0_MS_L_p10", "LP04_0", "MS", "L", 0.10, 96, 50, -1, False),
    ("LP04_0_MS_F_p05", "LP04_0", "MS", "F", 0.05, 96, 50, -1, False),
    ("LP04_0_MS_S_p05", "LP04_0", "MS", "S", 0.05, 48, 50, -1, False),
    ("LP04_0_NG_p02", "LP04_0", "NG", "F", 0.02, 100, 50, -1, True),
    ("LP04_0_BF_p02", "LP04_0", "BF", "F", 0.02, 100, 50, -1, True),
    ("LP04_0_BP_F_p05", "LP04_0", "BP", "F", 0.05, 48, 100, -1, False),
    ("LP04_0_BP_L_p08", "LP04_0", "BP", "L", 0.08, 32, 30, -1, False),
    ("LP04_0_MS_L_OSD0_p10", "LP04_0", "MS", "L", 0.10, 48, 8, 0, False),
    ("LP04_0_MS_L_OSD1_p10", "LP04_0", "MS", "L", 0.10, 24, 8, 1, False),
    ("LP04_0_MS_F_OSD2_p10", "LP04_0", "MS", "F", 0.10, 24, 8, 2, False),
    ("LP04_2_MS_L_p05", "LP04_2", "MS", "L", 0.05, 32, 50, -1, False),
    ("LP118_0_MS_L_p02", "LP118_0", "MS", "L", 0.02, 96, 50, -1, False),
    ("LP118_0_MS_L_p05", "LP118_0", "MS", "L", 0.05, 96, 50, -1, True),
    ("LP118_0_MS_L_p10", "LP118_0", "MS", "L", 0.10, 40, 50, -1, False),
    ("LP118_0_MS_F_p05", "LP118_0", "MS", "F", 0.05, 48, 50, -1, False),
    ("LP118_0_BP_F_p05", "LP118_0", "BP", "F", 0.05, 16, 100, -1, False),
    ("LP118_0_MS_L_OSD0_p10", "LP118_0", "MS", "L", 0.10, 8, 5, 0, False),
    # OSD at BASELINE config-3 size (n > 992: the two-words-per-lane OSD kernel); ~25 s per unconverged decode in the reference
    ("LP118_2_MS_S_OSD0_p05", "LP118_2", "MS", "S", 0.05, 6, 2, 0, False),
    ("T_MS_L_OSD0_p05", "T", "MS", "L", 0.05, 6, 2, 0, False),
    # order 10 (BASELINE config 3's order) where the reference can afford its 1024 REF calls per decode
    ("LP04_0_MS_L_OSD10_p10", "LP04_0", "MS", "L", 0.10, 12, 8, 10, False),
    ("LP118_1_MS_L_p05", "LP118_1", "MS", "L", 0.05, 24, 50, -1, False),
    ("LP118_2_MS_L_p05", "LP118_2", "MS", "L", 0.05, 16, 50, -1, False),
    ("LP118_2_MS_S_p05", "LP118_2", "MS", "S", 0.05, 4, 50, -1, False),
    ("T_MS_L_p03", "T", "MS", "L", 0.03, 16, 50, -1, False),
    ("bicycle_MS_L_p03", "bicycle", "MS", "L", 0.03, 48, 50, -1, False),
    ("bicycle_BP_F_p03", "bicycle", "BP", "F", 0.03, 16, 50, -1, False),
]


def ref_code(code):
    """(Hx, Hz) int8 through the reference's own load path (simulator.py:20-35 semantics)."""
    pcm = ref_loader.load("PCMlibrary")
    if code == "bicycle":
        Hx, Hz = pcm.bicycle_code()
    elif os.path.exists(f"{ref_loader.REF_DIR}/data/Hx_{code}.npy"):
        Hx = np.load(f"{ref_loader.REF_DIR}/data/Hx_{code}.npy")
        Hz = np.load(f"{ref_loader.REF_DIR}/data/Hz_{code}.npy")
    else:
        raise KeyError(code)
    return (Hx % 2).astype(np.int8), (Hz % 2).astype(np.int8)


def _decode_chunk(args):
    code, decType, sched, p, it, osd, rec, lo, hi = args
    warnings.filterwarnings("ignore")
    dec = ref_loader.load("decoders")
    layerize = ref_loader.load_layerize()
    Hx, Hz = ref_code(code)
    m_x, n = Hx.shape
    m_z = Hz.shape[0]
    if sched == "F":
        lX, lZ = [np.arange(m_x)], [np.arange(m_z)]
    else:
        lX, lZ = layerize(Hx, serial=sched == "S"), layerize(Hz, serial=sched == "S")
    captured = {}
    real_osd = dec.OSDdec

    def recording_osd(H, e_hat, syndrome, llr, order=0):
        captured["llr"] = np.array(llr, dtype=np.float64)
        captured["e_in"] = np.array(e_hat).astype(np.uint8)
        sat = np.where(np.abs(llr) < 100.0, llr, 100.0 * np.sign(llr))     # decoders.py:320-325
        prob = 1. / (1. + np.exp(sat))
        rel = np.where(prob > 0.5, prob, 1 - prob)
        captured["perm"] = np.argsort(rel).astype(np.int32)
        return real_osd(H, e_hat, syndrome, llr, order)

    dec.OSDdec = recording_osd
    out = []
    for s in range(lo, hi):
        row = rec[s]
        sy_z = row[:m_z].astype(int)
        sy_x = row[m_z:m_z + m_x].astype(int)
        res = []
        for H, sy, lay in ((Hz, sy_z, lX), (Hx, sy_x, lZ)):
            captured.clear()
            if decType == "NG":
                e, i = dec.NG_decoder(H, sy)
            elif decType == "BF":
                e, i = dec.BF_decoder(H, sy)
            elif decType == "MS":
                e, i = dec.MS_decoder(H, sy, p=p / 3, max_iter=it, layers=lay, OSDorder=osd)
            else:
                e, i = dec.BP_decoder(H, sy, p=p / 3, max_iter=it, layers=lay)
            res.append((np.asarray(e).astype(np.uint8), int(i), dict(captured)))
        out.append(res)
    return lo, out


def run_case(case, pool, jobs):
    name, code, decType, sched, p, shots, it, osd, counters = case
    Hx, Hz = ref_code(code)
    m_x, n = Hx.shape
    m_z = Hz.shape[0]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=SEED)
    step = max(1, -(-shots // (jobs * 4)))
    chunks = [(code, decType, sched, p, it, osd, rec, lo, min(shots, lo + step)) for lo in range(0, shots, step)]
    eX = np.zeros((shots, n), np.uint8)
    eZ = np.zeros((shots, n), np.uint8)
    itX = np.zeros(shots, np.int32)
    itZ = np.zeros(shots, np.int32)
    osd_rows = []
    for lo, out in pool.imap_unordered(_decode_chunk, chunks):
        for k, res in enumerate(out):
            s = lo + k
            eX[s], itX[s], cx = res[0]
            eZ[s], itZ[s], cz = res[1]
            for which, c in ((0, cx), (1, cz)):
                if c:
                    osd_rows.append((s, which, c["llr"], c["perm"], c["e_in"]))
    data = dict(code=code, decType=decType, sched=sched, p=p, shots=shots, decIterations=it, OSDorder=osd,
                seed=SEED, m_x=m_x, m_z=m_z, n=n,
                record=bitpack.pack_rows(rec), eX=bitpack.pack_rows(eX), eZ=bitpack.pack_rows(eZ), itX=itX, itZ=itZ)
    if osd_rows:
        osd_rows.sort(key=lambda r: (r[0], r[1]))
        data["osd_shot"] = np.array([r[0] for r in osd_rows], np.int32)
        data["osd_which"] = np.array([r[1] for r in osd_rows], np.int8)      # 0: X decode (Hz), 1: Z decode (Hx)
        data["osd_llr"] = np.stack([r[2] for r in osd_rows])
        data["osd_perm"] = np.stack([r[3] for r in osd_rows])
        data["osd_e_in"] = bitpack.pack_rows(np.stack([r[4] for r in osd_rows]))
    if counters:
        sim = ref_loader.load_simulator(rec)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = sim.simulate_p(Hx, Hz, p=p, shots=shots, decType=decType, decIterations=it, decSchedule=sched,
                               OSDorder=osd, rngSeed=SEED)
        data["counters"] = np.array([r["DecFailures_X"], r["DecFailures_Z"], r["decSuccessExact"], r["decSuccessDegen"],
                                     round(r["Avg_number_of_iterations_X"] * shots),
                                     round(r["Avg_number_of_iterations_Z"] * shots)], np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    return data


def write_codes():
    """Non-zero coordinates of every library matrix, from the reference's data files / generators."""
    out = {}
    for code in ["steane", "shor", "LP04_0", "LP04_1", "LP04_2", "LP04_3", "LP118_0", "LP118_1", "LP118_2", "T", "bicycle"]:
        Hx, Hz = ref_code(code)
        for tag, H in (("x", Hx), ("z", Hz)):
            r, c = np.nonzero(H)
            out[f"{code}_H{tag}_shape"] = np.array(H.shape, np.int32)
            out[f"{code}_H{tag}_rc"] = np.stack([r, c]).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "codes.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    assert ref_loader.available(), "reference tree not reachable"
    write_codes()
    with mp.Pool(a.jobs) as pool:
        for case in CASES:
            if a.only and a.only not in case[0]:
                continue
            t0 = time.time()
            d = run_case(case, pool, a.jobs)
            print(f"{case[0]:28s} shots={case[5]:5d} itX={int(d['itX'].sum()):6d} itZ={int(d['itZ'].sum()):6d} "
                  f"counters={d.get('counters')}  {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
