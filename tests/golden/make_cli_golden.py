#!/usr/bin/env python
"""Golden result TABLES of the unmodified reference's simulate() (simulator.py:319-347): the text a user of the reference sees.

    python tests/golden/make_cli_golden.py        (needs /root/reference; ~1 min)

simulate() is run verbatim behind the inert stim of oracle/ref_loader.py whose sampler hands it, for every p of the sweep, the
record of the deterministic sampler (seed 1234) -- the same records qldpcsim_b200.simulator.simulate draws
itself when called with rngSeed=1234.  Stored: the arguments and the printed table (everything from the 'SIMULATION RESULTS'
banner on; the reference also prints two progress lines per shot, simulator.py:245,:304, which are not part of the contract).
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from qldpcsim_b200 import pcmlibrary, sampler  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: Steane, MS flooding, 50 iterations, p sweep, 1000 shots
    "steane_MS_F_sweep": dict(code="steane", p=[0.01, 0.02, 0.05, 0.1], shots=1000, decType="MS", decIterations=50, decSchedule="F", OSDorder=-1),
    "LP04_0_MS_L": dict(code="LP04_0", p=[0.02, 0.05], shots=200, decType="MS", decIterations=50, decSchedule="L", OSDorder=-1),
    "steane_NG": dict(code="steane", p=[0.05], shots=500, decType="NG", decIterations=99, decSchedule="F", OSDorder=-1),
    "LP04_0_BF": dict(code="LP04_0", p=[0.02], shots=100, decType="BF", decIterations=99, decSchedule="F", OSDorder=-1),
    "steane_BP_S": dict(code="steane", p=[0.03, 0.1], shots=300, decType="BP", decIterations=20, decSchedule="S", OSDorder=-1),
}
SEED = 1234


def main():
    assert ref_loader.available()
    out = {}
    for name, c in CASES.items():
        Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(c["code"])]
        recs = [sampler.sample_record(Hx, Hz, p, c["shots"], seed=SEED) for p in c["p"]]
        sim = ref_loader.load_simulator(recs)
        with tempfile.TemporaryDirectory() as td:
            fx, fz = os.path.join(td, "Hx.npy"), os.path.join(td, "Hz.npy")
            np.save(fx, Hx.astype(np.int64))
            np.save(fz, Hz.astype(np.int64))
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sim.simulate(fx, fz, p=c["p"], shots=c["shots"], decType=c["decType"], decIterations=c["decIterations"],
                             decSchedule=c["decSchedule"], OSDorder=c["OSDorder"], rngSeed=SEED)
        text = buf.getvalue()
        table = text[text.index("\n                             ===          SIMULATION RESULTS"):]
        out[name] = dict(c, seed=SEED, table=table)
        print(name)
        print(table)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_tables.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
