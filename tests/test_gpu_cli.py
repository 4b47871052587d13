"""End-to-end front end: qldpcsim_b200.simulator.main / simulate with the reference's flags on .npy
and whitespace-text matrices print the same result table, line for line, as the unmodified reference's simulate()
(simulator.py:319-347; tables captured by tests/golden/make_cli_golden.py on the records of the deterministic sampler)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(GOLDEN_DIR, "cli_tables.json")))


def _table(text):
    t = text[text.index("\n                             ===          SIMULATION RESULTS"):]
    return [ln.rstrip() for ln in t.strip("\n").splitlines()]


def _write(tmp_path, code, as_text):
    from qldpcsim_b200 import pcmlibrary
    Hx, Hz = pcmlibrary.by_name(code)
    if as_text:
        fx, fz = str(tmp_path / "Hx.txt"), str(tmp_path / "Hz.txt")
        np.savetxt(fx, np.asarray(Hx) % 2, fmt="%d")
        np.savetxt(fz, np.asarray(Hz) % 2, fmt="%d")
    else:
        fx, fz = str(tmp_path / "Hx.npy"), str(tmp_path / "Hz.npy")
        np.save(fx, np.asarray(Hx).astype(np.int64))          # the reference's data files are dense int64
        np.save(fz, np.asarray(Hz).astype(np.int64))
    return fx, fz


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("as_text", [False, True])
def test_main_prints_the_reference_table(name, as_text, tmp_path, capsys, cuda_device):
    from qldpcsim_b200 import simulator
    c = CASES[name]
    fx, fz = _write(tmp_path, c["code"], as_text)
    argv = ["--Hx", fx, "--Hz", fz, "--p"] + [repr(p) for p in c["p"]] + ["--shots", str(c["shots"]), "--rngSeed", str(c["seed"]),
            "--decType", c["decType"], "--decIterations", str(c["decIterations"]), "--decSchedule", c["decSchedule"],
            "--OSDorder", str(c["OSDorder"])]
    simulator.main(argv)
    out = capsys.readouterr().out
    assert "Command line arguments" in out                               # simulator.py:367-369
    assert _table(out) == _table(c["table"]), "\n" + out


def test_simulate_function_defaults_and_return(tmp_path, capsys, cuda_device):
    """simulate() keeps the reference's signature and defaults (shots=1000, decType='MS', decIterations=99, decSchedule='F',
    OSDorder=-1) and returns None (simulator.py:319-327)."""
    import inspect
    from qldpcsim_b200 import simulator
    sig = inspect.signature(simulator.simulate)
    want = {"shots": 1000, "decType": "MS", "decIterations": 99, "decSchedule": "F", "OSDorder": -1, "rngSeed": None}
    for k, v in want.items():
        assert sig.parameters[k].default == v
    assert list(sig.parameters)[:3] == ["HxFile", "HzFile", "p"]
    c = CASES["steane_MS_F_sweep"]
    fx, fz = _write(tmp_path, c["code"], False)
    r = simulator.simulate(fx, fz, p=c["p"], shots=c["shots"], decType=c["decType"], decIterations=c["decIterations"],
                           decSchedule=c["decSchedule"], OSDorder=c["OSDorder"], rngSeed=c["seed"])
    assert r is None
    assert _table(capsys.readouterr().out) == _table(c["table"])
    with pytest.raises(AssertionError):
        simulator.simulate(fx, fz, p=[1.5], shots=10)                      # simulator.py:332
    with pytest.raises(SystemExit):
        simulator.main(["--Hx", fx, "--Hz", fz, "--p", "0.1", "--decType", "XX"])
