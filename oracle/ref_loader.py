"""
Loads the UNMODIFIED reference (albertogp71/qLDPCsim) when its tree is reachable -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python and lives read-only under /root/reference in the build container; it does
not exist on the GPU box.  Anything that needs it (tests/golden/make_golden.py and the live differential
tests) goes through `available()` first and skips otherwise.  No reference source is copied into this
repository.

 * `import qLDPCsim` itself fails (qLDPCsim/__init__.py -> version.py imports tomlkit, not installed), so a
   stub package object with the right __path__ is registered and the sub-modules are imported through it.
 * simulator.py imports stim (not installed): `load_simulator(record)` installs an inert stand-in whose
   sampler returns a caller-supplied measurement record, so simulate_p runs verbatim (layerize, the X/Z
   wiring, dispatch, classification).
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

import numpy as np

def _find_ref_dir() -> str:
    """QLDPC_REF_DIR, else the read-only tree of the build container, else the pip-installed copy that travels to the GPU box
    (baseline/_ref, produced by baseline/install_reference.sh; the package only -- no data/ directory)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in (os.environ.get("QLDPC_REF_DIR"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "qLDPCsim", "decoders.py")):
            return d
    return os.environ.get("QLDPC_REF_DIR", "/root/reference")


REF_DIR = _find_ref_dir()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "qLDPCsim", "decoders.py"))


def _stub_package():
    if "qLDPCsim" not in sys.modules or not hasattr(sys.modules["qLDPCsim"], "__path__"):
        pkg = types.ModuleType("qLDPCsim")
        pkg.__path__ = [os.path.join(REF_DIR, "qLDPCsim")]
        sys.modules["qLDPCsim"] = pkg
    return sys.modules["qLDPCsim"]


def load(name: str):
    """Import qLDPCsim.<name> (decoders, gf2math, PCMlibrary) from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_DIR}")
    _stub_package()
    return importlib.import_module(f"qLDPCsim.{name}")


def load_layerize():
    """The nested function simulator.py:212-224, extracted with ast (simulator.py itself needs stim)."""
    src = open(os.path.join(REF_DIR, "qLDPCsim", "simulator.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "layerize")
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), "layerize", "exec"), ns)
    return ns["layerize"]


class _Inert:
    """Stands in for stim.Circuit / Tableau / sampler."""
    record = None

    def __init__(self, *a, **k):
        pass

    def __add__(self, o):
        return self

    def __iadd__(self, o):
        return self

    def to_circuit(self, **k):
        return _Inert()

    def compile_sampler(self, **k):
        return self

    def sample(self, shots):
        rec = _Inert.record
        if isinstance(rec, list):            # a sweep (simulate() calls simulate_p once per p): one record per call, in order
            rec = rec.pop(0)
        assert rec is not None and rec.shape[0] == shots
        return rec


def load_simulator(record: np.ndarray):
    """Import the reference simulator behind an inert stim whose sampler returns `record`."""
    os.environ["PYTHONBREAKPOINT"] = "0"      # simulator.py:297 has a stray breakpoint()
    stim = types.ModuleType("stim")
    stim.Circuit = _Inert
    stim.PauliString = types.SimpleNamespace(from_numpy=lambda **k: None)
    stim.Tableau = types.SimpleNamespace(from_stabilizers=lambda *a, **k: _Inert())
    sys.modules["stim"] = stim
    _Inert.record = [np.asarray(r, dtype=bool) for r in record] if isinstance(record, list) else np.asarray(record, dtype=bool)
    return load("simulator")
