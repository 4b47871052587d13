import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not f.endswith("codes.npz") and not os.path.basename(f).startswith(("big_", "numpy_", "cli_")))


def load_golden(name):
    from qldpcsim_b200 import bitpack, pcmlibrary
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g = {k: (v.item() if v.ndim == 0 else v) for k, v in d.items()}
    Hx, Hz = pcmlibrary.by_name(g["code"])
    g["Hx"], g["Hz"] = (Hx % 2).astype(np.int8), (Hz % 2).astype(np.int8)
    mz, mx, n = g["m_z"], g["m_x"], g["n"]
    rec = bitpack.unpack_rows(g["record"], mz + mx + 2 * n)
    g["rec"] = rec.astype(bool)
    g["sy_z"], g["sy_x"] = rec[:, :mz], rec[:, mz:mz + mx]
    g["errX"], g["errZ"] = rec[:, mz + mx:mz + mx + n], rec[:, mz + mx + n:]
    g["eX_ref"] = bitpack.unpack_rows(g["eX"], n)
    g["eZ_ref"] = bitpack.unpack_rows(g["eZ"], n)
    return g


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
