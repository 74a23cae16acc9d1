import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(g) if f.endswith(".npz") and not f.startswith(("pool_", "plc_")))


def plc_golden_names():
    """Base names of the packet-loss fixtures (tests/golden/plc_<base>.npz, made by make_golden_plc.py)."""
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[4:-4] for f in os.listdir(g) if f.endswith(".npz") and f.startswith("plc_"))


def load_plc_golden(base):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "plc_" + base + ".npz"))
    return {k: z[k] for k in z.files}


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    ch, fs, br, vbr, dec_ch = [int(v) for v in z["meta"]]
    return dict(packets=z["packets"], lens=z["lens"], enc_rng=z["enc_rng"], dec_rng=z["dec_rng"], pcm=z["pcm"],
                channels=ch, frame_size=fs, bitrate=br, vbr=vbr, dec_channels=dec_ch, name=name)


@pytest.fixture(scope="session")
def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libopus_ref.so"))
