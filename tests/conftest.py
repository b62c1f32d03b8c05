import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    return np.load(GOLD / name, allow_pickle=False)


def seeded_vqgan(cfg_name, seed):
    """(cfg, torch state_dict, numpy state_dict) of the seeded synthetic weights."""
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg[cfg_name]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=seed)
    return cfg, sd, {k: v.numpy() for k, v in sd.items()}


def check_weight_checksums(gold, sd):
    """The fixtures were produced from seeded weights; make sure this machine regenerates the same ones."""
    for k, s in zip(gold["weight_keys"].tolist(), gold["weight_sums"].tolist()):
        got = float(sd[k].double().abs().sum())
        assert abs(got - s) <= 1e-9 * max(1.0, abs(s)), f"seeded weights differ from the fixture's ({k}): {got} vs {s}"


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
