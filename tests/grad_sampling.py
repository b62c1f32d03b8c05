"""Which gradient entries the backward-path fixtures keep (shared by the generator and the tests)."""
import zlib

import numpy as np

SMALL = 4096        # tensors up to this many elements are stored whole
SAMPLES = 512


def grad_sample_positions(name, numel):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    return np.sort(rng.choice(numel, size=SAMPLES, replace=False)).astype(np.int64)
