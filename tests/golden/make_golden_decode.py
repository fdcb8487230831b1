#!/usr/bin/env python
"""tests/golden/decode.npz from the REAL reference's recover_from_ric (build container only)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
np.float, np.int = float, int  # noqa: NPY001
sys.path.insert(0, os.environ.get("MST_REFERENCE_ROOT", "/root/reference"))

from data_loaders.humanml.scripts.motion_process import recover_from_ric  # noqa: E402
from oracle import decode as OD  # noqa: E402


def inputs(F, T, B=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn(B, F, 1, T, generator=g)
    mean = torch.randn(F, generator=g) * 0.3
    std = torch.rand(F, generator=g) * 0.5 + 0.05
    return sample, mean, std


def main():
    arrays = {}
    for name, F, T, J in [("stylexia", 181, 76, 20), ("humanml", 263, 196, 22), ("bandai", 190, 60, 21)]:
        sample, mean, std = inputs(F, T)
        ref = recover_from_ric((sample.permute(0, 2, 3, 1) * std + mean).float(), J)
        got = OD.decode_motion(sample, mean, std, J)
        err = float((got - ref).abs().max() / ref.abs().max())
        print(f"[{name}] oracle vs reference recover_from_ric: {err:.2e}")
        assert err < 1e-6 and tuple(ref.shape) == (3, 1, T, J, 3)
        arrays[f"{name}/joints"] = ref.numpy()
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "decode.npz"), **arrays)
    print("wrote tests/golden/decode.npz")


if __name__ == "__main__":
    main()
