"""Helpers shared by the CPU (oracle) and GPU (parity) tests."""
import os

import numpy as np

from oracle import cases as C

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def regenerate(case):
    """Seeded inputs/weights of a case; asserts the checksums recorded when the golden was made."""
    g = load_golden(case.name)
    feats, pp, fp = C.make_features(case), C.make_projector_params(case), C.make_fusion_params(case)
    in_ck = sum(float(np.abs(f.astype(np.float64)).sum()) for f in feats)
    p_ck = sum(float(np.abs(v.astype(np.float64)).sum()) for p in pp for v in p.values()) + sum(
        float(np.abs(v.astype(np.float64)).sum()) for v in fp.values()
    )
    assert abs(in_ck - float(g["input_checksum"])) <= 1e-9 * abs(in_ck), "numpy generator drift: inputs"
    assert abs(p_ck - float(g["param_checksum"])) <= 1e-9 * abs(p_ck), "numpy generator drift: params"
    return g, feats, pp, fp
