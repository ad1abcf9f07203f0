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


def load_variant(name):
    """(settings, inputs, params, golden arrays of one oracle/variants.py entry); asserts the recorded checksum."""
    from oracle import variants as V

    g = np.load(os.path.join(GOLDEN_DIR, "variants.npz"))
    inputs, params = V.make_variant(name)
    ck = sum(float(np.abs(a.astype(np.float64)).sum()) for a in inputs) + sum(float(np.abs(a.astype(np.float64)).sum()) for a in params.values())
    assert abs(ck - float(g[f"{name}.checksum"])) <= 1e-9 * abs(ck), "numpy generator drift"
    gold = {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith(name + ".")}
    return V.VARIANTS[name], inputs, params, gold
