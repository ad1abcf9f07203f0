"""Load the UNMODIFIED reference file ``/root/reference/merv/util/nn_utils.py`` as a module.

TEST INFRASTRUCTURE ONLY, and usable only where ``/root/reference`` exists (the build container —
not the GPU box).  ``import merv`` fails here (draccus/timm/decord/... absent), but the hot-path
classes only need torch + einops; ``timm`` is imported at nn_utils.py:15-16 for out-of-scope classes,
so a stub package is placed in ``sys.modules`` first.  Nothing is copied: the file is exec'd in place.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MERV_REFERENCE_ROOT", "/root/reference")
_NN_UTILS = os.path.join(REFERENCE_ROOT, "merv", "util", "nn_utils.py")
_cached = None


def reference_available() -> bool:
    return os.path.isfile(_NN_UTILS)


def load_reference_nn_utils():
    """Returns the reference module (``ref.AveragePooling3DProjector``, ``ref.CrossAttentionAdapterLearnableQuery`` ...)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"{_NN_UTILS} not found: the reference only exists in the build container")
    import torch

    if "timm" not in sys.modules:
        timm, layers, models, regnet = (
            types.ModuleType(n) for n in ("timm", "timm.layers", "timm.models", "timm.models.regnet")
        )
        layers.LayerNorm2d, layers.trunc_normal_, regnet.RegStage = torch.nn.LayerNorm, torch.nn.init.trunc_normal_, object
        timm.layers, timm.models, models.regnet = layers, models, regnet
        sys.modules.update(
            {"timm": timm, "timm.layers": layers, "timm.models": models, "timm.models.regnet": regnet}
        )
    spec = importlib.util.spec_from_file_location("ref_nn_utils", _NN_UTILS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    _cached = ref
    return ref
