"""Load the UNMODIFIED reference file ``merv/util/nn_utils.py`` as a module.

TEST INFRASTRUCTURE ONLY.  ``import merv`` fails here (draccus/timm/decord/... absent), but the hot-path classes only
need torch + einops; ``timm`` is imported at nn_utils.py:15-16 for out-of-scope classes, so a stub package is placed in
``sys.modules`` first.  The file is exec'd from where it lies:

* ``/root/reference/merv/util/nn_utils.py`` in the build container, or
* ``oracle/_ref/nn_utils.py`` — a byte-for-byte copy that ``__graft_entry__.build()`` makes there (``stage_reference()``
  below).  ``oracle/_ref/`` is git-ignored (the reference's sources never enter the history) but not gpurun-ignored, so
  the copy travels to the GPU box like a built ``.so``: the multi-GPU FSDP test compares against the real reference
  modules there, and ``bench.py --impl reference`` / ``cpu_baseline`` time them (``kind: "reference"``).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MERV_REFERENCE_ROOT", "/root/reference")
_SOURCE = os.path.join(REFERENCE_ROOT, "merv", "util", "nn_utils.py")
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "nn_utils.py")
_NN_UTILS = _SOURCE if os.path.isfile(_SOURCE) else _STAGED
_cached = None


def reference_available() -> bool:
    return os.path.isfile(_NN_UTILS)


def reference_origin() -> str:
    """Where the loaded file comes from (recorded in bench.py's cpu_baseline)."""
    return "/root/reference" if _NN_UTILS == _SOURCE else "oracle/_ref (staged copy of merv/util/nn_utils.py)"


def stage_reference() -> bool:
    """Copy the reference file into git-ignored ``oracle/_ref/`` (only where ``/root/reference`` exists).  Returns True if the
    staged copy exists afterwards."""
    if os.path.isfile(_SOURCE):
        import shutil

        os.makedirs(os.path.dirname(_STAGED), exist_ok=True)
        if not os.path.isfile(_STAGED) or open(_STAGED, "rb").read() != open(_SOURCE, "rb").read():
            shutil.copyfile(_SOURCE, _STAGED)
    return os.path.isfile(_STAGED)


def load_reference_nn_utils():
    """Returns the reference module (``ref.AveragePooling3DProjector``, ``ref.CrossAttentionAdapterLearnableQuery`` ...)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"neither {_SOURCE} nor {_STAGED} exists: run __graft_entry__.build() in the build container")
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
