"""Multi-GPU parity (-m gpu, skipped on a box with fewer than 2 GPUs; run with `gpurun --gpus 2`): the fused GEMM + all-gather
over NVLink (multicast and unicast transports) is bit-identical to compute-then-NCCL on every rank."""
import pytest
import torch

from tests.test_fsdp import run_ranks

pytestmark = pytest.mark.gpu


def test_fused_gemm_allgather_is_bit_identical_to_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    reports = run_ranks("gather_worker.py", world, timeout=600)
    for rep in reports:
        assert rep["ok"], rep
        assert len(rep["cases"]) == 4 and all(c["bit_identical"] for c in rep["cases"])
        assert any(c["transport"].startswith("unicast") for c in rep["cases"])
