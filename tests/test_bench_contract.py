"""CPU test of the bench.py contract: the reference arm runs on host cores only and prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-videos", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "merv-full fusion videos/sec" and d["unit"] == "videos/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    from oracle.ref_loader import reference_available

    # the unmodified reference modules where their file is reachable (build container, or oracle/_ref on the GPU box), else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["imports_merv_b200"] is False, "the reference arm must not load the product (libmerv_fusion.so)"
    assert d["e2e"] == {"value": d["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_host_pipeline_chunk_schedule_covers_the_batch():
    from merv_b200.pipeline import chunk_schedule

    for chunk in (1, 4, 8):
        for batch in list(range(0, 40)) + [64, 67, 128]:
            sched = chunk_schedule(batch, chunk)
            assert [lo for lo, _ in sched] == [0] + [hi for _, hi in sched[:-1]] if sched else batch == 0
            assert (sched[-1][1] if sched else 0) == batch
            assert all(0 < hi - lo <= chunk for lo, hi in sched)
    assert [hi - lo for lo, hi in chunk_schedule(64, 8)] == [2, 2, 4, 8, 8, 8, 8, 8, 8, 4, 2, 2]
