#!/usr/bin/env python
"""Benchmark of the MERV fusion hot path (BASELINE.json metric: merv-full fusion videos/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the whole path (pool -> projector -> score/softmax -> mix) over one batch of 64 synthetic
videos PER GPU at merv-full shapes (frames [16,16,32,16] -> 16 temporal tokens each; BASELINE.json configs[1]).
Videos are independent, so ranks shard by batch with no collective on the data path ("scaling": "weak").
Rank 0 prints ONE JSON line.  `value` is device-resident throughput; `e2e` is the same metric through the
host-buffer API (H2D of the inputs and D2H of the prefix inside the timed region).

--impl reference times the reference's CPU implementation of the path (oracle/torch_port.py: the same ATen ops
in the same order — the Python reference itself cannot be installed offline) on the host cores, rank 0 only.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

DIMS = [1024, 1024, 768, 768]
PATCHES = [256, 256, 196, 196]
TOKENS_T = [16, 16, 16, 16]  # temporal tokens after the backbones for frames [16,16,32,16] (ViViT tubelet 2)
MUS = [0.0, 0.5, -0.5, 0.25]
LLM_DIM, OUT_TOKENS = 4096, 1024
# algorithmic bytes / FLOPs per video (bf16)
BYTES_IN = sum(t * n * c * 2 for t, n, c in zip(TOKENS_T, PATCHES, DIMS))  # 26,411,008
BYTES_POOLED = sum(OUT_TOKENS * c * 2 for c in DIMS)  # 7,340,032
BYTES_Y = OUT_TOKENS * LLM_DIM * 2  # 8,388,608 per encoder
BYTES_OUT = OUT_TOKENS * LLM_DIM * 2
FLOPS_LINEAR = 2 * OUT_TOKENS * LLM_DIM * sum(DIMS)  # 30.065 GFLOP
FLOPS_GELU = FLOPS_LINEAR + 4 * 2 * OUT_TOKENS * LLM_DIM * LLM_DIM  # 167.5 GFLOP


def ncu_traffic_bytes(profile_txt: str):
    """dram read + write bytes per launch from the committed `ncu --set full` summary under profiles/ (None if absent)."""
    path = os.path.join(REPO, "profiles", profile_txt)
    if not os.path.isfile(path):
        return None
    for line in open(path):
        if line.strip().startswith("traffic (dram read + write)"):
            return float(line.split()[-2]) * 1e6
    return None


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")  # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def make_cpu_inputs(batch: int, seed: int = 7):
    import torch

    g = torch.Generator().manual_seed(seed)
    return [torch.randn((batch, t, n, c), generator=g) + mu for t, n, c, mu in zip(TOKENS_T, PATCHES, DIMS, MUS)]


def cpu_reference_run(mlp_type: str, sample_videos: int, steps: int, warmup: int, dtype_name: str = "bf16"):
    """Times oracle/torch_port.py (the reference's ATen op sequence) on all host cores. Returns (videos/s, ms/step, cores)."""
    import torch

    from merv_b200.nn_utils import MervFusion  # parameter container only (seeded like merv.py:87); no CUDA involved
    from oracle import torch_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    m = MervFusion.build(DIMS, LLM_DIM, TOKENS_T, 64, mlp_type, seed=1024, fused=False)
    with torch.no_grad():
        m.feature_fusion.Q.mul_(64.0)
    pp = [{k: v.to(dtype) for k, v in p.projector.state_dict().items()} for p in m.projectors]
    fp = {k: v.to(dtype) for k, v in m.feature_fusion.state_dict().items()}
    feats = [f.to(dtype) for f in make_cpu_inputs(sample_videos)]
    run = lambda: torch_port.fusion_forward(feats, pp, fp, TOKENS_T, 8, mlp_type, OUT_TOKENS)  # noqa: E731
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return sample_videos * steps / dt, dt / steps * 1e3, cores


def reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    sample = args.cpu_sample_videos
    value, ms, cores = cpu_reference_run(args.projector, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "merv-full fusion videos/sec", "value": value, "unit": "videos/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world, note=f"reference CPU path (torch ATen ops of merv/util/nn_utils.py via oracle/torch_port.py), {sample} videos per step"),
        "cpu_baseline": {"value": value, "unit": "videos/s", "cores": cores, "cpu": cpu_model(), "kind": "port", "sample": f"{sample} merv-full videos per step, bf16, {args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": value, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fused_tokens_per_s": value * OUT_TOKENS, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world: int, note: str = ""):
    return {
        "workload": f"merv-full fusion bf16 batch {args.batch} per GPU on {world}xB200 (3d-avg pool + {args.projector} projector + cross-attn mix)",
        "encoders": ["languagebind", "dinov2", "vivit", "siglip"], "frames": [16, 16, 32, 16], "feature_shapes": [[16, 256, 1024], [16, 256, 1024], [16, 196, 768], [16, 196, 768]],
        "prefix_tokens": OUT_TOKENS, "llm_dim": LLM_DIM, "projector": args.projector, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "pipeline": args.mode, "parallelism": f"dp{world} (batch-sharded, no data-path collective)",
        "l2": f"inputs larger than L2: {args.input_sets} rotating input sets of {BYTES_IN * args.batch / 1e9:.2f} GB each", "note": note,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="videos per GPU per step (weak scaling, the default)")
    ap.add_argument("--global-batch", type=int, default=0, help="if > 0: strong scaling, this many videos per step split over the ranks (SURVEY.md §8d config 3: 64)")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip timing the reference's op sequence in torch eager on the same GPU")
    ap.add_argument("--projector", default="linear", choices=["linear", "gelu-mlp"])
    ap.add_argument("--mode", default="fused", choices=["fused", "unfused"])
    ap.add_argument("--input-sets", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-videos", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL all-gather of the prefixes (N>1)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    scaling = "weak"
    if args.global_batch > 0:
        assert args.global_batch % world == 0, "--global-batch must be a multiple of the number of ranks"
        args.batch, scaling = args.global_batch // world, "strong"
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import merv_b200 as M
    from merv_b200 import ops
    from merv_b200.pipeline import HostPipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: merv_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_peaks()
    B = args.batch
    module = M.MervFusion.build(DIMS, LLM_DIM, TOKENS_T, 64, args.projector, seed=1024, fused=(args.mode == "fused"))
    with torch.no_grad():
        module.feature_fusion.Q.mul_(64.0)  # non-degenerate mixing weights
    module = module.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)

    sets = []
    for s in range(args.input_sets):
        g = torch.Generator(device=dev).manual_seed(7 + rank + 1000 * s)
        sets.append([(torch.randn((B, t, n, c), generator=g, device=dev) + mu).to(torch.bfloat16)
                     for t, n, c, mu in zip(TOKENS_T, PATCHES, DIMS, MUS)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.inference_mode():
        with ops.KernelTimer(timing=True):  # same code path (per-kernel events) as the timed region: no first-use allocations inside it
            for i in range(args.warmup):
                out, w = module(sets[i % len(sets)])
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.3)  # nvidia-smi start-up: make sure samples land inside the timed region
        sampler.lines.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ops.KernelTimer(timing=True) as kt:
            barrier()
            e0.record()
            for i in range(args.steps):
                out, w = module(sets[i % len(sets)])
            e1.record()
            barrier()
        clocks = sampler.stop()
        ms_total = e0.elapsed_time(e1)
        durations = kt.durations_ms()
        launches = kt.launches
        if world > 1:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total / 1e3)
        checksum = float(out.float().abs().mean().item())

        # optional: all-gather of the fused prefixes over NVLink (reported separately, never inside `value`)
        gather_ms = None
        if args.gather and world > 1:
            from merv_b200.parallel import all_gather_prefix

            for _ in range(3):
                all_gather_prefix(out, B * world)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                all_gather_prefix(out, B * world)
            g1.record()
            barrier()
            gather_ms = g0.elapsed_time(g1) / 5

        # ---- e2e: host buffers in, host buffers out, through the public host API ----
        e2e = None
        if not args.no_e2e:
            host_in = [torch.empty(f.shape, dtype=f.dtype).pin_memory() for f in sets[0]]
            for h, f in zip(host_in, sets[0]):
                h.copy_(f)
            host_out = torch.empty((B, OUT_TOKENS, LLM_DIM), dtype=torch.bfloat16).pin_memory()
            host_w = torch.empty((B, len(DIMS)), dtype=torch.bfloat16).pin_memory()
            pipe = HostPipeline(module, chunk_videos=8, device=dev)
            pipe(host_in, host_out, host_w)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                pipe(host_in, host_out, host_w)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            e2e = {"value": world * B * args.e2e_steps / dt, "unit": "videos/s", "h2d_bytes_per_step": BYTES_IN * B,
                   "d2h_bytes_per_step": BYTES_OUT * B + B * len(DIMS) * 2, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "steps": args.e2e_steps, "api": "merv_b200.pipeline.HostPipeline (pinned host tensors, 8-video chunks with ramp-up/-down, copy/compute overlap)"}
            # what the bus alone allows for the same bytes (not part of any reported throughput): explains e2e
            c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            dev_out = torch.empty((B, OUT_TOKENS, LLM_DIM), dtype=torch.bfloat16, device=dev)
            torch.cuda.synchronize()
            c0.record()
            for h, f in zip(host_in, sets[0]):
                f.copy_(h, non_blocking=True)
            c1.record()
            host_out.copy_(dev_out, non_blocking=True)
            c2.record()
            torch.cuda.synchronize()
            e2e["h2d_only_ms"], e2e["d2h_only_ms"] = c0.elapsed_time(c1), c1.elapsed_time(c2)
            e2e["h2d_GBps"] = BYTES_IN * B / e2e["h2d_only_ms"] / 1e6
            e2e["frac_of_h2d_limit"] = e2e["h2d_only_ms"] / e2e["ms_per_step"]
            del host_in, host_out, dev_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (SURVEY.md §8d algorithmic bytes/FLOPs per video x videos per launch / measured duration) ----
    def avg(name):
        d = durations.get(name, [])
        return (sum(d) / len(d)) if d else None

    kernels = {}
    pool_calls = durations.get("merv_pool3d", [])
    pool_ms = (sum(pool_calls) / args.steps) if pool_calls else None  # per step (module-by-module mode pools each encoder separately)
    if pool_ms:
        gbs = (BYTES_IN + BYTES_POOLED) * B / (pool_ms * 1e-3) / 1e9
        kernels["merv_pool3d"] = {"bound": "hbm", "ms": pool_ms, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                  "algorithmic_bytes": (BYTES_IN + BYTES_POOLED) * B,
                                  "traffic": ncu_traffic_bytes("r1g_prof_pool3d_tma.txt") if B == 64 else None}
    flops = FLOPS_LINEAR if args.projector == "linear" else FLOPS_GELU
    if args.mode == "fused":
        gemm_name = "merv_fused_linear_mix"
        gemm_flops = (FLOPS_LINEAR if args.projector == "linear" else 4 * 2 * OUT_TOKENS * LLM_DIM * LLM_DIM) * B
    else:
        gemm_name = "merv_linear_bias_act"
        gemm_flops = None
    roofline = None
    g_ms = avg(gemm_name)
    if args.mode == "fused" and g_ms:
        tf = gemm_flops / (g_ms * 1e-3) / 1e12
        roofline = {"kernel": "gemm_bf16_tcgen05_kernel (merv_fused_linear_mix)", "bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"],
                    "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"], "frac_of_burst_peak": tf / peaks["tf_burst"], "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                    "ms_per_launch": g_ms,
                    # ncu capture of this kernel at this exact configuration (B=64, linear, fused), committed under profiles/
                    "traffic": ncu_traffic_bytes("r1g_prof_gemm_bf16_tcgen05.txt") if (args.projector == "linear" and B == 64) else None,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                    "algorithmic_bytes": (BYTES_POOLED + BYTES_OUT) * B + sum(LLM_DIM * c * 2 for c in DIMS)}
        kernels[gemm_name] = roofline
    elif args.mode == "unfused":
        d = durations.get("merv_linear_bias_act", [])
        if d:
            per_step = sum(d) / args.steps
            tf = flops * B / (per_step * 1e-3) / 1e12
            roofline = {"kernel": "gemm_bf16_tcgen05_kernel (merv_linear_bias_act, all projector GEMMs of a step)", "bound": "tensor", "achieved": tf,
                        "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"], "frac_of_burst_peak": tf / peaks["tf_burst"],
                        "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "ms_per_step": per_step, "traffic": None}
        mix_ms = avg("merv_softmax_mix")
        if mix_ms:
            gbs = (4 * BYTES_Y + BYTES_OUT) * B / (mix_ms * 1e-3) / 1e9
            kernels["merv_softmax_mix"] = {"bound": "hbm", "ms": mix_ms, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"]}
    for name, d in durations.items():
        kernels.setdefault(name, {"ms": sum(d) / len(d)})
        kernels[name]["calls_per_step"] = len(d) / args.steps

    # second comparator (SURVEY.md §8d "the real bar"): the reference's op sequence in PyTorch eager on this same B200
    torch_eager = None
    if world == 1 and not args.no_torch_eager:
        try:
            from oracle import torch_port

            pp = [{k: v for k, v in p.projector.state_dict().items()} for p in module.projectors]
            fp = dict(module.feature_fusion.state_dict())
            run = lambda i: torch_port.fusion_forward(sets[i % len(sets)], pp, fp, TOKENS_T, 8, args.projector, OUT_TOKENS)  # noqa: E731
            for i in range(3):
                ref_out, _ = run(i)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(5):
                ref_out, _ = run(i)
            t1.record()
            torch.cuda.synchronize()
            ms_ref = t0.elapsed_time(t1) / 5
            with torch.inference_mode():
                ours_out, _ = module(sets[4 % len(sets)])
            diff = float((ours_out.float() - ref_out.float()).abs().max() / ref_out.float().abs().max())
            torch_eager = {"value": B / (ms_ref * 1e-3), "unit": "videos/s", "ms_per_step": ms_ref, "speedup_of_this_repo": ms_ref / ms_step,
                           "max_rel_diff_vs_this_repo": diff,
                           "what": "oracle/torch_port.py (the reference's ATen op sequence: permute, adaptive_avg_pool3d, F.linear, stack, mean, MHA, bmm) in torch eager bf16 on the same GPU, same inputs"}
            del ref_out, ours_out
        except Exception as e:  # a comparator, never a dependency
            torch_eager = {"error": repr(e)[:200]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload: ~10 s of CPU work (a probe pass sizes the number of timed passes)
        _, probe_ms, _ = cpu_reference_run(args.projector, args.cpu_sample_videos, 1, 1)
        passes = max(3, min(40, int(10e3 / max(probe_ms, 1.0))))
        v, ms, cores = cpu_reference_run(args.projector, args.cpu_sample_videos, passes, 1)
        cpu_baseline = {"value": v, "unit": "videos/s", "cores": cores, "cpu": cpu_model(), "kind": "port", "ms_per_video": 1e3 / v,
                        "sample": f"{args.cpu_sample_videos} merv-full videos per pass (bf16), {passes} passes (~{passes * ms / 1e3:.0f} s) after 1 warm-up, "
                                  "torch ATen op sequence of the reference (oracle/torch_port.py)"}

    line = {
        "metric": "merv-full fusion videos/sec", "value": value, "unit": "videos/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "fused_tokens_per_s": value * OUT_TOKENS,
        "path_effective_GBps_per_gpu": (BYTES_IN + BYTES_OUT) * B / (ms_step * 1e-3) / 1e9,
        "path_TFLOPs_per_gpu": flops * B / (ms_step * 1e-3) / 1e12,
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "torch_eager_same_gpu": torch_eager, "e2e": e2e, "gpu_launches": launches,
        "gpu_launches_per_step": launches / args.steps, "clocks": clocks, "allgather_ms": gather_ms, "output_abs_mean": checksum,
        "peaks": peaks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
