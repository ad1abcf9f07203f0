"""Summarise .ncu-rep captures (read here, no GPU needed) into profiles/<name>.txt.

    python scripts/ncu_summary.py gpurun_out/r1_prof_gemm_bf16_tcgen05.ncu-rep [...]
"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def summarise(path: str) -> str:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none capture: {os.path.basename(path)}"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')}  block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] not in ("", "no data"):
                out.append(f"  {k:85s} {d[k]:>18s} {u.get(k, '')}")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = float(rd) * scale[u["dram__bytes_read.sum"]] + float(wr) * scale[u["dram__bytes_write.sum"]]
            out.append(f"  traffic (dram read + write)                                                            {tot / 1e6:18.1f} MB")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    os.makedirs("profiles", exist_ok=True)
    for p in sys.argv[1:]:
        text = summarise(p)
        dst = os.path.join("profiles", os.path.basename(p).replace(".ncu-rep", ".txt"))
        open(dst, "w").write(text)
        print(text)
