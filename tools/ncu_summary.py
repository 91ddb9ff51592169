"""Turns an `ncu --set full` report of the replay kernel into the small JSON/markdown summary that is
committed under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).
    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME --filters N --timesteps T
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("out_prefix")
ap.add_argument("--filters", type=int, default=1 << 20)
ap.add_argument("--timesteps", type=int, default=200)
ap.add_argument("--bytes-per-step", type=int, default=36, help="algorithmic bytes per filter-step (36 in; +16 with the trajectory stored)")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def get(name, cast=float):
    try:
        return cast(vals[col[name]].replace(",", ""))
    except Exception:
        return None


def unit(name):
    return units[col[name]] if name in col else None


def to_bytes(name):
    v, u = get(name), unit(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    return None if v is None else v * scale


def to_ms(name):
    v, u = get(name), unit(name)
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
    return None if v is None else v * scale


steps = a.filters * a.timesteps
dur_ms = to_ms("gpu__time_duration.sum")
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
inst = get("smsp__inst_executed.sum")
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(vals[i])
          for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench      # noqa: E402  (one recipe for the source stamp: bench.kernel_source_sha)
summary = {
    "report": a.rep,
    # stamp: bench.py quotes this capture's DRAM traffic only while the device sources still hash to this value
    # (run this script before touching the sources the capture was taken from)
    "kernel_source_sha": bench.kernel_source_sha(),
    "kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None,
    "workload": {"filters": a.filters, "timesteps": a.timesteps, "filter_steps": steps},
    "duration_ms_under_ncu": dur_ms,
    "registers_per_thread": get("launch__registers_per_thread"),
    "local_memory_ld_st_inst": [get("smsp__inst_executed_op_local_ld.sum"), get("smsp__inst_executed_op_local_st.sum")],
    "occupancy_limit_blocks": {"registers": get("launch__occupancy_limit_registers"),
                               "shared_mem": get("launch__occupancy_limit_shared_mem")},
    "warps_active_pct_of_peak": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes_read": rd, "dram_bytes_write": wr,
    "dram_bytes_per_launch": None if rd is None else rd + wr,
    "algorithmic_bytes_per_launch": steps * a.bytes_per_step,
    "dram_traffic_over_algorithmic": None if rd is None else (rd + wr) / (steps * a.bytes_per_step),
    "dram_throughput_pct_of_peak": get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "achieved_dram_gbs": None if rd is None else (rd + wr) / (dur_ms * 1e-3) / 1e9,
    "warp_inst_executed": inst,
    "inst_per_filter_step": None if inst is None else inst * 32 / steps,
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "eligible_warps_per_cycle": get("smsp__warps_eligible.avg.per_cycle_active"),
    "active_warps_per_scheduler": get("smsp__warps_active.avg.per_cycle_active"),
    "sm_cycles_elapsed_avg": get("sm__cycles_elapsed.avg"),
    "sm_mhz_under_ncu": None if dur_ms is None else get("sm__cycles_elapsed.avg") / (dur_ms * 1e-3) / 1e6,
    "thread_inst_per_cycle_elapsed": {k: get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed")
                                      for k in ("ffma", "fmul", "fadd")},
    "stall_reasons_warps_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10]),
}
fm = summary["thread_inst_per_cycle_elapsed"]
if all(v is not None for v in fm.values()) and summary["sm_cycles_elapsed_avg"]:
    cyc = summary["sm_cycles_elapsed_avg"]
    summary["fp32_thread_inst_per_filter_step"] = {k: v * cyc / steps for k, v in fm.items()}
    f = summary["fp32_thread_inst_per_filter_step"]
    summary["flops_per_filter_step_from_counters"] = 2 * f["ffma"] + f["fmul"] + f["fadd"]
json.dump(summary, open(a.out_prefix + ".json", "w"), indent=1)
with open(a.out_prefix + ".md", "w") as fh:
    fh.write(f"# ncu summary: {summary['kernel']}\n\nsource report: `{a.rep}` (scratch, not committed)\n\n| metric | value |\n|---|---|\n")
    for k, v in summary.items():
        if k not in ("report", "kernel"):
            fh.write(f"| {k} | {json.dumps(v)} |\n")
print(json.dumps(summary, indent=1))
