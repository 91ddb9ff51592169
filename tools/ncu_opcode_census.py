"""Executed-instruction census of one kernel from an ncu report's source page (dynamic counts, per opcode),
normalised per filter-step.  The packed kernel's FFMA2/FMUL2/FADD2 do not show up in ncu's per-op FP32 thread
counters, so this is how the executed flop count quoted in DESIGN.md / bench.py is measured.
    python tools/ncu_opcode_census.py gpurun_out/X.ncu-rep --filters N --timesteps T [--out profiles/NAME.json]
"""
import argparse
import collections
import csv
import io
import json
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--filters", type=int, default=1 << 20)
ap.add_argument("--timesteps", type=int, default=200)
ap.add_argument("--out")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernel = rows[0][1]
ci = {h: i for i, h in enumerate(rows[1])}
warp, thread, samples = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    toks = r[ci["Source"]].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).rstrip(";").split(".")[0]
    warp[op] += int(r[ci["Instructions Executed"]])
    thread[op] += int(r[ci["Predicated-On Thread Instructions Executed"]])
    samples[op] += int(r[ci["# Samples"]])
steps = a.filters * a.timesteps
per = {op: thread[op] / steps for op in thread}      # thread instructions per filter-step (a packed thread serves 2 filters)
packed = {k: per.get(k, 0.0) for k in ("FFMA2", "FMUL2", "FADD2")}
scalar = {k: per.get(k, 0.0) for k in ("FFMA", "FMUL", "FADD", "MUFU")}
# lane operations per filter-step: a packed instruction is two lane operations of one thread = one per filter... the
# thread owns TWO filters, so thread-inst per filter-step already counts one lane-op pair per two filters:
lane_ops = {"fma": 2 * packed["FFMA2"] + scalar["FFMA"], "mul": 2 * packed["FMUL2"] + scalar["FMUL"],
            "add": 2 * packed["FADD2"] + scalar["FADD"], "mufu": scalar["MUFU"]}
flops = 2 * lane_ops["fma"] + lane_ops["mul"] + lane_ops["add"] + lane_ops["mufu"]
out = {"kernel": kernel, "workload": {"filters": a.filters, "timesteps": a.timesteps, "filter_steps": steps},
       "warp_inst_total": sum(warp.values()), "thread_inst_per_filter_step_total": sum(thread.values()) / steps,
       "thread_inst_per_filter_step": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:24]},
       "fp32_lane_ops_per_filter_step": {k: round(v, 2) for k, v in lane_ops.items()},
       "fp32_lane_ops_total_per_filter_step": round(lane_ops["fma"] + lane_ops["mul"] + lane_ops["add"], 2),
       "flops_per_filter_step_executed": round(flops, 1),
       "stall_samples_by_opcode": dict(samples.most_common(12))}
print(json.dumps(out, indent=1))
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
