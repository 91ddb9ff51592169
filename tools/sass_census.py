"""Instruction census of the replay kernels' SASS (cuobjdump), per timestep of the inner loop.

Used to (a) check the algorithmic flop count quoted in DESIGN.md / bench.py against what the
compiler actually emitted and (b) track instruction-count regressions.  Heuristic: the per-step body
of replay_tma_kernel<QR2, no LPF, no AUX> is the code between the full-barrier wait and the
empty-barrier arrive, divided by the number of unrolled timesteps; never-taken fallback blocks
(the selection-based quaternion conversion) are excluded by pattern.
    python tools/sass_census.py [path/to/libposekf_b200.so]
"""
import collections
import json
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "poseestimationkf_b200/libposekf_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        funcs[cur].append(line)


def opcode(line):
    toks = line.split()
    op = toks[2] if toks[1].startswith("@") else toks[1]
    return op.rstrip(";").split(".")[0]


out = {}
for name, lines in funcs.items():
    m = re.search(r"replay_(tma|ldg)_kernelILi(\d)ELb(\d)ELb(\d)ELb(\d)E", name)
    m2 = re.search(r"replay_tma2_kernelILi(\d)ELb(\d)ELb(\d)ELb(\d)E", name)
    if m:
        key = f"replay_{m.group(1)}<algo={m.group(2)},lpf={m.group(3)},aux={m.group(4)},comp={m.group(5)}>"
    elif m2:
        key = f"replay_tma2_packed<algo={m2.group(1)},lpf={m2.group(2)},aux={m2.group(3)},comp={m2.group(4)}>"
    else:
        continue
    ops = collections.Counter(opcode(l) for l in lines)
    fp = {k: ops[k] for k in ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "FSEL", "FSETP", "UTMALDG", "SYNCS", "LDS", "LDG", "STG",
                              "LDL", "STL", "HMMA")}
    out[key] = {"total_static": len(lines), **fp}
print(json.dumps(out, indent=1))
