"""Basic-block view of a kernel's SASS with the register-file read model (developer tool, CPU only).

    python tools/sass_blocks.py file.cubin|file.so [kernel-substring] [min_instructions]

Splits the kernel at branch instructions and branch targets, and for every block of at least `min_instructions`
prints: instructions, packed FP32 instructions, how many of them carry an operand that the PREVIOUS instruction left
in the reuse cache (same slot, `.reuse`), register operands read (64-bit for packed operands) and the model cycles
    cycles = sum over instructions of max(issue cycles, 64-bit-equivalent register reads)
which is what the round-2 micro-benchmarks (profiles/microbench/r02_pipes.cu) say the SM sustains: one 64-bit operand
per lane per cycle from the register file, whatever pipe the instruction goes to.
"""
import collections
import re
import subprocess
import sys

path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "replay_tma2_kernel"
min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
cur, lines = None, []
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if cur and pat in cur and mm:
        lines.append((int(mm.group(1), 16), mm.group(2).strip()))
targets = set()
for addr, body in lines:
    m = re.search(r"\b(BRA|BSSY|BSYNC|CALL|JMP)\b.*?(0x[0-9a-f]+)", body)
    if m and "BRA" in body:
        targets.add(int(m.group(2), 16))
blocks, cur_b = [], []
for addr, body in lines:
    if addr in targets and cur_b:
        blocks.append(cur_b)
        cur_b = []
    cur_b.append((addr, body))
    op = body.split()[1] if body.startswith("@") else body.split()[0]
    if op.split(".")[0] in ("BRA", "EXIT", "RET", "BRX", "JMP"):
        blocks.append(cur_b)
        cur_b = []
if cur_b:
    blocks.append(cur_b)

PACKED = ("FFMA2", "FMUL2", "FADD2")


def analyse(block):
    ops = collections.Counter()
    prev_reuse = {}
    cycles = reads64 = reuse_hits = 0.0
    for addr, body in block:
        toks = body.split(None, 1)
        if toks[0].startswith("@"):
            toks = toks[1].split(None, 1)
        op = toks[0].split(".")[0]
        ops[op] += 1
        args = [a.strip() for a in toks[1].split(",")] if len(toks) > 1 else []
        srcs = args[1:]
        this_reuse, r = {}, 0.0
        for slot, a in enumerate(srcs):
            m = re.match(r"[-|~!]*\|?(R\d+)", a)
            if not m:
                continue
            reg = m.group(1)
            wide = 1.0 if ("F32x2" in a or op in ("DFMA", "DMUL", "DADD") or ".64" in toks[0]) else 0.5
            if prev_reuse.get(slot) == reg:
                reuse_hits += 1
            else:
                r += wide
            if ".reuse" in a:
                this_reuse[slot] = reg
        prev_reuse = this_reuse
        issue = 2.0 if op in PACKED else 1.0
        reads64 += r
        cycles += max(issue, r)
    return ops, reads64, reuse_hits, cycles


for b in blocks:
    if len(b) < min_n:
        continue
    ops, reads64, hits, cyc = analyse(b)
    packed = sum(ops[o] for o in PACKED)
    print(f"block {b[0][0]:#06x}..{b[-1][0]:#06x}: {len(b)} instr, packed {packed} (FFMA2 {ops['FFMA2']} FMUL2 {ops['FMUL2']} FADD2 {ops['FADD2']}), "
          f"FSEL {ops['FSEL']} MUFU {ops['MUFU']} LDS {ops['LDS']} other {len(b) - packed - ops['FSEL'] - ops['MUFU'] - ops['LDS']}; "
          f"reuse hits {hits:.0f}, 64-bit reads {reads64:.0f}, model cycles {cyc:.0f} (pipe-only {2 * packed + len(b) - packed})")
