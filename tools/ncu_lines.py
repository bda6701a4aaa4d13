#!/usr/bin/env python
"""Per-source-line instruction counts of one profiled kernel: joins ncu's per-SASS-instruction metrics
(`ncu -i X.ncu-rep --page source --csv --print-source sass`) with nvdisasm's line table of the same build.

  tools/ncu_lines.py gpurun_out/ncu_X.ncu-rep <mangled kernel name substring> [cubin module: cloud|lsq|ndt|filters] [top N]

Output: for the innermost (first) "File/line" annotation of each instruction, summed warp instructions, thread instructions,
lane utilisation and stall samples — the numbers quoted in profiles/*.md.  The .so must be the build that was profiled.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(module, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mrg_slam_b200", "libb2r.so")], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(module + ".")][0]
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table = {}  # offset -> (innermost "file:line", outermost "file:line")
    in_k = False
    pend = []
    for line in dis.split("\n"):
        if line.startswith("\t.section\t.text."):
            in_k = kernel_sub in line
            pend = []
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            pend.append(f"{os.path.basename(m.group(1))}:{m.group(2)}")
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            if pend:
                cur = (pend[0], pend[-1])
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
            pend = []
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    module = sys.argv[3] if len(sys.argv) > 3 else "cloud"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    table = line_table(module, ksub)
    base = int(rows[0]["Address"], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    agg_outer = collections.defaultdict(lambda: [0, 0, 0, 0])
    ops = collections.defaultdict(lambda: [0, 0])
    tot = [0, 0, 0]
    for r in rows:
        off = int(r["Address"], 16) - base
        (inner, outer), sass = table.get(off, (("?", "?"), r["Source"]))
        wi, ti, ss = int(r["Instructions Executed"]), int(r["Thread Instructions Executed"]), int(r["# Samples"])
        for a, key in ((agg, inner), (agg_outer, outer)):
            a[key][0] += wi; a[key][1] += ti; a[key][2] += ss; a[key][3] += 1
        op = re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip()).split()[0].split(".")[0]
        ops[op][0] += wi; ops[op][1] += ti
        tot[0] += wi; tot[1] += ti; tot[2] += ss
    print(f"total warp instr {tot[0]:,}  thread instr {tot[1]:,}  lanes/instr {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]:,}\n")
    for title, a in (("innermost source line", agg), ("outermost (kernel-level) source line", agg_outer)):
        print(f"| {title} | SASS | warp instr | share | lanes/instr | stall samples share |\n|---|---|---|---|---|---|")
        for key, v in sorted(a.items(), key=lambda x: -x[1][0])[:top]:
            print(f"| {key} | {v[3]} | {v[0]:,} | {100 * v[0] / tot[0]:.1f}% | {v[1] / max(v[0], 1):.1f} | {100 * v[2] / max(tot[2], 1):.1f}% |")
        print()
    print("| opcode | warp instr | share | lanes/instr |\n|---|---|---|---|")
    for op, v in sorted(ops.items(), key=lambda x: -x[1][0])[:20]:
        print(f"| {op} | {v[0]:,} | {100 * v[0] / tot[0]:.1f}% | {v[1] / max(v[0], 1):.1f} |")


if __name__ == "__main__":
    main()
