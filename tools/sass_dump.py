#!/usr/bin/env python
"""Per-kernel SASS listings of libb2r.so (north star: "each kernel ships with a committed SASS listing").

  tools/sass_dump.py profiles/sass      writes <kernel>.sass.gz for every b2r kernel + INDEX.md (registers, instruction count,
                                        mnemonic histogram: LDG/STG/ATOM/SHFL/DFMA/FFMA..., no local-memory spills check)
"""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mrg_slam_b200", "libb2r.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    blocks = re.split(r"(?=\t\tFunction : )", sass)
    rows = []
    names = []
    for b in blocks:
        m = re.match(r"\t\tFunction : (\S+)", b)
        if not m:
            continue
        names.append(m.group(1))
    dm = demangle(names)
    for b in blocks:
        m = re.match(r"\t\tFunction : (\S+)", b)
        if not m:
            continue
        name = m.group(1)
        if "3b2r" not in name:
            continue  # CUB's radix-sort kernels are library code
        short = re.sub(r"[^A-Za-z0-9_]+", "_", dm[name].split("(")[0].replace("b2r::", "").replace("void ", "")).strip("_")
        ops = collections.Counter()
        n = 0
        for line in b.split("\n"):
            mm = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if mm:
                n += 1
                ops[mm.group(1).split(".")[0]] += 1
        with open(os.path.join(outdir, short + ".sass.gz"), "wb") as raw:  # mtime=0: unchanged kernels give unchanged files
            with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as f:
                f.write(b.encode())
        u = usage.get(name, "")
        reg = re.search(r"REG:(\d+)", u)
        stack = re.search(r"STACK:(\d+)", u)
        local = re.search(r"LOCAL:(\d+)", u)
        smem = re.search(r"SHARED:(\d+)", u)
        top = ", ".join(f"{k} {v}" for k, v in ops.most_common(8))
        has_local = ops.get("LDL", 0) + ops.get("STL", 0)
        rows.append((short, n, reg.group(1) if reg else "?", smem.group(1) if smem else "?", stack.group(1) if stack else "?",
                     local.group(1) if local else "?", has_local, top))
    with open(os.path.join(outdir, "INDEX.md"), "w") as f:
        f.write("# SASS listings of libb2r.so (sm_100a, `cuobjdump -sass`), one gzip per kernel\n\n"
                "Regenerate: `python tools/sass_dump.py profiles/sass` after `python __graft_entry__.py`.\n"
                "LDL/STL = local-memory loads/stores in the listing (spills or dynamically indexed arrays).\n\n"
                "| kernel | SASS instr | regs | smem B | stack B | local B | LDL+STL | most frequent opcodes |\n|---|---|---|---|---|---|---|---|\n")
        for r in sorted(rows):
            f.write("| `%s` | %d | %s | %s | %s | %s | %d | %s |\n" % r)
    print(f"{len(rows)} kernels -> {outdir}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass"))
