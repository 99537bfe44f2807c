#!/usr/bin/env python
"""Summarises the `-Xptxas -v` logs the Makefile leaves next to the sources (csrc/*.ptxas.log) as a markdown table:
registers, spills, stack and static shared memory per kernel and build.

  python tools/ptxas_summary.py > profiles/rNN_ptxas.md
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    rows = []
    for path in sorted(glob.glob(os.path.join(ROOT, "husky-rover-mppi-isaacsim_b200", "csrc", "*.ptxas.log"))):
        build = os.path.basename(path).replace(".ptxas.log", "")
        t = open(path).read()
        for m in re.finditer(r"Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers(?:, used (\d+) barriers)?"
                             r"(?:, (\d+) bytes smem)?", t):
            rows.append((build, m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(2)), int(m.group(7) or 0)))
    dm = demangle(sorted({r[1] for r in rows}))
    print("# ptxas -v summary (sm_100a)\n")
    print("| build | kernel | registers | spill stores (B) | stack (B) | static smem (B) |\n|---|---|---|---|---|---|")
    for b, n, regs, spill, stack, smem in rows:
        name = re.sub(r"\(.*", "", dm[n]).replace("void ", "").replace("mppi::", "")
        print(f"| {b} | `{name}` | {regs} | {spill} | {stack} | {smem} |")


if __name__ == "__main__":
    main()
