#!/usr/bin/env python
"""Where the Philox instructions of a rollout loop sit relative to the chain's FP32 instructions (SASS of a fused kernel).

  cuobjdump -sass -fun <mangled kernel> kernels_strict.o > k.sass;  python tools/sass_loop_mix.py k.sass

For every backward branch spanning >= 400 instructions (the rollout loops) prints the basic blocks of the body
(split at BRA / BSSY / BSYNC) with their instruction count and the number of IMAD.WIDE (Philox multiplies) and MUFU
in each: noise that precedes the chain shows up as a block of its own, noise that fills the chain's stall slots shares a
block with FMUL / FADD / FFMA.  CPU only."""
import collections
import re
import sys


def main(path):
    ins = []
    for l in open(path):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)', t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        n = (a - tgt) // 16
        if tgt >= a or n < 400 or n > 1500:
            continue
        body = [(x, y) for x, y in ins if tgt <= x <= a]
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instructions")
        blk, start = [], body[0][0]
        targets = set()
        for x, y in body:
            mm = re.search(r'(?:BRA|BSSY)\S*\s+(?:\S+,\s*)?(?:B\d+,\s*)?`?\(?0x([0-9a-f]+)', y)
            if mm:
                targets.add(int(mm.group(1), 16))
        for x, y in body:
            if x in targets and blk:
                report(start, blk)
                blk, start = [], x
            blk.append(y)
            if re.search(r'\b(BRA|BSYNC)\b', y):
                report(start, blk)
                blk, start = [], x + 16
        if blk:
            report(start, blk)


def report(start, blk):
    ops = collections.Counter((t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0] for t in blk)
    wide = sum(1 for t in blk if 'IMAD.WIDE' in t)
    fp = ops['FMUL'] + ops['FADD'] + ops['FFMA']
    print(f"  block {start:#x}: {len(blk):4d} instr   fp32 {fp:4d}   IMAD.WIDE {wide:3d}   MUFU {ops['MUFU']:2d}   LDG {ops['LDG']:2d}   LDS {ops['LDS']:2d}")


if __name__ == "__main__":
    main(sys.argv[1])
