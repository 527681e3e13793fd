#!/usr/bin/env python
"""tools/sass_cost.py <lib.so|cubin> <function-regex> [--loops N] -- static dispatch-cost estimate of a kernel's
innermost loops from its SASS (no GPU needed).

Cost model measured on B200 with tools/micro/fp32_mix.cu and pipe_mix.cu: one sub-partition dispatches
one warp instruction per cycle at best, and an instruction holds the dispatch port for
max(1, ceil-ish(#32-bit REGISTER source operands / 2)) cycles (packed F32x2 operands count as two
registers; uniform registers, constants and immediates are free); FMA-pipe and ALU-pipe instructions
do NOT overlap (FADD2 + LOP3 + IADD3 = 4.8 cycles, not 2.4).  So the estimate is simply the sum.

Prints, for each of the N largest loops (backward branches): instructions, estimated cycles per
iteration, and the opcode histogram.
"""
import collections
import re
import subprocess
import sys


def operands_cost(op, args):
    """(#register source operands in 32-bit units, cost in cycles)."""
    base = op.split(".")[0]
    parts = [a.strip() for a in args.split(",")] if args else []
    if not parts:
        return 0, 1.0
    stores = base in ("STS", "STG", "ST", "RED", "ATOMS", "ATOMG", "STL", "UTMALDG", "SYNCS", "BAR", "BRA", "EXIT",
                      "BSSY", "BSYNC", "WARPSYNC", "NANOSLEEP", "MEMBAR", "ERRBAR", "CCTL", "ISETP", "FSETP", "UISETP",
                      "PLOP3", "DSETP")
    srcs = parts if stores else parts[1:]
    if base in ("ISETP", "FSETP", "DSETP", "PLOP3", "UISETP"):
        srcs = parts[2:]
    n = 0
    for a in srcs:
        for m in re.finditer(r"(?<![UP\w])R(\d+)(\.64|\.F32x2[.\w]*)?", a):
            wide = bool(m.group(2)) or ".64" in a
            n += 2 if wide else 1
    if base in ("DADD", "DFMA", "DMUL"):
        n = max(n, 4)
    cost = max(1.0, n / 2.0)
    if base in ("FADD2", "FMUL2", "FFMA2"):
        cost = max(cost, 2.0)
    return n, cost


def main():
    if len(sys.argv) < 3:
        print(__doc__)
        sys.exit(1)
    lib, pat = sys.argv[1], re.compile(sys.argv[2])
    nloops = 3
    if "--loops" in sys.argv:
        nloops = int(sys.argv[sys.argv.index("--loops") + 1])
    rng = None
    if "--range" in sys.argv:                      # --range 0x3840:0x6cf0
        a, b = sys.argv[sys.argv.index("--range") + 1].split(":")
        rng = (int(a, 16), int(b, 16))
    show_back = "--branches" in sys.argv
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        if not (pat.search(name) or pat.search(dem)):
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?)\s*;", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
        print(f"== {dem[:150]}\n   {len(ins)} instructions")
        loops = []
        for addr, op, args in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", args)
                if t and int(t.group(1), 16) < addr:
                    loops.append((int(t.group(1), 16), addr))
        loops.sort(key=lambda ab: ab[0] - ab[1])
        if show_back:
            print("   backward branches:", ", ".join(f"{lo:#x}..{hi:#x}" for lo, hi in sorted(loops)))
        for lo, hi in ([rng] if rng else loops[:nloops]):
            body = [(a, o, g) for a, o, g in ins if lo <= a <= hi]
            hist = collections.Counter()
            cyc = 0.0
            for a, o, g in body:
                _, c = operands_cost(o, g)
                cyc += c
                hist[o.split(".")[0]] += 1
            top = ", ".join(f"{k} {v}" for k, v in hist.most_common(18))
            print(f"   loop {lo:#x}..{hi:#x}: {len(body)} instr, ~{cyc:.0f} dispatch cycles/iter\n      {top}")


if __name__ == "__main__":
    main()
