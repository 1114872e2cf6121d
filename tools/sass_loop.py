#!/usr/bin/env python
"""Static proxy for the per-plane instruction cost of a march kernel: find the hottest loop
(largest backward branch) in the SASS of every kernel matching a regex and print its length
and opcode mix.   python tools/sass_loop.py 'KCg<double, 2, 1>' """
import collections
import re
import subprocess
import sys

so = "darsia_b200/csrc/libtvdk.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "."
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
names = [f.split("\n", 1)[0].strip() for f in funcs]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
for f, d in zip(funcs, dem):
    short = re.sub(r"void tvdk::tvdk_kernel<tvdk::(.*)>\(.*", r"\1", d).replace("(int)", "").replace("(bool)", "")
    if not re.search(pat, short):
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, txt in ins:
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", txt)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < addr:
                loops.append((addr - tgt, tgt, addr))
    if not loops:
        print(short, ": no loop found", len(ins))
        continue
    # the march loop is the smallest loop containing the block barrier
    cand = []
    for span, tgt, end in loops:
        body = [t for a, t in ins if tgt <= a <= end]
        if any("BAR.SYNC" in t for t in body):
            cand.append((span, tgt, end))
    span, tgt, end = min(cand) if cand else max(loops)
    body = [t for a, t in ins if tgt <= a <= end]
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(f"{short}: total {len(ins)} instr; main loop {len(body)} instr: " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
