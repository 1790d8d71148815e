#!/usr/bin/env python
"""Opcode histogram of the hot kernels' SASS (cuobjdump -sass of the built library): which pipe the instructions of a kernel go to.
usage: python scripts/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bulletproofs-plus_b200/libbpp_b200.so"
want = ["k_msm_bucket_threadILi3", "k_decompress_proofsILi5", "k_fb_msmILi1", "k_vprep_proofE", "k_replay_sm", "k_msm_sort_segwILi9", "k_msm_reduce_warp", "k_msm_combine"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((w for w in want if w in m.group(1)), None)
        if cur:
            hist[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[cur][m.group(2)] += 1
print("# SASS opcode histograms (sm_100a, `cuobjdump -sass`), static instruction counts per kernel\n")
print("Integer multiplies are `IMAD.WIDE.U32` / `IMAD.WIDE.U32.X` (32x32 -> 64 with 64-bit accumulate, the FMA pipe), carries `IADD3` / `IADD3.X` /")
print("`IMAD.X` (ALU / FMA), nothing on the tensor or FP64 pipes; loads are 128-bit (`LDG.E.128`).\n")
for k, h in hist.items():
    tot = sum(h.values())
    mul = sum(v for o, v in h.items() if o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))
    print("## `%s`: %d instructions, %d wide multiplies (%.0f %%)\n" % (k, tot, mul, 100.0 * mul / max(1, tot)))
    print("| opcode | count |\n|---|---|")
    for o, v in h.most_common(14):
        print("| `%s` | %d |" % (o, v))
    print()
