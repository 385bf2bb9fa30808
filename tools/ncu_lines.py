#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep.
usage: python tools/ncu_lines.py report.ncu-rep [top_n [kernel_regex]]"""
import csv, io, subprocess, sys
kf = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", sys.argv[1]] + kf + ["--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d["Thread Instructions Executed"]), int(d["Instructions Executed"]), int(d["# Samples"]), int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
tot = sum(l[0] for l in lines); tots = sum(l[2] for l in lines)
print("total thread-instr %d, warp-instr %d, samples %d" % (tot, sum(l[1] for l in lines), tots))
print("--- by warp instructions"); 
for t, w, s, ln, src in sorted(lines, key=lambda l: -l[1])[:topn]: print("%5.1f%% winstr (lanes %4.1f) %5.1f%% samples  L%-4d %s" % (100. * w / max(1, sum(l[1] for l in lines)), t / max(1, w), 100. * s / max(1, tots), ln, src))
print("--- by stall samples")
for t, w, s, ln, src in sorted(lines, key=lambda l: -l[2])[:20]: print("%5.1f%% instr %5.1f%% samples  L%-4d %s" % (100. * t / tot, 100. * s / max(1, tots), ln, src))
