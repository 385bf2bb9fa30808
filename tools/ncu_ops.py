#!/usr/bin/env python
"""ncu --csv metric dump (one launch per kernel) -> profiles/fp64_ops.json entry.
usage: python tools/ncu_ops.py <workload> <ncu.csv> [profiles/fp64_ops.json]
Per kernel of one STEADY step (not the first step after an upload): FP64 thread operations (dfma, dmul, dadd), DRAM
bytes, duration; the launch with the longest duration is kept for kernels that run more than once per step (chunk loop)."""
import csv, json, os, sys
NAMES = {"k_face_states": "k4a_face_states", "k_face_setup": "k4b1_face_setup", "k_face_iterate": "k4b_face_riemann",
         "k_face_finish": "k4b3_face_finish", "k_flux_sum_update": "k4c_flux_sum_update", "k_face_index": "k2b_face_index",
         "k_neighbours": "k2_neighbours", "k_density_matrix": "k3_density_matrix", "k_gradient_limit": "k3b_gradient_limit",
         "k_gather_sorted": "k1_gather", "k_cell_key": "k1_cell_key", "k_sort_within_cells": "k1_sort_within_cells", "k_bbox<": "k0_bbox"}
wl, path = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fp64_ops.json")
rows = list(csv.reader(open(path)))
hdr = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
launches = {}
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    d = dict(zip(h, r))
    launches.setdefault(d["ID"], {"name": d["Kernel Name"]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
# One STEADY step: the launch list is cut at every k_cell_key (first kernel of a step); segments that are incomplete
# (no k_flux_sum_update), that contain an upload (k_iota) or that directly follow one (the first step from the initial
# condition runs more Riemann iterations than the steps bench.py times) are dropped, the last remaining one is used.
order = sorted(launches, key=lambda k: int(k))
segs, cur = [], []
for k in order:
    if "k_cell_key" in launches[k]["name"] and cur:
        segs.append(cur)
        cur = []
    cur.append(k)
segs.append(cur)
def has(seg, what):
    return any(what in launches[k]["name"] for k in seg)
good = [i for i, sg in enumerate(segs) if has(sg, "k_cell_key") and has(sg, "k_flux_sum_update") and not has(sg, "k_iota")
        and not (i > 0 and has(segs[i - 1], "k_iota")) and i > 0]
if not good:
    sys.exit("ncu_ops.py: no steady step in the capture (%d segments)" % len(segs))
step = segs[good[-1]]
print("# %d launches captured, %d step segments, steady candidates %s, using segment %d (%d launches)"
      % (len(order), len(segs), good, good[-1], len(step)))
res = {}
for k in step:
    L = launches[k]
    short = next((v for kk, v in NAMES.items() if kk in L["name"]), None)
    if not short:
        continue
    e = {"dfma": L.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0),
         "dmul": L.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0),
         "dadd": L.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0),
         "thread_inst": L.get("smsp__thread_inst_executed.sum", 0.0),
         "dram_bytes": L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0),
         "ncu_us": L.get("gpu__time_duration.sum", 0.0) / 1e3}
    if short not in res or e["ncu_us"] > res[short]["ncu_us"]:
        res[short] = e
allj = {}
if os.path.exists(out_path):
    allj = json.load(open(out_path))
allj[wl] = res
allj["_note"] = ("per launch, one step of the workload after warm-up; ncu --clock-control none; FP64 ops are thread-level "
                 "SASS counts (DFMA counts 2 flop in bench.py); dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum")
json.dump(allj, open(out_path, "w"), indent=1, sort_keys=True)
for k, v in sorted(res.items()):
    print("%-22s dfma %.3e dmul %.3e dadd %.3e dram %.1f MB  %.1f us" % (k, v["dfma"], v["dmul"], v["dadd"], v["dram_bytes"] / 1e6, v["ncu_us"]))
