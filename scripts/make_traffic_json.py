#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --page raw --csv` export of one step at 6.25e7 particles (scripts/r2_profile.sh):
DRAM bytes and FP64 work per particle of every kernel of the dominant phase.  usage: make_traffic_json.py raw.csv particles out.json"""
import csv, json, sys

raw, npart, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}

def val(r, k, scale=None):
    v = float(r[ix[k]].replace(",", ""))
    u = units[ix[k]]
    if k.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if k.startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v

kern = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    d = {"us": val(r, "gpu__time_duration.sum"), "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum")}
    # thread-level FP64 operations: (2 DFMA + DMUL + DADD) per elapsed cycle x elapsed SMSP cycles
    cyc = None
    for k in ("smsp__cycles_elapsed.max", "smsp__cycles_elapsed.avg", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg"):
        if k in ix:
            cyc = float(r[ix[k]].replace(",", ""))
            break
    f = 0.0
    for k, w in (("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", 2.0),
                 ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", 1.0),
                 ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", 1.0)):
        if k in ix and cyc:
            f += w * float(r[ix[k]].replace(",", "")) * cyc
    d["fp64_flop"] = f
    kern[name] = d
dom = [k for k in kern if k.startswith("k_bin_push") or k.startswith("k_far_walk") or k.startswith("k_far_hint")]
res = {"particles": npart, "kernels": {k: {kk: vv for kk, vv in v.items()} for k, v in kern.items()},
       "dram_bytes_per_particle_interp_push_track": sum(kern[k]["dram_read_bytes"] + kern[k]["dram_write_bytes"] for k in dom) / npart,
       "fp64_flop_per_particle_interp_push_track": sum(kern[k]["fp64_flop"] for k in dom) / npart,
       "dram_bytes_per_particle_step": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in kern.values()) / npart,
       "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of k_bin_push, k_far_hint and k_far_walk at 32^3 elements, "
                 "%.4g particles (1907 per element, as the flagship workload); profiles/r2_ncu_full_32cube_summary.txt" % npart}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: res[k] for k in res if k != "kernels"}, indent=1))
