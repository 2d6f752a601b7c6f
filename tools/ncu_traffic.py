#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` capture of one risk-stage launch group of configs[1] (k_rollouts<ROLL_OPT>, k_inner_cem_fast,
k_opt_risk; 200 episodes): dram__bytes_read.sum + dram__bytes_write.sum per kernel and their sum, which bench.py reports as roofline.traffic.
usage: ncu_traffic.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ci[name]].replace(",", "")); u = units[ci[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ks = {}
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
    ks[name] = {"dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                "duration_us": float(r[ci["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[ci["gpu__time_duration.sum"]].replace("second", "s")[:2].strip(), 1),
                "grid": r[ci["Grid Size"]] if "Grid Size" in ci else None}
tot = sum(k["dram_read"] + k["dram_write"] for k in ks.values())
json.dump({"dram_bytes_per_risk_launch": tot, "kernels": ks,
           "source": "ncu --set full --clock-control none, one launch each of the risk-stage kernels of configs[1] at 200 episodes (%s)" % sys.argv[1].split("/")[-1]},
          open(sys.argv[2], "w"), indent=1)
print(json.dumps(ks, indent=1)); print("total", tot)
