"""Writes profiles/r2_traffic.json — measured DRAM traffic per launch of the dominant kernel of each bench workload —
from `ncu --set full` captures (.ncu-rep) of the bench commands. bench.py reads the file for `roofline.traffic`
(a profiler cannot run inside the timed bench).

    python tools/ncu_traffic.py profiles/r2_traffic.json  <key>=<report.ncu-rep>:<kernel substring>:<workload text> ...

key = the name bench.py looks up ("k_union<COLLECT>", "c1_term_top10", ...). Needs ncu (reads reports; no GPU)."""
import csv
import io
import json
import subprocess
import sys


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


# the raw page prints each metric in the unit of its own column (row 2 of the csv): scale to bytes / ns
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "": 1.0}


def main():
    dst, specs = sys.argv[1], sys.argv[2:]
    out = {"capture": "ncu --set full --clock-control none (see profiles/r2_summary.md for the commands)", "kernels": {}}
    for spec in specs:
        key, rest = spec.split("=", 1)
        rep, sub, workload = rest.split(":", 2)
        hdr, units, rows = rows_of(rep)
        unit = dict(zip(hdr, units))
        tot, n, dur = 0.0, 0, 0.0
        for r in rows:
            d = dict(zip(hdr, r))
            if sub not in d.get("Kernel Name", ""):
                continue
            f = lambda k: float(d[k].replace(",", "")) * SCALE[unit.get(k, "")] if d.get(k) else 0.0
            tot += f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
            dur += f("gpu__time_duration.sum")
            n += 1
        if n:
            out["kernels"][key] = {"dram_bytes_per_launch": tot / n, "launches": n, "kernel": sub, "workload": workload,
                                   "report": rep.split("/")[-1], "ncu_duration_ns_per_launch": dur / n}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
