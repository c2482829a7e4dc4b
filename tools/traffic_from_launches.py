"""Dev tool: turns an ncu launch list (long CSV, one row per launch and metric) of a compose run into the per-step numbers of
profiles/traffic.json.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"blend|pyrdown|warp_tiles|seam|mirror" -c 80 --csv --log-file launches.csv python tools/ab_env.py --workload cfg4 --steps 1 --variants ISB_PDL=1
  python tools/traffic_from_launches.py launches.csv [--update cfg4]

A step starts at each seam_prep_kernel launch; the LAST complete step of the list is summed (the first ones include plan-time
work and cold tables).  Prints per-kernel rows and the totals; --update writes the totals into profiles/traffic.json.
"""
import argparse
import csv
import json
import os
import re
import sys


def read(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    launches = {}
    for r in rows:
        k = int(r["ID"])
        e = launches.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        e[m] = v
    return [launches[k] for k in sorted(launches)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--update", default=None, help="workload key of profiles/traffic.json to overwrite")
    ap.add_argument("--source", default=None, help="text for the entry's `source` field")
    a = ap.parse_args()
    ls = read(a.csv)
    starts = [i for i, e in enumerate(ls) if e["name"].startswith("seam_prep_kernel")]
    if len(starts) < 2:
        sys.exit("need at least two seam_prep_kernel launches to delimit one complete step")
    lo, hi = starts[-2], starts[-1]
    step = ls[lo:hi]
    tot_b = tot_us = 0.0
    for e in step:
        b = e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
        us = e.get("gpu__time_duration.sum", 0.0)
        tot_b += b
        tot_us += us
        short = re.sub(r"\(.*", "", e["name"])
        print(f"{short:48s} grid {e['grid']:>16s} {us:9.1f} us {b / 1e6:10.1f} MB")
    print(f"launches {len(step)}  serialised {tot_us:.1f} us  dram {tot_b / 1e9:.4f} GB")
    if a.update:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
        with open(p) as fh:
            t = json.load(fh)
        t["workloads"][a.update] = {"dram_bytes_per_step": int(tot_b), "n_gpus": 1, "round": 2, "serialised_us_per_step": round(tot_us, 1),
                                    "launches_per_step": len(step), "source": a.source or os.path.basename(a.csv)}
        with open(p, "w") as fh:
            json.dump(t, fh, indent=1)
            fh.write("\n")


if __name__ == "__main__":
    main()
