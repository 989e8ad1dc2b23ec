#!/usr/bin/env python
"""Summaries of ncu captures for profiles/ (run where `ncu` is on PATH; reading a report needs no GPU).

  python tools/ncu_summary.py launches <launches.csv> <out.json> [skip_launches]
      per-kernel-name totals of a whole-step metric pass (`ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,...`):
      launches, time, share of the step, DRAM bytes, time-weighted tensor-pipe / DRAM / SM utilisation.
  python tools/ncu_summary.py full <out.json> <a.ncu-rep> [<b.ncu-rep> ...]
      one row per captured launch of `ncu --set full` reports (duration, DRAM bytes, DRAM % / SM % / tensor-pipe %, issue
      slots, warps, registers, L2 hit rate) in the format bench.py reads for roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("fmmt::<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")
    return name.split("(")[0].replace("(int)", "")


def fnum(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def launches(path, out, skip=0):
    # ncu --csv log: some banner lines ("==PROF=="), then a header row starting with "ID"
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = OrderedDict()      # launch id -> {metric: value, "name": ...}
    for r in rows:
        lid = int(r["ID"])
        d = per.setdefault(lid, {"name": short(r["Kernel Name"])})
        v = fnum(r["Metric Value"])
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if v is None:
            continue
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    ids = sorted(per)[skip:]
    agg = OrderedDict()
    for lid in ids:
        d = per[lid]
        a = agg.setdefault(d["name"], {"kernel": d["name"], "launches": 0, "us": 0.0, "dram_bytes": 0.0, "_w": {}})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["us"] += t
        a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        for m, v in d.items():
            if "pct_of_peak" in m:
                a["_w"][m] = a["_w"].get(m, 0.0) + v * t
    total = sum(a["us"] for a in agg.values())
    ks = []
    for a in sorted(agg.values(), key=lambda a: -a["us"]):
        row = {"kernel": a["kernel"], "launches": a["launches"], "us": round(a["us"], 1), "share": round(a["us"] / total, 4),
               "dram_MB_per_launch": round(a["dram_bytes"] / a["launches"] / 1e6, 2),
               "dram_GBps": round(a["dram_bytes"] / a["us"] / 1e3, 1) if a["us"] else None}
        for m, w in a["_w"].items():
            key = {"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
                   "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
                   "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct"}.get(m, m)
            row[key] = round(w / a["us"], 2) if a["us"] else None
        ks.append(row)
    json.dump({"source": f"ncu metric pass over one step of tools/profile_step.py (--clock-control none; cold-cache, serialised "
                         f"launches: compare SHARES with the CUDA-event profile, not absolute times); first {skip} launches "
                         f"(warm-up step) skipped",
               "launches": len(ids), "total_us": round(total, 1), "kernels": ks}, open(out, "w"), indent=1)
    print(f"{len(ids)} launches, {total / 1e3:.2f} ms, {len(ks)} kernels -> {out}")


WANT = OrderedDict([
    ("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
])


def full(out, reps):
    res = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            print(f"{rep}: empty", file=sys.stderr)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"kernel": short(r[hdr.index("Kernel Name")]), "report": rep.split("/")[-1]}
            for m, key in WANT.items():
                if m not in hdr:
                    continue
                i = hdr.index(m)
                v = fnum(r[i])
                if v is None:
                    continue
                u = units[i]
                if key == "dur_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if key in ("dram_rd_MB", "dram_wr_MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                d[key] = round(v, 3)
            res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{len(res)} launches -> {out}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3:])
