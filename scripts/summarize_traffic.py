#!/usr/bin/env python
"""Join an ncu launch list (--csv, metrics gpu__time_duration.sum / dram__bytes_read.sum / dram__bytes_write.sum, one
forward) with the library's per-launch family tags (IRB_PROFILE_DUMP) -> launch CSV + per-family JSON.
Usage: summarize_traffic.py ncu.csv tags.csv out_launches.csv out_traffic.json"""
import collections, csv, json, sys

ncu_csv, tags_csv, out_csv, out_json = sys.argv[1:5]
rows = [r for r in csv.reader(open(ncu_csv)) if r and not r[0].startswith("==")]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()          # ncu ID -> {name, metrics}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    d = launches.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[ix["Metric Name"]]] = v * scale
tags = [r for r in csv.reader(open(tags_csv)) if r]
# the tagged pass is the LAST len(launches) rows of the dump (the dump file may hold earlier passes)
L = list(launches.values())
tags = tags[-len(L):] if len(tags) >= len(L) else tags
ok = len(tags) == len(L)
fam = collections.OrderedDict()
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["seq", "family", "kernel", "ncu_us", "dram_read_bytes", "dram_write_bytes", "algorithmic_bytes", "event_ms"])
    for i, d in enumerate(L):
        t = tags[i] if ok else ["", "unmatched", "0", "0", "0"]
        us = d.get("gpu__time_duration.sum", 0.0)
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        w.writerow([i, t[1], d["name"][:90], f"{us:.2f}", f"{rd:.0f}", f"{wr:.0f}", t[3], t[2]])
        a = fam.setdefault(t[1], {"launches": 0, "ncu_us": 0.0, "dram_bytes": 0.0, "algorithmic_bytes": 0.0})
        a["launches"] += 1; a["ncu_us"] += us; a["dram_bytes"] += rd + wr; a["algorithmic_bytes"] += float(t[3])
tot = sum(a["ncu_us"] for a in fam.values()) or 1.0
out = {"matched_with_tags": ok, "launches": len(L), "ncu_total_us": tot, "families": {}}
for k, a in fam.items():
    out["families"][k] = {"launches": a["launches"], "ncu_us": round(a["ncu_us"], 1), "share_of_step": round(a["ncu_us"] / tot, 4),
                          "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
                          "algorithmic_bytes_per_launch": a["algorithmic_bytes"] / a["launches"],
                          "dram_over_algorithmic": round(a["dram_bytes"] / a["algorithmic_bytes"], 3) if a["algorithmic_bytes"] else None}
json.dump(out, open(out_json, "w"), indent=1)
print(json.dumps({k: (v["launches"], v["share_of_step"], v["dram_over_algorithmic"]) for k, v in out["families"].items()}))
