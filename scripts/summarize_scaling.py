#!/usr/bin/env python
"""Collect bench JSON lines from gpurun_out/ into profiles/scaling.md."""
import glob
import json
import os
import sys

rows = []
for path in sorted(glob.glob("gpurun_out/b1[3-9]_*.json") + glob.glob("gpurun_out/b6_default.json")):
    try:
        line = [l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception:
        continue
    c = d.get("config", {})
    rows.append((d["n_gpus"], d.get("impl"), bool(c.get("item_cache")), c.get("sync_every"), d["value"] / 1e9,
                 d["ms_per_step"], d["e2e"]["value"] / 1e9, ",".join(d["clocks"]["reasons"]) or "-",
                 os.path.basename(path)))
rows.sort(key=lambda r: (r[1] != "fps_b200", r[0], not r[2], r[3] or 0))
base = next((r[4] for r in rows if r[0] == 1 and r[1] == "fps_b200" and not r[2]), None)
out = ["| N | impl | item cache | sync every | G updates/s (device) | ms/step | G updates/s (e2e) | weak-scaling eff. vs N=1 | clocks | file |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    eff = f"{r[4] / (base * r[0]):.2f}" if base and r[1] == "fps_b200" else "-"
    out.append(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]:.2f} | {r[5]:.3f} | {r[6]:.2f} | {eff} | {r[7]} | {r[8]} |")
print("\n".join(out))
