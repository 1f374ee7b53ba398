#!/usr/bin/env python
"""scripts/r2/configs_table.py -- profiles/r2_configs.md from the committed bench.py lines (profiles/r2_bench_n{1,2,4,8}.json)"""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lines = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")
    if os.path.exists(p):
        for l in open(p):
            if l.startswith("{"):
                lines[n] = json.loads(l)
out = ["# Round 2 -- every BASELINE.json config through `bench.py` (builder-run lines committed next to this file)", "",
       "Payload GB/s aggregate over all ranks (fraction of the roofline: HBM copy peak 6571 GB/s x 1/2 at N=1; 770 GB/s NVLink payload in per GPU at N>1,",
       "see `r2_nvlink_counters.md` for why 686 is the real ceiling of all-to-all pulls). Every entry's last timed batch is verified on the device.", ""]
names = []
for n in sorted(lines):
    for c in lines[n]["configs"]:
        if c["name"] not in names and "roofline" in c:
            names.append(c["name"])
hdr = "| config | " + " | ".join(f"N={n}" for n in sorted(lines)) + " | reference loop (N=1 box, CPU) |"
out += [hdr, "|---|" + "---|" * (len(lines) + 1)]
def cell(d):
    return f"{d['value']:.0f} ({d['roofline']['frac']:.2f})"
row = "| config 2 headline (B=65536 x 4 KiB) | " + " | ".join(cell(lines[n]) for n in sorted(lines)) + f" | {lines[1]['cpu_baseline']['value']:.0f} ({lines[1]['cpu_baseline']['cores']} threads) |" if 1 in lines else ""
out.append(row)
for nm in names:
    cells = []
    ref = ""
    for n in sorted(lines):
        c = next((c for c in lines[n]["configs"] if c["name"] == nm), None)
        cells.append(cell(c) if c else "")
        if c and c.get("reference"):
            ref = f"{c['reference']['value']:.0f}"
    out.append(f"| {nm} | " + " | ".join(cells) + f" | {ref} |")
out += ["", "e2e (host buffers in / out, GB/s aggregate): " + "; ".join(
    f"N={n}: pinned {lines[n]['e2e']['value']:.0f}, pageable {lines[n]['e2e']['pageable_dst_value']:.0f}, plain pinned D2H copy {lines[n]['e2e']['plain_pinned_d2h_copy_value']:.0f}"
    for n in sorted(lines) if lines[n].get("e2e"))]
for n in sorted(lines):
    if lines[n].get("push"):
        out.append(f"N={n} collective push fetch: {lines[n]['push']['value']:.0f} GB/s ({lines[n]['push']['ms_per_step']:.3f} ms/step) vs pull {(lines[n]['pull'] or lines[n])['value']:.0f}")
x = [c for c in lines.get(1, {}).get("configs", []) if c["name"] in ("per_sample_loop", "ingest", "prefetch_overlap")]
for c in x:
    out.append("")
    out.append(f"`{c['name']}`: " + ", ".join(f"{k} = {v:.3g}" if isinstance(v, float) else f"{k} = {v}" for k, v in c.items()
                                             if k not in ("workload", "reference", "name", "unit", "n_gpus") and not k.startswith("ms_") and not k.startswith("host_ms")))
open(os.path.join(ROOT, "profiles", "r2_configs.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
