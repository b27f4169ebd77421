"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--kernel-name ... --launch-skip N --launch-count 1]`
into the hottest source lines: warp-stall samples per line (rows with a line number and no SASS address are the per-line
aggregates) with the dominant stall reasons.  usage: ncu ... | python profiles/source_hotspots.py [top_n]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
cur, hdr, out, func = None, None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1], None
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":          # per-line aggregate row (no SASS address)
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            samples = int(d["# Samples"] or 0)
        except ValueError:
            samples = 0
        if samples:
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
            out.append((samples, cur.split("/")[-1], int(d["Line No"]), r[1].strip()[:100], d.get("Instructions Executed", ""), top))
total = sum(o[0] for o in out) or 1
print(f"kernel: {func}\ntotal warp-stall samples: {total}")
for s, f, ln, src, ie, top in sorted(out, reverse=True)[:top_n]:
    print(f"{s:6d} {100.0 * s / total:5.1f}%  {f}:{ln:<4d} inst={ie:<8s} {', '.join(f'{k} {v}' for k, v in top):40s} | {src}")
