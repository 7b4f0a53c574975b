"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output:
prints the hottest CUDA-C lines over all source files of the (first) kernel."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = {}
fname = "?"
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0]); samp = int(r[4]); inst = int(r[7])
    except ValueError:
        continue
    stalls = {hdr[i]: int(r[i]) for i in range(31, 48) if r[i].isdigit() and int(r[i]) > 0}
    a = agg.setdefault((fname, ln), [r[1], 0, 0, {}])
    a[1] += samp; a[2] += inst
    for k, v in stalls.items():
        a[3][k] = a[3].get(k, 0) + v
tot = sum(a[1] for a in agg.values()) or 1
print("total samples", tot)
for (fn, ln), (src, samp, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{fn[:14]:14s}{ln:5d} {100*samp/tot:5.1f}% inst={inst:9d} {src.strip()[:64]:64s} {s}")
