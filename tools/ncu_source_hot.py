"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output
(first kernel section that has CUDA-C line rows): prints the hottest lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# find first header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
agg = {}
for r in rows[hi + 1:]:
    if not r or r[0] in ("Line No", "File Path", "Function Name", "File Name"):
        if r and r[0] == "Line No":
            break
        continue
    if len(r) < len(hdr) or r[2] != "-":     # only CUDA-C lines (address column "-")
        continue
    try:
        ln = int(r[0]); samp = int(r[4]); inst = int(r[7])
    except ValueError:
        continue
    stalls = {hdr[i]: int(r[i]) for i in range(31, 48) if r[i].isdigit() and int(r[i]) > 0}
    a = agg.setdefault(ln, [r[1], 0, 0, {}])
    a[1] += samp; a[2] += inst
    for k, v in stalls.items():
        a[3][k] = a[3].get(k, 0) + v
tot = sum(a[1] for a in agg.values()) or 1
print("total samples", tot)
for ln, (src, samp, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{ln:5d} {100*samp/tot:5.1f}% inst={inst:9d} {src.strip()[:70]:70s} {s}")
