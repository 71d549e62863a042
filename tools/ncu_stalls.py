"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
isrc = h.index("Source"); isamp = h.index("# Samples"); iex = h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
top = sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:top_n]
for idx, r in sorted(top):
    st = sorted(((int(r[i]), h[i]) for i in stall_cols), reverse=True)[:2]
    print(idx, r[isamp], r[iex], r[isrc].strip()[:64], st)
