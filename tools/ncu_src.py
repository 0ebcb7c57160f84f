"""Summarise `ncu --page source --csv` output: per kernel section, stall totals and hottest SASS lines."""
import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = secs[which]
h = sec["rows"][0]; ci = {n: i for i, n in enumerate(h)}
data = [r for r in sec["rows"][1:] if len(r) == len(h)]
S = ci["# Samples"]
tot = sum(int(r[S]) for r in data)
print(len(secs), "sections; using", which, sec["name"][:80], "total samples", tot)
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]]) for r in data) for n in stall_cols}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -int(r[S]))[:ntop]:
    st = sorted(((n, int(r[ci[n]])) for n in stall_cols if int(r[ci[n]]) > 0), key=lambda x: -x[1])[:2]
    print(f"{int(r[S]):6d} {100*int(r[S])/max(tot,1):5.1f}% exec={r[ci['Instructions Executed']]:>9} {r[ci['Source']].strip()[:66]:66s} {st}")
