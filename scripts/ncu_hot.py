"""Top stall sites of one kernel from an ncu report's source page.
   python scripts/ncu_hot.py <report.ncu-rep> <kernel regex> [which occurrence, default 0] [top N]"""
import csv, io, re, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sel = [b for b in blocks if re.search(rx, b["name"])]
b = sel[which]
hdr, data = b["hdr"], b["data"]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r, h: int(r[ix[h]] or 0)
tot = sum(S(r, "# Samples") for r in data)
print(b["name"][:100], "| occurrences:", len(sel), "| samples", tot, "| sass lines", len(data))
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(S(r, h) for r in data) for h in stall}
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:7])
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:topn]:
    st = {h: S(r, h) for h in stall if S(r, h) > 0}
    print("%5d %9s  %-78s %s" % (S(r, "# Samples"), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:78],
                                sorted(st.items(), key=lambda kv: -kv[1])[:2]))
