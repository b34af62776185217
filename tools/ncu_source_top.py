"""Summarise an `ncu --page source --csv` export: per kernel, the hottest SASS lines by stall samples and the stall mix.
usage: python tools/ncu_source_top.py file.csv [ntop]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ks, cur, seen = [], None, set()
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:70], "hdr": None, "rows": []}
        if r[1] in seen:
            cur = None
        else:
            seen.add(r[1]); ks.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
for k in ks:
    h = k["hdr"]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp] or 0) for r in k["rows"] if len(r) > isamp) or 1
    mix = {h[i]: sum(int(r[i] or 0) for r in k["rows"] if len(r) > i) for i in stall_cols}
    print("=====", k["name"], "sass rows", len(k["rows"]), "samples", tot)
    print("   stall mix:", ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for n, v in sorted(mix.items(), key=lambda x: -x[1])[:9]))
    for r in sorted(k["rows"], key=lambda r: -int(r[isamp] or 0))[:ntop]:
        st = sorted([(int(r[i] or 0), h[i][6:]) for i in stall_cols], reverse=True)[:2]
        print("%6s %5.1f%% ex=%8s %-72s %s" % (r[isamp], 100 * int(r[isamp] or 0) / tot, r[iex], r[isrc][:72], st))
