"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel calls / time / share of the LAST
training step in the capture (a step = the launches between two consecutive fused-AdamW kernels).
usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/...summary.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v
    rows.append((r[ik], us))
# step boundaries: the multi-tensor AdamW kernel closes a step
marks = [i for i, (k, _) in enumerate(rows) if "multi_tensor_apply" in k and "Adam" in k]
ends = [i for n, i in enumerate(marks) if n + 1 == len(marks) or marks[n + 1] - i > 50]
if len(ends) >= 2:
    lo, hi = ends[-2] + 1, ends[-1] + 1
else:
    lo, hi = 0, len(rows)
step = rows[lo:hi]


def short(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"<.*", "", k)
    k = re.sub(r"\(.*", "", k)
    return k[:72]


agg = collections.OrderedDict()
for k, us in step:
    a = agg.setdefault(short(k), [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(us for _, us in step)
print(f"# one training step (launches {lo}..{hi - 1} of the ncu launch list, gpu__time_duration.sum, cold-cache/serialised: compare SHARES)")
print(f"# total {tot:.1f} us over {len(step)} launches")
print(f"{'kernel':74s}{'calls':>6s}{'us':>11s}{'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{k:74s}{n:6d}{us:11.1f}{100 * us / tot:6.1f}%")
vil = sum(us for k, (n, us) in agg.items() if "vil" in k or "simt_" in k or "layernorm" in k)
print(f"# vil:: kernels total {vil:.1f} us = {100 * vil / tot:.1f}% of the step")
