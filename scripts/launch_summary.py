"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals and the last proof's launch sequence."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
def name(r):
    return re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("Fp<FqParams>", "Fq")
def us(r):
    return float(r["Metric Value"].replace(",", "")) / 1e3
ends = [i for i, r in enumerate(rows) if "join_abc" in r["Kernel Name"]]
# one proof = from the launch after the previous proof's last reduce to this proof's last reduce
last_join = ends[-1]
prev_join = ends[-2]
# proof boundary: first msm_digits after prev proof's final reduce_level on stream of join
def proof_slice():
    j = last_join
    # walk forward to the end of H msm (last reduce_level on that stream)
    st = rows[j]["Stream"]
    e = j
    for k in range(j, len(rows)):
        if rows[k]["Stream"] == st and "msm_weighted_kernel" in rows[k]["Kernel Name"]:
            e = k
        if "intpipe" in rows[k]["Kernel Name"]:
            break
    # walk back to the previous proof's end
    pj = prev_join
    pe = pj
    for k in range(pj, j):
        if rows[k]["Stream"] == st and "msm_weighted_kernel" in rows[k]["Kernel Name"]:
            pe = k
            break
    # previous H msm end = last reduce on st before last_join
    pe = max(k for k in range(pj, j) if rows[k]["Stream"] == st and "msm_weighted_kernel" in rows[k]["Kernel Name"])
    return rows[pe + 1:e + 1]
sl = proof_slice()
agg = collections.OrderedDict()
for r in sl:
    a = agg.setdefault(name(r), [0, 0.0])
    a[0] += 1
    a[1] += us(r)
tot = sum(a[1] for a in agg.values())
print("one proof: %d launches, %.1f us serialised" % (len(sl), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("  %-40s n=%3d total=%9.1f us share=%5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
if "-v" in sys.argv:
    for r in sl:
        print(r["ID"], r["Stream"], "%-40s" % name(r), r["Grid Size"], r["Block Size"], "%9.1f us" % us(r))
