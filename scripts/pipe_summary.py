"""Whole-proof multiplier-pipe time from an ncu launch list that carries, per launch, gpu__time_duration.sum and
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed: sum(duration x fmaheavy%) over the launches of ONE
proof = the time the IMAD.WIDE pipe is busy per proof; divided by the measured ms/proof of the un-profiled bench it is
the whole-step pipe utilisation.  Usage: pipe_summary.py launches_pipe.csv [ms_per_proof [kernel-name-substring the proof must contain]]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
ms_per_proof = float(sys.argv[2]) if len(sys.argv) > 2 else None
must_contain = sys.argv[3] if len(sys.argv) > 3 else None     # e.g. "msm_pair": pick a throughput-mode proof
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("Fp<FqParams>", "Fq"),
                                    "stream": r["Stream"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ls = list(launch.values())
starts = [i for i, l in enumerate(ls) if "r1cs_eval" in l["name"]]
# one proof = the launches from one r1cs evaluation up to the next (the last complete interval of the capture)
proof = ls[starts[-2]:starts[-1]]
if must_contain:
    for a, b in reversed(list(zip(starts, starts[1:]))):
        if any(must_contain in l["name"] for l in ls[a:b]):
            proof = ls[a:b]
            break
agg = collections.OrderedDict()
for l in proof:
    t = l["gpu__time_duration.sum"] / 1e3
    pct = l.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    a = agg.setdefault(l["name"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] += t * pct / 100.0
tot_t = sum(a[1] for a in agg.values())
tot_p = sum(a[2] for a in agg.values())
print("one proof: %d launches, %.1f us serialised kernel time, %.1f us of multiplier-pipe time (fmaheavy-active)" % (len(proof), tot_t, tot_p))
for k, a in sorted(agg.items(), key=lambda x: -x[1][2]):
    print("  %-40s n=%3d time=%8.1f us  pipe=%8.1f us  (%4.1f%% active, %4.1f%% of pipe time)" %
          (k, a[0], a[1], a[2], 100 * a[2] / a[1] if a[1] else 0, 100 * a[2] / tot_p))
if ms_per_proof:
    print("whole-step multiplier-pipe utilisation at %.3f ms/proof (un-profiled bench): %.1f %%" % (ms_per_proof, 100 * tot_p / 1e3 / ms_per_proof))
