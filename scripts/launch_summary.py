"""Per-kernel time of one timestep from an ncu launch list (--metrics gpu__time_duration.sum --csv):
python scripts/launch_summary.py gpurun_out/launches.csv [step_index]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[h + 2:] if len(r) > vi]
idx = [i for i, (n, _) in enumerate(L) if "k_sample" in n]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seg = L[idx[k]:idx[k + 1]]
d = collections.OrderedDict()
for n, v in seg:
    n = n.split("(")[0][-60:]
    d.setdefault(n, [0, 0.0]); d[n][0] += 1; d[n][1] += v
tot = sum(v[1] for v in d.values())
for n, v in sorted(d.items(), key=lambda x: -x[1][1]):
    print("%-62s %3d %9.1f us %5.1f%%  avg %.1f" % (n, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
print("one step (k_sample to k_sample), %d launches: %.1f us" % (len(seg), tot / 1e3))
