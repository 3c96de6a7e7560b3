"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python profiles/launch_summary.py file.csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "ns ": 1e-6}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).replace("sla::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:72]:72s} n={cnt[k]:3d} total={v:9.3f} ms  avg={v / cnt[k] * 1e3:9.1f} us  share={100 * v / total:5.1f}%")
print(f"total (serialised) {total:.3f} ms   ({sum(cnt.values())} launches)")
