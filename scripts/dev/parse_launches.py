"""Print an ncu --csv launch list (gpu__time_duration + dram bytes + tensor-pipe share) one line per launch."""
import collections, csv, sys
for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        per.setdefault((r["ID"], r["Kernel Name"][:44]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    print(f)
    tot = 0.0
    for (i, k), m in per.items():
        t = m.get("gpu__time_duration.sum", 0) / 1000
        tot += t
        print(f"{i:>4} {k:44s} {t:8.1f} us  rd {m.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB  wr {m.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB"
              f"  tc {m.get('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):5.1f} %")
    print(f"total {tot:.1f} us")
