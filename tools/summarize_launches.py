"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the last N launches (N = the bench line's
gpu_launches, i.e. one timed step) grouped by kernel.  Durations are cold-cache and serialised: compare SHARES, not absolutes.
usage: python tools/summarize_launches.py launches.csv[.gz] N [label]"""
import collections
import csv
import gzip
import io
import re
import sys


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    label = sys.argv[3] if len(sys.argv) > 3 else path
    raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
    start = raw.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(raw[start:])))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    step = rows[-n:]
    agg, cnt = collections.Counter(), collections.Counter()
    for r in step:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("vb::", "").replace("g2::", "").strip()
        unit, val = r["Metric Unit"], float(r["Metric Value"].replace(",", ""))
        ms = val * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[unit]
        agg[name] += ms; cnt[name] += 1
    tot = sum(agg.values())
    print(f"ncu launch list, {label}: last {len(step)} of {len(rows)} profiled launches (one timed step), "
          f"sum of serialised cold-cache durations {tot:.1f} ms")
    for k, v in agg.most_common():
        print(f"{v:10.3f} ms {100 * v / tot:5.1f}%  n={cnt[k]:5d}  {k}")


if __name__ == "__main__":
    main()
