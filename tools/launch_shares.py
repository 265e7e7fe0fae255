"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows:
        if r is hdr or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r[ci["Kernel Name"]]).strip()
        unit = r[ci["Metric Unit"]]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        d = agg.setdefault(name, [0.0, 0])
        d[0] += ms
        d[1] += 1
    tot = sum(v[0] for v in agg.values())
    n = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {sys.argv[2] if len(sys.argv) > 2 else 'one forward pass of the hot path at BASELINE configs[1] (64 x 4 s)'}")
    print(f"# total {tot:.2f} ms over {n} launches (cold-cache, serialised: compare SHARES)")
    for name, (ms, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:9.3f} ms  {ms / tot * 100:5.1f}%  x{k:<3d} {name}")


if __name__ == "__main__":
    main(sys.argv[1])
