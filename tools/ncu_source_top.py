"""Top SASS instructions of an `ncu --page source --csv` export by executed count and by stall samples."""
import csv
import sys


def main(path, topn=25):
    rows = list(csv.reader(open(path)))
    # one section per profiled launch: a "Kernel Name" row, a header row, then the SASS lines; summarise the first launch of each kernel
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    seen = set()
    for si, st in enumerate(starts):
        name = rows[st][1].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        end = starts[si + 1] if si + 1 < len(starts) else len(rows)
        print(f"===== {name}")
        section(rows[st:end], topn)


def section(rows, topn):
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    fl = lambda r, k: float(r[ci[k]] or 0)
    tot_inst = sum(fl(r, "Instructions Executed") for r in data)
    tot_samp = sum(fl(r, "# Samples") for r in data)
    print(f"total warp-instructions {tot_inst:.3e}, samples {tot_samp:.0f}, SASS lines {len(data)}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(fl(r, h) for r in data) for h in stall_cols}
    print("stall samples:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    # opcode histogram
    ops = {}
    for r in data:
        op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[ci["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + fl(r, "Instructions Executed")
    print("by opcode:", ", ".join(f"{k}={v / tot_inst * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
    print("--- top by samples")
    for r in sorted(data, key=lambda r: -fl(r, "# Samples"))[:topn]:
        top = sorted(((h[6:], fl(r, h)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"{fl(r, '# Samples'):7.0f} {fl(r, 'Instructions Executed'):12.0f}  {r[ci['Source']][:90]:90s} {top}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
