"""Per-kernel shares of ONE step from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

    python profiles/tools/launch_summary.py launches.csv --marker fbank_kernel [--segment -1]

A step = the launches from one `--marker` kernel up to the next; --segment picks which one (default: the last complete
one).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event tables."""
import argparse
import csv
import re
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("mm::", "")[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--marker", default="fbank_kernel")
    ap.add_argument("--segment", type=int, default=-1)
    a = ap.parse_args()
    rows, hdr = [], None
    for r in csv.reader(open(a.csv, errors="ignore")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                unit = d.get("Metric Unit", "ns")
                us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
                rows.append((d["Kernel Name"], us))
    marks = [i for i, (n, _) in enumerate(rows) if a.marker in n]
    if len(marks) < 2:
        raise SystemExit(f"{len(rows)} launches, {len(marks)} markers: need two")
    seg = a.segment if a.segment >= 0 else (len(marks) - 1) + a.segment
    lo, hi = marks[seg], marks[seg + 1]
    step = rows[lo:hi]
    total = sum(us for _, us in step)
    agg = OrderedDict()
    for n, us in step:
        k = short(n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    print(f"{len(rows)} launches in the list, {len(marks)} steps (marker {a.marker}); step {seg}: {len(step)} launches, "
          f"{total:.1f} us summed")
    print(f"{'kernel':<72}{'launches':>9}{'us':>11}{'share':>8}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<72}{c:>9}{t:>11.1f}{100 * t / total:>7.1f}%")


if __name__ == "__main__":
    main()
