"""Markdown table (share of summed kernel time, launches, average duration) from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file <csv> ...`).  usage: python scripts/ncu_launch_table.py <csv> [top-n]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mi:
            continue
        name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[mi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| share | launches | avg | kernel |\n|---:|---:|---:|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {100 * t / tot:.1f} % | {c} | {t / c:.1f} us | `{n[:100]}` |")


if __name__ == "__main__":
    main()
