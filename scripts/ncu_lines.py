"""Per-source-line summary of an ncu report (stall samples and executed warp instructions).

usage: python scripts/ncu_lines.py gpurun_out/<name>.ncu-rep [kernel-regex] [top-n]
Reads `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo at compile time and
--import-source on at capture time) and prints the hottest source lines per kernel.
"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, func, hdr = None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue  # SASS rows carry an address; aggregated source rows carry "-"
        if pat and not pat.search(func or ""):
            continue
        d = dict(zip(hdr, r))
        key = (func, fname, int(r[0]), r[1].strip()[:90])
        samples = int(d.get("# Samples", "0") or 0)
        inst = int(d.get("Instructions Executed", "0") or 0)
        stalls = {k: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += samples
        a[1] += inst
        for k, v in stalls.items():
            a[2][k] = a[2].get(k, 0) + v
    by_func = {}
    for (func, fname, line, src), v in agg.items():
        by_func.setdefault(func, []).append((v[0], v[1], fname, line, src, v[2]))
    for func, items in by_func.items():
        ts, ti = sum(i[0] for i in items), sum(i[1] for i in items)
        print(f"== {func[:110]}\n   samples {ts}  warp-instructions {ti}")
        for s, i, fname, line, src, st in sorted(items, reverse=True)[:top]:
            top_st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
            print(f"  {100 * s / max(ts, 1):5.1f}% smp {100 * i / max(ti, 1):5.1f}% inst  {fname}:{line:<4d} {src}\n"
                  f"            [{top_st}]")


if __name__ == "__main__":
    main()
