"""Markdown summary of an ncu report: key raw metrics per captured launch, stall mix, hottest source lines.

usage: python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep [top-n] >> profiles/<round>_<what>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible / scheduler"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe CYCLES ACTIVE % (of elapsed)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor pipe instructions (HMMA/UTCHMMA) % of issue peak"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe instructions %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe: tensor-core operand fetch %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe: LSU loads/stores %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "DRAM active %"),
]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit_of = dict(zip(hdr, units))
    print(f"### `{rep.split('/')[-1]}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"* **{d['Kernel Name'][:110]}**")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"  * {label}: {d[k]} {unit_of.get(k, '')}")
        stalls = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(d[k].replace(",", "")), k.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        print("  * stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))
    src = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, ".", str(top)],
                         capture_output=True, text=True).stdout
    print("\n```\n" + "\n".join(l[:170] for l in src.splitlines()) + "\n```\n")


if __name__ == "__main__":
    main()
