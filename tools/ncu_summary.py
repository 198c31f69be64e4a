"""Summarise an .ncu-rep (read here, no GPU): key roofline metrics, stall reasons and the dynamic SASS mix.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [cell_steps_per_launch]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "local_load", "local_store",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local")


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    cell_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")]
        print(f"=== launch {k}: {name}")
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if any(h == x or (x in h and x.startswith("local")) for x in KEEP):
                print(f"  {h} [{u}] = {v}")
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(v), h.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(x for x, _ in stalls) or 1
        print("  stalls: " + ", ".join(f"{n} {x / tot * 100:.1f}%" for x, n in sorted(stalls, reverse=True)[:7]))
    src = run([rep, "--page", "source", "--csv"])
    blocks = src.split('"Kernel Name"')[1:]
    for k, b in enumerate(blocks):
        rows = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
        hdr = rows[1]
        ci, e = hdr.index("Source"), hdr.index("Instructions Executed")
        c, tot = Counter(), 0
        for r in rows[2:]:
            try:
                n = int(r[e])
            except (ValueError, IndexError):
                continue
            toks = r[ci].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            if op == "IMAD" and "MOV" in r[ci]:
                op = "IMAD.MOV"
            c[op] += n
            tot += n
        print(f"=== launch {k}: {tot} warp instructions, static {len(rows) - 2}")
        if cell_steps:
            print("  " + "  ".join(f"{op} {n * 32 / cell_steps:.2f}" for op, n in c.most_common(22)) +
                  f"   | total {tot * 32 / cell_steps:.1f} thread-inst per cell-step")
        else:
            print("  " + "  ".join(f"{op} {n / tot * 100:.1f}%" for op, n in c.most_common(22)))


if __name__ == "__main__":
    main()
