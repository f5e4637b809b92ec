"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`, no GPU needed) into a markdown table.
    python tools/ncu_summary.py gpurun_out/r02_*.ncu-rep > profiles/r02_ncu_summary.md"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (any)"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor INT8 pipe %"),
    ("sm__inst_executed_pipe_tensor_op_utcimma.sum", "UTCIMMA inst"),
    ("sm__ops_path_tensor_op_utcimma_src_int8.avg.pct_of_peak_sustained_elapsed", "UTCIMMA int8 ops % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


print("# ncu --set full captures, round 2 (one launch per kernel; cold cache, clocks not locked)\n")
print("Source: `tools/profile_r02.sh` on one B200 (`tools/rns_one.py`: one 8192^3 FP64 product, 16 moduli; "
      "`tools/profile_step.py 16384 2`: one training iteration + prediction, n = 16384, q = 2, d = 21, Matern-5/2).\n")
for path in sys.argv[1:]:
    recs, units = load(path)
    for d in recs:
        name = d.get("Kernel Name", "?").split("(")[0]
        print(f"## `{name}`  ({path.split('/')[-1]})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"| {label} (`{k}`) | {d[k]} | {units.get(k, '')} |")
        # top stall reasons
        stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_")
                         and k.endswith("_per_issue_active.ratio") and v not in ("", "nan")), reverse=True)[:4]
        if stalls:
            print("| top stalls (warps stalled per issue) | " + "; ".join(
                f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls) + " | |")
        print()
