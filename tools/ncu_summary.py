"""Compact per-launch summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/r01_xxx.csv]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_pct"),
    ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem_bank_rd_pct"),
    ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem_bank_wr_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__waves_per_multiprocessor", "waves"), ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math_throttle"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleeping"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_no_inst"),
]


def main():
    rep = sys.argv[1]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    table = [[short for _, short in METRICS]]
    for r in rows[2:]:
        line = []
        for name, short in METRICS:
            v = r[idx[name]] if name in idx else ""
            if short == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("vrcoc::", "")[:48]
            elif short.endswith("_MB") and v:
                u = units[idx[name]]
                f = float(v)
                v = f"{f / 1e6 if u == 'byte' else f * (1e-3 if u == 'Kbyte' else 1.0 if u == 'Mbyte' else 1e3):.2f}"
            elif short == "time_us" and v:
                u = units[idx[name]]
                f = float(v)
                v = f"{f * (1e-3 if u in ('ns', 'nsecond') else 1.0 if u in ('us', 'usecond') else 1e3):.1f}"
            else:
                try:
                    v = f"{float(v):.2f}" if "." in v else v
                except ValueError:
                    pass
            line.append(v)
        table.append(line)
    w = csv.writer(open(out, "w", newline="") if out else sys.stdout)
    w.writerows(table)
    if out:
        for line in table[1:]:
            print(" ".join(f"{k}={v}" for k, v in zip(table[0], line)))


if __name__ == "__main__":
    main()
