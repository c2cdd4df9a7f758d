"""Summarise an ncu report: python tools/ncu_summary.py report.ncu-rep [out.csv] — one row per kernel launch with the
counters the DESIGN/profiles tables quote (run where ncu is installed; the .ncu-rep comes back from the GPU box)."""
import csv, subprocess, sys, io
WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_selected"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}   # -> us, MB
out = [[n + ("_MB" if n.startswith("dram") else "") for _, n in WANT]]
for r in rows[2:]:
    line = []
    for k, n in WANT:
        v = r[hdr.index(k)] if k in hdr else ""
        if n == "kernel":
            v = v.replace("<unnamed>::", "").replace("void ", "")[:48]
        else:
            try:
                x = float(v)
                if n in ("us", "dram_rd", "dram_wr"):
                    x *= SCALE.get(units[hdr.index(k)], 1.0)
                v = f"{x:.4g}"
            except ValueError:
                pass
        line.append(v)
    out.append(line)
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
