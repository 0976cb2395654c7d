"""Markdown table of the key per-launch metrics from `ncu -i X.ncu-rep --page raw --csv`.  usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm thr %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible/cyc"), ("launch__grid_size", "grid"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sectors_op_atom.sum", "L2 atom sect"), ("lts__t_sectors_op_red.sum", "L2 red sect"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 wavefronts"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld requests"),
        ("launch__occupancy_limit_registers", "occ lim regs"), ("launch__occupancy_limit_shared_mem", "occ lim smem")]
cols = [(hdr.index(k), t) for k, t in want if k in hdr]
print("| " + " | ".join(t for _, t in cols) + " |")
print("|" + "---|" * len(cols))
for r in body:
    out = []
    for i, t in cols:
        v = r[i]
        if t == "kernel":
            v = v.replace("void ", "").split("(NmfScene")[0].split("(const")[0].replace("(int)", "")
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if t != "warp inst" else f"{f:.3e}"
            except ValueError:
                pass
            if units[i] and t in ("time", "dram rd", "dram wr", "L2 bytes"):
                v += " " + units[i].replace("byte", "B")
        out.append(v)
    print("| " + " | ".join(out) + " |")
