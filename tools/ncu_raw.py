"""Prints key metrics + top stall reasons for every kernel in an .ncu-rep (via ncu --page raw --csv)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
col = {h: i for i, h in enumerate(H)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("====", r[col["Kernel Name"]][:70], "grid", r[col.get("Grid Size", 0)] if "Grid Size" in col else "")
    for w in want:
        for h in H:
            if h.startswith(w) and h in col and r[col[h]] != "":
                print(f"   {h} = {r[col[h]]} {rows[1][col[h]]}")
                break
    st = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, H[i]) for i in range(len(H))
                 if "smsp__average_warps_issue_stalled" in H[i] and H[i].endswith("_per_issue_active.ratio")), reverse=True)[:6]
    for v, h in st:
        print(f"      stall {v:8.2f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
