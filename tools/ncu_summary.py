"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of metrics the roofline uses."""
import csv, subprocess, sys, json
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.max"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"report": rep, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            d[h] = f"{v} {u}".strip()
    print(json.dumps(d, indent=1))
