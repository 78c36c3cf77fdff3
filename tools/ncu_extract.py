"""tools/ncu_extract.py <rawpage.csv> <out.csv> — selected metrics of one `ncu -i X.ncu-rep --page raw --csv` export in the
`metric,value,unit` form that bench.py (roofline.traffic) and profiles/README.md read; warp-stall sampling as percentages."""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "sm__icc_request_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
out = [("metric", "value", "unit")]
for k in KEEP:
    if k in d:
        out.append((k, d[k][0].replace(",", ""), d[k][1]))
st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v[0].replace(",", "")) for k, v in d.items()
      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v[0] not in ("", "n/a")}
tot = sum(st.values()) or 1.0
for k, v in sorted(st.items(), key=lambda x: -x[1]):
    out.append(("stall_" + k, f"{100 * v / tot:.2f}", "% of warp-stall samples"))
with open(sys.argv[2], "w") as f:
    for r in out:
        f.write(",".join(r) + "\n")
print(open(sys.argv[2]).read())
