"""Selected metrics per kernel out of `ncu -i report.ncu-rep --page raw --csv` (stdin or file): the summary format of
profiles/*_ncu_full_kernels.txt.  usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_extract.py"""
import csv, sys
KEEP = ["dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
f = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
rows = list(csv.reader(l for l in f if l.startswith('"')))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"== {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for k in KEEP:
        if k in col:
            print(f"   {k} [{units[col[k]]}] = {r[col[k]]}")
