import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
keep = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic','sm__cycles_elapsed.avg','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_atom.sum','smsp__inst_executed_op_global_ld.sum']
keep += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('#', r[hdr.index('Kernel Name')][:70])
    for w in keep:
        if w in hdr:
            v = r[hdr.index(w)]
            try:
                if float(v) == 0: continue
            except: pass
            print(f'{w:82s} {v:>18s} {units[hdr.index(w)]}')
