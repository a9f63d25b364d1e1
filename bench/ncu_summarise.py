import csv,sys,json
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sector_hit_rate.pct']
res=[]
for r in rows[2:]:
    d={}
    for w in want:
        if w in hdr:
            i=hdr.index(w); d[w]=(r[i]+' '+units[i]).strip()
    res.append(d)
print(json.dumps(res,indent=1))
