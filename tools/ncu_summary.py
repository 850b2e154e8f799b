"""Summarises an .ncu-rep (first profiled kernel) into a markdown table.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" >> profiles/y.md"""
import csv, io, subprocess, sys
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'sm__inst_executed.avg.per_cycle_elapsed']
txt = (open(sys.argv[1]).read() if sys.argv[1].endswith('.csv') else
       subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
print(f"## {sys.argv[2]}\n\n`{vals[hdr.index('Kernel Name')]}`\n\n| metric | unit | value |\n|---|---|---|")
for i, h in enumerate(hdr):
    if h in want or ('warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')):
        print(f"| {h} | {units[i]} | {vals[i]} |")
print()
