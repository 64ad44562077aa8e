"""Print the handful of ncu metrics that matter for these kernels from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit', 'sm__cycles_elapsed.max', 'sm__cycles_elapsed.avg.per_second',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed_op_local', 'smsp__inst_executed_op_shared',
        'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index('Kernel Name')
    for r in data:
        print('==', r[name_col][:90])
        for i, h in enumerate(hdr):
            if any(h == w or h.startswith(w + '.') or h.startswith(w) and w.endswith('stalled') for w in WANT):
                if 'pct_of_peak' in h and not h.endswith('elapsed') and not h.endswith('active'):
                    continue
                print(f'   {h:85s} {r[i]:>18s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1])
