"""Text summary of an ncu report (the metrics the profiles/*_summary.txt files quote):

    python tools/ncu_summary.py report.ncu-rep ["header line"] > profiles/rNN_..._summary.txt

Runs `ncu -i report --page raw --csv` and prints, per captured kernel, launch shape, duration, DRAM bytes, instruction
count, issue / pipe utilisation, shared-memory wavefronts and bank conflicts, and the stall reasons per issued instruction.
"""
import csv
import io
import subprocess
import sys

KEYS = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print('# ' + sys.argv[2])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('----')
        rec = dict(zip(hdr, zip(vals, units)))
        print('Kernel Name =', rec['Kernel Name'][0])
        for k in KEYS:
            if k in rec:
                print('%s = %s %s' % (k, rec[k][0], rec[k][1]))
        for k in hdr:
            if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
                print('%s = %s %s' % (k, rec[k][0], rec[k][1]))


if __name__ == '__main__':
    main()
