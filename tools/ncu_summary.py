"""Condense an .ncu-rep (ncu --set full) into a small committed summary under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__inst_executed_pipe_tensor',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warp_issue_stalled',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'smsp__average_warp', 'sm__sass_inst_executed_op_shared', 'smsp__pcsamp_warps_issue_stalled',
]


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# condensed from %s (ncu --set full --clock-control none); one block per profiled launch\n' % rep)
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write('\n== %s\n' % name[:160])
            for h, u, v in zip(hdr, units, vals):
                if any(k in h for k in KEYS):
                    f.write('%-90s %-10s %s\n' % (h, u, v))
    print('wrote', out)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
