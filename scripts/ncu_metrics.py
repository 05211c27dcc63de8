"""Print selected metrics of every kernel in an .ncu-rep (prefix match on metric names)."""
import csv
import subprocess
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput',
           'l1tex__throughput', 'sm__throughput.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__occupancy_limit', 'launch__registers_per_thread', 'smsp__issue_active.avg.pct',
           'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts', 'lts__t_sector_hit_rate',
           'l1tex__t_sector_hit_rate', 'launch__waves', 'smsp__average_warps_issue_stalled',
           'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
           'sm__inst_executed_pipe', 'launch__grid_size', 'launch__block_size', 'smsp__average_warp_latency']


def main():
    rep = sys.argv[1]
    keys = sys.argv[2:] or DEFAULT
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print('kernels:', [r[ki][:50] for r in rows[2:]])
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in keys):
            print(f'{h} [{units[i]}]', [r[i] for r in rows[2:]])


if __name__ == '__main__':
    main()
