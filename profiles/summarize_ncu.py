"""Key counters of `ncu --set full` reports as a markdown table (one column per captured launch).

    python profiles/summarize_ncu.py gpurun_out/r02f_*.ncu-rep > profiles/r02_ncu_c3.md
"""
import csv
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('smsp__cycles_active.avg', 'SMSP cycles active'), ('sm__cycles_elapsed.max', 'cycles elapsed'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %'),
    ('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'DMMA pipe %'),
    ('sm__pipe_tensor_op_gmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor (UTCMMA) pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
]
STALLS = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'


def main():
    print('# ncu --set full --clock-control none: key counters per captured launch\n')
    print('(cold caches: ncu flushes L2 between replays, so DRAM bytes of kernels whose operands live in L2 inside the '
          'real pass are upper bounds; durations are serialised launches, not the in-graph times bench.py reports)\n')
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print('## `%s`  (%s)\n' % (r[hdr.index('Kernel Name')][:70], path.split('/')[-1]))
            print('| counter | value |\n|---|---|')
            for key, label in WANT:
                if key in hdr and r[hdr.index(key)] not in ('', 'n/a'):
                    print('| %s | %s %s |' % (label, r[hdr.index(key)], units[hdr.index(key)]))
            st = []
            for i, h in enumerate(hdr):
                if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                    try:
                        st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print('| top stalls (warps per issue) | %s |' % ', '.join('%s %.2f' % (n, v) for v, n in st[:5]))
            print()


if __name__ == '__main__':
    main()
