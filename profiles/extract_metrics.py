"""Reduce `ncu --page raw --csv` exports (gpurun_out/prof_*_<tag>*.raw.csv) to the handful of metrics the rooflines use:
duration, DRAM bytes read / written, DRAM and tensor-pipe utilisation, L2->SM (lts) read bytes, registers.  One line per
captured launch.  usage: extract_metrics.py <dir> <tag>"""
import csv
import glob
import os
import sys

WANT = {
    'gpu__time_duration.sum': 'dur',
    'dram__bytes_read.sum': 'dram_rd',
    'dram__bytes_write.sum': 'dram_wr',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
    'lts__t_sectors_srcunit_tex_op_read.sum': 'l2_read_sectors',
    'lts__t_sectors_srcunit_tex_op_write.sum': 'l2_write_sectors',
    'launch__registers_per_thread': 'regs',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occupancy_pct',
    'smsp__cycles_active.avg': 'cycles',
}


def main(d, tag):
    for path in sorted(glob.glob(os.path.join(d, f'prof_*{tag}*.raw.csv'))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            print(os.path.basename(path), ': empty')
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print('==', os.path.basename(path))
        for r in rows[2:]:
            name = r[idx['Kernel Name']].split('(')[0][-60:] if 'Kernel Name' in idx else '?'
            parts = []
            for m, short in WANT.items():
                if m in idx and r[idx[m]]:
                    parts.append(f'{short}={r[idx[m]]}{units[idx[m]] if units[idx[m]] not in ("", "%") else ("%" if units[idx[m]] == "%" else "")}')
            print(f'  {name:60s} ' + ' '.join(parts))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
