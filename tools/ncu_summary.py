"""Summarise an .ncu-rep (read here, on the GPU-less box) into a small JSON for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.json"""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for w in WANT:
            if w in hdr:
                v = r[hdr.index(w)]
                try:
                    v = float(v)
                except ValueError:
                    pass
                d[w] = {'value': v, 'unit': units[hdr.index(w)]}
        res.append(d)
    json.dump({'source': rep, 'note': 'ncu --set full --clock-control none; per-launch values (cold-cache, serialised)',
               'launches': res}, open(out, 'w'), indent=1)
    print('wrote', out, len(res), 'launches')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
