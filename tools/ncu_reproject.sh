#!/bin/bash
# per-kernel durations + DRAM bytes of the fused re-projection kernels under ncu (clean caches, clocks untouched)
# usage: tools/ncu_reproject.sh <tag> [env assignments...]
tag=$1; shift
mkdir -p gpurun_out
env "$@" ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  -k regex:'reproject_(loss|rows)' -c 26 --csv --log-file gpurun_out/ncu_reproject_$tag.csv python tools/bench_reproject.py 64 $SMOOTH > gpurun_out/ncu_reproject_$tag.log 2>&1
python - "$tag" <<'PY'
import csv,sys,collections
tag=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/ncu_reproject_%s.csv'%tag)))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; kn,mn,mv,idc=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value'),hdr.index('ID')
d=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    d.setdefault((r[idc],r[kn][:60]),{})[r[mn]]=r[mv]
seen=set()
for (i,k),m in d.items():
    if k in seen: continue
    seen.add(k)
    print(tag,k,{a.split('.')[0].replace('__','_'):b for a,b in m.items()})
PY
