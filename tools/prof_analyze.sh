#!/bin/bash
# after gpu_prof_mesh.sh <tag>: per-line / per-opcode stall summary of raster_kernel<0,2> (run here, not on the GPU box)
TAG=$1
cd /root/repo/gpurun_out/$TAG
cuobjdump -xelf raster.sm_100a.cubin ../../kazan_b200/libkazan_b200.so >/dev/null 2>&1
nvdisasm -g raster.sm_100a.cubin > raster.disasm 2>/dev/null
ncu -i prof_mesh.ncu-rep --page source --csv --kernel-name regex:raster_kernel 2>/dev/null > src_raster.csv
ncu -i prof_mesh.ncu-rep --page raw --csv 2>/dev/null > raw.csv
python - <<'P'
import csv
rows=list(csv.reader(open('raw.csv'))); hdr=rows[0]
for w in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem']:
    if w in hdr: i=hdr.index(w); print(w,[r[i] for r in rows[2:]][-1])
for i,h in enumerate(hdr):
    if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
        v=float([r[i] for r in rows[2:]][-1])
        if v>0.3: print('  stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2))
P
python /root/repo/tools/sass_regions.py raster.disasm src_raster.csv raster_kernelILj0ELi2E ${3:-16200} /root/repo/kazan_b200/csrc/raster.cu 2>&1 | cut -c1-170 | head -${2:-40}
