#!/bin/bash
mkdir -p gpurun_out
timeout 300 tests/native/slab_selftest bench > gpurun_out/r2c_plain.txt 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:slab_tc_kernel --csv --log-file gpurun_out/r2c_durations.csv tests/native/slab_selftest bench > gpurun_out/r2c_ncu0.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_durations.csv')) if len(r)>5]
h=rows[0]; iv=h.index('Metric Value'); 
vals=[float(r[iv].replace(',','')) for r in rows[1:]]
# 10 cases x (1 + 6*21) launches
per=1+6*21
for c in range(len(vals)//per):
    blk=vals[c*per:(c+1)*per]
    out=[]
    for d in range(6):
        seg=blk[1+d*21+1:1+(d+1)*21]
        out.append(sum(seg)/len(seg)/1000)
    print(c, ' '.join(f'{x:7.1f}' for x in out))
PY
