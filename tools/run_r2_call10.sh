#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_p2p.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2c10_tests.log
for n in 10000 30000 100000 1000000; do python tools/variant_bench.py $n 2>&1 | grep "base\|closeoff"; done | tee gpurun_out/r2c10_variants.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dram_r02_n1e6.csv python tools/prof_step.py --n 1000000 --steps 1 > gpurun_out/r2c10_ncu_dram.log 2>&1
python - <<'PY'
import csv
rd=wr=t=0.0; n=0
for r in csv.reader(open('gpurun_out/dram_r02_n1e6.csv')):
    if len(r)>14 and ('k_pair_sym' in r[4] or 'k_sym_reduce' in r[4]):
        v=float(r[14].replace(',','')); u=r[13]
        if r[12]=='dram__bytes_read.sum': rd+=v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
        if r[12]=='dram__bytes_write.sum': wr+=v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
        if r[12]=='gpu__time_duration.sum': t+=v*{'ns':1e-9,'us':1e-6,'usecond':1e-6,'ms':1e-3,'msecond':1e-3,'nsecond':1e-9,'second':1}[u]; n+=1
print('dram read GB',rd/1e9,'write GB',wr/1e9,'launches',n,'time s',t)
PY
