#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu_r2s.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2s.log
tail -5 gpurun_out/pytest_gpu_r2s.log
for v in o_old default; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/onevn_check.py 2>&1 | tail -2
done
unset B200_RMSD_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_onevn_r2s.csv python tools/onevn_check.py > gpurun_out/onevn_under_ncu.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_onevn_r2s.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); k=d['Kernel Name'][:70]
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        u=d['Metric Unit']; v = v/1000 if u=='ns' else v*1000 if u=='ms' else v
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print('%-72s %4d avg %8.1f us'%(k,n,t/n))
P
