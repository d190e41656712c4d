#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"cov_tc|k_center|k_chan_sums|k_cov_finalize|k_rayleigh|k_eig_post|k_outer|k_finalize|k_mean|k_maxpool|k_upsample" --csv --log-file gpurun_out/wct_launches.csv python tools/profile_step.py 16 > gpurun_out/ncu_wct.log 2>&1
echo rc=$?
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(l for l in open('gpurun_out/wct_launches.csv') if l.startswith('"'))]
h=rows[0]; rows=rows[1:]
ik,im,iv,ig=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("Grid Size")
iid=h.index("ID")
d={}
for r in rows:
    d.setdefault(r[iid],{'k':re.sub(r"\(.*","",r[ik]).strip(),'g':r[ig]})[r[im]]=float(r[iv].replace(',',''))
agg={}
for v in d.values():
    key=(v['k'],v['g'])
    a=agg.setdefault(key,[0,0.0,0.0,0.0]); a[0]+=1; a[1]+=v.get('gpu__time_duration.sum',0); a[2]+=v.get('dram__bytes_read.sum',0); a[3]+=v.get('dram__bytes_write.sum',0)
for (k,g),a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-28s grid %-18s n=%3d  total %9.1f us  avg %8.1f us  rd %8.1f MB wr %8.1f MB per launch" % (k[:28], g, a[0], a[1]/1e3, a[1]/a[0]/1e3, a[2]/a[0]/1e6 if a[2]>1e3 else a[2]/a[0], a[3]/a[0]/1e6 if a[3]>1e3 else a[3]/a[0]))
PY
