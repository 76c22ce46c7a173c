#!/bin/bash
mkdir -p gpurun_out
for pair in 1 0; do for bn in 128 176 256; do
  ASN_PAIR=$pair ASN_GEMM_BN=$bn timeout 120 python tools/gemm_probe.py > gpurun_out/gemm_probe_pair${pair}_bn${bn}.json 2> gpurun_out/gemm_probe.err || tail -5 gpurun_out/gemm_probe.err
done; done
python - <<'PY'
import json,glob
rows={}
for f in sorted(glob.glob('gpurun_out/gemm_probe_pair*_bn*.json')):
    d=json.load(open(f)); tag=f.split('gemm_probe_')[1][:-5]
    for k,v in d.items():
        rows.setdefault(k,{})[tag]=v['asn_tflops']; rows[k]['cublas']=v['cublas_tflops']
tags=sorted({t for r in rows.values() for t in r})
print('shape', *tags)
for k,r in rows.items(): print(k, *[r.get(t,'-') for t in tags])
PY
