#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "clique or mirror or sharded or batch or planted" > gpurun_out/n1d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n1d_pytest.log
tail -5 gpurun_out/n1d_pytest.log
python tools/clique_probe.py 50000
python tools/clique_probe.py 200000
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extras planted,config4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','k3_ms','allgather_mirror_degree_ms','max_clique_ms']})
for k,v in d['extra'].items(): print(k, json.dumps(v)[:1500])"
