#!/bin/bash
# single-GPU: full gpu test suite (no -x: see every failure), then a bench line
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q -x --deselect tests/test_multigpu.py > gpurun_out/n1b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n1b_pytest.log
tail -30 gpurun_out/n1b_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n1b_bench.json 2> gpurun_out/n1b_bench.err
echo "bench rc=$?" | tee -a gpurun_out/n1b_bench.err
tail -5 gpurun_out/n1b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n1b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','k3_ms','allgather_mirror_degree_ms','max_clique_ms','max_clique_size','gpu_launches']})
print('e2e', d['e2e']['ms_per_step'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:900])
PY
