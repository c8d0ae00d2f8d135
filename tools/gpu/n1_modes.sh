#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "simple or other_modes or tiled or config2_synthetic or config3 or scenarios or mirror or g2o" > gpurun_out/n1c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n1c_pytest.log
tail -15 gpurun_out/n1c_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extras modes,planted > gpurun_out/n1c_bench.json 2> gpurun_out/n1c_bench.err
echo "bench rc=$?" | tee -a gpurun_out/n1c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n1c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','k3_ms','allgather_mirror_degree_ms','max_clique_ms']})
for k,v in d['extra'].items(): print(k, json.dumps(v)[:1500])
PY
