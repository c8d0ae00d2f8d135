#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_round2.py::test_config4_stated_size_groups_and_fmc 2>&1 | tail -5) > gpurun_out/s2_pytest.log
cat gpurun_out/s2_pytest.log
RPGO_CLIQUE_TRACE=1 timeout 200 python tools/clique_probe.py 50000 2>&1 | tail -3
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/s2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','k3_ms','max_clique_ms']}, d['e2e']['ms_per_step'], d['e2e']['ms_all_rank0'])
P
