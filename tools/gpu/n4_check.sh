#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu.py tests/test_gpu_round2.py -m gpu -x -q -k "torchrun or cpp or two_devices or mirror" > gpurun_out/n4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n4_pytest.log
tail -6 gpurun_out/n4_pytest.log
export NCCL_DEBUG=INFO NCCL_DEBUG_FILE=/tmp/nccl.%p.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err
echo "bench rc=$?" | tee -a gpurun_out/n4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n4_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['n_gpus','value','ms_per_step','k3_ms','allgather_mirror_degree_ms','max_clique_ms','max_clique_size','gpu_launches']})
print('roofline', d['roofline']['frac'], d['roofline']['frac_nominal'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_all_rank0'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:1200])
PY
tail -5 gpurun_out/n4_bench.err
