#!/bin/bash
# the driver's own invocation at N=8 and N=4 (NCCL_DEBUG=INFO to a file, --steps 20 --warmup 5) + the sharded parity check
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_FILE=/tmp/nccl.%p.log
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench N=$N rc=$?" | tee -a gpurun_out/n${N}_bench.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open('gpurun_out/n%s_bench.json'%N).read().strip().splitlines()[-1])
print({k:d[k] for k in ['n_gpus','value','ms_per_step','k3_ms','allgather_mirror_degree_ms','max_clique_ms','max_clique_size','gpu_launches']})
print('roofline', d['roofline']['frac'], d['roofline']['frac_nominal'], 'clocks', d['clocks'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_all_rank0'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:700])
PY
tail -3 gpurun_out/n${N}_bench.err
done
grep -h "NVLS\|Connected all" /tmp/nccl.*.log 2>/dev/null | sed 's/.*NCCL INFO//' | sort | uniq -c | sort -rn | head -4
unset NCCL_DEBUG NCCL_DEBUG_FILE
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29589 tests/mgpu_check.py > gpurun_out/n8_mgpu.log 2>&1
echo "mgpu rc=$?"; grep -v "^$" gpurun_out/n8_mgpu.log | tail -6
