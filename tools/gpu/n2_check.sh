#!/bin/bash
# multi-GPU check at N=2: pytest multi-GPU file, then the driver's own bench invocation
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/n2_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/n2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n2_pytest.log
tail -5 gpurun_out/n2_pytest.log
export NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl.%p.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "bench rc=$?" | tee -a gpurun_out/n2_bench.err
tail -c 3000 gpurun_out/n2_bench.json
tail -20 gpurun_out/n2_bench.err
rm -f gpurun_out/nccl.*.log.keep; ls gpurun_out/nccl.*.log 2>/dev/null | head -3
grep -h "NVLS\|via P2P\|Connected all" gpurun_out/nccl.*.log 2>/dev/null | head -5
rm -f gpurun_out/nccl.*.log
