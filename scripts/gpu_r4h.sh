#!/bin/bash
# end-of-session verification on two GPUs: peer-memory fused step tests, sharded pipeline, 2-rank bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pipeline.py -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4h_bench_n2.json 2> gpurun_out/r4h_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r4h_bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'train', d['train_step']['ms_per_step'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-160
