#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tensorcore.py -m gpu -x -q -k "multi or jitter or p2p" > gpurun_out/r2o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2o_bench_n2.err
tail -3 gpurun_out/r2o_bench_n2.err; tail -c 1500 gpurun_out/r2o_bench_n2.json
