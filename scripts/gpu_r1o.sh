#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_o.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lstm.py tests/test_gpu_pipeline.py -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
