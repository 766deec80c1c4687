#!/bin/bash
# end-of-session verification: GPU suite, smoke, bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('roofline frac', d['roofline']['frac'], 'whole', d['roofline']['whole_path']['frac_of_bf16_burst'], d['roofline']['phase_ms_per_step'])
print('train', d['train_step']['ms_per_step'], 'fp32', d['fp32_mode']['value'], 'h256', d['h256_bf16']['value'], 'ode', d['ode']['rk4_full_trajectory']['value'], 'cpu', d['cpu_baseline']['value'])"
timeout 200 python scripts/time_fp32.py 256 2>&1 | tail -2
