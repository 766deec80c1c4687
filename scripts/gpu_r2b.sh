#!/bin/bash
# round 2: topology diagnostics + 2-GPU bench line (multi-rank checks) + the 2-GPU test
mkdir -p gpurun_out
{
  echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== nodes"; ls /sys/devices/system/node/ 2>&1 | head; cat /sys/devices/system/node/node*/cpulist 2>&1
  echo "== topo"; nvidia-smi topo -m 2>&1
  echo "== pci numa"; for d in /sys/bus/pci/devices/*; do if [ -f $d/vendor ] && [ "$(cat $d/vendor)" = "0x10de" ]; then echo "$d class=$(cat $d/class) numa=$(cat $d/numa_node 2>/dev/null) cpus=$(cat $d/local_cpulist 2>/dev/null)"; fi; done
  echo "== affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])"
  echo "== nvml"; python - <<'PY'
import pynvml as nv
nv.nvmlInit()
for i in range(nv.nvmlDeviceGetCount()):
    h = nv.nvmlDeviceGetHandleByIndex(i)
    pci = nv.nvmlDeviceGetPciInfo(h)
    try: numa = nv.nvmlDeviceGetNumaNodeId(h)
    except Exception as e: numa = repr(e)
    try: aff = list(nv.nvmlDeviceGetCpuAffinity(h, 4))
    except Exception as e: aff = repr(e)
    try: maff = list(nv.nvmlDeviceGetMemoryAffinity(h, 2, nv.NVML_AFFINITY_SCOPE_NODE))
    except Exception as e: maff = repr(e)
    print(i, pci.busId, "numa", numa, "cpuaff", [hex(a) for a in aff] if isinstance(aff, list) else aff, "memaff", maff)
PY
  echo "== meminfo"; grep -E "MemTotal|MemFree" /proc/meminfo
} > gpurun_out/r2b_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2b_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2b_bench_n2.err
tail -3 gpurun_out/r2b_pytest_multi.log; tail -5 gpurun_out/r2b_bench_n2.err; tail -c 1800 gpurun_out/r2b_bench_n2.json
