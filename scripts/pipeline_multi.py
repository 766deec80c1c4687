"""Config 5 on N GPUs: 60 subjects x 3 sessions x 2 tasks of synthetic EEG = 421 200 windows, sharded by contiguous
window range; LSTM -> coupling -> ODE (06 path) and 08-style forecast; final gather.  torchrun --nproc-per-node N."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from lstm_ode_bci_b200 import integration, lstm, ode, parallel, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 421200
b, e = parallel.shard_range(N, rank, world)
model = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3, logit_gain=20.0), precision="bf16", device=f"cuda:{local}")
integ = integration.LSTMODEIntegration(model, ode.CognitiveStateODE(), 0.5, device=f"cuda:{local}")
g = torch.Generator(device="cuda").manual_seed(1000 + rank)
x = torch.randn((e - b, 256, 61), device="cuda", generator=g)
for it in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = parallel.forecast_pipeline_sharded(integ, x, N)
    full_final = parallel.gather_shards(res["final"], N)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "5: LSTM->coupling->ODE + forecast", "windows": N, "n_gpus": world, "seconds": dt,
                      "windows_per_s": N / dt, "final_shape": list(full_final.shape),
                      "mean_final": [float(v) for v in full_final.mean(dim=0)]}))
if world > 1: dist.destroy_process_group()
