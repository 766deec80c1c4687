"""Two-rank GPU test of the data-parallel training step (BASELINE config 3): the peer-memory fused step
(`bci_fused_step`: all-reduce inside the clip + AdamW kernels over NVLink) against (a) the NCCL all-reduce +
bci_adamw_step path and (b) a single-process step on the concatenated batch with the CPU port of the reference
(oracle/torch_port.py).  Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a one-GPU box."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from lstm_ode_bci_b200 import lstm, synth, train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        H, B, T = 128, 6, 24
        params = synth.make_lstm_params(45, 61, H, 3, logit_gain=8.0)
        x = synth.make_windows(10, B * world, T, 61)
        y = (np.arange(B * world) % 2).astype(np.int64)
        xs = torch.from_numpy(x[rank * B:(rank + 1) * B]).cuda()
        ys = torch.from_numpy(y[rank * B:(rank + 1) * B]).cuda()
        res = {}
        for mode in ("p2p", "nccl"):
            m = lstm.from_params(params, precision="fp32", dropout=0.0).train()
            tr = train.FusedTrainer(m, lr=3e-3, weight_decay=1e-2, max_norm=0.05, collective=mode)
            norms = []
            for _ in range(3):
                _loss, norm = tr.step(xs, ys)
                norms.append(float(norm))
            torch.cuda.synchronize()
            res[mode] = (tr.flat.detach().cpu().numpy().copy(), norms)
            if tr.comm is not None:
                dist.barrier()
                tr.comm.close()
        # the collective + optimizer step in isolation, on IDENTICAL (seeded) gradients: the fused peer-memory step against
        # NCCL all-reduce + bci_adamw_step.  (Whole training steps cannot be compared this tightly: the backward pass
        # accumulates with atomics, and Adam turns 1e-9-level gradient noise on near-zero entries into +-lr steps.)
        import ctypes as C
        from lstm_ode_bci_b200 import _native as N, ops, parallel
        n = 300007
        gen = torch.Generator(device="cuda").manual_seed(5)
        p0 = torch.randn(n, device="cuda", generator=gen) * 0.1
        comm = parallel.P2PComm(n)
        iso = {}
        for mode in ("p2p", "nccl"):
            pp, mm, vv = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
            norm = torch.zeros(2, device="cuda")
            for step in range(1, 4):
                gg = torch.Generator(device="cuda").manual_seed(100 * step + rank)
                grad = torch.randn(n, device="cuda", generator=gg) * (0.01 * step)
                if mode == "p2p":
                    comm.bucket.copy_(grad)
                    comm.fused_step(pp, mm, vv, 3e-3, (0.9, 0.999), 1e-8, 1e-2, step, 0.5, norm)
                else:
                    dist.all_reduce(grad)
                    N.check(N.lib().bci_adamw_step(ops._ptr(pp), ops._ptr(grad), ops._ptr(mm), ops._ptr(vv), n, 3e-3, 0.9, 0.999, 1e-8,
                                                   1e-2, step, 1.0 / world, 0.5, ops._ptr(norm), ops._stream()))
            torch.cuda.synchronize()
            iso[mode] = (pp.cpu().numpy().copy(), float(norm[1]))
        dist.barrier()
        comm.close()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), p2p=res["p2p"][0], nccl=res["nccl"][0],
                 p2p_norms=np.array(res["p2p"][1]), nccl_norms=np.array(res["nccl"][1]),
                 iso_p2p=iso["p2p"][0], iso_nccl=iso["nccl"][0], iso_norms=np.array([iso["p2p"][1], iso["nccl"][1]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_fused_step_matches_nccl_and_single_process(tmp_path):
    import torch.multiprocessing as mp
    from lstm_ode_bci_b200 import synth
    from oracle import torch_port
    world = 2
    mp.spawn(_worker, args=(world, 29711, str(tmp_path)), nprocs=world, join=True)
    r = [dict(np.load(tmp_path / ("rank%d.npz" % i))) for i in range(world)]
    # replicas stay bit-identical across ranks (rank-ordered sums, deterministic norm)
    assert np.array_equal(r[0]["p2p"], r[1]["p2p"])
    # identical gradients in -> the fused peer-memory step equals NCCL all-reduce + bci_adamw_step (only the reduction order of
    # the gradient norm differs: 1e-7 relative on the clip coefficient), and is bit-identical across ranks
    assert np.array_equal(r[0]["iso_p2p"], r[1]["iso_p2p"])
    assert np.abs(r[0]["iso_p2p"] - r[0]["iso_nccl"]).max() <= 2e-7
    assert abs(r[0]["iso_norms"][0] - r[0]["iso_norms"][1]) <= 1e-6 * r[0]["iso_norms"][1]
    # whole training steps through both collectives agree on average (see the comment in the worker)
    assert np.mean(np.abs(r[0]["p2p"] - r[0]["nccl"])) <= 1e-6
    assert np.abs(r[0]["p2p_norms"] - r[0]["nccl_norms"]).max() <= 1e-5 * np.abs(r[0]["nccl_norms"]).max()
    # and as one process on the concatenated batch through the reference's own loop (CPU port, unweighted CE = mean)
    H, B, T = 128, 6, 24
    params = synth.make_lstm_params(45, 61, H, 3, logit_gain=8.0)
    x = torch.from_numpy(synth.make_windows(10, B * world, T, 61))
    y = torch.from_numpy((np.arange(B * world) % 2).astype(np.int64))
    port = torch_port.build_port(params, dropout=0.0).train()
    opt = torch.optim.AdamW(port.parameters(), lr=3e-3, weight_decay=1e-2)
    for step in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(port(x), y).backward()
        norm_ref = float(torch.nn.utils.clip_grad_norm_(port.parameters(), 0.05))
        opt.step()
        assert abs(r[0]["p2p_norms"][step] - norm_ref) <= 3e-4 * norm_ref
    flat_ref = np.concatenate([p.detach().numpy().reshape(-1) for _, p in port.named_parameters()])
    names = [(k, p.numel()) for k, p in port.named_parameters()]
    off = 0
    for k, n in names:
        if k != "attention.attention.2.bias":     # d loss / d b2 == 0 exactly (see test_gpu_train.py)
            assert np.mean(np.abs(r[0]["p2p"][off:off + n] - flat_ref[off:off + n])) <= 3e-6, k
            assert np.abs(r[0]["p2p"][off:off + n] - flat_ref[off:off + n]).max() <= 3e-3, k      # at most one lr-sized Adam step
        off += n
