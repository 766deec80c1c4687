#!/usr/bin/env python
"""bench.py -- BiLSTM 256x61 windows/s (+ ODE trajectories/s) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference ...                     # the reference's CPU path (torch port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1 (driver launches it this way)

A step = one pass of the hot path (EnhancedLSTMModel.forward + softmax -> P(open)/P(closed),
04_lstm_model.py:206-222, 06:351) over one batch of synthetic windows per GPU.  Workload at every N:
BASELINE.json configs[1] (inference sweep point: one or two full waves of the recurrence kernel per GPU --
16 896 windows = 33 four-CTA clusters x 4 tiles x 128 for the fused bf16 path on a 148-SM B200 -- bf16
tensor-core mode), weak scaling (per-GPU batch fixed; windows are independent -> no data-path collective).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW = 557_793_536          # SURVEY.md §8 d (H=128, T=256, C=61, 3 layers, forward)
FLOP_PHASE = {"input_proj": 3_997_696, "proj_gemm": 335_544_320, "recurrence": 201_326_592,
              "pool_head": 16_842_752 + 82_176}
# fused bf16 path (csrc/lstm_bf16_fused.cu): projection and recurrence are ONE kernel, reported under "recurrence"
FLOP_PHASE_FUSED = dict(FLOP_PHASE, proj_gemm=0, recurrence=335_544_320 + 201_326_592)
ODE_SUBSTEPS = 8
ODE_FLOP_PER_TRAJ = 12 + 19 * ODE_SUBSTEPS * 123 + 20 * 11      # SURVEY.md §8 d: 18 928 at S=8


def load_traffic():
    """ncu-measured DRAM bytes per launch (profiles/r1_traffic.json); None if the file is absent."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin, self.t_end = 0.0, float("inf")

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples taken while the timed region ran (a line is stamped when it is read: up to one period after it was taken);
        # if the region was shorter than the sampling period, fall back to the nearest samples around it
        inside = [r for ts, r in self.rows if self.t_begin <= ts <= self.t_end + 0.06]
        self.rows = inside if inside else [r for ts, r in self.rows if ts >= self.t_begin - 0.1][:2]
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_lstm_baseline(budget_s=20.0, sample=128):
    """Reference CPU path (torch port of the reference module, all host threads) on a bounded sample."""
    import torch
    from lstm_ode_bci_b200 import synth
    from oracle import torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    params = synth.make_lstm_params(42, 61, 128, 3)
    port = torch_port.build_port(params).eval()
    x = torch.from_numpy(synth.make_windows(7, sample, 256, 61))
    with torch.no_grad():
        port(x[:8])
        t0 = time.perf_counter()
        port(x)
        one = time.perf_counter() - t0
        reps = max(1, min(10, int(budget_s / max(one, 1e-3)) - 1))
        best = one
        for _ in range(reps):
            t0 = time.perf_counter()
            port(x)
            best = min(best, time.perf_counter() - t0)
    return {"value": sample / best, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample} windows x 256 x 61 fp32, best of {reps + 1} (oracle/torch_port.py = torch CPU path of the reference module)"}


def cpu_ode_baseline(n=1500):
    from lstm_ode_bci_b200 import synth
    from oracle import ode_oracle
    sw = synth.make_ode_sweep(42, n)
    t0 = time.perf_counter()
    ode_oracle.reference_style_loop(sw["p_open"], sw["p_closed"], dict(synth.DEFAULT_RATES), 0.5, 20)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "trajectories/s", "cores": 1, "kind": "port",
            "sample": f"{n} trajectories, per-sample scipy.odeint loop as 06:372-401 (serial by construction)"}


def run_reference(args, rank, world, json_out):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    import torch
    from lstm_ode_bci_b200 import synth
    from oracle import torch_port
    if args.batch <= 0:
        args.batch = 16896      # the b200 arm's default per-GPU batch on a 148-SM B200 (bci_lstm_chunk_windows, fused bf16 path)
    torch.set_num_threads(os.cpu_count() or 1)
    sample = args.ref_sample
    port = torch_port.build_port(synth.make_lstm_params(42, 61, 128, 3)).eval()
    x = torch.from_numpy(synth.make_windows(7, sample, 256, 61))
    with torch.no_grad():
        for _ in range(warm):
            torch.softmax(port(x), 1)
        t0 = time.perf_counter()
        for _ in range(steps):
            torch.softmax(port(x), 1)
        dt = time.perf_counter() - t0
    v = sample * steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "bilstm_windows_per_s", "value": v, "unit": "windows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} windows per step (bounded sample of the per-GPU batch), torch CPU path of the reference module"},
            "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=json_out, flush=True)


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: BiLSTM(3x128,T=256,C=61)+attention pooling inference -> P(open)/P(closed)",
            "windows_per_gpu": args.batch, "global_windows_per_step": args.batch * world, "hidden": 128, "layers": 3,
            "seq_len": 256, "channels": 61, "precision_mode": args.precision,
            "l2_policy": "inputs larger than L2 (%.2f GB per step per GPU)" % (args.batch * 256 * 61 * 4 / 1e9),
            "parallelism": f"window-sharded x{world}, no data-path collective"}


def _claim_stdout():
    """Keep the real stdout for the ONE JSON line; anything a library prints to fd 1 (NCCL's version banner, ...) is
    redirected to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="windows per GPU per step; 0 = the smallest multiple >= 16384 of the forward's internal pass size "
                         "(bci_lstm_chunk_windows: 16896 for the fused bf16 path on 148 SMs)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ode-n", type=int, default=1 << 24)
    ap.add_argument("--ref-sample", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ode", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--train-batch", type=int, default=512)
    ap.add_argument("--no-extras", action="store_true", help="skip the preprocessing / ablation-variant measurements")
    ap.add_argument("--preproc-recordings", type=int, default=36)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world, json_out)

    import numpy as np
    import torch
    import torch.distributed as dist
    from lstm_ode_bci_b200 import _native, integration, lstm, ode, ops, synth

    torch.cuda.set_device(local)
    _native.require_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = load_peaks()
    K, W = args.steps, args.warmup
    params = synth.make_lstm_params(42, 61, 128, 3)
    model = lstm.from_params(params, precision=args.precision, device=f"cuda:{local}")
    hid = model._engine(args.precision)
    chunk = ops.lstm_chunk_windows(hid)
    if args.batch <= 0:
        args.batch = chunk * max(1, -(-16384 // chunk))
    B = args.batch
    gen = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = torch.randn((B, 256, 61), device="cuda", generator=gen)           # N(0,1): z-scored EEG (02:134-152)

    # ---- device-resident throughput (value) --------------------------------------------------
    # the clock sampler starts before the warm-up so that nvidia-smi is already reporting when the timed region begins (a K-step
    # region lasts a few hundred ms); samples are stamped and only those taken inside the region are kept
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        model.predict_proba(x)
    ops.lstm_set_profiling(hid, True)
    ops.lstm_get_profile(hid)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.t_begin = time.time()
    l0 = ops.launch_count()
    e0.record()
    for _ in range(K):
        probs = model.predict_proba(x)
    e1.record()
    barrier()
    sampler.t_end = time.time()
    launches = ops.launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    prof = ops.lstm_get_profile(hid)
    ops.lstm_set_profiling(hid, False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ---------------------------------
    x_host = torch.empty((B, 256, 61), dtype=torch.float32).pin_memory()
    x_host.copy_(x)
    d2h = [torch.empty((B, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    def e2e_run(n_steps):
        """n_steps API steps, each = one batch of B host windows in, B x 2 probabilities out (pinned host);
        the H2D copy of step i+1 overlaps the kernels of step i (integration.stream_lstm_probs)."""
        done = []
        for i, (p, _) in enumerate(integration.stream_lstm_probs(model, (x_host for _ in range(n_steps)), f"cuda:{local}")):
            d2h[i & 1].copy_(p, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(); done.append(ev)
            if i >= 1:
                done[i - 1].synchronize()              # result of step i-1 is on the host
        done[-1].synchronize()
        return d2h[(n_steps - 1) & 1]
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(K, 10))
    ph = e2e_run(e2e_steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e = {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "windows/s",
           "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(ph.numel() * 4),
           "api": "integration.stream_lstm_probs: pinned host windows -> H2D on a copy stream (one-wave pieces), overlapped with "
                  "the kernels of the previous step -> probabilities D2H to pinned host memory; every step's copies are inside the timed region",
           "steps": e2e_steps, "h2d_gbs": x_host.numel() * 4 * e2e_steps / (e2e_ms * 1e-3) / 1e9}

    # ---- roofline of the dominant kernel ------------------------------------------------------
    fused = args.precision == "bf16" and prof["proj_gemm"][1] == 0
    flop_phase = FLOP_PHASE_FUSED if fused else FLOP_PHASE
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_launches = prof[dom]
    per_launch_flop = flop_phase[dom] * B * K / max(dom_launches, 1)
    achieved = flop_phase[dom] * B * K / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else None
    kernel_name = {"recurrence": "lstm_fused_bf16 (projection + recurrence, 4-CTA clusters)" if fused else "lstm_rec_bf16",
                   "proj_gemm": "proj_gemm_bf16", "input_proj": "input_proj_bf16", "pool_head": "attn_score_bf16 + attn_pool_finish_bf16"}
    roof = {"bound": "tensor", "kernel": dom, "kernel_name": kernel_name.get(dom, dom) if args.precision == "bf16" else dom,
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (achieved / peak_tf) if peak_tf else None, "traffic": None,
            "peak_source": "bf16_tflops_sustained of %s (kernel timed inside a long step)" % peaks["source"],
            "flop_per_launch": per_launch_flop, "avg_launch_ms": dom_ms / max(dom_launches, 1),
            "phase_ms_per_step": {k: v[0] / K for k, v in prof.items()},
            "phase_share": {k: v[0] / max(sum(p[0] for p in prof.values()), 1e-9) for k, v in prof.items()},
            "whole_path": {"achieved": value / world * FLOP_PER_WINDOW / 1e12, "unit": "TFLOP/s per GPU",
                           "frac_of_bf16_burst": value / world * FLOP_PER_WINDOW / 1e12 / peaks["bf16_tflops"],
                           "frac_of_bf16_sustained": value / world * FLOP_PER_WINDOW / 1e12 / peaks["bf16_tflops_sustained"]}}
    traffic = load_traffic()
    tkey = "recurrence_fused" if (fused and dom == "recurrence") else dom
    if traffic and args.precision == "bf16" and tkey in traffic:
        wpl = traffic[tkey].get("windows_per_launch", traffic.get("windows_per_launch"))
        if wpl and B % wpl == 0:
            roof["traffic"] = traffic[tkey]["bytes_per_launch"]
            roof["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, %d-window launch); algorithmic %d" % (
                wpl, traffic[tkey]["algorithmic_bytes_per_launch"])
    if args.precision == "fp32":
        fp32_peak = ops.fp32_peak_probe()
        roof.update({"bound": "fp32", "peak": fp32_peak, "frac": achieved / fp32_peak,
                     "peak_source": "FP32 FMA micro-benchmark measured in this run (bci_fp32_peak_probe)"})

    line = {"metric": "bilstm_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "clocks": clocks}

    # ---- ODE ensemble (second half of the metric: trajectories/s) -----------------------------
    if not args.no_ode:
        n = args.ode_n
        sw = synth.make_ode_sweep(42 + rank, n)
        dev = {k: torch.from_numpy(v).cuda() for k, v in sw.items()}
        def ode_step(want_traj=True, mode="rk4"):
            return ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                      alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, style="ref06", mode=mode,
                                      t_end=20.0, n_points=20, substeps=ODE_SUBSTEPS, want_traj=want_traj)
        def time_ode(**kw):
            for _ in range(3):
                ode_step(**kw)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ode_step(**kw)
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / 5
        t_traj, t_final, t_rk45 = time_ode(want_traj=True), time_ode(want_traj=False), time_ode(want_traj=True, mode="rk45")
        fp32_peak = ops.fp32_peak_probe()
        bytes_traj = n * (36 + 240 + 12)
        line["ode"] = {
            "metric": "ode_trajectories_per_s", "unit": "trajectories/s", "n_per_gpu": n, "substeps": ODE_SUBSTEPS,
            "rk4_full_trajectory": {"value": world * n / (t_traj * 1e-3), "ms": t_traj,
                                    "roofline": {"bound": "fp32", "achieved": n * ODE_FLOP_PER_TRAJ / (t_traj * 1e-3) / 1e12,
                                                 "peak": fp32_peak, "unit": "TFLOP/s",
                                                 "frac": n * ODE_FLOP_PER_TRAJ / (t_traj * 1e-3) / 1e12 / fp32_peak,
                                                 "traffic": (traffic or {}).get("ode_rk4", {}).get("bytes_per_launch") if n == 1 << 24 else None,
                                                 "hbm_gbs": bytes_traj / (t_traj * 1e-3) / 1e9,
                                                 "hbm_frac": bytes_traj / (t_traj * 1e-3) / 1e9 / peaks["hbm_gbs"]}},
            "rk4_final_state_only": {"value": world * n / (t_final * 1e-3), "ms": t_final,
                                     "roofline": {"bound": "fp32", "achieved": n * ODE_FLOP_PER_TRAJ / (t_final * 1e-3) / 1e12,
                                                  "peak": fp32_peak, "unit": "TFLOP/s",
                                                  "frac": n * ODE_FLOP_PER_TRAJ / (t_final * 1e-3) / 1e12 / fp32_peak}},
            "rk45_full_trajectory": {"value": world * n / (t_rk45 * 1e-3), "ms": t_rk45, "rtol": 1e-3, "atol": 1e-6},
            "flop_per_trajectory": ODE_FLOP_PER_TRAJ, "fp32_peak_source": "FMA micro-benchmark in this run"}
        del dev

    # ---- training step (BASELINE configs[2]): fwd + BPTT + NCCL all-reduce + clip + AdamW, 512 windows per GPU -----
    if not args.no_train:
        from lstm_ode_bci_b200 import train
        tb = args.train_batch
        tmodel = lstm.from_params(params, precision="fp32", device=f"cuda:{local}", dropout=0.4).train()
        trainer = train.FusedTrainer(tmodel, lr=3e-4, weight_decay=1e-4, max_norm=1.0, class_weight=[0.8, 1.2])
        xt = x[:tb].contiguous()
        yt = (torch.arange(tb, device="cuda") % 2)
        for i in range(2):
            trainer.step(xt, yt, seed=i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tsteps = 5
        for i in range(tsteps):
            loss_t, norm_t = trainer.step(xt, yt, seed=10 + i)
        b.record()
        barrier()
        tms = max_over_ranks(a.elapsed_time(b)) / tsteps
        line["train_step"] = {"metric": "train_windows_per_s", "value": world * tb / (tms * 1e-3), "unit": "windows/s",
                              "ms_per_step": tms, "windows_per_gpu": tb, "precision": "fp32", "dropout": 0.4,
                              "gemms": "split-precision 3xTF32 tcgen05 (gemm_tf32x3.cu)" if os.environ.get("BCI_FP32_GEMM", "")[:1] != "s"
                              else "CUDA-core FFMA (BCI_FP32_GEMM=simt)",
                              "optimizer": "AdamW(3e-4, wd 1e-4) + clip 1.0 with the gradient all-reduce fused into the optimizer kernels "
                                           "over NVLink peer memory (bci_fused_step)" if world > 1
                              else "AdamW(3e-4, wd 1e-4) + clip 1.0, fused",
                              "flop_per_window": 3 * FLOP_PER_WINDOW, "achieved_tflops_per_gpu": tb * 3 * FLOP_PER_WINDOW / (tms * 1e-3) / 1e12,
                              "loss": float(loss_t), "grad_norm": float(norm_t)}
        del trainer
        # fp32 parity mode of the inference forward (BASELINE configs[1] lists fp32 next to bf16): 2048 windows, one chunk
        tmodel.eval()
        xf = x[:2048]
        with torch.no_grad():
            for _ in range(2):
                tmodel.predict_proba(xf)
            barrier()
            a.record()
            for _ in range(3):
                tmodel.predict_proba(xf)
            b.record()
        barrier()
        fms = max_over_ranks(a.elapsed_time(b)) / 3
        line["fp32_mode"] = {"metric": "windows_per_s", "value": world * int(xf.shape[0]) / (fms * 1e-3), "unit": "windows/s", "ms": fms,
                             "windows_per_gpu": int(xf.shape[0]),
                             "tflops_per_gpu": int(xf.shape[0]) * FLOP_PER_WINDOW / (fms * 1e-3) / 1e12,
                             "tolerance": "logits/probabilities <= 1e-5, attention <= 1e-6 vs the reference's fp32 CPU path"}
        del tmodel

    # ---- SURVEY §8 f rows 3-4: preprocessing of raw recordings and one ablation variant (measured, not part of `value`) ----
    if not args.no_extras:
        from lstm_ode_bci_b200 import preprocessing as pp
        R, Cc, n = args.preproc_recordings, 61, 150000
        raw = torch.randn((R, Cc, n), device="cuda", generator=gen) * 1e-5 + 1e-4      # fp32 stand-in for mne's raw.get_data()
        b_, a_, zi_, padlen = pp.design_bandpass()
        for _ in range(2):
            out = pp.preprocess_recordings(raw, b_, a_, zi_, padlen)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = pp.preprocess_recordings(raw, b_, a_, zi_, padlen)
        b.record()
        barrier()
        pms = max_over_ranks(a.elapsed_time(b)) / 3
        nwin = int(out["X"].shape[0])
        # the recursion is FP64-pipe bound: 33 unfused operations (17 mul + 16 add/sub, scipy's evaluation order) per sample and
        # pass; the chunk-parallel form adds 8192 warm-up samples per 16384-sample chunk.  Peak = measured DFMA issue rate.
        fp64_peak = ops.fp64_peak_probe()                  # TFLOP/s counting 2 flop per DFMA
        useful_ops = R * Cc * (n + 54) * 2 * 33.0
        pbytes = R * Cc * n * (4 + 8 + 8 + 8 + 8 + 8)
        line["preprocess"] = {"metric": "preprocessed_windows_per_s", "value": world * nwin / (pms * 1e-3), "unit": "windows/s",
                              "ms": pms, "recordings_per_gpu": R, "samples_per_recording": n, "channels": Cc, "windows_per_gpu": nwin,
                              "roofline": {"bound": "fp64", "achieved": useful_ops / (pms * 1e-3) / 1e12, "peak": fp64_peak / 2.0,
                                           "unit": "T fp64 instr/s (useful filter operations vs measured DFMA issue rate)",
                                           "frac": useful_ops / (pms * 1e-3) / 1e12 / (fp64_peak / 2.0),
                                           "hbm_gbs": pbytes / (pms * 1e-3) / 1e9, "hbm_frac": pbytes / (pms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                           "note": "whole call (filter passes + statistics + windowing); warm-up work (x1.5) not counted as useful"}}
        del raw, out
        abl = lstm.AblationLSTMModel(input_size=61, hidden_size=256, num_layers=1, bidirectional=False, use_attention=False).cuda().eval()
        xa = x[:2048]
        with torch.no_grad():
            for _ in range(2):
                abl(xa)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                abl(xa)
            b.record()
        barrier()
        ams = max_over_ranks(a.elapsed_time(b)) / 3
        # hidden_size 256 = the reference's trained checkpoint on 61 channels (04:876-877): bf16 tensor-core mode
        m256 = lstm.from_params(synth.make_lstm_params(44, 61, 256, 3), precision="bf16", device=f"cuda:{local}")
        b256 = ops.lstm_chunk_windows(m256._engine("bf16"))
        x256 = x[:b256]
        for _ in range(2):
            m256.predict_proba(x256)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            m256.predict_proba(x256)
        b.record()
        barrier()
        hms = max_over_ranks(a.elapsed_time(b)) / 3
        line["h256_bf16"] = {"metric": "windows_per_s", "value": world * int(x256.shape[0]) / (hms * 1e-3), "unit": "windows/s", "ms": hms,
                             "windows_per_gpu": int(x256.shape[0]), "flop_per_window": 2_223_047_168,
                             "tflops_per_gpu": int(x256.shape[0]) * 2.223047168e9 / (hms * 1e-3) / 1e12,
                             "config": "EnhancedLSTMModel(61, hidden 256, 3 layers): projection GEMM + cluster recurrence (lstm_bf16_h256.cu)"}
        del m256
        line["ablation_minimal"] = {"metric": "windows_per_s", "value": world * 2048 / (ams * 1e-3), "unit": "windows/s", "ms": ams,
                                    "config": "09:342-349 'Minimal': H=256, 1 layer, unidirectional, mean pooling, fp32", "windows_per_gpu": 2048}
        del abl
        # the reference's own call, unmodified: LSTMODEIntegration.predict_batch(X_numpy, forecast_steps=20, batch_size=512)
        # (06_lstm_ode_integration.py:801-806) -- pageable numpy in, numpy out, LSTM + coupling + ODE + classification.  Rank 0 only.
        if rank == 0:
            import numpy as np
            integ = integration.LSTMODEIntegration(model, ode.CognitiveStateODE(), coupling_strength=0.5, device=f"cuda:{local}")
            nd = 2 * B
            xd = np.random.default_rng(0).standard_normal((nd, 256, 61), dtype=np.float32)
            integ.predict_batch(xd, forecast_steps=20, batch_size=512, show_progress=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            trj, _pp, _pd = integ.predict_batch(xd, forecast_steps=20, batch_size=512, show_progress=False)
            dsec = time.perf_counter() - t0
            line["dropin_predict_batch"] = {"metric": "windows_per_s", "value": nd / dsec, "unit": "windows/s", "windows": nd, "seconds": dsec,
                                            "api": "LSTMODEIntegration.predict_batch(X: pageable numpy (N,256,61), forecast_steps=20, batch_size=512) "
                                                   "-> (trajectories (N,20,3) f64, probs (N,2), predictions (N,)) as numpy; one process, one GPU",
                                            "h2d_bytes": int(xd.nbytes), "d2h_bytes": int(trj.nbytes + _pp.nbytes + _pd.nbytes)}
            del xd, trj, integ
        barrier()

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_lstm_baseline()
        if not args.no_ode:
            line["ode"]["cpu_baseline"] = cpu_ode_baseline()
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
