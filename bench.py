#!/usr/bin/env python
"""bench.py -- BiLSTM 256x61 windows/s (+ ODE trajectories/s) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference ...                     # the reference's CPU path (torch port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1 (driver launches it this way)

A step = one pass of the hot path (EnhancedLSTMModel.forward + softmax -> P(open)/P(closed),
04_lstm_model.py:206-222, 06:351) over one batch of synthetic windows per GPU.  Workload at every N:
BASELINE.json configs[1] (inference sweep point: one full wave of the recurrence kernel per GPU -- 16 896 windows
= 33 four-CTA clusters x 4 tiles x 128 for the fused bf16 path on a 148-SM B200 -- bf16 tensor-core mode), weak
scaling (per-GPU batch fixed; windows are independent -> no data-path collective).  Prints ONE JSON line on rank 0;
the long per-kernel dictionaries come first and the short headline keys (e2e, cpu_baseline, checks, train / ODE /
fp32 / config-5 numbers) last, so a truncated tail of the line still shows them.
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
REC_SAMPLES, REC_CHANNELS, SEQ_LEN, SEQ_STEP = 150_000, 61, 256, 128     # 300 s x 500 Hz recordings (01:51-52), 02:49-51
WIN_PER_REC = (REC_SAMPLES - SEQ_LEN) // SEQ_STEP + 1                     # 1170 (02:169)


def load_traffic():
    """ncu-measured DRAM bytes per launch (profiles/r2_traffic.json, else r1_traffic.json); None if absent."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons every 50 ms while the timed region runs."""
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


# ---- CPU legs (the only places bench.py executes oracle/) ------------------------------------------------------------------
def cpu_lstm_baseline(budget_s=15.0, sample=128):
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
            "sample": f"{sample} fp32 windows, best of {reps + 1}, oracle/torch_port.py"}


def cpu_train_baseline(batch=512):
    """One training step of the reference loop (04:486-507: fwd + bwd + clip + AdamW) on the host cores at the config-3
    batch (SURVEY §8 d: B = 512), after a small warm-up step."""
    import numpy as np
    import torch
    from lstm_ode_bci_b200 import synth
    from oracle import torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    port = torch_port.build_port(synth.make_lstm_params(42, 61, 128, 3), dropout=0.4).train()
    opt = torch.optim.AdamW(port.parameters(), lr=3e-4, weight_decay=1e-4)
    w = torch.tensor([0.8, 1.2])

    def one(n):
        x = torch.from_numpy(synth.make_windows(8, n, 256, 61))
        y = torch.from_numpy((np.arange(n) % 2).astype(np.int64))
        t0 = time.perf_counter()
        opt.zero_grad()
        torch.nn.functional.cross_entropy(port(x), y, weight=w).backward()
        torch.nn.utils.clip_grad_norm_(port.parameters(), 1.0)
        opt.step()
        return time.perf_counter() - t0
    one(32)
    dt = one(batch)
    return {"value": batch / dt, "unit": "windows/s", "ms_per_step": dt * 1e3, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 step, {batch} windows"}


def cpu_ode_baseline(n=1500):
    from lstm_ode_bci_b200 import synth
    from oracle import ode_oracle
    sw = synth.make_ode_sweep(42, n)
    t0 = time.perf_counter()
    ode_oracle.reference_style_loop(sw["p_open"], sw["p_closed"], dict(synth.DEFAULT_RATES), 0.5, 20)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "trajectories/s", "cores": 1, "kind": "port",
            "sample": f"{n} trajectories, per-sample odeint loop as 06:372-401 (serial by construction)"}


def run_cpu_legs(args, json_out):
    """child process of the b200 arm: the CPU baselines on every host core (affinity restored from BCI_BENCH_CPUS first)."""
    cpus = [int(c) for c in os.environ.get("BCI_BENCH_CPUS", "").split(",") if c]
    if cpus and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            pass
    out = {}
    for leg in [l for l in args.cpu_legs.split(",") if l]:
        if leg == "lstm":
            out["lstm"] = cpu_lstm_baseline()
        elif leg == "ode":
            out["ode"] = cpu_ode_baseline()
        elif leg.startswith("train"):
            out["train"] = cpu_train_baseline(int(leg.split(":")[1]) if ":" in leg else 512)
    print(json.dumps(out), file=json_out, flush=True)


def run_reference(args, rank, world, json_out):
    """--impl reference: the reference's CPU implementation of the same path (torch CPU port of the reference module -- the
    reference is loose scripts that do not exist on the GPU box) on the box's host cores.  Each step is a BOUNDED SAMPLE of the
    b200 arm's per-GPU batch: `ref_sample` fp32 windows (CPU throughput is flat in the batch size)."""
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
    cfg = workload_config(args, world)
    # what this arm actually timed: the same workload, `sample` windows per step, in the reference's own fp32
    cfg.update({"precision_mode": "fp32 (the reference's CPU path)", "windows_per_step_timed": sample,
                "b200_arm_windows_per_gpu": args.batch})
    line = {"impl": "reference", "metric": "bilstm_windows_per_s", "value": v, "unit": "windows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} fp32 windows per step, torch CPU path of the reference module"},
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu_legs"])
    ap.add_argument("--cpu-legs", default="", help="internal: comma list of CPU baselines to time in this (child) process")
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
    ap.add_argument("--no-extras", action="store_true", help="skip the preprocessing / ablation / H=256 / drop-in measurements")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process and its staging buffers to the GPU's NUMA node")
    ap.add_argument("--preproc-recordings", type=int, default=36)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world, json_out)
    if args.impl == "cpu_legs":
        return run_cpu_legs(args, json_out)

    import numpy as np
    import torch
    import torch.distributed as dist
    from lstm_ode_bci_b200 import _native, hostmem, integration, lstm, ode, ops, parallel, synth

    torch.cuda.set_device(local)
    _native.require_device(local)
    # host placement BEFORE any pinned buffer exists: this rank's CPUs and pages on its GPU's NUMA node
    all_cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None
    numa = {"node": None} if args.no_numa else hostmem.bind_to_gpu_node(local)
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

    def all_ranks(v):
        """per-rank scalar -> list over ranks (rank order)"""
        if world == 1:
            return [float(v)]
        t = torch.zeros(world, device="cuda", dtype=torch.float64)
        t[rank] = float(v)
        dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def timed(fn, reps, warm=2):
        """device time per call in ms (CUDA events on the current stream, barrier + synchronize on both sides, max over ranks)"""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

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
    # Headline e2e: integration.stream_recordings -- the host holds the band-passed, z-scored RECORDINGS (sample-major (S, C),
    # bf16), the windows are cut on the device by the input projection (bci_lstm_forward_view): 15 616 B per window cross PCIe
    # instead of the 62 464 B of the reference's materialised fp32 windows, with bit-identical results in the bf16 mode.
    # R recordings per step so that R x 1170 windows fill (not overflow) one pass of the recurrence kernel.
    d2h = [torch.empty((2 * B, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = max(3, min(K, 10))

    def e2e_loop(stream_fn, host, n_steps):
        done, last = [], None
        for i, (p, _) in enumerate(stream_fn(model, (host for _ in range(n_steps)), device=f"cuda:{local}")):
            d2h[i & 1][:p.shape[0]].copy_(p, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(); done.append(ev)
            if i >= 1:
                done[i - 1].synchronize()              # result of step i-1 is on the host
            last = p
        done[-1].synchronize()
        return int(last.shape[0])

    def e2e_measure(stream_fn, host, n_steps):
        e2e_loop(stream_fn, host, 2)
        barrier()
        t0 = time.perf_counter()
        n_win = e2e_loop(stream_fn, host, n_steps)
        barrier()
        sec = max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3
        nbytes = host.numel() * host.element_size()
        return {"value": world * n_win * n_steps / sec, "windows_per_step_per_gpu": n_win, "h2d_bytes_per_step": int(nbytes),
                "h2d_gbs_per_gpu": nbytes * n_steps / sec / 1e9}

    R = max(1, B // WIN_PER_REC)
    rec_host = torch.empty((R, REC_SAMPLES, REC_CHANNELS), dtype=torch.bfloat16, pin_memory=True)
    rec_host.copy_(torch.randn((R, REC_SAMPLES, REC_CHANNELS), device="cuda", generator=gen).to(torch.bfloat16))
    buf_nodes = hostmem.page_nodes(rec_host)
    m_rec = e2e_measure(integration.stream_recordings, rec_host, e2e_steps)
    e2e = {"value": m_rec["value"], "unit": "windows/s", "h2d_bytes_per_step": m_rec["h2d_bytes_per_step"],
           "d2h_bytes_per_step": m_rec["windows_per_step_per_gpu"] * 8, "steps": e2e_steps,
           "windows_per_step_per_gpu": m_rec["windows_per_step_per_gpu"], "h2d_gbs_per_gpu": m_rec["h2d_gbs_per_gpu"],
           "api": "integration.stream_recordings: pinned bf16 normalised recordings (R,150000,61) -> H2D (copy stream) -> windows "
                  "cut in place by the input projection -> probs D2H (pinned); every copy inside the timed region",
           "host_format": f"{R} recordings x {REC_SAMPLES} x {REC_CHANNELS} bf16 = {WIN_PER_REC} windows each (02:157-180)"}
    del rec_host
    # the same call on fp32 recordings, and the reference's literal host format (materialised fp32 windows, round 1's e2e path)
    rec32 = torch.empty((R, REC_SAMPLES, REC_CHANNELS), dtype=torch.float32, pin_memory=True)
    rec32.copy_(torch.randn((R, REC_SAMPLES, REC_CHANNELS), device="cuda", generator=gen))
    m32 = e2e_measure(integration.stream_recordings, rec32, 3)
    del rec32
    x_host = torch.empty((B, 256, 61), dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    mwin = e2e_measure(lambda m, it, device: integration.stream_lstm_probs(m, it, device), x_host, 3)
    del x_host
    e2e["other_host_formats"] = {"fp32_recordings": {k: m32[k] for k in ("value", "h2d_bytes_per_step", "h2d_gbs_per_gpu")},
                                 "fp32_windows_reference_format": {k: mwin[k] for k in ("value", "h2d_bytes_per_step", "h2d_gbs_per_gpu")}}
    h2d_ranks = all_ranks(m_rec["h2d_gbs_per_gpu"])
    e2e["numa"] = {"gpu_node": numa.get("node"), "bound": bool(numa.get("affinity") or numa.get("mempolicy")),
                   "staging_pages_on_node": buf_nodes, "h2d_gbs_by_rank": [round(v, 1) for v in h2d_ranks]}

    # ---- roofline of the dominant kernel ------------------------------------------------------
    fused = args.precision == "bf16" and prof["proj_gemm"][1] == 0
    flop_phase = FLOP_PHASE_FUSED if fused else FLOP_PHASE
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_launches = prof[dom]
    per_launch_flop = flop_phase[dom] * B * K / max(dom_launches, 1)
    achieved = flop_phase[dom] * B * K / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else None
    kernel_name = {"recurrence": "lstm_fused_bf16" if fused else "lstm_rec_bf16", "proj_gemm": "proj_gemm_bf16",
                   "input_proj": "input_proj_bf16", "pool_head": "attn_pool_stream_bf16"}
    roof = {"bound": "tensor", "kernel": kernel_name.get(dom, dom) if args.precision == "bf16" else dom,
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (achieved / peak_tf) if peak_tf else None, "traffic": None,
            "peak_source": "bf16_tflops_sustained of %s MEASURED_PEAKS (kernel timed inside a long step)" % peaks["source"],
            "flop_per_launch": per_launch_flop, "avg_launch_ms": dom_ms / max(dom_launches, 1),
            "phase_ms_per_step": {k: round(v[0] / K, 4) for k, v in prof.items()},
            "whole_path": {"achieved": value / world * FLOP_PER_WINDOW / 1e12, "unit": "TFLOP/s per GPU",
                           "frac_of_bf16_burst": value / world * FLOP_PER_WINDOW / 1e12 / peaks["bf16_tflops"],
                           "frac_of_bf16_sustained": value / world * FLOP_PER_WINDOW / 1e12 / peaks["bf16_tflops_sustained"]}}
    traffic = load_traffic()
    tkey = "recurrence_fused" if (fused and dom == "recurrence") else dom
    if traffic and args.precision == "bf16" and tkey in traffic:
        wpl = traffic[tkey].get("windows_per_launch", traffic.get("windows_per_launch"))
        if wpl and B % wpl == 0:
            # per launch like flop_per_launch: the MEAN over the launches of one step (layer 0 reads a 128-wide input, layers 1-2 a
            # 256-wide one), not one layer's figure
            roof["traffic"] = traffic[tkey]["bytes_per_launch"]
            roof["traffic_algorithmic"] = traffic[tkey]["algorithmic_bytes_per_launch"]
    if args.precision == "fp32":
        fp32_peak = ops.fp32_peak_probe()
        roof.update({"bound": "fp32", "peak": fp32_peak, "frac": achieved / fp32_peak,
                     "peak_source": "FP32 FMA micro-benchmark measured in this run (bci_fp32_peak_probe)"})

    line = {"metric": "bilstm_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
    tail = {}          # short headline keys, appended after the long dictionaries
    checks = {}

    # ---- ODE ensemble (second half of the metric: trajectories/s) -----------------------------
    if not args.no_ode:
        n = args.ode_n
        sw = synth.make_ode_sweep(42 + rank, n)
        dev = {k: torch.from_numpy(v).cuda() for k, v in sw.items()}

        def ode_step(want_traj=True, mode="rk4", substeps=ODE_SUBSTEPS):
            return ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                      alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, style="ref06", mode=mode,
                                      t_end=20.0, n_points=20, substeps=substeps, want_traj=want_traj)
        t_traj = timed(lambda: ode_step(True), 5, 3)
        t_final = timed(lambda: ode_step(False), 5, 3)
        t_rk45 = timed(lambda: ode_step(True, "rk45"), 5, 3)
        t_adapt = timed(lambda: ode_step(True, substeps=0), 5, 3)
        fp32_peak = ops.fp32_peak_probe()
        bytes_traj = n * (36 + 240 + 12)

        def roof_ode(t):
            tf = n * ODE_FLOP_PER_TRAJ / (t * 1e-3) / 1e12
            return {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak}
        r_traj = roof_ode(t_traj)
        r_traj.update({"traffic": (traffic or {}).get("ode_rk4", {}).get("bytes_per_launch") if n == 1 << 24 else None,
                       "hbm_gbs": bytes_traj / (t_traj * 1e-3) / 1e9, "hbm_frac": bytes_traj / (t_traj * 1e-3) / 1e9 / peaks["hbm_gbs"]})
        line["ode"] = {
            "unit": "trajectories/s", "n_per_gpu": n, "substeps": ODE_SUBSTEPS, "flop_per_trajectory": ODE_FLOP_PER_TRAJ,
            "rk4_full_trajectory": {"value": world * n / (t_traj * 1e-3), "ms": t_traj, "roofline": r_traj},
            "rk4_final_state_only": {"value": world * n / (t_final * 1e-3), "ms": t_final, "roofline": roof_ode(t_final)},
            "rk4_adaptive_substeps": {"value": world * n / (t_adapt * 1e-3), "ms": t_adapt,
                                      "note": "substeps=0: S per trajectory for <= 2e-7 truncation"},
            "rk45_full_trajectory": {"value": world * n / (t_rk45 * 1e-3), "ms": t_rk45, "rtol": 1e-3, "atol": 1e-6},
            "fp32_peak_source": "FMA micro-benchmark in this run"}
        tail["ode_rk4_gtraj_s"] = round(world * n / (t_traj * 1e-3) / 1e9, 4)
        tail["ode_rk4_frac"] = round(r_traj["frac"], 4)
        tail["ode_rk4_final_frac"] = round(line["ode"]["rk4_final_state_only"]["roofline"]["frac"], 4)
        tail["ode_rk45_gtraj_s"] = round(world * n / (t_rk45 * 1e-3) / 1e9, 4)
        del dev

    # ---- training step (BASELINE configs[2]): fwd + BPTT + gradient all-reduce + clip + AdamW, 512 windows per GPU -----
    if not args.no_train:
        from lstm_ode_bci_b200 import train
        tb = args.train_batch
        xt = x[:tb].contiguous()
        yt = (torch.arange(tb, device="cuda") % 2)

        def make_trainer(collective, precision="fp32"):
            tm = lstm.from_params(params, precision="fp32", device=f"cuda:{local}", dropout=0.4).train()
            return tm, train.FusedTrainer(tm, lr=3e-4, weight_decay=1e-4, max_norm=1.0, class_weight=[0.8, 1.2], collective=collective,
                                          precision=precision)
        tmodel, trainer = make_trainer("auto")          # p2p fused step when world > 1
        seeds = iter(range(10_000))
        tms = timed(lambda: trainer.step(xt, yt, seed=next(seeds)), 5, 2)
        loss_t, norm_t = trainer.step(xt, yt, seed=next(seeds))
        line["train_step"] = {"value": world * tb / (tms * 1e-3), "unit": "windows/s", "ms_per_step": tms, "windows_per_gpu": tb,
                              "precision": "fp32", "dropout": 0.4,
                              "collective": "bci_fused_step (peer-memory all-reduce inside clip+AdamW)" if world > 1 else "none",
                              "tflops_per_gpu": tb * 3 * FLOP_PER_WINDOW / (tms * 1e-3) / 1e12,
                              "loss": float(loss_t), "grad_norm": float(norm_t)}
        tail["train_ms_per_step"] = round(tms, 3)
        tail["train_windows_s"] = round(world * tb / (tms * 1e-3), 1)
        if world > 1:
            # replicas must be bit-identical after the timed steps (rank-ordered sums, deterministic norm): compare a hash
            hsh = trainer.flat.view(torch.int32).to(torch.int64).sum()
            hs = all_ranks(float(hsh % (1 << 40)))
            checks["replicas_bit_identical"] = bool(len(set(hs)) == 1)
            # the same step without any collective = the N=1 step on this box, timed back to back
            barrier()
            trainer.close()
            lmodel, ltrainer = make_trainer("none")
            lms = timed(lambda: ltrainer.step(xt, yt, seed=next(seeds)), 5, 2)
            tail["train_ms_per_step_local"] = round(lms, 3)
            tail["train_eff_vs_n1"] = round(lms / tms, 4)
            del ltrainer, lmodel
            # the collective + optimizer in isolation on IDENTICAL seeded gradients: fused peer-memory step vs NCCL all-reduce +
            # bci_adamw_step (max-abs over all parameters after 3 steps), and bit-identity of the fused result across ranks
            nflt = 1_137_731
            gsd = torch.Generator(device="cuda").manual_seed(5)
            p0 = torch.randn(nflt, device="cuda", generator=gsd) * 0.1
            comm = parallel.P2PComm(nflt)
            iso = {}
            for mode in ("p2p", "nccl"):
                pp_, mm_, vv_ = p0.clone(), torch.zeros(nflt, device="cuda"), torch.zeros(nflt, device="cuda")
                nrm = torch.zeros(2, device="cuda")
                for st in range(1, 4):
                    gg = torch.Generator(device="cuda").manual_seed(100 * st + rank)
                    grad = torch.randn(nflt, device="cuda", generator=gg) * (0.01 * st)
                    if mode == "p2p":
                        comm.bucket.copy_(grad)
                        comm.fused_step(pp_, mm_, vv_, 3e-3, (0.9, 0.999), 1e-8, 1e-2, st, 0.5, nrm)
                    else:
                        dist.all_reduce(grad)
                        _native.check(_native.lib().bci_adamw_step(ops._ptr(pp_), ops._ptr(grad), ops._ptr(mm_), ops._ptr(vv_), nflt,
                                                                   3e-3, 0.9, 0.999, 1e-8, 1e-2, st, 1.0 / world, 0.5, ops._ptr(nrm),
                                                                   ops._stream()))
                iso[mode] = pp_
            checks["fused_step_vs_nccl_adamw_maxabs"] = float((iso["p2p"] - iso["nccl"]).abs().max())
            hs = all_ranks(float(iso["p2p"].view(torch.int32).to(torch.int64).sum() % (1 << 40)))
            checks["fused_step_bit_identical_across_ranks"] = bool(len(set(hs)) == 1)
            barrier()
            comm.close()
        else:
            trainer.close()
        del trainer
        # the same step in the mixed mode (BCI_TRAIN_MIXED: the reference's own GPU training runs under autocast + GradScaler,
        # 04:486-490): 16-bit tensor-core recurrences with W_hh resident in tensor memory, single-pass TF32 GEMMs
        barrier()
        mmodel, mtrainer = make_trainer("auto", "mixed")
        mms = timed(lambda: mtrainer.step(xt, yt, seed=next(seeds)), 5, 2)
        loss_m, norm_m = mtrainer.step(xt, yt, seed=next(seeds))
        line["train_step_mixed"] = {"value": world * tb / (mms * 1e-3), "unit": "windows/s", "ms_per_step": mms, "windows_per_gpu": tb,
                                    "precision": "mixed (fp16 / bf16 recurrence operands, TF32 GEMMs, fp32 accumulation and state)",
                                    "tflops_per_gpu": tb * 3 * FLOP_PER_WINDOW / (mms * 1e-3) / 1e12,
                                    "loss": float(loss_m), "grad_norm": float(norm_m)}
        tail["train_mixed_ms_per_step"] = round(mms, 3)
        tail["train_mixed_windows_s"] = round(world * tb / (mms * 1e-3), 1)
        if world > 1:
            hs = all_ranks(float(mtrainer.flat.view(torch.int32).to(torch.int64).sum() % (1 << 40)))
            checks["mixed_replicas_bit_identical"] = bool(len(set(hs)) == 1)
            barrier()
        mtrainer.close()
        del mtrainer, mmodel
        # fp32 parity mode of the inference forward (BASELINE configs[1] lists fp32 next to bf16): one full pass of its pair recurrence
        # (bci_lstm_chunk_windows: 9472 windows on 148 SMs); the whole LSTM stack runs on the tensor cores in split fp16 precision
        tmodel.eval()
        nf = min(ops.lstm_chunk_windows(tmodel._engine("fp32")), B)
        xf = x[:nf]
        with torch.no_grad():
            fms = timed(lambda: tmodel.predict_proba(xf), 3, 2)
        line["fp32_mode"] = {"value": world * nf / (fms * 1e-3), "unit": "windows/s", "ms": fms, "windows_per_gpu": nf,
                             "tflops_per_gpu": nf * FLOP_PER_WINDOW / (fms * 1e-3) / 1e12,
                             "tolerance": "logits/probs <= 1e-5, attention <= 1e-6 vs the reference's fp32 CPU path"}
        tail["fp32_windows_s"] = round(world * nf / (fms * 1e-3), 1)
        del tmodel

    # ---- config 5: recordings on the host -> preprocessing -> LSTM -> coupling -> ODE -> forecast -> gather ----------------------
    if not args.no_config5:
        # 60 subjects x 3 sessions x 2 tasks = 360 recordings = 421 200 windows (SURVEY §8 d), ranks own contiguous recording ranges.
        # The host holds RAW fp32 recordings (R, 61, 150000) as mne returns them (02:200); one pinned 12-recording batch is re-sent
        # for every batch of the rank's share (synthetic data: the bytes copied and the work done are those of distinct recordings)
        # Batches of 14 recordings = 16 380 windows fill one pass of the recurrence kernel (15 would spill 654 windows into a second,
        # nearly empty pass), and every rank runs whole batches: 364 / 364 / 392 / 448 recordings at N = 1 / 2 / 4 / 8 instead of 360;
        # the figure reported is windows actually processed per second.
        n_rec_total, per_batch = 360, max(1, chunk // WIN_PER_REC)
        rb, re_ = parallel.shard_range(n_rec_total, rank, world)
        n_batches = -(-(re_ - rb) // per_batch)
        my_recs = n_batches * per_batch
        tot_recs = int(sum(all_ranks(my_recs)))
        n_total = tot_recs * WIN_PER_REC
        raw_host = torch.empty((per_batch, REC_CHANNELS, REC_SAMPLES), dtype=torch.float32, pin_memory=True)
        raw_host.copy_(torch.randn((per_batch, REC_CHANNELS, REC_SAMPLES), device="cuda", generator=gen) * 1e-5 + 1e-4)
        integ = integration.LSTMODEIntegration(model, ode.CognitiveStateODE(), coupling_strength=0.5, device=f"cuda:{local}")

        def cfg5():
            # ranks own equal whole-batch shares, so the contiguous split of the probability gather matches
            return parallel.forecast_pipeline_from_recordings(integ, (raw_host for _ in range(n_batches)), n_total, raw=True,
                                                              want_traj=False)
        cfg5(); barrier()
        t0 = time.perf_counter()
        res5 = cfg5()
        torch.cuda.synchronize()
        barrier()
        sec5 = max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3
        line["config5"] = {"value": n_total / sec5, "unit": "windows/s", "seconds": sec5, "windows": n_total, "recordings": tot_recs,
                           "pipeline": "host raw fp32 recordings -> H2D -> bci_preprocess (filtfilt+zscore+windows) -> BiLSTM bf16 -> "
                                       "coupling+ODE RK4 -> NCCL gather of probs -> 08 forecast (h=5,10,20)",
                           "h2d_bytes_per_window": REC_CHANNELS * SEQ_STEP * 4}
        tail["config5_windows_s"] = round(n_total / sec5, 1)
        if world > 1:
            # sharded == unsharded on a 4096-window subsample (same seed on every rank): each rank runs its shard through the
            # sharded pipeline and all 4096 windows alone; the fp32 engine is used because its per-window results do not depend
            # on the batch a window arrives in (bit-identical), so any difference would be a sharding / gather error
            m32 = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3, logit_gain=20.0), precision="fp32", device=f"cuda:{local}")
            integ32 = integration.LSTMODEIntegration(m32, ode.CognitiveStateODE(), coupling_strength=0.5, device=f"cuda:{local}")
            sub = torch.randn((4096, 256, 61), device="cuda", generator=torch.Generator(device="cuda").manual_seed(777))
            sb, se = parallel.shard_range(4096, rank, world)
            r_sh = parallel.forecast_pipeline_sharded(integ32, sub[sb:se].contiguous(), 4096)
            r_un = integ32.predict_batch_device(sub, want_traj=False)        # (traj, probs, final, pred, cls) unsharded
            d_pr = float((r_sh["probs"] - r_un[1]).abs().max())
            d_fi = float((r_sh["final"] - r_un[2][sb:se]).abs().max())
            checks["config5_sharded_vs_unsharded_probs_maxabs"] = max(all_ranks(d_pr))
            checks["config5_sharded_vs_unsharded_final_maxabs"] = max(all_ranks(d_fi))
            fb, fe = r_sh["forecast_range"]
            fc_un = integration._forecast_device(r_un[1][:4096 - 20, 1].contiguous(), integ32.base_params, 20, [5, 10, 20], sub.device, 8)
            d_fc = float((r_sh["forecast"] - fc_un[fb:fe]).abs().max()) if fe > fb else 0.0
            checks["config5_sharded_vs_unsharded_forecast_maxabs"] = max(all_ranks(d_fc))
            del sub, m32, integ32
        del raw_host, res5, integ

    # ---- SURVEY §8 f rows 3-4 and the H = 256 checkpoint size (measured, not part of `value`) ----
    if not args.no_extras:
        from lstm_ode_bci_b200 import preprocessing as pp
        Rp, Cc, ns = args.preproc_recordings, 61, 150000
        raw = torch.randn((Rp, Cc, ns), device="cuda", generator=gen) * 1e-5 + 1e-4      # fp32 stand-in for mne's raw.get_data()
        b_, a_, zi_, padlen = pp.design_bandpass()
        out = pp.preprocess_recordings(raw, b_, a_, zi_, padlen)
        nwin = int(out["X"].shape[0])
        pms = timed(lambda: pp.preprocess_recordings(raw, b_, a_, zi_, padlen), 3, 1)
        fp64_peak = ops.fp64_peak_probe()                  # TFLOP/s counting 2 flop per DFMA
        useful_ops = Rp * Cc * (ns + 54) * 2 * 33.0
        line["preprocess"] = {"value": world * nwin / (pms * 1e-3), "unit": "windows/s", "ms": pms, "recordings_per_gpu": Rp,
                              "roofline": {"bound": "fp64", "achieved": useful_ops / (pms * 1e-3) / 1e12, "peak": fp64_peak / 2.0,
                                           "unit": "T fp64 instr/s", "frac": useful_ops / (pms * 1e-3) / 1e12 / (fp64_peak / 2.0)}}
        del raw, out
        abl = lstm.AblationLSTMModel(input_size=61, hidden_size=256, num_layers=1, bidirectional=False, use_attention=False).cuda().eval()
        xa = x[:2048]
        with torch.no_grad():
            ams = timed(lambda: abl(xa), 3, 2)
        line["ablation_minimal"] = {"value": world * 2048 / (ams * 1e-3), "unit": "windows/s", "ms": ams,
                                    "config": "09:342-349 'Minimal': H=256, 1 layer, unidirectional, mean pooling, fp32"}
        del abl
        # hidden_size 256 = the reference's trained checkpoint on 61 channels (04:876-877): bf16 tensor-core mode
        m256 = lstm.from_params(synth.make_lstm_params(44, 61, 256, 3), precision="bf16", device=f"cuda:{local}")
        b256 = ops.lstm_chunk_windows(m256._engine("bf16"))
        x256 = x[:b256]
        hms = timed(lambda: m256.predict_proba(x256), 3, 2)
        line["h256_bf16"] = {"value": world * int(x256.shape[0]) / (hms * 1e-3), "unit": "windows/s", "ms": hms,
                             "windows_per_gpu": int(x256.shape[0]), "tflops_per_gpu": int(x256.shape[0]) * 2.223047168e9 / (hms * 1e-3) / 1e12}
        tail["h256_bf16_windows_s"] = round(world * int(x256.shape[0]) / (hms * 1e-3), 1)
        del m256
        # ... and its training step in the mixed mode (CTA-pair tensor-core recurrences, lstm_rec_swap.cu), 512 windows per GPU, no collective
        if not args.no_train:
            from lstm_ode_bci_b200 import train as _train
            t256 = lstm.from_params(synth.make_lstm_params(44, 61, 256, 3), precision="fp32", device=f"cuda:{local}", dropout=0.4).train()
            tr256 = _train.FusedTrainer(t256, lr=3e-4, weight_decay=1e-4, max_norm=1.0, class_weight=[0.8, 1.2], collective="none",
                                        precision="mixed")
            xt256, yt256 = x[:args.train_batch].contiguous(), (torch.arange(args.train_batch, device="cuda") % 2)
            h256ms = timed(lambda: tr256.step(xt256, yt256, seed=7), 3, 2)
            line["h256_train_mixed"] = {"ms_per_step": h256ms, "windows_per_gpu": args.train_batch, "value": world * args.train_batch / (h256ms * 1e-3),
                                        "unit": "windows/s", "library_bar_ms": {"cudnn_autocast_fp16": 38.9, "cudnn_tf32": 40.0,
                                                                                 "source": "profiles/r1_library_bar_torch_cuda.json"}}
            tail["h256_train_mixed_ms_per_step"] = round(h256ms, 3)
            tr256.close()
            del tr256, t256
        # 07_explainability.py:287-361 at its own defaults (1 000 test windows, 61 channels x 5 permutations + the baseline = 306
        # sweeps): the subset stays resident, bci_permute_channels gathers the variants, the bf16 forward runs them in full passes
        from lstm_ode_bci_b200 import explain
        n_pi, reps_pi = 1000, 5
        ch_pi = [-1] + [c for c in range(61) for _ in range(reps_pi)]
        rng_pi = np.random.default_rng(3)
        perms_pi = np.stack([rng_pi.permutation(n_pi) for _ in ch_pi])
        y_pi = rng_pi.integers(0, 2, n_pi)
        x_pi = x[:n_pi]
        explain.permuted_channel_accuracy(model, x_pi, y_pi, ch_pi[:40], perms_pi[:40])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        explain.permuted_channel_accuracy(model, x_pi, y_pi, ch_pi, perms_pi)
        torch.cuda.synchronize()
        sec_pi = time.perf_counter() - t0
        line["permutation_importance"] = {"value": world * len(ch_pi) * n_pi / sec_pi, "unit": "windows/s", "seconds": sec_pi,
                                          "sweeps": len(ch_pi), "windows_per_sweep": n_pi, "precision": "bf16",
                                          "config": "07:287 defaults: n_samples=1000, n_permutations=5, 61 channels"}
        tail["permutation_importance_windows_s"] = round(world * len(ch_pi) * n_pi / sec_pi, 1)
        # the reference's own call, unmodified: LSTMODEIntegration.predict_batch(X_numpy, forecast_steps=20, batch_size=512)
        # (06:801-806) -- pageable fp32 numpy windows in, numpy out, LSTM + coupling + ODE + classification; a model built
        # with precision="auto" runs its bf16 engine there because that is where the reference autocasts (06:348-351).  Rank 0 only.
        if rank == 0:
            auto = lstm.from_params(params, precision="auto", device=f"cuda:{local}")
            integ = integration.LSTMODEIntegration(auto, ode.CognitiveStateODE(), coupling_strength=0.5, device=f"cuda:{local}")
            nd = 4 * B      # four passes of the recurrence wave: the call is a pipeline (stage k+1 on the host | kernels of k), fill included
            xd = np.concatenate([np.random.default_rng(0).standard_normal((B, 256, 61), dtype=np.float32)] * 4)
            # this leg runs on rank 0 alone (the other ranks wait at the barrier, one core each): the staging copy may use the rest
            # of the host cores instead of rank 0's 1/world share
            if world > 1 and not os.environ.get("BCI_STAGING_THREADS"):
                os.environ["BCI_STAGING_THREADS"] = str(max(1, min(32, (len(all_cpus) if all_cpus else (os.cpu_count() or 1)) - (world - 1))))
                staging_env_set = True
            else:
                staging_env_set = False
            integ.predict_batch(xd, forecast_steps=20, batch_size=512, show_progress=False)
            torch.cuda.synchronize()
            dsec = float("inf")
            for _ in range(3):   # host-bound leg on a shared VM: best of three calls (0.1-0.2 s each)
                t0 = time.perf_counter()
                trj, _pp, _pd = integ.predict_batch(xd, forecast_steps=20, batch_size=512, show_progress=False)
                dsec = min(dsec, time.perf_counter() - t0)
            line["dropin_predict_batch"] = {"value": nd / dsec, "unit": "windows/s", "windows": nd, "seconds": dsec, "timing": "best of 3 calls",
                                            "precision": "auto -> bf16 under the method's own autocast",
                                            "host_bytes_in": int(xd.nbytes), "staging_threads": integration._staging_threads(),
                                            # the native staging copy narrows the pageable fp32 windows to bf16 (the rounding the
                                            # input projection applies on load): half the bytes cross the link, no result bit changes
                                            "h2d_bytes": int(xd.nbytes) // 2, "d2h_bytes": int(trj.nbytes + _pp.nbytes + _pd.nbytes)}
            tail["dropin_predict_batch_windows_s"] = round(nd / dsec, 1)
            if staging_env_set:
                os.environ.pop("BCI_STAGING_THREADS", None)
            del xd, trj, integ, auto
        barrier()

    # ---- CPU baselines: rank 0, at every N (the other ranks wait at the barrier).  They run in a child process that starts
    # with the affinity this process had BEFORE it bound itself to the GPU's NUMA node (threads created since inherit the
    # binding), so the reference's CPU path gets every host core ----
    # The other ranks must SLEEP meanwhile (a wait on the rendezvous store), not spin in an NCCL barrier / cudaDeviceSynchronize:
    # OpenMP's active-wait barriers degrade by an order of magnitude when spinning threads of other processes hold cores
    # (measured: 68 instead of 1058 windows/s on the 2-GPU box with rank 1 parked in dist.barrier()).
    cpu = None
    barrier()
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if rank == 0 and not args.no_cpu_baseline:
        ncpu = len(all_cpus) if all_cpus else (os.cpu_count() or 1)
        env = dict(os.environ, BCI_BENCH_CPUS=",".join(str(c) for c in (all_cpus or [])), OMP_NUM_THREADS=str(ncpu))
        env.pop("MKL_NUM_THREADS", None)
        legs = ["lstm"] + ([] if args.no_ode else ["ode"]) + ([] if args.no_train else ["train:%d" % args.train_batch])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu_legs", "--cpu-legs", ",".join(legs)],
                           env=env, capture_output=True, text=True, timeout=600)
        got = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"error": r.stderr[-300:]}
        cpu = got.get("lstm")
        if "ode" in got:
            tail["ode_cpu_baseline"] = got["ode"]
        if "train" in got:
            tail["train_cpu_baseline"] = got["train"]
        if "error" in got:
            tail["cpu_baseline_error"] = got["error"]
    if store is not None:
        if rank == 0:
            store.set("bci_bench_cpu_legs_done", "1")
        else:
            import datetime
            store.wait(["bci_bench_cpu_legs_done"], datetime.timedelta(seconds=1200))
    barrier()
    line["e2e"] = e2e
    line["cpu_baseline"] = cpu
    if world > 1:
        line["checks"] = checks
    line.update(tail)
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
