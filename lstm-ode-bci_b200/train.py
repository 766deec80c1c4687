"""Training-step host side: autograd bridge and a fused data-parallel trainer.

  lstm_attn_autograd   EnhancedLSTMModel.forward in .train() mode with grad enabled: forward saves its
                       activations in a workspace, backward is bci_lstm_backward (BPTT).  Gradients come
                       back through torch.autograd, so the reference's own loop (04_lstm_model.py:482-507:
                       loss.backward(); clip_grad_norm_; optimizer.step()) runs unchanged, and so does the
                       backward-to-input of 07_explainability.py:242-258.
  FusedTrainer         config 3 of BASELINE.json: per-GPU batch, weighted cross-entropy, flat fp32 gradient
                       bucket; collective="p2p" (default on >1 GPU ranks): the all-reduce is fused into the clip +
                       AdamW kernels over NVLink peer memory (bci_fused_step, csrc/comm_p2p.cu); collective="nccl":
                       one torch.distributed all-reduce per step, then bci_adamw_step.  Windows are independent, so
                       data parallelism over ranks is exact.
"""
import ctypes as C

import torch

from . import _native as N
from . import ops


def _grad_struct(model, grads):
    return ops.fill_pointer_struct(N.LstmGrads(), grads, model.num_layers)


class _LstmAttnFn(torch.autograd.Function):
    """autograd glue between torch and the two registered ops `bci::lstm_attn_forward_train` / `bci::lstm_attn_backward`."""

    @staticmethod
    def forward(ctx, model, x, want_attn, dropout, seed, *params):
        hid = model._engine("fp32")          # the training step lives on the fp32 engine (fp32 master weights and layouts)
        xc = x.contiguous()
        with torch.cuda.device(xc.device):
            ops.lstm_set_train_mode(hid, model._train_precision_now())
            logits, attn, ws = ops.lstm_attn_forward_train(xc, hid, float(dropout), int(seed))
        ctx.hid, ctx.ws, ctx.x = hid, ws, xc
        ctx.names = [k for k, _ in model.named_parameters()]
        ctx.need_dx = x.requires_grad
        ctx.mark_non_differentiable(attn)
        return logits, attn

    @staticmethod
    def backward(ctx, dlogits, _dattn):
        with torch.cuda.device(ctx.x.device):
            dx, flat = ops.lstm_attn_backward(ctx.x, dlogits.contiguous().float(), ctx.ws, ctx.hid, bool(ctx.need_dx))
        # the workspace is kept: backward(retain_graph=True) may be called again on the same forward (07:252)
        by_key = {k: flat[o:o + n].view(shape) for k, o, n, shape in ops.param_layout(ctx.hid)}
        return (None, dx if ctx.need_dx else None, None, None, None) + tuple(by_key[k] for k in ctx.names)


def lstm_attn_autograd(model, x, return_attention=False, seed=None):
    """Differentiable forward (fp32 path).  Dropout follows the module's `dropout_p` when model.training."""
    p = float(model.dropout_p) if model.training else 0.0
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
    params = tuple(pp for _, pp in model.named_parameters())
    logits, attn = _LstmAttnFn.apply(model, x, bool(return_attention), p, seed, *params)
    return (logits, attn) if return_attention else logits


class FusedTrainer:
    """One optimizer step = forward (train) -> weighted CE -> BPTT -> [all-reduce] -> clip + AdamW (fused).

    Mirrors the loop body of 04_lstm_model.py:482-507 (AdamW lr 3e-4, wd 1e-4 04:438; weighted CrossEntropyLoss 04:456-458;
    loss / accumulation_steps and an optimizer step every `accumulation_steps` micro-batches 04:489,497-507; clip 1.0).  The
    reference's GradScaler (04:490,499-503) exists because it trains in fp16; this step computes in fp32, where scaling the loss
    by s and the gradients by 1/s is the identity, so there is no scaler.  Parameters live in ONE flat fp32 bucket (views are
    handed back to the module) so the all-reduce and the update are one launch each; the loss and its gradient are one launch
    (bci_ce_loss_grad); after every optimizer step the module is told that its parameters changed behind torch's back
    (mark_weights_changed), so a following model.eval()(x) -- fp32 or bf16 engine -- re-packs and sees the new weights."""

    def __init__(self, model, lr=3e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, class_weight=None,
                 process_group=None, collective="auto", accumulation_steps=1, precision="fp32"):
        """precision: "fp32" (parity step) or "mixed" -- the reduced-precision step corresponding to the reference's autocast +
        GradScaler training (04:486-490,499-503): 16-bit tensor-core recurrences and single-pass TF32 GEMMs, fp32 master weights,
        loss and optimizer; bf16 carries the gradients through time, so there is no loss scale to maintain."""
        self.model = model
        if precision not in ops.TRAIN_MODES:
            raise N.BciError(-1, "precision must be fp32 or mixed")
        self.precision = precision
        self.lr, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.pg = process_group
        self.step_count = 0          # optimizer steps taken
        self.micro_count = 0         # micro-batches since the last optimizer step
        self.accum = int(accumulation_steps)
        if self.accum < 1:
            raise N.BciError(-1, "accumulation_steps must be >= 1")
        ps = [p for _, p in model.named_parameters()]
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        import torch.distributed as dist
        world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if collective == "auto":
            collective = "p2p" if world > 1 else "none"
        if collective not in ("p2p", "nccl", "none"):
            raise N.BciError(-1, "collective must be auto, p2p, nccl or none")
        self.collective, self.world = collective, world
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.comm = None
        if collective == "p2p":
            from .parallel import P2PComm
            self.comm = P2PComm(n, group=process_group, device=dev)
            self.grad = self.comm.bucket                      # backward writes straight into the peer-visible bucket
        else:
            self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        # with accumulation the micro-batch gradients are written to a scratch bucket and summed into self.grad
        self.micro = torch.empty(n, device=dev, dtype=torch.float32) if self.accum > 1 else self.grad
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.norm = torch.zeros(2, device=dev, dtype=torch.float32)
        off = 0
        self.grad_views, self.micro_views = {}, {}
        with torch.no_grad():
            for k, p in model.named_parameters():
                m = p.numel()
                self.flat[off:off + m].copy_(p.reshape(-1))
                p.data = self.flat[off:off + m].view_as(p)          # module parameters alias the bucket
                self.grad_views[k] = self.grad[off:off + m].view_as(p)
                self.micro_views[k] = self.micro[off:off + m].view_as(p)
                off += m
        self.class_weight = None if class_weight is None else torch.as_tensor(class_weight, dtype=torch.float32, device=dev)
        self.hid = model._engine("fp32")

    def step(self, x, y, seed=0):
        """One micro-batch: x (B,T,C) CUDA fp32, y (B,) int64.  Returns (loss of this micro-batch -- unscaled, as the reference
        reports it, 04:509 -- and the pre-clip gradient norm of the last optimizer step) as device tensors.  The optimizer runs
        on every `accumulation_steps`-th call."""
        import torch.distributed as dist
        model = self.model
        with torch.cuda.device(x.device):
            self.hid = model._engine("fp32")          # re-packs the kernel layouts iff the parameters changed since the last pack
            h = ops._handles[self.hid]
            ops.lstm_set_train_mode(self.hid, self.precision)
            B, T = int(x.shape[0]), int(x.shape[1])
            p_drop = float(model.dropout_p) if model.training else 0.0
            logits = torch.empty((B, model.num_classes), device=x.device, dtype=torch.float32)
            nbytes = ops.lstm_workspace_bytes(self.hid, B, T, 1)
            ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
            xc = x.contiguous()
            N.check(N.lib().bci_lstm_forward(h.ptr, ops._ptr(xc), B, T, 1, p_drop, int(seed), ops._ptr(logits), C.c_void_p(0),
                                             C.c_void_p(0), ops._ptr(ws), nbytes, ops._stream()))
            # weighted cross-entropy / accumulation_steps and its gradient wrt the logits: one launch
            loss, dlogits = ops.ce_loss_grad(logits, y, self.class_weight, 1.0 / self.accum)
            gs = _grad_struct(model, self.micro_views)
            N.check(N.lib().bci_lstm_backward(h.ptr, ops._ptr(xc), ops._ptr(dlogits), B, T, C.c_void_p(0), C.byref(gs),
                                              ops._ptr(ws), nbytes, ops._stream()))
            if self.accum > 1:
                ops.grad_accumulate(self.grad, self.micro, first=self.micro_count == 0)
            self.micro_count += 1
            loss_out = loss[0] * float(self.accum) if self.accum > 1 else loss[0]
            if self.micro_count < self.accum:
                return loss_out, self.norm[1]
            self.micro_count = 0
            self.step_count += 1
            if self.comm is not None:
                self.comm.fused_step(self.flat, self.m, self.v, self.lr, self.betas, self.eps, self.wd, self.step_count,
                                     self.max_norm, self.norm)
            else:
                scale = 1.0
                if self.collective == "nccl" and self.world > 1:
                    dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)      # NCCL over NVLink on GPUs
                    scale = 1.0 / self.world
                N.check(N.lib().bci_adamw_step(ops._ptr(self.flat), ops._ptr(self.grad), ops._ptr(self.m), ops._ptr(self.v),
                                               self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                               self.step_count, scale, self.max_norm, ops._ptr(self.norm), ops._stream()))
            model.mark_weights_changed()
        return loss_out, self.norm[1]

    def close(self):
        """Release the peer-memory communicator (the gradient views die with it)."""
        if self.comm is not None:
            self.grad = self.micro = None
            self.grad_views = self.micro_views = {}
            self.comm.close()
            self.comm = None
