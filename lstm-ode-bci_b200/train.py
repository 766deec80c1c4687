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
    @staticmethod
    def forward(ctx, model, x, want_attn, dropout, seed, *params):
        hid = model._engine("fp32")          # training runs the fp32 path
        h = ops._handles[hid]
        B, T = int(x.shape[0]), int(x.shape[1])
        logits = torch.empty((B, model.num_classes), device=x.device, dtype=torch.float32)
        attn = torch.empty((B, T), device=x.device, dtype=torch.float32)
        nbytes = ops.lstm_workspace_bytes(hid, B, T, 1)
        ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
        xc = x.contiguous()
        N.check(N.lib().bci_lstm_forward(h.ptr, ops._ptr(xc), B, T, 1, float(dropout), int(seed), ops._ptr(logits),
                                         C.c_void_p(0), ops._ptr(attn), ops._ptr(ws), nbytes, ops._stream()))
        ctx.model, ctx.hid, ctx.ws, ctx.nbytes, ctx.shape = model, hid, ws, nbytes, (B, T)
        ctx.x = xc
        ctx.need_dx = x.requires_grad
        ctx.mark_non_differentiable(attn)
        return logits, attn

    @staticmethod
    def backward(ctx, dlogits, _dattn):
        model, (B, T) = ctx.model, ctx.shape
        h = ops._handles[ctx.hid]
        names = [k for k, _ in model.named_parameters()]
        grads = {k: torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                 for k, p in model.named_parameters()}
        gs = _grad_struct(model, grads)
        dx = torch.empty_like(ctx.x) if ctx.need_dx else None
        dl = dlogits.contiguous().float()
        N.check(N.lib().bci_lstm_backward(h.ptr, ops._ptr(ctx.x), ops._ptr(dl), B, T, ops._ptr(dx), C.byref(gs),
                                          ops._ptr(ctx.ws), ctx.nbytes, ops._stream()))
        # the workspace is kept: backward(retain_graph=True) may be called again on the same forward (07:252)
        return (None, dx, None, None, None) + tuple(grads[k] for k in names)


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

    Mirrors 04_lstm_model.py:438,486-507 (AdamW lr 3e-4, wd 1e-4, clip 1.0) without AMP/accumulation; parameters
    live in ONE flat fp32 bucket (views are handed back to the module) so the all-reduce and the update are one
    launch each."""

    def __init__(self, model, lr=3e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, class_weight=None,
                 process_group=None, collective="auto"):
        self.model = model
        self.lr, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.pg = process_group
        self.step_count = 0
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
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.norm = torch.zeros(2, device=dev, dtype=torch.float32)
        off = 0
        self.grad_views = {}
        with torch.no_grad():
            for k, p in model.named_parameters():
                m = p.numel()
                self.flat[off:off + m].copy_(p.reshape(-1))
                p.data = self.flat[off:off + m].view_as(p)          # module parameters alias the bucket
                self.grad_views[k] = self.grad[off:off + m].view_as(p)
                off += m
        self.class_weight = None if class_weight is None else torch.as_tensor(class_weight, dtype=torch.float32, device=dev)
        self.hid = model._engine("fp32")

    def step(self, x, y, seed=0):
        """x (B,T,C) CUDA fp32, y (B,) int64.  Returns (loss, pre-clip grad norm) as device tensors."""
        import torch.distributed as dist
        model = self.model
        h = ops._handles[self.hid]
        B, T = int(x.shape[0]), int(x.shape[1])
        ops.lstm_load_weights(self.hid, {k: v for k, v in model.state_dict().items()})
        model._loaded["fp32"] = model._signature()
        p_drop = float(model.dropout_p) if model.training else 0.0
        logits = torch.empty((B, model.num_classes), device=x.device, dtype=torch.float32)
        nbytes = ops.lstm_workspace_bytes(self.hid, B, T, 1)
        ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
        xc = x.contiguous()
        N.check(N.lib().bci_lstm_forward(h.ptr, ops._ptr(xc), B, T, 1, p_drop, int(seed), ops._ptr(logits), C.c_void_p(0),
                                         C.c_void_p(0), ops._ptr(ws), nbytes, ops._stream()))
        # weighted cross-entropy and its gradient wrt the logits (tiny (B,2) tensors: torch elementwise ops)
        logp = torch.log_softmax(logits, dim=1)
        w = self.class_weight[y] if self.class_weight is not None else torch.ones(B, device=x.device)
        wsum = w.sum()
        loss = -(w * logp.gather(1, y[:, None])[:, 0]).sum() / wsum
        dlogits = (torch.exp(logp) - torch.nn.functional.one_hot(y, model.num_classes).float()) * (w / wsum)[:, None]
        gs = _grad_struct(model, self.grad_views)
        N.check(N.lib().bci_lstm_backward(h.ptr, ops._ptr(xc), ops._ptr(dlogits.contiguous()), B, T, C.c_void_p(0), C.byref(gs),
                                          ops._ptr(ws), nbytes, ops._stream()))
        self.step_count += 1
        if self.comm is not None:
            self.comm.fused_step(self.flat, self.m, self.v, self.lr, self.betas, self.eps, self.wd, self.step_count,
                                 self.max_norm, self.norm)
            return loss.detach(), self.norm[1]
        scale = 1.0
        if self.collective == "nccl" and self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)      # NCCL over NVLink on GPUs
            scale = 1.0 / self.world
        N.check(N.lib().bci_adamw_step(ops._ptr(self.flat), ops._ptr(self.grad), ops._ptr(self.m), ops._ptr(self.v),
                                       self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                       self.step_count, scale, self.max_norm, ops._ptr(self.norm), ops._stream()))
        return loss.detach(), self.norm[1]
