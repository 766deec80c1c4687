"""Library bar on the same B200: stock torch.nn (cuDNN RNN + cuBLAS) running the reference's architecture
(04_lstm_model.py:153-222) on cuda:0 -- what a user of the reference gets on this GPU today (SURVEY.md 2.3:
"the bar is PyTorch's own cuDNN/cuBLAS path").  Not a product path and not the oracle: a plain torch module,
timed with CUDA events.  fp32 and the reference's own autocast(fp16) mode (04:486-490, 06:348-351).
Writes one JSON line to stdout."""
import json
import sys

import torch
from torch import nn


class Ref(nn.Module):
    def __init__(self, C=61, H=128, L=3):
        super().__init__()
        self.input_proj = nn.Sequential(nn.Linear(C, H), nn.LayerNorm(H), nn.GELU(), nn.Dropout(0.2))
        self.lstm = nn.LSTM(H, H, L, batch_first=True, dropout=0.4, bidirectional=True)
        self.layer_norm = nn.LayerNorm(2 * H)
        self.att = nn.Sequential(nn.Linear(2 * H, H), nn.Tanh(), nn.Linear(H, 1))
        self.classifier = nn.Sequential(nn.Linear(2 * H, H), nn.GELU(), nn.Dropout(0.4), nn.Linear(H, H // 2), nn.GELU(),
                                        nn.Dropout(0.4), nn.Linear(H // 2, 2))

    def forward(self, x):
        y = self.layer_norm(self.lstm(self.input_proj(x))[0])
        w = torch.softmax(self.att(y), dim=1)
        return self.classifier((w * y).sum(dim=1))


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.manual_seed(42)
    dev = torch.device("cuda:0")
    m = Ref(H=H).to(dev)
    out = {"what": "stock torch.nn (cuDNN/cuBLAS) on the same GPU", "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "hidden": H, "gpu": torch.cuda.get_device_name(0)}
    # inference, batch sizes of the reference (512, 06:339) and one wave of ours (16896)
    m.eval()
    for B in (512, 4096, 16896):
        x = torch.randn(B, 256, 61, device=dev)
        for tag, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("fp32_strict", torch.autocast("cuda", enabled=False)),
                         ("autocast_fp16", torch.autocast("cuda", dtype=torch.float16))):
            # "fp32" = torch defaults (cuDNN RNN may use TF32); "fp32_strict" = TF32 disabled everywhere (an fp32-parity mode)
            torch.backends.cudnn.allow_tf32 = tag != "fp32_strict"
            torch.backends.cuda.matmul.allow_tf32 = False
            if B > 4096 and tag == "fp32_strict":
                continue

            def f():
                with torch.no_grad(), ctx:
                    return torch.softmax(m(x), 1)
            try:
                ms = timed(f)
                out["fwd_B%d_%s" % (B, tag)] = {"ms": ms, "windows_per_s": B / ms * 1e3}
            except RuntimeError as e:   # out of memory at the large batch
                out["fwd_B%d_%s" % (B, tag)] = {"error": str(e)[:80]}
        del x
    # training step B=512: forward + backward + clip + AdamW (04:482-507 without accumulation)
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4, weight_decay=1e-4)
    x = torch.randn(512, 256, 61, device=dev); y = torch.arange(512, device=dev) % 2
    w = torch.tensor([0.8, 1.2], device=dev)
    scaler = torch.amp.GradScaler("cuda")
    for tag in ("fp32", "fp32_strict", "autocast_fp16"):
        torch.backends.cudnn.allow_tf32 = tag != "fp32_strict"

        def step():
            opt.zero_grad(set_to_none=True)
            if tag != "autocast_fp16":
                loss = nn.functional.cross_entropy(m(x), y, weight=w)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                opt.step()
            else:
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = nn.functional.cross_entropy(m(x), y, weight=w)
                scaler.scale(loss).backward()
                scaler.unscale_(opt)
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                scaler.step(opt); scaler.update()
        ms = timed(step)
        out["train_B512_%s" % tag] = {"ms": ms, "windows_per_s": 512 / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
