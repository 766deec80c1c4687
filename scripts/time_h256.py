"""bf16 mode at hidden size 256 (the reference's checkpoint size): windows/s and per-phase device time per chunk
(BCI_GEMM_PAIR=off keeps the one-CTA projection GEMM for comparison)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, ops, synth
m = lstm.from_params(synth.make_lstm_params(44, 61, 256, 3), precision="bf16")
hid = m._engine("bf16")
chunk = ops.lstm_chunk_windows(hid)
print("pair=%s chunk=%d" % (os.environ.get("BCI_GEMM_PAIR", "on"), chunk))
for B in (1024, chunk, 2 * chunk):
    x = torch.randn(B, 256, 61, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            m.predict_proba(x)
        ops.lstm_set_profiling(hid, True); ops.lstm_get_profile(hid)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(3):
            m.predict_proba(x)
        b.record(); torch.cuda.synchronize()
    prof = ops.lstm_get_profile(hid); ops.lstm_set_profiling(hid, False)
    ms = a.elapsed_time(b) / 3
    print("B=%6d: %8.2f ms = %8.0f windows/s | %s" % (B, ms, B / ms * 1e3, "  ".join("%s %.2f" % (k, v[0] / 3) for k, v in prof.items())))
    del x
