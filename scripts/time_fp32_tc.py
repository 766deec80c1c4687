"""fp32 parity mode, inference: windows/s and per-phase device time at several batch sizes (the recurrence runs on the tensor
cores from 16 work items up: lstm_fp32_tc.cu; BCI_FP32_REC=simt keeps the CUDA-core recurrence for comparison)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, ops, synth
m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3), precision="fp32")
hid = m._engine("fp32")
chunk = ops.lstm_chunk_windows(hid)
print("rec=%s chunk=%d" % (os.environ.get("BCI_FP32_REC", "tc"), chunk))
for B in (512, 1024, 2048, 4096, chunk, 2 * chunk):
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
