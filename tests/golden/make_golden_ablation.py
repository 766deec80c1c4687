"""Golden vectors of the reference's AblationLSTMModel (09_sensitivity_analysis.py:176-240) -- see make_golden_next.py."""
import os

import numpy as np
import torch

from oracle import ref_loader
from lstm_ode_bci_b200 import synth

OUT = os.path.dirname(os.path.abspath(__file__))

# (tag, H, layers, bidirectional, use_attention, use_layer_norm): the six configurations of run_architecture_ablation
# (09:342-349, at a reduced hidden size for the fixture; the H = 256 'Minimal' one is included as is) + LayerNorm off
CASES = [
    ("full", 128, 3, True, True, True),
    ("noattn", 128, 3, True, False, True),
    ("unidir", 128, 3, False, True, True),
    ("layers1", 128, 1, True, True, True),
    ("layers2", 128, 2, True, True, True),
    ("minimal256", 256, 1, False, False, True),
    ("unidir256", 256, 2, False, True, True),
    ("noln", 128, 2, True, True, False),
    ("noln_unidir_mean", 128, 2, False, False, False),
]


def ablation_cases():
    ref09 = ref_loader.load("ref09")
    out = {}
    B, T, C = 5, 48, 61
    for i, (tag, H, L, bidir, att, ln) in enumerate(CASES):
        params = synth.make_lstm_params(60 + i, C, H, L, bidirectional=bidir, logit_gain=4.0, use_attention=att, use_layer_norm=ln)
        x = synth.make_windows(70 + i, B, T, C)
        y = (np.arange(B) % 2).astype(np.int64)
        torch.manual_seed(0)
        m = ref09.AblationLSTMModel(input_size=C, hidden_size=H, num_layers=L, num_classes=2, dropout=0.0,
                                    bidirectional=bidir, use_attention=att, use_layer_norm=ln).to("cpu")
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()}, strict=True)
        m.train()   # dropout = 0: train mode only selects the differentiable CPU LSTM path
        xt = torch.from_numpy(x).requires_grad_(True)
        logits = m(xt)
        loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(y))   # plain CE as in 09:277,300
        loss.backward()
        out[tag + ":cfg"] = np.array([60 + i, 70 + i, H, L, int(bidir), int(att), int(ln), B, T, C])
        out[tag + ":logits"] = logits.detach().numpy()
        out[tag + ":loss"] = np.float64(loss.item())
        out[tag + ":dx_norm"] = np.float64(xt.grad.norm().item())
        for k, p in m.named_parameters():
            g = p.grad.detach().numpy()
            out[tag + ":gnorm:" + k] = np.float64(np.linalg.norm(g.astype(np.float64)))
            out[tag + ":ghead:" + k] = g.reshape(-1)[:8].copy()
        print("ablation", tag, "loss", float(loss), "logits[0]", logits[0].detach().numpy())
    np.savez_compressed(os.path.join(OUT, "ablation_ref09.npz"), **out)
