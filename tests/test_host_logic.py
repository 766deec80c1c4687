"""Host-side logic that needs no GPU: NUMA placement helpers, the precision the drop-in model picks under the reference's
autocast, engine ownership on copies, the view arithmetic of bci_lstm_input."""
import copy
import pickle

import numpy as np
import torch

from lstm_ode_bci_b200 import hostmem, lstm


def test_cpulist_and_page_query():
    assert hostmem.parse_cpulist("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    assert hostmem.parse_cpulist("") == [] and hostmem.parse_cpulist(None) == []
    nodes = hostmem.page_nodes(torch.zeros(1 << 18))
    assert nodes is None or (sum(nodes.values()) > 0 and all(n >= 0 for n in nodes))
    # without a GPU there is no node to bind to: the call reports that and changes nothing
    info = hostmem.bind_to_gpu_node(0)
    assert set(info) == {"node", "cpus", "affinity", "mempolicy"}


def test_copies_do_not_share_engine_handles():
    m = lstm.EnhancedLSTMModel(61, 128, 3, 2)
    m._engines[("fp32", 0)] = 12345            # pretend an engine exists (no GPU here)
    m._loaded[("fp32", 0)] = m._signature()
    c = copy.deepcopy(m)
    assert c._engines == {} and c._loaded == {} and m._engines == {("fp32", 0): 12345}
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(m.parameters(), c.parameters()))
    p = pickle.loads(pickle.dumps(m))
    assert p._engines == {} and p._loaded == {}
    m._engines.clear()                         # nothing real to destroy
    sig = m._signature()
    m.mark_weights_changed()
    assert m._signature() != sig


def test_view_offsets_enumerate_create_sequences_order():
    """bci_lstm_input's offset rule == the order create_sequences (02:157-180) emits windows, recording after recording."""
    R, S, C, T, step = 3, 1000, 5, 256, 128
    n_seq = (S - T) // step + 1
    rec = np.arange(R * S * C, dtype=np.int64).reshape(R, S, C)
    want = np.stack([rec[r, i * step:i * step + T] for r in range(R) for i in range(n_seq)])
    flat = rec.reshape(-1)
    for w in (0, 1, n_seq - 1, n_seq, 2 * n_seq + 3):
        off = (w // n_seq) * (S * C) + (w % n_seq) * (step * C)
        assert np.array_equal(flat[off:off + T * C].reshape(T, C), want[w])
