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


def test_host_stage_is_bit_exact_round_to_nearest_even():
    """bci_host_stage (pure host code of the library: pageable -> pinned staging of the drop-in callers): the fp32 copy is exact, the
    bf16 narrowing equals round-to-nearest-even on the bit pattern (what cvt.rn.bf16.f32 does on the device), for every alignment of
    the destination, thread counts that do not divide the length, ties, denormals, infinities and signed zeros."""
    from lstm_ode_bci_b200 import _native as N
    rng = np.random.default_rng(5)
    n = (1 << 20) + 37
    src = rng.standard_normal(n).astype(np.float32)
    src[:10] = np.array([0.0, -0.0, 1e-40, -1e-39, np.inf, -np.inf, 3.3895314e38, 1.00390625, 1.01171875, -1.00390625], np.float32)
    u = src.view(np.uint32).astype(np.uint64)
    want = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    assert want[7] == 0x3F80 and want[8] == 0x3F82            # the two ties go to the even neighbour
    for off in (0, 1, 5, 16):
        for threads in (1, 3, 8):
            d16 = np.zeros(n + off, np.uint16)
            N.check(N.lib().bci_host_stage(d16[off:].ctypes.data, src.ctypes.data, n, 1, threads))
            assert np.array_equal(d16[off:], want) and not d16[:off].any()
            d32 = np.zeros(n + off, np.float32)
            N.check(N.lib().bci_host_stage(d32[off:].ctypes.data, src.ctypes.data, n, 0, threads))
            assert np.array_equal(d32[off:].view(np.uint32), src.view(np.uint32)) and not d32[:off].any()
    assert N.lib().bci_host_stage(None, src.ctypes.data, 4, 1, 1) != 0
    N.check(N.lib().bci_host_stage(None, None, 0, 1, 4))


def test_permutation_importance_host_side_contract():
    """explain.compute_permutation_importance without a GPU: a CPU model is refused loudly (no CPU fallback), and the oracle's
    host-side gather -- what bci_permute_channels replaces -- is the reference's `X_permuted[:, :, ch] = X_subset[perm, :, ch]`."""
    import pytest
    from lstm_ode_bci_b200 import _native as N
    from lstm_ode_bci_b200 import explain
    from oracle import explain_oracle
    m = lstm.EnhancedLSTMModel(5, 128, 1, 2)
    X = np.random.default_rng(0).standard_normal((6, 8, 5)).astype(np.float32)
    y = np.zeros(6, dtype=np.int64)
    np.random.seed(0)
    with pytest.raises(N.BciError):
        explain.compute_permutation_importance(m, X, y, n_permutations=1, n_samples=4)
    perm = np.array([3, 0, 5, 1, 2, 4])
    Xp = explain_oracle.permuted_copy(X, perm, 2)
    assert np.array_equal(Xp[:, :, 2], X[perm][:, :, 2]) and np.array_equal(np.delete(Xp, 2, axis=2), np.delete(X, 2, axis=2))
    assert not np.shares_memory(Xp, X)
    # the oracle's importance of a classifier that looks at channel 2 only: every other channel scores exactly 0
    yy = (X[:, :, 2].sum(axis=1) > 0).astype(np.int64)
    np.random.seed(1)
    imp, base = explain_oracle.permutation_importance(lambda Z: (Z[:, :, 2].sum(axis=1) > 0).astype(np.int64), X, yy, 3, 100)
    assert base == 1.0 and np.all(np.delete(imp, 2) == 0.0) and imp[2] >= 0.0
