"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/cfrb200.h declares, its
host-only tree enumeration is bit-exact with the oracle / the reference's golden trees, it refuses to run without a GPU,
and the product never reaches into oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.oracle import game_dims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cfrb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfrb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rebel_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cfrb200.h but not exported by libcfrb200.so"


def test_host_tree_enumeration_bit_exact(port, golden):
    from rebel_b200 import capi
    g = golden("trees.npz")
    for key in g.files:
        _, D, F, lb, pl, md = key.split("_")
        t = capi.unroll_tree(int(D), int(F), int(lb), int(pl), int(md))
        assert t.shape == g[key].shape and (t == g[key]).all(), key
    for (D, F) in [(1, 4), (1, 6), (2, 3), (1, 5), (2, 2)]:
        A, H, Q = game_dims(D, F)
        for lb in range(-1, A - 1):
            for md in (0, 1, 2, 3):
                assert (capi.unroll_tree(D, F, lb, lb & 1, md) == port.unroll_tree(D, F, lb, lb & 1, md)).all()
    assert len(capi.unroll_tree(1, 2)) == 31          # tree_test.cc:27


def test_no_cpu_fallback():
    from rebel_b200 import capi
    if capi.lib().cfrb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.CfrbError, match="no CUDA device"):
        capi.WaveSolver(1, 4, 8)


def test_product_does_not_touch_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "rebel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|libcfr_oracle|libref_", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_net2_mirror_state_dict_layout():
    import torch
    from rebel_b200.models import FLAT_ORDER, Net2, flatten_state_dict, input_size, make_selfplay_net
    net = make_selfplay_net(1, 6)
    sd = net.state_dict()
    assert tuple(sd.keys()) == FLAT_ORDER          # same names/order as the reference's Net2 (models.py:64-94)
    A, H, Q = game_dims(1, 6)
    assert input_size(6, 1) == Q
    flat = flatten_state_dict(sd)
    assert flat.size == 256 * Q + 3 * 256 + 256 * 256 + 3 * 256 + H * 256 + H == 75526
    x = torch.rand(5, Q)
    assert net(x).shape == (5, H)
    assert torch.jit.script(Net2(num_faces=6, num_dice=1, n_layers=2, use_layer_norm=True))(x).shape == (5, H)
