"""CPU suite, part 3: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls: there is no GPU here), and the host mirror keeps the reference's interface."""
import glob
import inspect
import os
import re

import pytest
import torch

from conftest import ROOT, Golden, golden_cases, make_decoder
from gnn_decode_b200 import _cabi


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(gd_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_header_declares_functions():
    assert {"gd_graph_create", "gd_decode_fwd", "gd_propagate_fwd", "gd_decode_host", "gd_sample"} <= set(_declared())


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    for name in _declared():
        assert hasattr(lib, name), "%s declared in include/ but not exported" % name
    assert lib.gd_abi_version() == _cabi.ABI_VERSION


def test_every_declared_symbol_is_bound():
    bound = set(_cabi._SIGNATURES) | set(_cabi._OPTIONAL)
    assert set(_declared()) <= bound


def test_weights_size():
    lib = _cabi.lib()
    import ctypes as C
    assert lib.gd_weights_size(C.byref(_cabi.GdModel(_cabi.PROG_V2_4, 128, 15, 0))) == 1283
    assert lib.gd_weights_size(C.byref(_cabi.GdModel(_cabi.PROG_CGNNI, 10, 25, 0))) == 62
    assert lib.gd_weights_size(C.byref(_cabi.GdModel(_cabi.PROG_BP_QUANTUM, 0, 10, 0))) == 0
    assert lib.gd_weights_size(C.byref(_cabi.GdModel(99, 1, 1, 0))) == -1
    assert b"invalid model" in lib.gd_last_error()


@pytest.mark.parametrize("name", golden_cases())
def test_state_dict_keys_match_reference_checkpoints(name):
    g = Golden(name)
    mod, dec = make_decoder(g)                     # load_state_dict(strict=True) inside
    assert list(dec.state_dict().keys()) == list(g.weights.keys())
    n = sum(p.numel() for p in dec._gd_params())
    import ctypes as C
    assert n == _cabi.lib().gd_weights_size(C.byref(dec.gd_model()))


def test_signatures_match_reference():
    from gnn_decode_b200.quantum import decoder_v2_4 as q
    from gnn_decode_b200.classical import CGNNI as c
    assert list(inspect.signature(q.MessagePassing.__init__).parameters) == ["self", "aggr", "flow"]
    assert list(inspect.signature(q.MessagePassing.propagate).parameters) == ["self", "edge_index", "extra", "size", "kwargs"]
    assert list(inspect.signature(c.MessagePassing.propagate).parameters) == ["self", "edge_index", "post", "size", "kwargs"]
    assert list(inspect.signature(q.GraphConv.__init__).parameters) == ["self", "flow", "aggr", "bias"]
    assert list(inspect.signature(q.GraphConv.forward).parameters) == ["self", "m", "edge_index", "x", "prev"]
    assert list(inspect.signature(c.GatedGraphConv.forward).parameters) == ["self", "m", "edge_index", "x"]
    assert list(inspect.signature(q.GNNI.forward).parameters) == ["self", "data"]
    with pytest.raises(AssertionError):
        q.MessagePassing(aggr="sum")
    with pytest.raises(AssertionError):
        q.MessagePassing(flow="sideways")


def test_no_cpu_fallback():
    """CPU tensors must raise, not silently run somewhere else."""
    from gnn_decode_b200.quantum import decoder_v2_4 as q
    g = Golden("v2_4_toricL4_epoch1")
    dec = q.GNNI(2)

    class D(object):
        x = g.x.reshape(-1, 1)
        edge_index = g.batched_edge_index()
    with pytest.raises(_cabi.GdError):
        dec(D())
    conv = q.GraphConv("source_to_target")
    with pytest.raises(_cabi.GdError):
        conv(torch.zeros(g.B * g.E, 1), D.edge_index, D.x)


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import gnn_decode_b200, gnn_decode_b200.quantum.decoder_v2_4, "
            "gnn_decode_b200.classical.CGNNI, gnn_decode_b200.codes; "
            "bad=[m for m in sys.modules if m.split('.')[0]=='oracle']; assert not bad, bad")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for f in glob.glob(os.path.join(ROOT, "gnn_decode_b200", "**", "*.py"), recursive=True):
        assert "oracle" not in re.sub(r"#.*", "", open(f).read()).replace("no CPU fallback", ""), f


def test_p2p_allreduce_argument_checks():
    import ctypes as C
    lib = _cabi.lib()
    assert lib.gd_p2p_buffer_floats(1283) == 2 * 1312 + 32 and lib.gd_p2p_buffer_floats(0) == -1
    ptrs = (C.c_uint64 * 2)(0, 0)
    assert lib.gd_p2p_allreduce(ptrs, 2, 5, None, None, 10, 1, C.c_float(1.0), None, None) == _cabi.GD_ERR_INVALID
    assert b"gd_p2p_allreduce" in lib.gd_last_error()
