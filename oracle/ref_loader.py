"""Oracle harness: execute the UNMODIFIED class bodies of a reference script on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing under gnn_decode_b200/ imports this file; only tests/,
oracle/make_golden.py and (never at run time on the GPU box) developers do.  It needs
/root/reference, which exists in the build container only -- the GPU box uses the committed
fixtures in tests/golden/ plus oracle/restate.py instead.

How the reference is made to run:
  * oracle/shim/ supplies stand-ins for the absent third-party packages torch_geometric,
    torch_scatter (both unpinned by the reference) and matplotlib;
  * Tensor.cuda / Module.cuda become identity, torch.load is forced to map_location='cpu';
  * module-level CONSTANTS (L, BATCH_SIZE, run1, run2, P1, P2, H selection ...) are substituted
    textually so the import-time dataset generation stays small, and the one internal API drift
    (generate_PCM returns a tuple since error_generate.py:132 but decoder_v2_4.py:189,
    QGNNI.py:157, BP.py:165 still use the old single-array return) is patched with `[0]`;
  * the script is exec'd with __name__ != '__main__' from a scratch cwd that holds symlinks to
    the reference's checkpoint directories and `BCH(63,45).txt`.
The class bodies (MessagePassing / GraphConv / GatedGraphConv / GNNI / LossFunc) are untouched.
"""
import contextlib
import os
import random
import re
import sys
import tempfile
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("GNN_DECODE_REFERENCE", "/root/reference/GNN-decode")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "quantum"))


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


@contextlib.contextmanager
def _patched_env(subdir):
    """cwd with symlinks, shim on sys.path, .cuda() -> identity, torch.load -> cpu."""
    saved_path = list(sys.path)
    saved_cwd = os.getcwd()
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.load)
    saved_anomaly = torch.is_anomaly_enabled()
    saved_mods = {k: sys.modules.get(k) for k in
                  ("torch_geometric", "torch_geometric.data", "torch_geometric.utils",
                   "torch_scatter", "matplotlib", "matplotlib.pyplot", "error_generate")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    tmp = tempfile.mkdtemp(prefix="gd_oracle_")
    src_dir = os.path.join(REF_ROOT, subdir)
    for name in os.listdir(src_dir):
        if name.endswith(".py") or name == "__pycache__":
            continue
        os.symlink(os.path.join(src_dir, name), os.path.join(tmp, name))
    try:
        sys.path.insert(0, src_dir)      # for `import error_generate`
        sys.path.insert(0, _SHIM)
        os.chdir(tmp)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _orig_load = saved[2]
        torch.load = lambda f, *a, **k: _orig_load(f, map_location="cpu", weights_only=True)
        yield tmp
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.load = saved
        torch.autograd.set_detect_anomaly(saved_anomaly)
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def _sub_const(src, name, value):
    """Replace the first module-level assignment `name = ...` (single line)."""
    pat = re.compile(r"^%s\s*=.*$" % re.escape(name), re.M)
    assert pat.search(src), "constant %s not found" % name
    return pat.sub("%s = %s" % (name, value), src, count=1)


def load_reference(script, consts=None, seed=1234, raw_subs=(), truncate_at=None):
    """Exec reference `script` (e.g. 'quantum/decoder_v2_4.py') and return its namespace.

    consts   : {name: python-literal-string} module-level constant substitutions
    raw_subs : [(old, new)] literal text substitutions (must each match)
    truncate_at : literal text; the module source is cut there (scripts whose TAIL cannot run -- a training driver that
               needs another script's checkpoint -- keep their class bodies, which is all the oracle uses)
    The namespace stays usable after return ONLY inside `reference_session()`, because the
    class bodies call `.cuda()` at forward time.
    """
    assert reference_available(), "reference tree not mounted at %s" % REF_ROOT
    subdir = os.path.dirname(script)
    with open(os.path.join(REF_ROOT, script)) as f:
        src = f.read()
    # internal API drift patch (see module docstring)
    src = src.replace("generate_PCM(2 * L * L - 2, L))", "generate_PCM(2 * L * L - 2, L)[0])")
    for old, new in raw_subs:
        assert old in src, "raw substitution %r did not match" % old
        src = src.replace(old, new)
    if truncate_at is not None:
        assert truncate_at in src, "truncation marker %r not found" % truncate_at
        src = src[:src.index(truncate_at)]
    for k, v in (consts or {}).items():
        src = _sub_const(src, k, v)
    mod = types.ModuleType("ref_" + os.path.basename(script)[:-3])
    mod.__file__ = os.path.join(REF_ROOT, script)
    seed_all(seed)
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    torch.autograd.set_detect_anomaly(False)
    return mod


@contextlib.contextmanager
def reference_session(subdir):
    """Context in which reference modules can be loaded AND run (keeps .cuda() patched)."""
    with _patched_env(subdir) as tmp:
        yield tmp


def load_error_generate():
    """Import quantum/error_generate.py (needs only the matplotlib stub)."""
    with _patched_env("quantum"):
        import importlib
        eg = importlib.import_module("error_generate")
    return eg
