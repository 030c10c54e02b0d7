"""TEST INFRASTRUCTURE ONLY - import the UNMODIFIED reference (TwoWL/utils.py,
TwoWL/operators/datasets.py, TwoWL/model/model.py) from /root/reference in the build
container, with oracle/pyg_shim standing in for torch_scatter / torch_geometric.

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or
bench.py calls this. It is used by oracle/gen_golden.py and by the CPU tests that pin
oracle/twowl_oracle.py to the reference (they skip when the reference is absent).
"""
import contextlib
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("TWOWL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "TwoWL", "utils.py"))


@contextlib.contextmanager
def _paths():
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k == "TwoWL" or k.startswith("TwoWL.") or k == "constant"}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIM)
    try:
        yield
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules
                  if k == "TwoWL" or k.startswith("TwoWL.") or k == "constant"]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


class _Ref:
    """Namespace holding the reference's modules, imported once."""
    utils = None
    datasets = None
    model = None


def load_reference():
    if _Ref.utils is not None:
        return _Ref
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    with _paths():
        _Ref.utils = importlib.import_module("TwoWL.utils")
        _Ref.datasets = importlib.import_module("TwoWL.operators.datasets")
        _Ref.model = importlib.import_module("TwoWL.model.model")
    return _Ref


@contextlib.contextmanager
def reference_cwd():
    """The reference reads raw_data/... relative to cwd (constant.py:7)."""
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)
