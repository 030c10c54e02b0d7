"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/twowl.h declares; host-only entry points answer; the product path refuses CPU tensors (no
fallback); the drop-in module keeps the reference's signatures and state_dict keys."""
import inspect
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import state_dict_from
from oracle import twowl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge._load_build_module().build()          # nvcc cross-compiles sm_100a without a GPU
    import twowl_b200._lib as lib_mod
    return lib_mod


def test_library_exports_every_declared_symbol(L):
    protos = L.parse_header()
    assert len(protos) >= 40
    for name in protos:
        assert hasattr(L.lib, name), f"{name} declared in include/twowl.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T twowl_" in ln}
    assert exported == set(protos), f"undeclared exports: {exported - set(protos)}; missing: {set(protos) - exported}"


def test_host_only_entry_points(L):
    assert L.lib.twowl_version() == 100
    assert isinstance(L.lib.twowl_last_error(), bytes)
    small = L.lib.twowl_csr_build_workspace_bytes(1000, 100)
    big = L.lib.twowl_csr_build_workspace_bytes(10_000_000, 1_000_000)
    assert 0 < small < big
    assert L.lib.twowl_graphnorm_stats_workspace_bytes(10, 64) >= 148 * 4 * 2 * 64 * 8
    assert L.lib.twowl_linear_bwd_weight_workspace_bytes(1 << 20, 64, 64) >= 64 * 64 * 4
    assert L.lib.twowl_select_workspace_bytes(0) > 0 and L.lib.twowl_ei2_count_workspace_bytes(0) > 0


def test_argument_errors_do_not_touch_the_device(L):
    """TWOWL_EINVAL paths return before any CUDA call: checkable without a GPU."""
    rc = L.lib.twowl_linear_fwd(None, None, 10, 6, 8, None, 0, None)      # Ci % 4 != 0
    assert rc == -22 and b"multiples of 4" in L.lib.twowl_last_error()
    rc = L.lib.twowl_graphnorm_stats(None, 10, 64, None, 1e-5, None, None, 0, None)   # short workspace
    assert rc == -28 and b"workspace" in L.lib.twowl_last_error()
    with pytest.raises(RuntimeError, match="code -22"):
        L.check(L.lib.twowl_readout_fwd(None, None, 1, 5, 30, None, None, None, None), "readout_fwd")


def test_no_cpu_fallback(L):
    import TwoWL.utils as U
    ei = torch.tensor([[0, 1], [1, 0]])
    for call in (lambda: U.degree(ei, 2), lambda: U.get_ei2(2, ei, ei), lambda: U.reverse(ei),
                 lambda: U.double(ei), lambda: U.sample_block(torch.tensor([0]), 2, ei, None)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_signatures_match_the_reference(L):
    import TwoWL.model.model as model
    import TwoWL.utils as U
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(U.degree) == ["ei", "num_node"] and sig(U.set_mul) == ["a", "b"]
    assert sig(U.check_in_set) == ["target", "set"] and sig(U.get_ei2) == ["n_node", "pos_edge", "pred_edge"]
    assert sig(U.blockei2) == ["ei2", "blocked_idx"] and sig(U.idx2mask) == ["num", "idx"]
    assert sig(U.sample_block) == ["sample_idx", "size", "ei", "ei2"] and sig(U.reverse) == ["edge_index"]
    assert sig(U.double) == ["x", "for_index"]
    assert sig(U.random_split_edges) == ["data", "val_ratio", "test_ratio"]
    assert sig(model.LocalWLNet.forward) == ["self", "x", "edge1", "pos", "idx", "ei2", "test"]
    init = inspect.signature(model.LocalWLNet.__init__).parameters
    assert list(init)[1:] == ["max_x", "use_node_feat", "node_feat", "channels_1wl", "channels_2wl", "depth1", "depth2",
                              "dp_lin0", "dp_lin1", "dp_emb", "dp_1wl0", "dp_2wl", "dp_1wl1", "act0", "act1"]
    assert init["channels_1wl"].default == 256 and init["channels_2wl"].default == 32 and init["dp_lin0"].default == 0.7


def test_state_dict_keys_match_the_reference(L, fb):
    import TwoWL.model.model as model
    sd = state_dict_from(fb)                                  # keys as the reference's module produced them
    mod = model.LocalWLNet(sd["emb.0.weight"].shape[0] - 1, False, None, 64, 24, 2, 2)
    assert set(mod.state_dict().keys()) == set(sd.keys())
    mod.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    sd2 = O.init_state_dict(9, 32, 16, 1, 3)
    mod2 = model.LocalWLNet(9, False, None, 32, 16, 1, 3)
    assert set(mod2.state_dict().keys()) == set(sd2.keys())
    # default init of the drop-in follows the reference's initialisers (glorot for lin, zeros bias, 1/0/1 norm)
    w = mod2.conv2s[0].modlist[0].lin.weight
    assert float(w.abs().max()) <= np.sqrt(6.0 / 32) + 1e-6
    assert float(mod2.conv2s[0].modlist[0].bias.abs().max()) == 0.0
    assert torch.equal(mod2.emb[1].mean_scale.detach(), torch.ones(32))


def test_seed_provider_and_tagged_seed_addresses():
    """Host-side plumbing of the dropout seeds: ops.next_seed() draws from torch's CPU generator (reproducible, below 2^62) unless a
    provider is installed; a device-resident seed travels as its address with bit 63 set, written as the signed 64-bit integer a torch
    custom op accepts, and ctypes hands the same 64 bits to the uint64_t argument of the C ABI."""
    import ctypes
    import torch
    from twowl_b200 import ops
    torch.manual_seed(3)
    a = [ops.next_seed() for _ in range(4)]
    torch.manual_seed(3)
    assert a == [ops.next_seed() for _ in range(4)] and all(0 <= v < 2 ** 62 for v in a)
    it = iter([11, 22])
    prev = ops.set_seed_provider(lambda: next(it))
    try:
        assert prev is None and ops.next_seed() == 11 and ops.next_seed() == 22
    finally:
        assert ops.set_seed_provider(prev) is not None
    assert ops.next_seed() not in (11, 22) or True
    addr = 0x7F12_3456_7000
    tagged = addr - (1 << 63)                                  # what graphed.DeviceSeeds.provider returns
    assert -(1 << 63) <= tagged < 0                            # fits torch's int64 schema
    as_u64 = ctypes.c_uint64(tagged).value
    assert as_u64 >> 63 == 1 and as_u64 & 0x7FFF_FFFF_FFFF_FFFF == addr
