"""BASELINE configs[3] at FULL size (R-MAT 1 M nodes / 16 M edge samples -> E = 30 M observed edge rows, R = 60 M pair rows,
3 M target links per step) at the BENCHMARKED width (hidden 64) and at 32, plus configs[2]'s graph at the widths of the
configs[4] sweep: the CUDA path against an independent float64 evaluation of the whole train step.

The CPU oracle cannot run here (the explicit wedge index has 6.5e10 columns), so the reference is tests/ref64.py: a float64
evaluation of LocalWLNet.forward + BCE + backward written with plain torch ops and torch autograd, chunked over the pair rows,
which tests/test_ref64_cpu.py pins to the oracle (explicit index, plain autograd) at sizes the oracle can run. Checked:
  * logits, loss and EVERY element of EVERY parameter gradient against float64 (tests/helpers.parity: north_star band
    |a-b| <= 1e-6 + 1e-5*|b| against float64 or against the same program evaluated by torch in fp32; else within 4 x that fp32
    evaluation's own worst error on the tensor; a gradient element may instead be within 3e-5 (5e-5 on the collab graph) x the
    largest magnitude of its tensor - a weight gradient here is a sum over up to 6e7 rows whose terms cancel, so its error
    scales with the terms, not with the result; the ledger reports how many elements needed which clause),
  * bit-identical logits and gradients on a second run (no atomics on data anywhere).
"""
import gc
import os
import sys

import pytest
import torch

import ref64
from helpers import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,hidden", [("rmat", 64), ("rmat", 32), ("collab", 64), ("collab", 128), ("collab", 256)])
def test_full_size_step_vs_fp64(workload, hidden):
    """("rmat", 64): the configuration bench.py reports, at full size. ("collab", 64 / 128 / 256): configs[2]'s graph (R = 5 M
    pair rows) at the widths of the configs[4] sweep - the column-window / multi-launch paths of the wide layers."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 150 * 2 ** 30:
        pytest.skip("needs a 180 GB device")
    sys.path.insert(0, ROOT)
    import bench
    import TwoWL.model.model as model
    import TwoWL.utils as U
    from twowl_b200 import graph as G
    G.clear_cache()
    gc.collect()
    torch.cuda.empty_cache()
    dev = torch.device("cuda", 0)
    g = bench.make_graph(workload, 0, dev)
    n, pos, pred, pos1 = g["n"], g["pos"], g["pred"], g["pos1"]
    E, P = pos.shape[1], pred.shape[1]
    if workload == "rmat":
        assert n == 1 << 20 and E > 29_000_000 and E + P > 59_000_000       # the BASELINE configs[3] sizes
    else:
        assert n == 1 << 18 and E + P > 4_900_000                           # configs[2]: ogbl-collab scale
    ei2 = U.get_ei2_implicit(n, pos, pred)
    nb = g["und"] // 10
    i1, i2, y = (t.to(dev) for t in bench.draw_batch(g["und"], P // 2, nb, 0))
    idx1 = U.double(i1, for_index=True)
    idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
    ei_new, x_new, ei2_new = U.sample_block(idx1, n, pos, ei2)

    torch.manual_seed(3)
    mod = model.LocalWLNet(int(U.degree(pos, n).max().item()), False, None, hidden, hidden, 1, 1, 0., 0., 0., 0., 0., 0.)
    with torch.no_grad():
        for p in mod.parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    mod = mod.to(dev).train()

    def ours():
        for p in mod.parameters():
            p.grad = None
        out = mod(x_new, ei_new, pos1, idx, ei2_new)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return out.detach().clone(), loss.detach().clone(), {k: p.grad.detach().clone() for k, p in mod.named_parameters()}

    out, loss, grads = ours()
    out2, loss2, grads2 = ours()
    assert torch.equal(out, out2) and torch.equal(loss, loss2), "forward is not deterministic"
    for k in grads:
        assert torch.equal(grads[k], grads2[k]), f"gradient {k} is not deterministic"
    del out2, grads2

    blocked = torch.zeros(E, dtype=torch.bool, device=dev)
    blocked[idx1] = True
    assert ei_new.shape[1] == E - idx1.numel()
    sd = {k: v.detach() for k, v in mod.state_dict().items()}
    ei_plain = ei_new.clone()          # no tags: the reference sees a plain tensor
    del ei2_new, ei2, ei_new
    G.clear_cache()                    # the cached CSRs / regrouped tables of the product path: the reference needs the room
    gc.collect()
    torch.cuda.empty_cache()

    lg64, l64, g64 = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y)
    # the same program in fp32: what plain torch fp32 arithmetic - the reference's kind - makes of these formulas at this size.
    # It is the `ref32` of helpers.parity: an element outside the band of both evaluations must be no farther from float64
    # than this fp32 evaluation's own worst element of the tensor.
    lg32, l32, g32 = ref64.step(sd, x_new, ei_plain, pos1, idx, E, blocked, y, dtype=torch.float32)
    tag = f"{workload}/hidden{hidden} "
    # Logits: at most 3 % of them may need a relaxation clause (observed at rmat / hidden 64: 0.55 % of the 3.0 M logits lie
    # outside the float64 band - the torch-fp32 evaluation of the same program leaves 0.54-0.63 % outside; before the column sums
    # of GraphNorm were folded into double every 8 rows it was 10.6 %, tools/diag_stages.py), and every one of those within
    # 1e-5 x the largest logit (~1e-3 absolute at |logit| up to 98; observed worst 1.5e-4, torch-fp32's own 1.6-2.3e-4 - its
    # atomics move its worst element from run to run, so "4 x its error" alone is not a stable yardstick).
    # (the wide models on the collab graph are worse conditioned in ANY fp32 evaluation: 7.5 % of the hidden-256 logits are outside
    #  both bands, all of them within 2 x the torch-fp32 evaluation's own worst error - 15 % allowed there, as in round 1)
    parity(out, lg32, lg64, tag + "logits", scale_floor=1e-5, allow_relaxed=out.numel() * (3 if workload == "rmat" else 15) // 100)
    parity(loss, l32, l64, tag + "loss")
    for k in sorted(grads):
        # gradient floor 3e-5 x the tensor's largest magnitude: a weight / GraphNorm gradient here is a sum over up to 6e7 rows of
        # products of fp32 activations that each carry ~1e-6 of relative error, with ~10x cancellation between the terms
        # (observed worst over all tensors: 2.0e-5 x max; the torch-fp32 evaluation's own worst: 1.7e-4 x max on emb.0.weight)
        # No limit on HOW MANY elements of a gradient tensor take the floor (the ledger reports it): for a 64..256-element
        # GraphNorm gradient whose terms cancel, a third of the elements sit outside the 1e-5 band of float64 in every fp32
        # evaluation, and which of them the torch-fp32 run happens to land next to changes with the order of its atomics -
        # a count limit made this test pass or fail from run to run. Every element must still be within the band or the floor.
        # Floor: 3e-5 at the benchmarked graph (observed worst 2.0e-5); 5e-5 on the collab graph, whose worst is the conv2s_r bias
        # gradient at hidden 128 (3.5e-5): the fused backward gets it in closed form, P * colsum(O) + M * Q + colsum(G), three
        # large terms that cancel.
        parity(grads[k], g32[k], g64[k], tag + "grad " + k, scale_floor=3e-5 if workload == "rmat" else 5e-5,
               allow_relaxed=grads[k].numel())
