"""Row-block sharding of the pair table ("G-rank logits/grads == 1-rank within the same tolerance").

World sizes 2, 3 and 8 (uneven blocks; blocks that hold only prediction pairs and no observed edge; node blocks of 78 nodes) of
one fb-pages-food train step (depth1 = 2: two node layers cut into node blocks; depth2 = 1 or 2) run as separate processes on
cuda:0 over gloo; rank 0's logits, loss and rank-summed parameter gradients are compared
with (a) the CPU oracle in fp32 / fp64 and (b) this package's single-GPU path, at the north_star tolerance.

The weight seeds are chosen so that no pre-ReLU value of the pair layer at a selected row lies within 1e-5 of zero (checked
with the fp64 oracle): with ~50 k such values per step one of them sits within fp32 rounding of the ReLU kink for roughly every
tenth seed, and then ANY fp32 evaluation order flips that unit's mask against exact arithmetic (seed 1234, C = 32: one value at
1.3e-7 moved every conv2s gradient by 1e-3 relative, on one GPU and sharded alike)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import twowl_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(world, c2, seed, out, depth2=1):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "rowshard_worker.py"), str(r), str(world), str(port), out, str(c2),
                               str(seed), str(depth2)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)[-4000:]
    return dict(np.load(out))


@pytest.mark.parametrize("world,c2,seed,depth2", [(2, 32, 7, 1), (3, 64, 5, 1), (2, 32, 4, 2), (8, 32, 7, 1)])
def test_row_sharded_step_matches_single_gpu_and_oracle(tmp_path, world, c2, seed, depth2):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, HERE)
    from rowshard_worker import CFG, build_step
    from test_gpu_model import close
    got = _run_world(world, c2, seed, str(tmp_path / "out.npz"), depth2)

    mod, args, y = build_step(c2, seed, depth2)
    out = mod(*args)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    x, e1, pos, idx, ei2 = args
    acts = (CFG["act0"], CFG["act1"])
    ei2_np = ei2.materialize().cpu().numpy() if hasattr(ei2, "materialize") else ei2.cpu().numpy()
    ref = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ref[name] = O.fwd_bwd({k: v.to(dt) for k, v in sd.items()}, x.cpu(), e1.cpu(), pos.cpu(), idx.cpu(), ei2_np, y.cpu().to(dt), *acts)
    close(torch.from_numpy(got["logits"]), ref["f32"][0], ref["f64"][0], f"world {world} logits")
    close(torch.from_numpy(got["loss"]), ref["f32"][1], ref["f64"][1], f"world {world} loss")
    for k, p in mod.named_parameters():
        close(torch.from_numpy(got["grad/" + k]), ref["f32"][2][k], ref["f64"][2][k], f"world {world} grad {k}", grad=True)
        # and against this package's own single-GPU evaluation of the same step
        one = p.grad.detach().cpu().double()
        tol = 1e-6 + 2e-5 * float(one.abs().max())
        assert float((torch.from_numpy(got["grad/" + k]).double() - one).abs().max()) <= tol, f"world {world} grad {k} vs 1 GPU"
    assert float((torch.from_numpy(got["logits"]).double() - out.detach().cpu().double()).abs().max()) <= 2e-5
