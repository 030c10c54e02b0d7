"""One rank of the row-sharded TwoWL step (launched by tests/test_gpu_rowshard.py, not collected by pytest).

usage: rowshard_worker.py <rank> <world> <port> <out.npz> <channels_2wl> <seed> <depth2>
Every rank rebuilds the same fb-pages-food seed-0 train step from tests/golden (train.py:16-38), runs forward + BCE +
backward on ITS block of pair rows (LocalWLNet.row_shard), sums the parameter gradients over the ranks, and rank 0 writes
logits / loss / gradients. The ranks share cuda:0 and talk through gloo (NCCL refuses two ranks on one device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "link-prediction-gnn_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CFG = dict(channels_1wl=64, depth1=2, dp_lin0=0., dp_lin1=0., dp_emb=0., dp_1wl0=0., dp_2wl=0., dp_1wl1=0.,
           act0=True, act1=True)


def build_step(c2: int, seed: int, depth2: int = 1):
    """(module, args of forward, y) - shared with the single-GPU side of the test."""
    import TwoWL.model.model as model
    import TwoWL.utils as U
    from helpers import fb_split
    g = np.load(os.path.join(ROOT, "tests", "golden", "fb_pages_food_seed0.npz"), allow_pickle=False)
    fb = {k: g[k] for k in g.files}
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    n = int(fb["num_nodes"][0])
    ei, pred, pos1 = fb_split(fb, 0)
    dei, dpred, dpos = dev(ei), dev(pred), dev(pos1)
    ei2 = U.get_ei2(n, dei, dpred)
    idx1, idx2 = dev(fb["idx1"].astype(np.int64)), dev(fb["idx2"].astype(np.int64))
    ei_new, x_new, ei2_new = U.sample_block(idx1, n, dei, ei2)
    pos2 = torch.cat((idx1, idx2))
    bs = idx1.numel() // 2
    y = torch.cat((torch.ones(bs), torch.zeros(bs))).unsqueeze(-1).cuda()
    torch.manual_seed(seed)
    mod = model.LocalWLNet(int(x_new.max().item()), False, None, channels_2wl=c2, depth2=depth2, **CFG)
    # GraphNorm / bias parameters away from their (1, 0, 1) initial values so that every gradient path is exercised
    with torch.no_grad():
        for k, p in mod.named_parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    mod = mod.cuda().train()
    return mod, (x_new, ei_new, dpos, pos2, ei2_new), y


def main():
    rank, world, port, out, c2, seed, depth2 = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]),
                                                int(sys.argv[6]), int(sys.argv[7]))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from twowl_b200 import dist as D
    from twowl_b200.rowshard import RowShard
    mod, args, y = build_step(c2, seed, depth2)
    mod.row_shard = RowShard()
    logits = mod(*args)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    D.allreduce_grads(mod.parameters())
    if rank == 0:
        res = {"logits": logits.detach().cpu().numpy(), "loss": loss.detach().cpu().numpy()}
        for k, p in mod.named_parameters():
            res["grad/" + k] = p.grad.cpu().numpy()
        np.savez(out, **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
