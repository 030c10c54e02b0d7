"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in twowl_b200/dist.py: disjoint target-link slices, the
flat gradient all-reduce, max-over-ranks timing, even/balanced row blocks, Chan merge of GraphNorm column statistics."""
import importlib.util
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dist():
    # twowl_b200/__init__ pulls the CUDA library in; dist.py itself is pure host logic, so load it by path
    spec = importlib.util.spec_from_file_location("twowl_dist", os.path.join(ROOT, "link-prediction-gnn_b200", "twowl_b200", "dist.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = _load_dist()
    try:
        assert D.world() == (rank, world)
        # 1. disjoint slices of one global batch
        mine = D.shard_batch(1000, 37, step=5, seed=3)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        allids = torch.cat(gathered)
        assert allids.unique().numel() == world * 37
        assert torch.equal(mine, D.shard_batch(1000, 37, step=5, seed=3))           # reproducible
        assert not torch.equal(mine, D.shard_batch(1000, 37, step=6, seed=3))       # a new draw per step
        # 2. gradient all-reduce: same model, rank-dependent gradients; one parameter without a gradient on rank 1
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        extra = torch.nn.Parameter(torch.zeros(4))
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.arange(3.0) * (rank + 1)
        if rank == 0:
            extra.grad = torch.ones(4)
        nred = D.allreduce_grads(list(lin.parameters()) + [extra])
        assert nred == 15 + 3 + 4
        tot = sum(range(1, world + 1))
        assert torch.equal(lin.weight.grad, torch.full_like(lin.weight, float(tot)))
        assert torch.equal(lin.bias.grad, torch.arange(3.0) * tot)
        assert torch.equal(extra.grad, torch.ones(4))
        # 3. slowest rank's time
        assert D.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        # 4. per-rank column statistics merged with Chan's formula == statistics of the concatenation
        torch.manual_seed(7)
        full = torch.randn(101, 6, dtype=torch.float64) * 3 + 5
        blocks = D.row_blocks([1] * 100, world)
        lo, hi = blocks[rank]
        part = full[lo:hi] if rank < world - 1 else full[lo:]
        stat = torch.cat((torch.tensor([float(part.shape[0])], dtype=torch.float64), part.mean(0), ((part - part.mean(0)) ** 2).sum(0)))
        stats = [torch.empty_like(stat) for _ in range(world)]
        dist.all_gather(stats, stat)
        S = torch.stack(stats)
        N, mu, M2 = D.merge_column_stats(S[:, 0], S[:, 1:7], S[:, 7:])
        assert N == 101
        assert torch.allclose(mu, full.mean(0), rtol=1e-12, atol=1e-12)
        assert torch.allclose(M2 / N, full.var(0, unbiased=False), rtol=1e-12, atol=1e-12)
        # 5. row-sharded readout plumbing (twowl_b200.rowshard): every rank contributes the logits of the target links inside
        #    its row block, ends with the full [L,1] tensor, and gets back the gradient of its own links only; the complete
        #    GraphNorm / readout parameter gradients stay on rank 0 so that the caller's SUM over ranks is exact
        sys.path.insert(0, os.path.join(ROOT, "link-prediction-gnn_b200"))
        from twowl_b200 import rowshard as RS
        shard = RS.RowShard()
        assert (shard.rank, shard.world) == (rank, world)
        L = 11
        # the readout rows of a batch: link l selects the mates (2*r_l, 2*r_l + 1); a rank keeps the links whose rows fall into
        # its block as block-local ids and masks the others with -1 (no compaction, hence no device->host size read)
        # (a block = the rank's share of the observed rows [0, E) followed by its share of the prediction rows [E, R))
        E, R = 24, 40
        ranges = RS.blocks_of(E, R, rank, world)
        (olo, ohi), (plo, phi) = ranges
        rows_of_link = torch.tensor([3, 17, 0, 19, 8, 12, 5, 11, 15, 2, 9])
        idx = torch.stack((2 * rows_of_link, 2 * rows_of_link + 1), dim=1).reshape(-1)
        idx_l = RS.mask_links(idx, ranges)
        in_o = (2 * rows_of_link >= olo) & (2 * rows_of_link < ohi)
        in_p = (2 * rows_of_link >= plo) & (2 * rows_of_link < phi)
        mine = in_o | in_p
        # block-local ids = positions in take_ranges(table, ranges)
        local_of = torch.full((R,), -1, dtype=torch.int64)
        local_of[RS.take_ranges(torch.arange(R), ranges)] = torch.arange((ohi - olo) + (phi - plo))
        assert torch.equal(idx_l, local_of[idx]) and bool((idx_l.reshape(L, 2)[~mine] == -1).all()) and bool((idx_l.reshape(L, 2)[mine] >= 0).all())
        counted = mine.to(torch.int64).clone()
        dist.all_reduce(counted)
        assert bool((counted == 1).all())                                       # every link belongs to exactly one rank
        # logits: own links filled, 0 elsewhere -> one all-reduce gives every rank the full [L,1]; the gradient comes back whole
        pred_l = (torch.where(mine, torch.arange(L) * 10 + rank, torch.zeros(L, dtype=torch.int64))).double().reshape(-1, 1).requires_grad_(True)
        full = RS._SumLogits.apply(pred_l, shard)
        owner = torch.tensor([next(r for r in range(world) if any(lo_ <= 2 * int(v) < hi_ for lo_, hi_ in RS.blocks_of(E, R, r, world)))
                              for v in rows_of_link])
        assert torch.equal(full.detach().reshape(-1), (torch.arange(L) * 10 + owner).double())
        w = torch.arange(1.0, L + 1).double().reshape(-1, 1)
        (full * w).sum().backward()
        assert torch.equal(pred_l.grad, w)
        # node blocks: equal size B, cover [0, N), the last may be short
        N = 37
        blocks = [RS.node_block(N, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == N and all(b[2] * world >= N for b in blocks)
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        # all_gather_blocks: every rank fills its own block of a [world*B, C] buffer; afterwards all blocks are filled everywhere
        B = blocks[0][2]
        buf = torch.full((B * world, 3), -7.0)
        buf[rank * B:(rank + 1) * B] = float(rank + 1)
        shard.all_gather_blocks(buf, B)
        assert torch.equal(buf, torch.arange(1.0, world + 1).repeat_interleave(B).reshape(-1, 1).expand(-1, 3))
        # a gradient that is complete on every rank is kept on rank 0 only (the caller sums over ranks)
        prm = torch.ones(4, dtype=torch.float64, requires_grad=True)
        (RS._Rank0Grad.apply(prm, shard) * torch.arange(4.0).double()).sum().backward()
        tot = prm.grad.clone()
        dist.all_reduce(tot)
        assert torch.equal(tot, torch.arange(4.0).double())
        C = 3
        dpf, dpr = torch.arange(4.0 * C), torch.arange(4.0 * C) + 100
        gf, gr = RS._param_grads(shard, C, dpf.clone(), dpr.clone())
        tot = torch.cat(gf + gr).clone()
        dist.all_reduce(tot)
        assert torch.equal(tot, torch.cat((dpf[3 * C:], dpf[:3 * C], dpr[3 * C:], dpr[:3 * C])))   # counted once
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_row_blocks_even_and_balanced():
    D = _load_dist()
    w = [1] * 10 + [50, 50] + [1] * 28                      # a hub pair in the middle
    blocks = D.row_blocks(w, 4)
    assert blocks[0][0] == 0 and blocks[-1][1] == len(w)
    assert all(lo % 2 == 0 and hi % 2 == 0 and lo <= hi for lo, hi in blocks)
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(3))
    sums = [sum(w[lo:hi]) for lo, hi in blocks]
    assert max(sums) <= 100 + 2                              # no block holds more than the hub + its share
    with pytest.raises(ValueError):
        D.row_blocks([1, 2, 3], 2)
    assert D.row_blocks([], 3) == [(0, 0)] * 3


def test_single_process_defaults():
    D = _load_dist()
    assert D.world() == (0, 1)
    assert D.max_over_ranks(3.5) == 3.5
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert D.allreduce_grads([p]) == 3 and torch.equal(p.grad, torch.full((3,), 2.0))


def test_row_shard_blocks_cover_the_pair_table_with_even_boundaries():
    """twowl_b200.rowshard.block_of: contiguous, disjoint, even boundaries (mates 2k / 2k+1 stay together), sizes differ by
    at most one pair; replicate=True hands every rank the same batch (row-sharded steps cut ONE batch)."""
    from twowl_b200 import dist as D
    from twowl_b200.rowshard import block_of
    for R in (2, 6, 6576, 60007608):
        for world in (1, 2, 3, 8):
            blocks = [block_of(R, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == R
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(lo % 2 == 0 and hi % 2 == 0 and hi >= lo for lo, hi in blocks)
            sizes = [(hi - lo) // 2 for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        block_of(7, 0, 2)
    # a rank's block of the pair table: its share of the observed rows [0, E) + its share of the prediction rows [E, R) - the
    # ranks' blocks tile both parts, every rank carries the same number of each kind (+- one pair)
    from twowl_b200.rowshard import blocks_of
    for E, R in ((0, 2), (4, 4), (24, 40), (30003804, 60007608)):
        for world in (1, 2, 3, 8):
            blocks = [blocks_of(E, R, r, world) for r in range(world)]
            assert blocks[0][0][0] == 0 and blocks[-1][0][1] == E and blocks[0][1][0] == E and blocks[-1][1][1] == R
            for k in (0, 1):
                assert all(a[k][1] == b[k][0] for a, b in zip(blocks, blocks[1:]))
                assert all(b[k][0] % 2 == 0 and b[k][1] % 2 == 0 and b[k][1] >= b[k][0] for b in blocks)
                sizes = [(b[k][1] - b[k][0]) // 2 for b in blocks]
                assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        blocks_of(3, 8, 0, 2)
    # own_links_first: the rank's links packed at the front in their original order, the masked ones behind them; dest maps a
    # link to its new position (a permutation), so results gathered by dest come back in the caller's order
    from twowl_b200.rowshard import own_links_first, work_cuts
    idx_l = torch.tensor([-1, -1, 4, 5, -1, -1, 0, 1, 8, 9, -1, -1])
    packed, dest = own_links_first(idx_l)
    assert packed.tolist() == [4, 5, 0, 1, 8, 9, -2, -2, -2, -2, -2, -2] and sorted(dest.tolist()) == list(range(6))
    assert torch.equal(packed.view(6, 2)[dest].clamp_min(-1), idx_l.view(6, 2))       # the masked tail carries -2 ("and all later")
    packed, dest = own_links_first(torch.full((8,), -1))
    assert bool((packed == -2).all()) and sorted(dest.tolist()) == [0, 1, 2, 3]
    # work_cuts: consecutive node ranges covering [0, M) with about equal work; a range is flagged iff it holds a long row
    lens = torch.tensor([200, 3, 1, 0, 0, 2, 70, 1, 1, 1, 5, 5, 5, 0, 9, 1])
    ptr = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)))
    for K in (1, 2, 4, 7):
        cuts = work_cuts(ptr, ptr, lens.numel(), K)
        assert cuts[0][0] == 0 and cuts[-1][1] == lens.numel() and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        for lo, hi, has_long in cuts:
            assert has_long or not bool((lens[lo:hi] > 64).any())
    assert work_cuts(ptr, ptr, lens.numel(), 4)[0][:2] == (0, 1)          # the 200-entry row is a range of its own
    # the bounds come from the first (whole-graph) CSR only - identical on every rank; the flags from the rank's own
    local = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(torch.minimum(lens, torch.tensor(3)), 0)))
    assert [c[:2] for c in work_cuts(ptr, local, lens.numel(), 4)] == [c[:2] for c in work_cuts(ptr, ptr, lens.numel(), 4)]
    assert not any(c[2] for c in work_cuts(ptr, local, lens.numel(), 4))
    a = D.shard_batch(1000, 100, step=3, seed=1, replicate=True)
    b = D.shard_batch(1000, 100, step=3, seed=1)            # world size 1: the same draw
    assert torch.equal(a, b) and a.unique().numel() == 100
