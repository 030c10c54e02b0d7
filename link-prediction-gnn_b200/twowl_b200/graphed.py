"""One training step of TwoWL/model/train.py:16-38 (block the batch's edges -> forward -> BCE-with-logits -> backward) captured
ONCE as a CUDA graph and replayed for every batch of the same size.

On the small configurations (fb-pages-food, Cora scale) a step is ~90 kernel launches of a few microseconds each and the host
(Python dispatch, ctypes, allocator) sets the pace; replaying the captured graph removes that cost. Nothing in forward / backward
reads the device from the host, so the whole step is capturable; the only data-dependent SIZE of the reference's step,
`ei_new = ei[:, ~mask]` (utils.py:62-63), is never needed as a tensor: the node-level GCNConv reads the cached CSRs of the whole
graph with a per-entry mask (graph.MaskedEdges), and the degree feature is the full degree minus the blocked edges' sources.

Replay-time inputs are copied into static buffers (blocked edge ids, readout row ids, labels); outputs are the static loss /
logits tensors and the parameters' `.grad`. Dropout: scalar kernel arguments are baked into a captured graph, so during capture
the seeds are handed out as ADDRESSES of slots of a device buffer (the kernels read a seed with bit 63 set through that address,
include/twowl.h), and every replay first refills the buffer with fresh random numbers on the device: new masks per step, forward
and backward of one step still see the same seeds.
"""
from __future__ import annotations

import torch

from . import graph as G
from . import ops


class DeviceSeeds:
    """Dropout seeds in device memory: `provider` hands out the tagged address of the next slot (as the signed int a torch custom
    op accepts); `refresh()` redraws every slot on the device (no host sync)."""
    SLOTS = 256

    def __init__(self, device):
        self.buf = torch.zeros(self.SLOTS, dtype=torch.int64, device=device)
        self.used = 0
        self.refresh()

    def provider(self) -> int:
        if self.used >= self.SLOTS:
            raise RuntimeError("more dropout seeds in one step than DeviceSeeds.SLOTS")
        addr = self.buf.data_ptr() + 8 * self.used
        self.used += 1
        return addr - (1 << 63)          # bit 63 set, as a signed 64-bit integer

    def refresh(self):
        self.buf.random_(0, 2 ** 62)

    def values(self):
        """Host copy of the current seeds (tests: replaying the same step eagerly)."""
        return [int(v) for v in self.buf.cpu().tolist()]


class GraphedTrainStep:
    def __init__(self, mod, n_node: int, ei: torch.Tensor, pos1: torch.Tensor, ei2, n_block: int, n_links: int,
                 loss_fn=None, warmup: int = 3, optimizer=None):
        """mod: LocalWLNet; ei: int64 [2,E] observed edges; pos1: int64 [R,2]; ei2: what get_ei2 / get_ei2_implicit returned for
        them; n_block: blocked edge ids per step (= 2 x positive undirected links); n_links: target links per step.
        loss_fn: default = twowl::bce_with_logits (train.py:37 in two launches). optimizer: a twowl_b200.optim.FusedAdam whose
        update (train.py:39) is then part of the captured step - the caller must not call its step() again."""
        if loss_fn is None:
            from .functional import bce_with_logits as loss_fn
        from TwoWL.utils import _struct_of
        self.mod, self.n, self.ei, self.pos1, self.loss_fn = mod, int(n_node), ei, pos1, loss_fn
        self.struct = ei2.struct if isinstance(ei2, G.WedgeIndex) else _struct_of(ei2)
        if self.struct is None:
            raise RuntimeError("GraphedTrainStep needs an ei2 made by TwoWL.utils.get_ei2 / get_ei2_implicit")
        dev = ei.device
        self.E = ei.shape[1]
        self.deg_src = ops.degree(ei[0], self.n)                      # utils.py:66-67 counts by SOURCE
        self.s_block = torch.zeros(n_block, dtype=torch.int64, device=dev)
        self.s_idx = torch.zeros(2 * n_links, dtype=torch.int64, device=dev)
        self.s_y = torch.zeros((n_links, 1), dtype=torch.float32, device=dev)
        self.params = [p for p in mod.parameters() if p.requires_grad]
        self.optimizer = optimizer
        self.graph = None
        self.loss = self.logits = None
        self._warmup = warmup
        self.seeds = DeviceSeeds(dev)

    def _body(self):
        self.seeds.used = 0
        prev = ops.set_seed_provider(self.seeds.provider)
        try:
            self._step()
        finally:
            ops.set_seed_provider(prev)

    def _step(self):
        mask = ops.mask_from_idx(self.s_block, self.E)
        # degree by source of the edges the MASK keeps (utils.py:62-67): an id that occurs twice in the blocked list removes its
        # edge once, exactly like sample_block's ei[:, ~mask]
        x_new = self.deg_src - torch.zeros_like(self.deg_src).index_add_(0, self.ei[0], mask.to(self.deg_src.dtype))
        edges = G.MaskedEdges(self.ei, mask)
        wedges = G.WedgeIndex(self.struct.with_blocked(mask))
        self.logits = self.mod(x_new, edges, self.pos1, self.s_idx, wedges)
        self.loss = self.loss_fn(self.logits, self.s_y)
        self.loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()

    def _load(self, blocked_ids, idx, y):
        self.s_block.copy_(blocked_ids.reshape(-1), non_blocking=True)
        self.s_idx.copy_(idx.reshape(-1), non_blocking=True)
        self.s_y.copy_(y.reshape(self.s_y.shape), non_blocking=True)

    def capture(self, blocked_ids, idx, y):
        """Warm up on a side stream (builds every cached index structure), then capture one step."""
        self._load(blocked_ids, idx, y)
        # the warm-up iterations must leave no trace in the optimiser: its state and the parameters are restored afterwards
        snap = None if self.optimizer is None else {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.optimizer.state_dict().items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                for p in self.params:
                    p.grad = None
                self._body()
                self.loss = self.logits = None      # drop the autograd graph before the next iteration / the capture
        cur.wait_stream(side)
        if snap is not None:
            self.optimizer.load_state_dict(snap)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        launches0 = ops.launches()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = ops.launches() - launches0
        self.static_grads = [p.grad for p in self.params]     # the buffers every replay writes
        return self

    def __call__(self, blocked_ids, idx, y):
        """-> loss (a static tensor, overwritten by the next call). Gradients land in the parameters' .grad."""
        if self.graph is None:
            self.capture(blocked_ids, idx, y)
        self._load(blocked_ids, idx, y)
        self.seeds.refresh()
        self.graph.replay()
        ops.add_launches(self.launches_per_step)
        for p, g in zip(self.params, self.static_grads):       # the caller may have cleared .grad (optimizer.zero_grad)
            p.grad = g
        return self.loss
