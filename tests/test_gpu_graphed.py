"""CUDA-graph replay of the training step (twowl_b200.graphed) against the eager drop-in path
(sample_block -> LocalWLNet.forward -> BCE -> backward, train.py:16-38): bit-identical loss, logits and gradients on batches the
graph was NOT captured on."""
import numpy as np
import pytest
import torch


class _F2:
    def __getattr__(self, k):
        from twowl_b200 import functional
        return getattr(functional, k)


F2 = _F2()

from helpers import fb_split

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("implicit", [False, True])
def test_graphed_step_replays_bit_identical_to_eager(fb, implicit):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import TwoWL.model.model as model
    import TwoWL.utils as U
    from twowl_b200.graphed import GraphedTrainStep
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    n = int(fb["num_nodes"][0])
    ei, pred, pos1 = fb_split(fb, 0)
    dei, dpred, dpos = dev(ei), dev(pred), dev(pos1)
    E = dei.shape[1]
    ei2 = U.get_ei2_implicit(n, dei, dpred) if implicit else U.get_ei2(n, dei, dpred)
    torch.manual_seed(5)
    mod = model.LocalWLNet(int(U.degree(dei, n).max().item()), False, None, channels_1wl=64, channels_2wl=32, depth1=2, depth2=1,
                           dp_lin0=0., dp_lin1=0., dp_emb=0., dp_1wl0=0., dp_2wl=0., dp_1wl1=0.).cuda().train()
    nb = 192
    step = GraphedTrainStep(mod, n, dei, dpos, ei2, n_block=2 * nb, n_links=2 * nb)
    y = torch.cat((torch.ones(nb), torch.zeros(nb))).unsqueeze(-1).cuda()
    g = torch.Generator().manual_seed(0)
    for it in range(4):
        i1 = torch.randperm(E // 2, generator=g)[:nb].cuda()
        i2 = torch.randperm(dpred.shape[1] // 2, generator=g)[:nb].cuda()
        idx1 = U.double(i1, for_index=True)
        idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
        loss_g = step(idx1, idx, y).clone()                     # iteration 0 captures, 1..3 replay on new batches
        logits_g = step.logits.clone()
        grads_g = {k: p.grad.clone() for k, p in mod.named_parameters()}
        for p in mod.parameters():
            p.grad = None
        ei_new, x_new, ei2_new = U.sample_block(idx1, n, dei, ei2)
        out = mod(x_new, ei_new, dpos, idx, ei2_new)
        loss = F2.bce_with_logits(out, y)     # the loss the captured step uses (twowl::bce_with_logits, train.py:37)
        loss.backward()
        assert torch.equal(out, logits_g) and torch.equal(loss, loss_g), f"batch {it}"
        for k, p in mod.named_parameters():
            assert torch.equal(p.grad, grads_g[k]), f"batch {it} grad {k}"
        for p in mod.parameters():
            p.grad = None
    assert step.launches_per_step > 50


def test_graphed_step_with_dropout_redraws_its_masks_and_matches_eager(fb):
    """Dropout under CUDA-graph replay: the seeds live in device memory (a seed argument with bit 63 set is an address), every
    replay redraws them, and an eager step fed the same seed values reproduces the replay bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import TwoWL.model.model as model
    import TwoWL.utils as U
    from twowl_b200 import ops
    from twowl_b200.graphed import GraphedTrainStep
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    n = int(fb["num_nodes"][0])
    ei, pred, pos1 = fb_split(fb, 0)
    dei, dpred, dpos = dev(ei), dev(pred), dev(pos1)
    E = dei.shape[1]
    ei2 = U.get_ei2_implicit(n, dei, dpred)
    torch.manual_seed(9)
    mod = model.LocalWLNet(int(U.degree(dei, n).max().item()), False, None, channels_1wl=64, channels_2wl=32, depth1=2, depth2=2,
                           dp_lin0=0., dp_lin1=0., dp_emb=0.2, dp_1wl0=0.1, dp_2wl=0.3, dp_1wl1=0.1).cuda().train()
    nb = 192
    step = GraphedTrainStep(mod, n, dei, dpos, ei2, n_block=2 * nb, n_links=2 * nb)
    y = torch.cat((torch.ones(nb), torch.zeros(nb))).unsqueeze(-1).cuda()
    g = torch.Generator().manual_seed(0)
    i1 = torch.randperm(E // 2, generator=g)[:nb].cuda()
    i2 = torch.randperm(dpred.shape[1] // 2, generator=g)[:nb].cuda()
    idx1 = U.double(i1, for_index=True)
    idx = torch.cat((idx1, U.double(i2, for_index=True) + E))
    losses = []
    for it in range(3):
        loss_g = step(idx1, idx, y).clone()                 # the same batch every time: only the masks change
        logits_g = step.logits.clone()
        grads_g = {k: p.grad.clone() for k, p in mod.named_parameters()}
        losses.append(float(loss_g))
        assert step.seeds.used >= 6                         # emb + 2 node layers + 2 x 2 pair-layer branches drew seeds
        vals = iter(step.seeds.values()[: step.seeds.used])
        prev = ops.set_seed_provider(lambda: next(vals))
        try:
            for p in mod.parameters():
                p.grad = None
            ei_new, x_new, ei2_new = U.sample_block(idx1, n, dei, ei2)
            out = mod(x_new, ei_new, dpos, idx, ei2_new)
            loss = F2.bce_with_logits(out, y)     # the loss the captured step uses (twowl::bce_with_logits, train.py:37)
            loss.backward()
        finally:
            ops.set_seed_provider(prev)
        assert torch.equal(out, logits_g) and torch.equal(loss, loss_g), f"replay {it}"
        for k, p in mod.named_parameters():
            assert torch.equal(p.grad, grads_g[k]), f"replay {it} grad {k}"
        for p in mod.parameters():
            p.grad = None
    assert len(set(losses)) == 3, losses                    # fresh masks on every replay
