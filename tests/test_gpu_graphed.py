"""GPU: the CUDA-graph step of the public API (GraphedHeadStep) reproduces the eager autograd API call for call,
including the device-side seed counter of the sampler."""
import pytest
import torch

from balancedgroupsoftmax_b200.tables import synthetic_tables

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('n', [300, 1024])
def test_graphed_step_matches_eager(n):
    from balancedgroupsoftmax_b200 import ops
    from balancedgroupsoftmax_b200.api import GraphedHeadStep, bags_head_loss
    dev = torch.device('cuda', 0)
    t = synthetic_tables(1231, seed=0)
    dt = ops.DeviceTables.from_tables(t, dev)
    g = torch.Generator().manual_seed(7)
    W = torch.nn.Parameter((torch.randn(t.num_logits, 1024, generator=g) * 0.05).to(dev).bfloat16())
    b = torch.nn.Parameter((torch.randn(t.num_logits, generator=g) * 0.1).to(dev))
    lw = torch.tensor([1.0, 0.5, 0.5, 0.25, 2.0])
    step = GraphedHeadStep(W, b, dt, n, others_sample_ratio=8.0, loss_weights=lw, seed=1234)
    for it in range(3):
        x = torch.relu(torch.randn(n, 1024, generator=g)).to(dev).bfloat16()
        labels = torch.zeros(n, dtype=torch.int64)
        labels[: n // 4] = torch.randint(1, t.num_classes, (n // 4,), generator=g)
        labels = labels.to(dev)
        losses = step(x, labels).clone()
        gW, gb, gx = step.grad_weight.clone(), step.grad_bias.clone(), step.grad_x.clone()
        torch.cuda.synchronize()
        # eager call with the masks the graph's sampler drew at this replay (seed counter = it)
        cnt = torch.tensor([it], dtype=torch.int64, device=dev)
        wmask, avg = ops.sample_others(labels, dt, 8.0, 1234, seed_step=cnt)
        xe = x.clone().requires_grad_(True)
        W.grad = None
        b.grad = None
        le = bags_head_loss(xe, W, b, labels, dt, 8.0, wmask=wmask, avg=avg)
        (le * lw.to(dev)).sum().backward()
        torch.cuda.synchronize()
        assert torch.allclose(losses, le.detach(), rtol=1e-6, atol=1e-7), (it, losses, le)
        # Same kernels, same inputs: the only run-to-run difference is the order of the split-K red.add of dW (fp32),
        # which can flip the bf16 rounding of individual elements of the bf16-typed gradients (one bf16 ulp = 2^-8
        # relative on that element) -- hence bf16-ulp-sized bounds for gW / gx and an fp32 bound for gb.
        assert rel(gW.float(), W.grad.float()) < 2e-3
        assert rel(gb, b.grad) < 1e-4
        assert rel(gx.float(), xe.grad.float()) < 2e-3
        # restore the graph's static grads as the parameters' grads (the eager call replaced them)
        W.grad, b.grad = step.grad_weight, step.grad_bias
    # the sampler really draws a different subset per replay
    labels = torch.zeros(n, dtype=torch.int64, device=dev)
    labels[: n // 8] = 5
    m0, _ = ops.sample_others(labels, dt, 2.0, 1234, seed_step=torch.tensor([0], device=dev))
    m1, _ = ops.sample_others(labels, dt, 2.0, 1234, seed_step=torch.tensor([1], device=dev))
    mp, _ = ops.sample_others(labels, dt, 2.0, 1234)
    assert torch.equal(m0, mp)                  # counter 0 == the plain seed
    assert not torch.equal(m0, m1)
    assert torch.equal(m0.sum(1), m1.sum(1))    # same subset sizes (exact-k sampling)


def test_graph_cached_head_loss_matches_eager_for_recurring_roi_counts():
    """GraphCachedHeadLoss: a data-dependent RoI count inside an ordinary autograd graph -- recurring counts replay two
    CUDA graphs (forward / backward), new ones run eagerly first; results equal the eager API with the same masks."""
    from balancedgroupsoftmax_b200 import ops
    from balancedgroupsoftmax_b200.api import GraphCachedHeadLoss, bags_head_loss
    dev = torch.device('cuda', 0)
    t = synthetic_tables(1231, seed=0)
    dt = ops.DeviceTables.from_tables(t, dev)
    g = torch.Generator().manual_seed(3)
    W = torch.nn.Parameter((torch.randn(t.num_logits, 1024, generator=g) * 0.05).to(dev))      # fp32 master
    b = torch.nn.Parameter((torch.randn(t.num_logits, generator=g) * 0.1).to(dev))
    fn = GraphCachedHeadLoss(dt, 8.0, seed=77, capture_after=1, max_graphs=2)
    replays = {}
    for it, n in enumerate([512, 512, 300, 512, 300, 300, 1000, 1000, 512]):
        x = torch.relu(torch.randn(n, 1024, generator=g)).to(dev).requires_grad_(True)        # fp32, as a trunk gives it
        labels = torch.zeros(n, dtype=torch.int64)
        labels[: n // 4] = torch.randint(1, t.num_classes, (n // 4,), generator=g)
        labels = labels.to(dev)
        W.grad = b.grad = None
        cached_before = n in fn._pairs
        seen_before = fn._seen.get(n, 0)
        losses = fn(x, W, b, labels)
        (losses * torch.tensor([1.0, 0.5, 2.0, 1.0, 0.25], device=dev)).sum().backward()
        torch.cuda.synchronize()
        graphed = cached_before or seen_before >= 1
        if not graphed:
            continue
        # the same step through the eager API with the masks this replay's device sampler drew
        k = replays.get(n, 0)
        replays[n] = k + 1
        pair_seed = (77 + 0x9E3779B97F4A7C15 * n) & 0xFFFFFFFFFFFFFFFF
        if n not in fn._pairs:       # evicted and recaptured: its counter restarted
            continue
        wmask, avg = ops.sample_others(labels, dt, 8.0, pair_seed,
                                       seed_step=torch.tensor([int(fn._pairs[n].seed_step.item()) - 1], device=dev))
        gW, gb, gx = W.grad.clone(), b.grad.clone(), x.grad.clone()
        xe = x.detach().clone().requires_grad_(True)
        W.grad = b.grad = None
        le = bags_head_loss(xe, W, b, labels, dt, 8.0, wmask=wmask, avg=avg)
        (le * torch.tensor([1.0, 0.5, 2.0, 1.0, 0.25], device=dev)).sum().backward()
        torch.cuda.synchronize()
        assert torch.allclose(losses, le.detach(), rtol=1e-5, atol=1e-6), (it, n)
        assert rel(gW, W.grad) < 1e-4 and rel(gb, b.grad) < 2e-3 and rel(gx, xe.grad) < 2e-3, (it, n)
    assert fn.stats['captures'] >= 3 and fn.stats['eager'] >= 3 and len(fn._pairs) <= 2


def test_head_loss_with_graph_cache_trains_like_the_eager_head():
    """gs_config['graph_cache']=True: GSBBoxHeadWith0.loss replays CUDA graphs for a recurring RoI count; the losses of
    a few SGD steps track the eager head's (same initial weights, different "others" draws: statistical agreement)."""
    from balancedgroupsoftmax_b200.head import GSBBoxHeadWith0
    t = synthetic_tables(1231, seed=0)

    def make(graph_cache):
        torch.manual_seed(0)
        h = GSBBoxHeadWith0(num_fcs=2, in_channels=16, fc_out_channels=1024, roi_feat_size=2, num_classes=1231,
                            gs_config=dict(tables=t, others_sample_ratio=8.0, num_bins=5, graph_cache=graph_cache,
                                           loss_bin=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)))
        h.init_weights()
        return h.cuda().train()
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(640, 16, 2, 2, generator=g).cuda()
    labels = torch.zeros(640, dtype=torch.long)
    labels[:160] = torch.randint(1, 1231, (160,), generator=g)
    labels = labels.cuda()
    curves = []
    for gc in (False, True):
        head = make(gc)
        opt = torch.optim.SGD(head.parameters(), lr=0.05)
        curve = []
        for it in range(6):
            opt.zero_grad(set_to_none=True)
            cls, reg = head(feats)
            losses = head.loss(cls, None, labels, None, None, None)
            total = sum(losses.values())
            total.backward()
            assert head.fc_cls.weight.grad is not None and head.shared_fcs[0].weight.grad.abs().sum().item() > 0
            opt.step()
            curve.append(total.item())
        curves.append(curve)
        if gc:
            assert head._graph_loss.stats['replays'] >= 3 and head._graph_loss.stats['captures'] == 1
    a, b = curves
    assert all(abs(x - y) <= 0.05 * abs(x) + 0.05 for x, y in zip(a, b)), (a, b)
    assert b[-1] < b[0]        # it trains
