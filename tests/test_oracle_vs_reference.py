"""CPU, build container only: the oracle (and the host-side mirror of the head) against the reference's
OWN source executed in place through oracle/ref_shim.py.  Skipped where no checkout is reachable
(e.g. the GPU box) -- the committed fixtures in tests/golden/ carry the pinning there."""
import numpy as np
import pytest
import torch

from balancedgroupsoftmax_b200.tables import synthetic_tables, build_group_tables, synthetic_instance_counts
from oracle import bags_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference checkout not reachable')


@pytest.fixture(scope='module')
def ref():
    t = synthetic_tables(1231, seed=0)
    head = ref_shim.build_reference_head(t, fc_out_channels=128)
    head.init_weights()
    return t, head


@pytest.mark.parametrize('N,npos,seed', [(1, 1, 0), (1, 0, 1), (64, 16, 2), (300, 75, 3), (512, 128, 4), (40, 40, 5)])
def test_loss_and_grads_match_reference(ref, N, npos, seed):
    t, head = ref
    l2b, ps = torch.from_numpy(t.label2binlabel), torch.from_numpy(t.pred_slice)
    torch.manual_seed(seed)
    with torch.no_grad():
        head.fc_cls.weight.normal_(0, 0.2)
        head.fc_cls.bias.normal_(0, 0.1)
    x = torch.relu(torch.randn(N, 128))
    labels = torch.zeros(N, dtype=torch.long)
    labels[:npos] = torch.randint(1, 1231, (npos,))
    np.random.seed(seed)
    xr = x.clone().requires_grad_(True)
    head.zero_grad()
    z = head.fc_cls(xr)
    losses = head.loss(z, None, labels, None, None, None)
    sum(losses.values()).backward()
    assert set(losses.keys()) == {'loss_cls_bin%d' % g for g in range(5)}
    W, b = head.fc_cls.weight.detach(), head.fc_cls.bias.detach()
    np.random.seed(seed)   # same numpy state => identical sampled masks
    lo, dW, db, dX = O.head_step(x, W, b, labels, l2b, ps, 8.0)
    for k in losses:
        assert abs(lo[k].item() - losses[k].item()) <= 1e-6 * max(1.0, abs(losses[k].item())), k
    assert ((dW - head.fc_cls.weight.grad).norm() / head.fc_cls.weight.grad.norm().clamp_min(1e-20)).item() < 1e-6
    assert ((db - head.fc_cls.bias.grad).norm() / head.fc_cls.bias.grad.norm().clamp_min(1e-20)).item() < 1e-6
    assert ((dX - xr.grad).norm() / xr.grad.norm().clamp_min(1e-20)).item() < 1e-6


def test_merge_score_matches_reference(ref):
    t, head = ref
    ps = torch.from_numpy(t.pred_slice)
    torch.manual_seed(7)
    for _ in range(3):
        z = torch.randn(200, t.num_logits) * 3
        a = head._merge_score(z)
        b = O.merge_score(z, ps, [torch.from_numpy(s) for s in t.fg_splits], t.num_classes)
        assert (a - b).abs().max().item() == 0.0
        assert torch.equal(a.argmax(1), b.argmax(1))


def test_tables_load_into_reference_head(ref):
    """Files written by tables.save_reference_files are what the reference's constructor reads
    (gs_bbox_head_with0.py:37-49): shapes/dtypes/keys survive the round trip."""
    t, head = ref
    assert head.label2binlabel.dtype == torch.int64 and tuple(head.label2binlabel.shape) == (5, 1231)
    assert torch.equal(head.label2binlabel, torch.from_numpy(t.label2binlabel))
    assert torch.equal(head.pred_slice, torch.from_numpy(t.pred_slice))
    assert len(head.fg_splits) == 4
    for a, b in zip(head.fg_splits, t.fg_splits):
        assert torch.equal(a, torch.from_numpy(b))
    assert head.fc_cls.out_features == 1236


def test_host_mirror_numpy_sampler_is_the_reference_sampler(ref):
    """GSBBoxHeadWith0(sampler='numpy')._sample_others_numpy draws the same masks as the reference."""
    from balancedgroupsoftmax_b200.head import GSBBoxHeadWith0
    t, head = ref
    mine = GSBBoxHeadWith0(num_fcs=2, in_channels=4, fc_out_channels=128, roi_feat_size=2, num_classes=1231,
                           gs_config=dict(tables=t, others_sample_ratio=8.0, num_bins=5, sampler='numpy',
                                          loss_bin=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)))
    labels = torch.zeros(400, dtype=torch.long)
    labels[:90] = torch.randint(1, 1231, (90,), generator=torch.Generator().manual_seed(3))
    np.random.seed(11)
    _, ref_w, ref_avg = head._remap_labels(labels)
    np.random.seed(11)
    for g in range(1, 5):
        w = mine._sample_others_numpy(mine.label2binlabel[g][labels])
        assert torch.equal(w, ref_w[g])


def test_get_target_matches_reference_bbox_target():
    """The head's standalone target generator against the reference's bbox_target.py / transforms.py run in place."""
    from types import SimpleNamespace
    from balancedgroupsoftmax_b200.head import GSBBoxHeadWith0, bbox2delta, bbox_target
    ref_bbox_target, ref_bbox2delta = ref_shim.load_bbox_target()
    g = torch.Generator().manual_seed(11)

    def boxes(n):
        xy = torch.rand(n, 2, generator=g) * 600
        wh = torch.rand(n, 2, generator=g) * 200 + 1
        return torch.cat([xy, xy + wh], 1)

    imgs = []
    for npos, nneg in ((5, 20), (0, 12), (7, 0), (1, 1)):
        imgs.append(SimpleNamespace(pos_bboxes=boxes(npos), neg_bboxes=boxes(nneg), pos_gt_bboxes=boxes(npos),
                                    pos_gt_labels=torch.randint(1, 1231, (npos,), generator=g)))
    means, stds = [0., 0., 0., 0.], [0.1, 0.1, 0.2, 0.2]
    assert torch.equal(bbox2delta(imgs[0].pos_bboxes, imgs[0].pos_gt_bboxes, means, stds),
                       ref_bbox2delta(imgs[0].pos_bboxes, imgs[0].pos_gt_bboxes, means, stds))
    for pos_weight in (-1, 2.5):
        cfg = ref_shim.AttrDict(pos_weight=pos_weight)
        args = ([r.pos_bboxes for r in imgs], [r.neg_bboxes for r in imgs], [r.pos_gt_bboxes for r in imgs],
                [r.pos_gt_labels for r in imgs], cfg)
        want = ref_bbox_target(*args, reg_classes=1231, target_means=means, target_stds=stds)
        got = bbox_target(*args, reg_classes=1231, target_means=means, target_stds=stds)
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and torch.equal(a, b)
        # not concatenated
        want = ref_bbox_target(*args, target_means=means, target_stds=stds, concat=False)
        got = bbox_target(*args, target_means=means, target_stds=stds, concat=False)
        for la, lb in zip(got, want):
            assert len(la) == len(lb) and all(torch.equal(a, b) for a, b in zip(la, lb))
    # through the head method
    t = synthetic_tables()
    head = GSBBoxHeadWith0(num_fcs=1, in_channels=4, fc_out_channels=16, roi_feat_size=1, num_classes=t.num_classes,
                           target_means=means, target_stds=stds,
                           gs_config=dict(tables=t, others_sample_ratio=8.0, num_bins=5,
                                          loss_bin=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)))
    cfg = ref_shim.AttrDict(pos_weight=-1)
    got = head.get_target(imgs, None, None, cfg)
    want = ref_bbox_target([r.pos_bboxes for r in imgs], [r.neg_bboxes for r in imgs], [r.pos_gt_bboxes for r in imgs],
                           [r.pos_gt_labels for r in imgs], cfg, 1231, target_means=means, target_stds=stds)
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    assert got[0].dtype == torch.long and got[0][:5].tolist() == imgs[0].pos_gt_labels.tolist() and got[0][5:25].sum() == 0


def test_multiclass_nms_oracle_matches_reference_loop():
    """The oracle's multiclass_nms against the reference's bbox_nms.py run in place (its compiled NMS op replaced by
    the oracle's greedy "+1" NMS): thresholds, labels, class order and the top-k rule."""
    ref_mc_nms = ref_shim.load_multiclass_nms(O.nms_plus1)
    g = torch.Generator().manual_seed(5)
    n, classes = 60, 9
    xy = torch.rand(n, 2, generator=g) * 80
    wh = torch.rand(n, 2, generator=g) * 40 + 2
    boxes4 = torch.cat([xy, xy + wh], 1)
    boxes_pc = (boxes4[:, None, :] + torch.rand(n, classes, 4, generator=g)).reshape(n, classes * 4)
    scores = torch.rand(n, classes, generator=g) ** 3
    for mb in (boxes4, boxes_pc):
        for thr, iou, k in ((0.05, 0.5, 20), (0.0, 0.3, 1000), (0.9999, 0.5, 10), (0.2, 0.5, -1)):
            want = ref_mc_nms(mb, scores.clone(), thr, dict(type='nms', iou_thr=iou), k)
            got = O.multiclass_nms(mb, scores.clone(), thr, iou, k)
            assert got[0].shape == want[0].shape and torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_reweight_variant_matches_reference():
    """Reweight head variant (gs_bbox_head_with0_reweight.py): the oracle's weights / normalisers / per-bin losses
    against the reference class run in place."""
    t = synthetic_tables(1231, seed=0)
    g = torch.Generator().manual_seed(21)
    cls_weights = [torch.rand(int(t.pred_slice[b, 1]), generator=g) * 2 + 0.1 for b in range(1, t.num_bins)]
    head = ref_shim.build_reference_reweight_head(t, cls_weights, fc_out_channels=64)
    head.init_weights()
    l2b, ps = torch.from_numpy(t.label2binlabel), torch.from_numpy(t.pred_slice)
    for N, npos, seed in ((300, 75, 1), (64, 0, 2), (40, 40, 3), (512, 128, 4)):
        torch.manual_seed(seed)
        with torch.no_grad():
            head.fc_cls.weight.normal_(0, 0.2)
            head.fc_cls.bias.normal_(0, 0.1)
        x = torch.relu(torch.randn(N, 64))
        labels = torch.zeros(N, dtype=torch.long)
        labels[:npos] = torch.randint(1, 1231, (npos,))
        z = head.fc_cls(x).detach()
        np.random.seed(seed)
        want = head.loss(z, None, labels, None, None, None)
        np.random.seed(seed)
        rl, rw, ra = head._remap_labels(labels)
        np.random.seed(seed)
        remapped = O.remap_labels_reweight(labels, l2b, 8.0, cls_weights)
        for a, b in zip(remapped[1], rw):
            assert torch.allclose(a.float(), b.float(), rtol=0, atol=0)
        assert remapped[2] == ra
        got = O.bags_loss(z, labels, l2b, ps, remapped=remapped)
        for k in want:
            assert abs(got[k].item() - want[k].item()) <= 1e-6 * max(1.0, abs(want[k].item())), (k, got[k], want[k])
