"""GPU, EXPERIMENTAL: the one-launch class-aware NMS (bags_class_nms) against the oracle's per-class loop.
The kernel was written after round 1 ran out of GPU time, so this test only runs with BAGS_TEST_EXPERIMENTAL=1 until
it has been seen green once (then drop the gate)."""
import os

import pytest
import torch

from oracle import bags_oracle as O

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize('n,classes,per_class_boxes', [(60, 9, False), (300, 40, True), (1000, 6, False), (5, 3, True)])
def test_class_nms_matches_oracle(n, classes, per_class_boxes):
    from balancedgroupsoftmax_b200 import ops
    g = torch.Generator().manual_seed(n + classes)
    xy = torch.rand(n, 2, generator=g) * 200
    wh = torch.rand(n, 2, generator=g) * 80 + 2
    b4 = torch.cat([xy, xy + wh], 1)
    boxes = (b4[:, None, :] + torch.rand(n, classes, 4, generator=g)).reshape(n, classes * 4) if per_class_boxes else b4
    scores = torch.rand(n, classes, generator=g) ** 2
    for thr, iou, k in ((0.05, 0.5, 100), (0.0, 0.3, 300), (0.999999, 0.5, 10), (0.3, 0.7, -1)):
        want_b, want_l = O.multiclass_nms(boxes, scores, thr, iou, k)
        got_b, got_l = ops.multiclass_nms(boxes.cuda(), scores.cuda(), thr, iou, k)
        assert got_b.shape == want_b.shape
        # same detections; order inside equal scores may differ
        key_w = sorted(zip(want_l.tolist(), [tuple(r) for r in want_b.tolist()]))
        key_g = sorted(zip(got_l.cpu().tolist(), [tuple(r) for r in got_b.cpu().tolist()]))
        assert key_w == key_g
