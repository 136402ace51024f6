"""``Matcher`` with the reference's constructor, call signature and error behaviour
(detectron2/modeling/matcher.py:21-132), computed by ``fsg_matcher`` on the GPU."""
import torch

from . import ops

STRICT = False  # True: keep the reference's host-synchronising asserts (matcher.py:82)


class Matcher(object):
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix):
        """(M,N) quality matrix -> (matches int64 (N), match_labels int8 (N)); matcher.py:55-97."""
        assert match_quality_matrix.dim() == 2
        if STRICT and match_quality_matrix.numel() > 0:
            assert torch.all(match_quality_matrix >= 0)
        return ops.matcher(match_quality_matrix, self.thresholds[1:-1], self.labels, self.allow_low_quality_matches)

    def match_boxes(self, gt_boxes, anchors):
        """Fused form: IoU + matching for one image without materialising the matrix.
        gt_boxes (M,4), anchors (R,4) -> (matches int64 (R), match_labels int8 (R))."""
        gt = ops.PackedGT.from_lists([gt_boxes], [torch.zeros(gt_boxes.shape[0], dtype=torch.int64)], anchors.device)
        out = ops.match_anchors(anchors, gt, 1, self.thresholds[1:-1], self.labels, None, None,
                                want=("matches", "match_labels"),
                                allow_low_quality_matches=self.allow_low_quality_matches)
        return out["matches"][0], out["match_labels"][0]
