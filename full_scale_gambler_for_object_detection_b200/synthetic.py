"""Seeded synthetic inputs for the BASELINE.json configs.

Everything is generated on the CPU under ``torch.manual_seed(20261018 + config_id)`` so the CPU oracle and
the CUDA kernels see identical bits; callers copy to the GPU.
"""
import math

import torch

from .anchor_generator import retinanet_anchors

BASE_SEED = 20261018
PRIOR_LOGIT = -math.log((1 - 0.01) / 0.01)  # retinanet.py:582-583 prior-prob bias, -4.595


def _gt_boxes(gen, M, H, W):
    cx = torch.rand(M, generator=gen) * W
    cy = torch.rand(M, generator=gen) * H
    lo, hi = math.log(16.0), math.log(0.6 * min(H, W))
    w = torch.exp(torch.rand(M, generator=gen) * (hi - lo) + lo)
    h = torch.exp(torch.rand(M, generator=gen) * (hi - lo) + lo)
    b = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), dim=1)
    b[:, 0::2] = b[:, 0::2].clamp(0, W)
    b[:, 1::2] = b[:, 1::2].clamp(0, H)
    return b.to(torch.float32)


def train_inputs(config_id, N, height, width, K, M=8, empty_image=True, logits=True):
    """Inputs of the fused match + loss step: anchors (R,4), per-image GT (one image with M=0),
    logits ~ N(-4.595, 1), predicted deltas ~ N(0, 0.1), bets = sigmoid(N(-4.595, 1))."""
    gen = torch.Generator().manual_seed(BASE_SEED + config_id)
    anchors, offs, grids = retinanet_anchors(height, width)
    R = anchors.shape[0]
    gt_boxes, gt_classes = [], []
    for i in range(N):
        m = 0 if (empty_image and i == N - 1) else M
        gt_boxes.append(_gt_boxes(gen, m, height, width) if m else torch.zeros((0, 4)))
        gt_classes.append(torch.randint(0, K, (m,), generator=gen, dtype=torch.int64))
    out = dict(anchors=anchors, level_offsets=offs, grids=grids, gt_boxes=gt_boxes, gt_classes=gt_classes,
               N=N, R=R, K=K, A=3)
    if logits:
        out["logits"] = torch.randn((N, R, K), generator=gen) + PRIOR_LOGIT
        out["deltas"] = torch.randn((N, R, 4), generator=gen) * 0.1
        out["bets"] = torch.sigmoid(torch.randn((N, R), generator=gen) + PRIOR_LOGIT)
    return out


def inference_inputs(config_id, N, level_anchor_counts, K, image_hw=(800, 1344)):
    """Config 4: logits ~ N(-4.595, 1.5), deltas ~ N(0, 0.2), anchors = plausible boxes per level."""
    gen = torch.Generator().manual_seed(BASE_SEED + config_id)
    H, W = image_hw
    offs = [0]
    anchors = []
    for l, cnt in enumerate(level_anchor_counts):
        size = 32.0 * 2 ** l
        cx = torch.rand(cnt, generator=gen) * W
        cy = torch.rand(cnt, generator=gen) * H
        s = size * 2 ** (torch.randint(0, 3, (cnt,), generator=gen).float() / 3.0)
        anchors.append(torch.stack((cx - s / 2, cy - s / 2, cx + s / 2, cy + s / 2), dim=1))
        offs.append(offs[-1] + cnt)
    anchors = torch.cat(anchors).to(torch.float32).contiguous()
    R = anchors.shape[0]
    logits = torch.randn((N, R, K), generator=gen) * 1.5 + PRIOR_LOGIT
    deltas = torch.randn((N, R, 4), generator=gen) * 0.2
    return dict(anchors=anchors, level_offsets=offs, logits=logits, deltas=deltas, N=N, R=R, K=K)


def matcher_stress_inputs(config_id, N, R, M):
    """Config 5: free-form anchors and GT (x1,y1 ~ U(0,1200)/(0,1000), sizes U(16,316)/(20,220))."""
    gen = torch.Generator().manual_seed(BASE_SEED + config_id)
    a_xy = torch.rand((N, R, 2), generator=gen) * 1200
    a_wh = torch.rand((N, R, 2), generator=gen) * 300 + 16
    anchors = torch.cat((a_xy, a_xy + a_wh), dim=2).to(torch.float32).contiguous()
    gt_boxes = []
    for _ in range(N):
        g_xy = torch.rand((M, 2), generator=gen) * 1000
        g_wh = torch.rand((M, 2), generator=gen) * 200 + 20
        gt_boxes.append(torch.cat((g_xy, g_xy + g_wh), dim=1).to(torch.float32))
    gt_classes = [torch.randint(0, 80, (M,), generator=gen, dtype=torch.int64) for _ in range(N)]
    return dict(anchors=anchors, gt_boxes=gt_boxes, gt_classes=gt_classes, N=N, R=R, M=M)


def rpn_inputs(config_id, N, level_counts, image_hw=(800, 1344), ties=True):
    """RPN head outputs after decoding: per level proposals (N, S_l, 4) scattered over (and partly outside) the
    image, some degenerate / tiny, and objectness logits ~ N(0, 2) with a few exact ties."""
    gen = torch.Generator().manual_seed(BASE_SEED + config_id)
    H, W = image_hw
    props, logits = [], []
    for l, cnt in enumerate(level_counts):
        size = 32.0 * 2 ** l
        cx = torch.rand((N, cnt), generator=gen) * (W + 100) - 50
        cy = torch.rand((N, cnt), generator=gen) * (H + 100) - 50
        w = size * torch.exp(torch.randn((N, cnt), generator=gen) * 0.5)
        h = size * torch.exp(torch.randn((N, cnt), generator=gen) * 0.5)
        w[:, ::37] = 2.0            # tiny boxes: dropped by min_box_side_len
        h[:, 5::41] = 0.0           # empty boxes
        p = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), dim=2).to(torch.float32)
        lg = (torch.randn((N, cnt), generator=gen) * 2.0).to(torch.float32)
        if ties and cnt > 64:
            lg[:, 11] = lg[:, 3]
            lg[:, cnt // 2] = lg[:, 3]
            lg[:, 20:28] = lg[:, 20:21]
        props.append(p.contiguous())
        logits.append(lg.contiguous())
    return dict(proposals=props, logits=logits, image_sizes=[(H, W)] * N, N=N)


def fast_rcnn_inputs(config_id, R, K, class_specific=True, image_hw=(800, 1344)):
    """Box-head outputs of one image: boxes (R, K*4) or (R, 4) around the image, softmax scores (R, K+1)."""
    gen = torch.Generator().manual_seed(BASE_SEED + config_id)
    H, W = image_hw
    C = K if class_specific else 1
    cx = torch.rand((R, 1), generator=gen) * (W + 60) - 30 + torch.randn((R, C), generator=gen) * 4
    cy = torch.rand((R, 1), generator=gen) * (H + 60) - 30 + torch.randn((R, C), generator=gen) * 4
    w = torch.exp(torch.rand((R, 1), generator=gen) * 3 + 2.5) * torch.exp(torch.randn((R, C), generator=gen) * 0.1)
    h = torch.exp(torch.rand((R, 1), generator=gen) * 3 + 2.5) * torch.exp(torch.randn((R, C), generator=gen) * 0.1)
    boxes = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), dim=2).reshape(R, C * 4).to(torch.float32)
    scores = torch.softmax(torch.randn((R, K + 1), generator=gen) * 2.5, dim=1).to(torch.float32)
    return dict(boxes=boxes.contiguous(), scores=scores.contiguous(), image_shape=(H, W))
