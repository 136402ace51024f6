"""CPU: the call boundary is pinned against the reference's own callables.

``tests/golden/signatures.json`` holds (name, default, kind) of every parameter of the reference functions this
package replaces, read with ``inspect.signature`` from the reference's source files through ``oracle/ref_loader.py``
(``oracle/make_golden.py``).  The drop-ins must take the same parameters, in the same order, with the same defaults;
anything extra has to be keyword-only with a default (an extension a reference caller never sees).  Where the
reference tree is present (the build container) the JSON itself is re-checked against the live reference."""
import inspect
import json
import os

import pytest

import full_scale_gambler_for_object_detection_b200 as fsg
from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "signatures.json")))

OURS = {
    "structures.Boxes.__init__": fsg.Boxes.__init__,
    "structures.Instances.__init__": fsg.Instances.__init__,
    "structures.pairwise_iou": fsg.pairwise_iou,
    "modeling.matcher.Matcher.__init__": fsg.Matcher.__init__,
    "modeling.matcher.Matcher.__call__": fsg.Matcher.__call__,
    "modeling.box_regression.Box2BoxTransform.__init__": fsg.Box2BoxTransform.__init__,
    "modeling.box_regression.Box2BoxTransform.get_deltas": fsg.Box2BoxTransform.get_deltas,
    "modeling.box_regression.Box2BoxTransform.apply_deltas": fsg.Box2BoxTransform.apply_deltas,
    "layers.nms.batched_nms": fsg.batched_nms,
    "layers.nms.nms": fsg.nms,
    "meta_arch.retinanet.RetinaNet.losses": fsg.RetinaNetDensePath.losses,
    "meta_arch.retinanet.RetinaNet.get_ground_truth": fsg.RetinaNetDensePath.get_ground_truth,
    "meta_arch.retinanet.RetinaNet.get_picky_ground_truth": fsg.RetinaNetDensePath.get_picky_ground_truth,
    "meta_arch.retinanet.RetinaNet.inference": fsg.RetinaNetDensePath.inference,
    "meta_arch.retinanet.RetinaNet.inference_single_image": fsg.RetinaNetDensePath.inference_single_image,
    "gambler_heads.LayeredUnetGambler.gambler_loss": fsg.GamblerLoss.gambler_loss,
    "gambler_heads.get_loss_upper_bound": fsg.get_loss_upper_bound,
    "proposal_generator.rpn_outputs.find_top_rpn_proposals": fsg.find_top_rpn_proposals,
    "modeling.sampling.subsample_labels": fsg.subsample_labels,
    "modeling.postprocessing.detector_postprocess": fsg.detector_postprocess,
    "roi_heads.fast_rcnn.fast_rcnn_inference": fsg.fast_rcnn_inference,
    "roi_heads.fast_rcnn.fast_rcnn_inference_single_image": fsg.fast_rcnn_inference_single_image,
    "modeling.anchor_generator.DefaultAnchorGenerator.grid_anchors": fsg.DefaultAnchorGenerator.grid_anchors,
}


def describe(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, d, p.kind.name])
    return out


def test_every_pinned_reference_callable_has_a_drop_in():
    assert sorted(OURS) == sorted(PINS)


@pytest.mark.parametrize("name", sorted(OURS))
def test_drop_in_signature_matches_the_reference(name):
    want, got = PINS[name], describe(OURS[name])
    assert got[:len(want)] == want, "%s: reference %s, drop-in %s" % (name, want, got[:len(want)])
    for extra in got[len(want):]:      # extensions must be invisible to a reference caller
        assert extra[2] == "KEYWORD_ONLY" and extra[1] is not None, "%s: extra parameter %s" % (name, extra)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_pins_match_the_live_reference():
    from oracle.make_golden import reference_callables

    live = {k: describe(f) for k, f in reference_callables(ref_loader.load_reference()).items()}
    assert live == PINS
