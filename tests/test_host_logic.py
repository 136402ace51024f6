"""CPU: host-side pieces of the product (containers, config mapping, planning arithmetic, anchors)."""
import pytest
import torch

import full_scale_gambler_for_object_detection_b200 as fsg
from full_scale_gambler_for_object_detection_b200 import _lib, anchor_generator, sharded, synthetic


def test_boxes_container_semantics():
    from full_scale_gambler_for_object_detection_b200 import structures as st

    b = fsg.Boxes(torch.tensor([[0.0, 0.0, 2.0, 3.0], [1.0, 1.0, 2.0, 2.0]], dtype=torch.float64))
    assert b.tensor.dtype == torch.float32 and len(b) == 2           # boxes.py:91-95
    assert fsg.Boxes(torch.zeros(0)).tensor.shape == (0, 4)
    assert len(b[0]) == 1 and len(b[torch.tensor([True, False])]) == 1
    assert st.cat_tensors([b]).data_ptr() == b.tensor.data_ptr()      # single-element shortcut (wrappers.py:15-22)
    assert st.cat_tensors([b, b.tensor]).shape == (4, 4)              # Boxes-likes and tensors mix
    with pytest.raises(AssertionError):
        fsg.Boxes(torch.zeros(3, 5))


def test_containers_are_duck_typed_and_pluggable():
    """Anything with ``.tensor`` is a Boxes, results are built with whatever classes the integrator registers."""
    from full_scale_gambler_for_object_detection_b200 import structures as st

    class TheirBoxes:
        def __init__(self, tensor):
            self.tensor = tensor

        def __len__(self):
            return self.tensor.shape[0]

    class TheirInstances(fsg.Instances):
        pass

    t = torch.zeros(3, 4)
    assert st.as_tensor(TheirBoxes(t)) is t and st.as_tensor(t) is t
    try:
        st.use_containers(TheirBoxes, TheirInstances)
        inst = st.make_instances((4, 5), pred_boxes=st.make_boxes(t), scores=torch.zeros(3))
        assert isinstance(inst, TheirInstances) and isinstance(inst.pred_boxes, TheirBoxes) and len(inst) == 3
    finally:
        st.use_containers()
    assert isinstance(st.make_boxes(t), fsg.Boxes)


def test_instances_container():
    t = fsg.Instances((4, 5))
    t.gt_boxes = fsg.Boxes(torch.zeros(3, 4))
    t.gt_classes = torch.arange(3)
    assert len(t) == 3 and t.image_size == (4, 5) and t.has("gt_classes")
    with pytest.raises(AssertionError):
        t.bad = torch.zeros(2)
    with pytest.raises(AttributeError):
        t.missing
    assert len(t[torch.tensor([0, 2])]) == 2


def test_matcher_constructor_asserts_like_the_reference():
    m = fsg.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True)
    assert m.thresholds == [-float("inf"), 0.4, 0.5, float("inf")]     # matcher.py:44-47
    with pytest.raises(AssertionError):
        fsg.Matcher([0.0, 0.5], [0, -1, 1])
    with pytest.raises(AssertionError):
        fsg.Matcher([0.6, 0.5], [0, -1, 1])
    with pytest.raises(AssertionError):
        fsg.Matcher([0.4, 0.5], [0, 2, 1])
    with pytest.raises(AssertionError):
        fsg.Matcher([0.4, 0.5], [0, 1])


def test_loss_config_mapping():
    c = fsg.DenseLossConfig()
    assert c.norm_mode == _lib.NORM_IMAGE
    assert fsg.DenseLossConfig(gambler_output="L_BAHW_extendtobatch").norm_mode == _lib.NORM_BATCH
    assert fsg.DenseLossConfig(normalize=False).norm_mode == _lib.NORM_NONE
    p = c.loss_params(1.0, 0.5, -2.0)
    assert (p.num_classes, p.gambler_mode, p.norm_mode) == (80, 0, 1)
    assert abs(p.c_reg - 0.5) < 1e-7 and abs(p.c_gam + 2.0) < 1e-7 and abs(p.temperature - 0.1) < 1e-7
    with pytest.raises(ValueError):
        fsg.DenseLossConfig(gambler_output="L_B1HW").loss_params(1, 1, 1)
    with pytest.raises(NotImplementedError):
        fsg.GamblerLoss(gambler_output="L_B1HW")


def test_retinanet_grids_match_baseline_configs():
    """config 1 R = 16368, config 2 R = 67200 with level split 50400/12600/3150/819/231."""
    a1, offs1, g1 = anchor_generator.retinanet_anchors(512, 512)
    assert a1.shape == (16368, 4) and g1 == [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]
    a2, offs2, g2 = anchor_generator.retinanet_anchors(800, 1333)
    assert a2.shape == (67200, 4)
    assert [offs2[i + 1] - offs2[i] for i in range(5)] == [50400, 12600, 3150, 819, 231]
    assert g2[0] == (100, 168) and g2[-1] == (7, 11)


def test_synthetic_inputs_are_seeded_and_shaped():
    a = synthetic.train_inputs(1, 2, 128, 128, 5, M=3)
    b = synthetic.train_inputs(1, 2, 128, 128, 5, M=3)
    assert torch.equal(a["logits"], b["logits"]) and torch.equal(a["gt_boxes"][0], b["gt_boxes"][0])
    assert a["gt_boxes"][-1].shape == (0, 4)                           # one GT-free image
    assert a["logits"].shape == (2, a["R"], 5) and a["bets"].min() > 0 and a["bets"].max() < 1
    gb = a["gt_boxes"][0]
    assert (gb[:, 2] >= gb[:, 0]).all() and gb.min() >= 0 and gb.max() <= 128


def test_image_shard():
    assert sharded.image_shard(16, 4, 0) == slice(0, 4) and sharded.image_shard(16, 4, 3) == slice(12, 16)
    with pytest.raises(ValueError):
        sharded.image_shard(10, 4, 0)


def test_global_losses_single_process():
    scalars = torch.tensor([8.0, 2.0, 0.5, 3.0, 1.0, 0, 0, 0, 0, 0], dtype=torch.float64)
    stats = torch.tensor([4.0, 10.0], dtype=torch.float64)
    out = sharded.global_losses(scalars, stats, (1.0, 2.0, -1.0))
    assert torch.allclose(out, torch.tensor([2.0, 0.5, -0.5, 2.0 + 1.0 + 0.5], dtype=torch.float64))


def test_get_loss_upper_bound_matches_oracle_definition():
    torch.manual_seed(0)
    N = 2
    levels = [torch.rand(N, 3, h, h) for h in (8, 4, 2, 2, 2)]
    got = fsg.get_loss_upper_bound(levels, N, 0.1, 1.5)
    total = sum(3 * l.shape[2] * l.shape[3] for l in levels)
    per_img = torch.stack([torch.stack([l[n].max() for l in levels]).max() for n in range(N)])
    want = 1.5 * (1 + 0.1) / (total * 0.1 + 1) * N * per_img.sum()
    assert torch.allclose(got, want)
    with pytest.raises(AssertionError):
        fsg.get_loss_upper_bound(levels[:4], N, 0.1, 1.0)


def test_anchor_range_partitions_exactly():
    """sharded.anchor_range: contiguous, disjoint, covering, sizes differing by at most one."""
    for R, world in ((1000000, 8), (6001, 2), (7, 8), (0, 3), (67200, 5)):
        ranges = [sharded.anchor_range(R, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == R
        for (lo, hi), (lo2, _) in zip(ranges[:-1], ranges[1:]):
            assert hi == lo2 and lo <= hi
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_postprocess_rows_and_struct_sizes():
    """The (N,4) table detect() takes for the fused detector_postprocess (postprocessing.py:27: scale = out / in)."""
    import ctypes

    rows = fsg.ops.postprocess_rows([(800, 1344), (512, 512)], [(480, 640), (1024, 768)], "cpu")
    want = torch.tensor([[640 / 1344, 480 / 800, 640.0, 480.0], [768 / 512, 1024 / 512, 768.0, 1024.0]])
    assert torch.equal(rows, want.to(torch.float32))
    # ctypes mirrors of the header's structs (include/fsg_dense.h)
    assert ctypes.sizeof(_lib.HeadLevel) == 6 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.BetLevels) == 8 * 8 + 2 * 8 * 4 + 2 * 4
    assert ctypes.sizeof(_lib.PostLevel) == 3 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.AnchorLevel) == 4 * 4 + 16 * 4 * 4


def test_default_anchor_generator_host_side():
    """Cell anchors and broadcasting of sizes / ratios over levels (anchor_generator.py:83-96,131-168); the grid
    itself is produced on the device (GPU test)."""
    gen = fsg.DefaultAnchorGenerator([[32, 64]], [[0.25, 1, 4]], [4, 8], device="cpu")
    assert gen.num_cell_anchors == [6, 6] and gen.box_dim == 4
    want = torch.tensor([[-32.0, -8.0, 32.0, 8.0], [-16.0, -16.0, 16.0, 16.0], [-8.0, -32.0, 8.0, 32.0],
                         [-64.0, -16.0, 64.0, 16.0], [-32.0, -32.0, 32.0, 32.0], [-16.0, -64.0, 16.0, 64.0]])
    assert torch.allclose(gen.cell_anchors[0], want)          # tests/test_anchor_generator.py:14-43 (cell part)
    gen2 = fsg.DefaultAnchorGenerator(anchor_generator.RETINANET_SIZES, ((1.0,),), anchor_generator.RETINANET_STRIDES,
                                      device="cpu")
    assert gen2.num_cell_anchors == [3] * 5
