import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg
from tests.test_gpu_parity import _train_inputs
from oracle import dense_oracle as orc
cuda = torch.device("cuda:0")
N, K = 3, 80
base = _train_inputs(63, N, 256, 320, K, M=12)
R = base["R"]
x, d, b = (base[k].to(cuda) for k in ("logits", "deltas", "bets"))
anchors = base["anchors"].to(cuda)
cfg = fsg.DenseLossConfig(num_classes=K)
clean = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), max_total_gt=256)
boxes, classes = base["gt_boxes"], base["gt_classes"]
far = torch.tensor([[9000.0, 9000.0, 9100.0, 9050.0]])
rounds = [
    ([bx[:2] for bx in boxes], [c[:2] for c in classes]),
    (boxes, classes),
    ([torch.cat((bx[:1], far)) if bx.shape[0] else bx for bx in boxes],
     [torch.cat((c[:1], c[:1])) if c.shape[0] else c for c in classes]),
    ([bx[:0] for bx in boxes], [c[:0] for c in classes]),
    ([bx[:5] for bx in boxes], [c[:5] for c in classes]),
    (boxes, classes),
]
for rep in range(int(os.environ.get("REPS", "2"))):
  for i, (gb, gc) in enumerate(rounds):
    gt = fsg.ops.PackedGT.from_lists(gb, gc, cuda)
    fresh = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), max_total_gt=256)
    fresh._mc.workspace_is_clean = 0
    fresh.ws_step.fill_(0xAB)
    rc = clean.run(x, d, b, anchors, gt); gc_c = rc.gt_classes.clone(); st_c = rc.stats.clone()
    rf = fresh.run(x, d, b, anchors, gt); gc_f = rf.gt_classes.clone(); st_f = rf.stats.clone()
    torch.cuda.synchronize()
    want = orc.ground_truth(base["anchors"] if base["anchors"].dim() == 2 else list(base["anchors"]), gb, gc, K)["gt_classes"]
    print(rep, i, "M", [int(t.shape[0]) for t in gb], "clean!=oracle", int((gc_c.cpu() != want).sum()), "fresh!=oracle", int((gc_f.cpu() != want).sum()),
          "nf", float(st_c[0]), float(st_f[0]), flush=True)
