import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg
from full_scale_gambler_for_object_detection_b200 import synthetic
dev = torch.device('cuda')
N, K = 16, 80
inp = synthetic.train_inputs(2, N, 800, 1333, K, M=8, logits=False)
A, grids, R = inp["A"], inp["grids"], inp["R"]
g = torch.Generator(device='cuda').manual_seed(1)
cls_l = [torch.randn((N, A*K, h, w), device=dev, generator=g) - 4.595 for h, w in grids]
reg_l = [torch.randn((N, A*4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
bets = torch.sigmoid(torch.randn((N, R), device=dev, generator=g) - 4.595)
anchors = inp["anchors"].to(dev)
gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
cfg = fsg.DenseLossConfig(num_classes=K)
params = cfg.loss_params(1.0, 1.0, -1.0)
m = fsg.ops.match_anchors(anchors, gt, K, bets=bets, temperature=0.1)
def run():
    return fsg.ops.loss_main_levels(cls_l, m["gt_classes"], params, m["stats"], delta_levels=reg_l, anchors=anchors, gt=gt,
                                    matched_idx32=m["matched_idx32"], mask=m["mask"], bets=bets)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): run()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(10): keep = run()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): gr.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/50
print("levels main: %.1f us  %.0f GB/s (%.1f%% of 6461)" % (ms*1e3, (8*K+72)*N*R/ms/1e6, (8*K+72)*N*R/ms/1e6/64.612))
