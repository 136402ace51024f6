"""Loading of the committed reference outputs (tests/golden/*.npz, made by oracle/make_golden.py)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN_CASES = ["train_config1", "train_gphase", "train_extendtobatch", "train_sigmoid", "train_nonorm",
               "train_lvis_k1230"]
ORACLE_KW = {"GAMBLER_OUTPUT": "output", "GAMBLER_LOSS_MODE": "mode", "NORMALIZE": "normalize"}
CFG_KW = {"GAMBLER_OUTPUT": "gambler_output", "GAMBLER_LOSS_MODE": "gambler_loss_mode", "NORMALIZE": "normalize"}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" and z[k].ndim > 0 else z[k]) for k in z.files}


def train_case(name):
    """-> (inputs regenerated from the stored seed parameters, reference outputs, coeffs, detach, gcfg)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = load(name)
    cid, N, H, W, K, M = [int(v) for v in g["params"]]
    inp = synthetic.train_inputs(cid, N, H, W, K, M=M)
    coeffs = tuple(float(c) for c in g["coeffs"])
    gcfg = ast.literal_eval(str(g["gcfg"]))
    return inp, g, coeffs, bool(int(g["detach"])), gcfg, K
