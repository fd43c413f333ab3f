"""Helpers shared by the parity tests: load golden fixtures, rebuild their weights/inputs."""
import json
import os

import numpy as np
import torch

from hero_b200 import synth
from oracle import hero_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def dims_of(fx):
    return json.loads(str(fx["dims"]))


def weights_for(fx):
    d = dims_of(fx)
    shapes = orc.param_shapes(d["hidden"], d["inter"], d["f_layers"], d["c_layers"], d["vocab"],
                              514, 2, d["vfeat_dim"], d["max_img_len"])
    std = float(fx["weight_std"]) if "weight_std" in fx else 0.02
    return orc.seeded_weights(shapes, seed=int(fx["seed_weights"]), std=std)


def stored_batches(fx):
    vb = {k[3:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("vb.")}
    qb = {k[3:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("qb.")}
    vb["num_subs"] = json.loads(str(fx["num_subs"]))
    vb["sub_idx2frame_idx"] = [[(s, fr) for s, fr in clip]
                               for clip in json.loads(str(fx["sub_idx2frame_idx"]))]
    return vb, qb


def full_small_batches(fx):
    return synth.syn_tvr_ragged(batch_size=2, seed=int(fx["seed_batch"]), t_range=(10, 16),
                                s_range=(3, 5), l_range=(4, 12), q_range=(5, 9))
