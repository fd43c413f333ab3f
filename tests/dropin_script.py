"""Run by tests/test_dropin_reference_cpu.py in a subprocess (the reference's top-level package is
called `model`; keeping it out of the test process avoids polluting sys.modules).

argv: <reference dir> <work dir> <variant A|B>. Copies the reference's `model/` package into the
work dir, applies the import swap INTEGRATION.md §1 describes to its model/model.py (and nothing
else), builds the REFERENCE's HeroForVcmr (model/vcmr.py + model/pretrain.py, unmodified) on top of
it, and checks its outputs against the goldens of the unmodified reference. CPU: hero_b200.ops is
routed through tests/fake_ops.py.
"""
import json
import os
import shutil
import sys

ref, work, variant = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from baseline import ref_runner  # noqa: E402
from tests import fake_ops, golden_util as gu  # noqa: E402

pkg = os.path.join(work, "refpkg")
shutil.copytree(os.path.join(ref, "model"), os.path.join(pkg, "model"))
path = os.path.join(pkg, "model", "model.py")
src = open(path).read()
old_imports = """from .encoder import (
    RobertaModelConfig, RobertaPreTrainedModel)
from .encoder import CrossModalTrm
from .encoder import TemporalTrm
from .layers import (GELU, LinearLayer, MLPLayer)
"""
assert old_imports in src, "reference model/model.py import block changed"
new_imports = """from hero_b200.encoder import (
    RobertaModelConfig, RobertaPreTrainedModel)
from hero_b200.encoder import CrossModalTrm
from hero_b200.encoder import TemporalTrm
from hero_b200.layers import (GELU, LinearLayer, MLPLayer)
"""
src = src.replace(old_imports, new_imports)
if variant == "B":      # ... and re-export the packed-path classes
    src += "\nfrom hero_b200.model import HierarchicalVlModel, HeroModel, VideoModelConfig  # noqa\n"
open(path, "w").write(src)

ref_runner.REF_DIR = pkg
ref_runner.install(dist_backed=False)


class _MP:       # minimal monkeypatch stand-in for fake_ops.install
    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


fake_ops.install(_MP)
from model.vcmr import HeroForVcmr  # noqa: E402  (the reference's class)
from model.model import VideoModelConfig  # noqa: E402
import model.model as ref_model_py  # noqa: E402
import hero_b200.encoder as our_enc  # noqa: E402

assert ref_model_py.CrossModalTrm is our_enc.CrossModalTrm
fx, vx = gu.load("hier_tiny.npz"), gu.load("vsm_tiny.npz")
d = gu.dims_of(fx)


def cfg(n, v):
    c = {"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1,
         "hidden_size": d["hidden"], "initializer_range": 0.02, "intermediate_size": d["inter"],
         "max_position_embeddings": 514, "num_attention_heads": d["heads"],
         "num_hidden_layers": n, "type_vocab_size": 2}
    if v:
        c["vocab_size"] = d["vocab"]
    return c


cpath = os.path.join(work, "m.json")
json.dump({"f_config": cfg(d["f_layers"], True), "c_config": cfg(d["c_layers"], False),
           "q_config": json.loads(str(vx["q_config"]))}, open(cpath, "w"))
model = HeroForVcmr(VideoModelConfig(cpath), vfeat_dim=d["vfeat_dim"],
                    max_frm_seq_len=d["max_img_len"], lw_neg_ctx=8, lw_neg_q=8, lw_st_ed=0.01,
                    margin=0.1)
sd = {"v_encoder." + k: v for k, v in gu.weights_for(fx).items()}
sd.update({k[5:]: torch.from_numpy(v) for k, v in vx.items() if k.startswith("head.")})
missing, unexpected = model.load_state_dict(sd, strict=False)
assert not unexpected, unexpected
assert not [k for k in missing if k.startswith(("video_", "q_feat_attn", "v_encoder.f_encoder.enc",
                                                "v_encoder.c_encoder.enc"))], missing
model.eval()
vb, _ = gu.stored_batches(fx)
batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in vb.items()}
for k in ("query_input_ids", "query_pos_ids", "query_attn_masks", "targets", "q_vidx"):
    batch[k] = torch.from_numpy(vx[k])
with torch.no_grad():
    clip = model.v_encoder(batch, "repr")
    scores, st, ed = model(batch, "tvr", compute_loss=False)
cm = vb["c_attn_masks"].bool().numpy()
err_clip = float(np.abs(clip.float().numpy()[cm] - fx["clip_out"][cm]).max())
err_scores = float(np.abs(scores.float().numpy() - vx["scores"]).max())
print(json.dumps({"variant": variant, "clip_err": err_clip, "score_err": err_scores,
                  "encoder_class": type(model.v_encoder.f_encoder).__module__,
                  "model_class": type(model.v_encoder).__module__}))
assert err_clip < 6e-2 and err_scores < 2e-2, (err_clip, err_scores)
