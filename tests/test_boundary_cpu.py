"""Drop-in boundary checks that need no GPU: state_dict compatibility with the reference,
C-ABI symbol coverage, and the no-CPU-fallback rule."""
import ctypes
import json
import os
import re

import pytest
import torch

from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_json(tmp_path, f_layers=6, c_layers=3, hidden=768, inter=3072, heads=12, vocab=50272):
    def cfg(n, v):
        c = {"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1,
             "hidden_size": hidden, "initializer_range": 0.02, "intermediate_size": inter,
             "max_position_embeddings": 514, "num_attention_heads": heads,
             "num_hidden_layers": n, "type_vocab_size": 2}
        if v:
            c["vocab_size"] = vocab
        return c
    p = tmp_path / "model.json"
    p.write_text(json.dumps({"f_config": cfg(f_layers, True), "c_config": cfg(c_layers, False)}))
    return str(p)


def test_state_dict_keys_and_shapes_match_reference(tmp_path):
    """Fixture written by oracle/gen_golden.py from the reference's HierarchicalVlModel."""
    from hero_b200.model import HierarchicalVlModel, VideoModelConfig
    ref = json.load(open(os.path.join(gu.GOLDEN, "state_dict_keys.json")))
    model = HierarchicalVlModel(VideoModelConfig(_model_json(tmp_path)), vfeat_dim=4352,
                                max_frm_seq_len=100)
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert set(mine) == set(ref), (sorted(set(ref) - set(mine)), sorted(set(mine) - set(ref)))
    for k in ref:
        assert mine[k] == ref[k], k
    # tied LM head decoder <-> word embeddings, as in model/layers.py:345
    assert model.f_encoder.lm_head.decoder.weight is model.f_encoder.embeddings.word_embeddings.weight


def test_header_symbols_exported_and_bound():
    from hero_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hero_b200.h")).read()
    declared = set(re.findall(r"\b(hero_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hero_b200.h but not exported"
    bound = _lib.lib()
    for name in declared:
        fn = getattr(bound, name)
        assert fn.argtypes is not None or name in ("hero_version", "hero_sm_count"), name


def test_struct_mirrors_match_header_field_order():
    from hero_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hero_b200.h")).read()
    for cname, cls in (("hero_gemm_args", _lib.GemmArgs), ("hero_ln_args", _lib.LnArgs),
                       ("hero_layer_weights", _lib.LayerWeights),
                       ("hero_layer_acts", _lib.LayerActs), ("hero_layer_grads", _lib.LayerGrads),
                       ("hero_stack_args", _lib.StackArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [f[0] for f in cls._fields_], cname


def test_no_cpu_fallback(tmp_path):
    from hero_b200 import _lib
    from hero_b200.model import HierarchicalVlModel, VideoModelConfig
    from hero_b200 import synth
    model = HierarchicalVlModel(
        VideoModelConfig(_model_json(tmp_path, 1, 1, 128, 256, 2, 120)), vfeat_dim=64,
        max_frm_seq_len=20).eval()
    vb, _ = synth.syn_tvr_ragged(batch_size=2, seed=1, vfeat_dim=64, vocab=100, t_range=(6, 9),
                                 s_range=(2, 3), l_range=(3, 5))
    with pytest.raises(_lib.HeroError):
        model(vb, "repr")


def test_load_pretrained_weight_key_conventions(tmp_path):
    from hero_b200.encoder import CrossModalTrm, RobertaModelConfig, load_pretrained_weight
    cfg = RobertaModelConfig(120, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                             intermediate_size=256, max_position_embeddings=32)
    m = CrossModalTrm(cfg, vfeat_dim=64, max_img_seq_len=10)
    sd = {"roberta.embeddings.LayerNorm.gamma": torch.full((128,), 3.0),
          "roberta.embeddings.LayerNorm.beta": torch.full((128,), -2.0),
          "roberta.unknown.weight": torch.zeros(3)}
    load_pretrained_weight(m, sd)
    assert torch.all(m.embeddings.LayerNorm.weight == 3.0)
    assert torch.all(m.embeddings.LayerNorm.bias == -2.0)
    with pytest.raises(RuntimeError):
        load_pretrained_weight(m, {"embeddings.LayerNorm.weight": torch.zeros(5)})


def test_unknown_task_raises_value_error(tmp_path):
    from hero_b200.model import HierarchicalVlModel, VideoModelConfig
    model = HierarchicalVlModel(
        VideoModelConfig(_model_json(tmp_path, 1, 1, 128, 256, 2, 120)), vfeat_dim=64,
        max_frm_seq_len=20)
    with pytest.raises(ValueError):
        model({}, "nope")
    with pytest.raises(ValueError):
        model.f_encoder({}, "nope")
