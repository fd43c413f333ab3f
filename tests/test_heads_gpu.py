"""The heads that sit on the encoder on the GPU, with the real
kernels, against outputs of the unmodified reference: VSM / VCMR (tests/golden/vsm_tiny.npz) and
MFM regression / NCE, FOM, MLM (tests/golden/heads_tiny.npz). Same checks as the CPU tests
(tests/test_heads_cpu.py, which run the host logic on torch restatements of the C-ABI); here the
encoder, the fused VSM kernels (hero_b200/csrc/vsm.cu) and the fused LM-head cross-entropy run on
the device. Plus kernel-level forward/backward checks of the fused head ops against torch autograd.
"""
import json
import math

import numpy as np
import pytest
import torch

from hero_b200 import synth
from tests import golden_util as gu
from tests.test_heads_cpu import _clone, _model_with_heads, _vcmr_model

pytestmark = pytest.mark.gpu


def _cuda(b):
    return synth.to_device(b, "cuda")


def test_vsm_head_on_gpu_matches_reference_golden(tmp_path):
    model, vb, vx = _vcmr_model(tmp_path)
    model = model.cuda().eval()

    def batch():
        b = _clone(vb)
        for k in ("query_input_ids", "query_pos_ids", "query_attn_masks", "targets", "q_vidx"):
            b[k] = torch.from_numpy(vx[k])
        return _cuda(b)

    with torch.no_grad():
        scores, st, ed = model(batch(), "tvr", compute_loss=False)
        l_st_ed, l_ctx, l_q = model(batch(), "tvr", compute_loss=True)
        model.set_hard_negative(True, 2, 10)
        _, h_ctx, h_q = model(batch(), "tvr", compute_loss=True)
        model.set_hard_negative(False, 20, 10)
        model.ranking_loss_type = "lse"
        _, e_ctx, e_q = model(batch(), "tvr", compute_loss=True)
    assert np.abs(scores.cpu().numpy() - vx["scores"]).max() < 2e-2
    cm = vb["c_attn_masks"].bool().numpy()
    n = cm.sum(1)
    inner = cm.copy()
    for i, ni in enumerate(n):       # documented deviation: last two frames of shorter clips
        if ni < cm.shape[1]:
            inner[i, max(ni - 2, 0):] = False
    inner = np.broadcast_to(inner[None], st.shape)
    for got, ref in ((st, vx["st_prob"]), (ed, vx["ed_prob"])):
        err = np.abs(got.float().cpu().numpy() - ref)[inner]
        assert err.max() <= 2e-2 * np.abs(ref[inner]).max(), float(err.max())
    for got, key in ((l_ctx, "loss_neg_ctx"), (l_q, "loss_neg_q"), (h_ctx, "hard_neg_ctx"),
                     (h_q, "hard_neg_q"), (e_ctx, "lse_neg_ctx"), (e_q, "lse_neg_q")):
        ref = vx[key]
        assert np.abs(got.float().cpu().numpy() - ref).max() < 0.15 + 0.03 * np.abs(ref).max(), key

    # equal-length clips, one query per clip: the fused span-logit kernel (non-cross path)
    vbd = {k[4:]: torch.from_numpy(v) for k, v in vx.items() if k.startswith("vbd.")}
    qbd = {k[4:]: torch.from_numpy(v) for k, v in vx.items() if k.startswith("qbd.")}
    vbd["num_subs"] = json.loads(str(vx["d_num_subs"]))
    vbd["sub_idx2frame_idx"] = [[(s_, fr) for s_, fr in clip]
                                for clip in json.loads(str(vx["d_sub_idx2frame_idx"]))]
    b = _clone(vbd)
    b.update(query_input_ids=qbd["input_ids"], query_pos_ids=qbd["pos_ids"],
             query_attn_masks=qbd["attn_masks"], targets=torch.from_numpy(vx["d_targets"]))
    model.ranking_loss_type = "hinge"
    with torch.no_grad():
        d_scores, d_st, d_ed = model(_cuda(b), "tvr", compute_loss=False)
        d_l_st_ed, d_l_ctx, d_l_q = model(_cuda(_clone(b)), "tvr", compute_loss=True)
    assert np.abs(d_scores.cpu().numpy() - vx["d_scores"]).max() < 2e-2
    for got, ref in ((d_st, vx["d_st"]), (d_ed, vx["d_ed"])):
        assert np.abs(got.float().cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    assert abs(float(d_l_st_ed.sum()) - float(vx["d_loss_st_ed"].sum())) < 2e-2
    assert np.abs(d_l_ctx.cpu().numpy() - vx["d_loss_neg_ctx"]).max() < 0.15
    assert np.abs(d_l_q.cpu().numpy() - vx["d_loss_neg_q"]).max() < 0.15


def test_vsm_training_step_never_synchronises(tmp_path):
    """Training-mode VSM forward + backward on the GPU with collate-side plans: no device->host
    read anywhere (torch sync-debug mode raises on one)."""
    from hero_b200.plan import attach_plan
    model, vb, vx = _vcmr_model(tmp_path)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    vbd = {k[4:]: torch.from_numpy(v) for k, v in vx.items() if k.startswith("vbd.")}
    qbd = {k[4:]: torch.from_numpy(v) for k, v in vx.items() if k.startswith("qbd.")}
    vbd["num_subs"] = json.loads(str(vx["d_num_subs"]))
    vbd["sub_idx2frame_idx"] = [[(s_, fr) for s_, fr in clip]
                                for clip in json.loads(str(vx["d_sub_idx2frame_idx"]))]

    def make():
        b = _clone(vbd)
        b.update(query_input_ids=qbd["input_ids"], query_pos_ids=qbd["pos_ids"],
                 query_attn_masks=qbd["attn_masks"], targets=torch.from_numpy(vx["d_targets"]))
        return _cuda(attach_plan(b, kind="vsm"))

    losses = model(make(), "tvr", compute_loss=True)      # warm-up (flat buffers, caches)
    sum(l.sum() for l in losses).backward()
    torch.cuda.synchronize()
    b = make()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        losses = model(b, "tvr", compute_loss=True)
        sum(l.sum() for l in losses).backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert all(torch.isfinite(l).all() for l in losses)


def test_pretraining_heads_on_gpu_match_reference_golden(tmp_path):
    """MFM regression / NCE (model/model.py:239-289), FOM (:306-336) and MLM
    (model/encoder.py:355-374) with the encoder and the fused LM-head loss on the GPU."""
    model, vb, hx = _model_with_heads(tmp_path)
    model = model.cuda().eval()

    def mfm_batch():
        b = _clone(vb)
        b["c_v_masks"] = torch.from_numpy(hx["c_v_masks"])
        b["feat_targets"] = torch.from_numpy(hx["feat_targets"])
        return _cuda(b)

    with torch.no_grad():
        pred = model(mfm_batch(), "mffr", compute_loss=False)
        loss = model(mfm_batch(), "mffr", compute_loss=True)
        nce = model(mfm_batch(), "mfm-nce", compute_loss=True)
    assert np.abs(pred.float().cpu().numpy() - hx["mffr_pred"]).max() < 4e-2
    assert np.abs(loss.float().cpu().numpy() - hx["mffr_loss"]).max() < 4e-2
    assert np.abs(nce.float().cpu().numpy() - hx["nce_loss"]).max() < 6e-2

    b = _clone(vb)
    b["shuffled_orders"] = torch.from_numpy(hx["shuffled_orders"])
    b["targets"] = torch.from_numpy(hx["fom_targets"])
    with torch.no_grad():
        logits = model(_cuda(b), "fom", compute_loss=False)
        floss = model(_cuda(_clone(b)), "fom", compute_loss=True)
    valid = vb["c_attn_masks"].bool().reshape(-1).numpy()
    assert np.abs(logits.float().cpu().numpy()[valid] - hx["fom_logits"][valid]).max() < 6e-2
    assert abs(float(floss) - float(hx["fom_loss"])) < 2e-2

    mb = {"input_ids": vb["f_sub_input_ids"], "position_ids": vb["f_sub_pos_ids"],
          "v_feat": vb["f_v_feats"], "f_pos_ids": vb["f_v_pos_ids"],
          "attn_masks": vb["f_attn_masks"], "gather_index": vb["f_gather_index"],
          "txt_mask_tgt": torch.from_numpy(hx["txt_mask_tgt"]),
          "txt_labels": torch.from_numpy(hx["txt_labels"])}
    with torch.no_grad():
        scores = model.f_encoder(_cuda(mb), "mlm", compute_loss=False)
        mloss = model.f_encoder(_cuda(mb), "mlm", compute_loss=True)
    assert scores.shape == hx["mlm_scores"].shape
    assert np.abs(scores.float().cpu().numpy() - hx["mlm_scores"]).max() < 6e-2
    assert np.abs(mloss.float().cpu().numpy() - hx["mlm_loss"]).max() < 6e-2


# ------------------------------------------------------------------------------ fused head ops
def test_vsm_video_scores_kernels_forward_backward_vs_torch():
    from hero_b200 import functional as Fn
    g = torch.Generator().manual_seed(3)
    nq, nv, length, d = 37, 19, 23, 768
    q = torch.randn(nq, d, generator=g).cuda().requires_grad_(True)
    c = torch.randn(nv, length, d, generator=g).cuda().requires_grad_(True)
    mask = (torch.rand(nv, length, generator=g) < 0.8).long()
    mask[:, 0] = 1
    mask[3] = 0                       # a fully masked clip: score -1e4, no gradient
    mask = mask.cuda()
    w = torch.randn(nq, nv, generator=g).cuda()
    got = Fn.vsm_video_scores(q, c, mask)
    (got * w).sum().backward()
    gq, gc = q.grad.clone(), c.grad.clone()
    q.grad = c.grad = None
    qn = torch.nn.functional.normalize(q, dim=-1, eps=1e-5)
    cn = torch.nn.functional.normalize(c, dim=-1, eps=1e-5)
    s = torch.einsum("md,nld->mln", qn, cn)
    m = mask.transpose(0, 1).unsqueeze(0).float()
    ref = (s * m + (1 - m) * -1e4).max(dim=1).values
    (ref * w).sum().backward()
    assert (got - ref).abs().max().item() < 2e-4
    assert (gq - q.grad).abs().max().item() < 2e-3 * max(1.0, q.grad.abs().max().item())
    assert (gc - c.grad).abs().max().item() < 2e-3 * max(1.0, c.grad.abs().max().item())


def test_vsm_span_logits_kernels_forward_backward_vs_torch():
    from hero_b200 import functional as Fn
    g = torch.Generator().manual_seed(4)
    n, length, d, k = 9, 100, 768, 5
    q = torch.randn(n, d, generator=g).cuda().requires_grad_(True)
    c = (0.1 * torch.randn(n, length, d, generator=g)).cuda().requires_grad_(True)
    w_st = torch.randn(1, 1, k, generator=g).cuda().requires_grad_(True)
    w_ed = torch.randn(1, 1, k, generator=g).cuda().requires_grad_(True)
    mask = torch.ones(n, length, dtype=torch.long)
    for i in range(n):
        mask[i, 60 + 4 * i:] = 0
    mask = mask.cuda()
    a = torch.randn(n, length, generator=g).cuda()
    b = torch.randn(n, length, generator=g).cuda()
    st, ed = Fn.vsm_span_logits(q, c, mask, w_st, w_ed)
    ((st * a).sum() + (ed * b).sum()).backward()
    got = [t.grad.clone() for t in (q, c, w_st, w_ed)]
    for t in (q, c, w_st, w_ed):
        t.grad = None
    sim = torch.einsum("bd,bld->bl", q, c).unsqueeze(1)
    mf = mask.float()
    st_r = torch.nn.functional.conv1d(sim, w_st, padding=k // 2).squeeze(1) * mf + (1 - mf) * -1e4
    ed_r = torch.nn.functional.conv1d(sim, w_ed, padding=k // 2).squeeze(1) * mf + (1 - mf) * -1e4
    ((st_r * a).sum() + (ed_r * b).sum()).backward()
    assert (st - st_r).abs().max().item() < 1e-3 and (ed - ed_r).abs().max().item() < 1e-3
    for gt, t in zip(got, (q, c, w_st, w_ed)):
        assert (gt - t.grad).abs().max().item() < 2e-3 * max(1.0, t.grad.abs().max().item())


def test_video_corpus_embedding_pass_matches_per_batch_forward(tmp_path):
    """eval_vcmr.py:161-203: the forward-only corpus pass fills the same
    (n_videos, max_clip_len, H) tensor as calling the encoder batch by batch, with ragged clip
    lengths across batches and an explicit video order."""
    from hero_b200.evalpass import embed_video_corpus
    from tests.test_bench_path_gpu import DIMS, _build
    from oracle import hero_oracle as orc
    d = dict(DIMS, f_layers=2, c_layers=1)
    P = orc.seeded_weights(orc.param_shapes(f_layers=2, c_layers=1), seed=21)
    model = _build(tmp_path, d, P, train=False)
    batches, order, n = [], [], 0
    for i, (lo, hi) in enumerate([(20, 40), (10, 16), (30, 60)]):
        vb, _ = synth.syn_tvr_ragged(batch_size=3, seed=50 + i, t_range=(lo, hi), s_range=(3, 6),
                                     l_range=(4, 12))
        batches.append(vb)
        order.append([8 - n - j for j in range(3)])           # corpus rows in reverse order
        n += 3
    emb, masks = embed_video_corpus(model, batches, n_videos=9, max_clip_len=100,
                                    video_indices=order)
    longest = max(b["c_v_feats"].shape[1] for b in batches)
    assert emb.shape == (9, longest, 768) and masks.shape == (9, longest)
    with torch.no_grad():
        for vb, idx in zip(batches, order):
            want = model(synth.to_device(vb, "cuda"), "repr")
            T = want.shape[1]
            assert torch.equal(emb[torch.tensor(idx), :T], want)
            if T < emb.shape[1]:
                assert float(emb[torch.tensor(idx), T:].abs().max()) == 0.0
            assert torch.equal(masks[torch.tensor(idx), :T].cpu(), vb["c_attn_masks"])
    assert model.training is False


@pytest.mark.parametrize("n,vocab_pad", [(77, 7), (2400, 7), (300, 0)])
def test_fused_lm_head_cross_entropy_forward_backward_vs_torch(n, vocab_pad):
    """Vocabulary GEMM + online-softmax cross entropy (no logits tensor) against torch on
    materialised fp32 logits from the same bf16 operands: loss, d h, d E (tied embedding), d bias.
    V = 50272 is not a multiple of 64 (MN-major wgrad operand with a padded row stride) and the
    last `vocab_pad` columns are vocabulary padding (excluded from the softmax, zero gradient)."""
    from hero_b200 import functional as Fn
    g = torch.Generator().manual_seed(9)
    V, H = 50272, 768
    emb = (0.02 * torch.randn(V, H, generator=g)).cuda().requires_grad_(True)
    bias = (0.1 * torch.randn(V, generator=g)).cuda().requires_grad_(True)
    h = torch.randn(n, H, generator=g).cuda().requires_grad_(True)
    labels = torch.randint(0, V - vocab_pad, (n,), generator=g).cuda()
    w = torch.rand(n, generator=g).cuda()
    emb.grad = torch.zeros_like(emb)        # in-place sinks, like the flat gradient buffer
    bias.grad = torch.zeros_like(bias)
    cfg = {"labels": labels, "n_valid": V - vocab_pad, "emb_bf16": emb.detach().to(torch.bfloat16)}
    loss = Fn.lm_head_cross_entropy(h, emb, bias, cfg)
    (loss * w).sum().backward()
    got = (loss.detach().clone(), h.grad.clone(), emb.grad.clone(), bias.grad.clone())
    h.grad = None
    emb.grad = None
    bias.grad = None
    hb = h.to(torch.bfloat16).float()
    eb = emb.to(torch.bfloat16).float()
    logits = (hb @ eb.t() + bias)[:, :V - vocab_pad]
    ref = torch.nn.functional.cross_entropy(logits, labels, reduction="none")
    (ref * w).sum().backward()
    assert (got[0] - ref).abs().max().item() < 2e-3
    for name, a, b in (("dh", got[1], h.grad), ("dE", got[2], emb.grad), ("dbias", got[3], bias.grad)):
        rel = ((a - b).norm() / b.norm().clamp(min=1e-12)).item()
        assert rel < 2e-2, (name, rel)
    if vocab_pad:
        assert float(got[2][V - vocab_pad:].abs().max()) == 0.0
        assert float(got[3][V - vocab_pad:].abs().max()) == 0.0
