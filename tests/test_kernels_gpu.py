"""Per-kernel numerics on the GPU: every C-ABI entry point against a plain torch fp32
restatement of the same op on the same (bf16-rounded) inputs."""
import math

import numpy as np

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _dev():
    return torch.device("cuda:0")


def _rand(shape, scale=1.0, seed=0, dtype=BF16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(_dev())


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, (f"{what}: {bad}/{err.numel()} elements out of tolerance; "
                      f"max abs err {err.max().item():.4g}, ref max {ref.abs().max().item():.4g}")


# ------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 768, 768), (16000, 768, 768),
                                   (1000, 2304, 768), (515, 3072, 768), (640, 768, 3072),
                                   (777, 768, 4352), (8, 128, 64)])
@pytest.mark.parametrize("block_n,cta_pair", [(128, 1), (256, 1), (256, 2)])
def test_gemm_forward_kmajor(m, n, k, block_n, cta_pair):
    from hero_b200 import ops
    a, w = _rand((m, k), seed=1), _rand((n, k), 0.05, seed=2)
    bias = _rand((n,), 0.5, seed=3, dtype=torch.float32)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    ops.gemm(a, w, out, bias=bias, block_n=block_n, cta_pair=cta_pair)
    ref = a.float() @ w.float().t() + bias
    _close(out, ref, 2e-2, 1.6e-2, f"gemm fwd {m}x{n}x{k} bn{block_n} pair{cta_pair}")


@pytest.mark.parametrize("act", ["gelu", "relu", "resid"])
def test_gemm_epilogues(act):
    from hero_b200 import ops
    m, n, k = 1000, 768, 768
    a, w = _rand((m, k), seed=4), _rand((n, k), 0.05, seed=5)
    bias = _rand((n,), 0.5, seed=6, dtype=torch.float32)
    resid = _rand((m, n), seed=7)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    pre = a.float() @ w.float().t() + bias
    if act == "gelu":
        aux = torch.empty(m, n, dtype=BF16, device=_dev())
        ops.gemm(a, w, out, bias=bias, act=ops.ACT_GELU, aux_out=aux)
        dgelu = 0.5 * (1 + torch.erf(pre / math.sqrt(2))) + pre * torch.exp(-0.5 * pre * pre) / \
            math.sqrt(2 * math.pi)
        _close(aux, dgelu, 1e-2, 1e-2, "saved gelu derivative")
        ref = torch.nn.functional.gelu(pre)
    elif act == "relu":
        ops.gemm(a, w, out, bias=bias, act=ops.ACT_RELU, resid=resid)
        ref = torch.relu(pre) + resid.float()
    else:
        ops.gemm(a, w, out, bias=bias, resid=resid)
        ref = pre + resid.float()
    _close(out, ref, 2e-2, 1.6e-2, f"epilogue {act}")


@pytest.mark.parametrize("m,n,k", [(1000, 768, 3072), (16000, 768, 2304), (300, 3072, 768)])
@pytest.mark.parametrize("block_n,cta_pair", [(128, 1), (256, 1), (256, 2)])
def test_gemm_dgrad_b_mnmajor(m, n, k, block_n, cta_pair):
    """dX[m, n] = dY[m, k] @ W[k, n]  (W stored [k, n]: the nn.Linear weight [out=k, in=n])."""
    from hero_b200 import ops
    dy, w = _rand((m, k), seed=8), _rand((k, n), 0.05, seed=9)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    ops.gemm(dy, w, out, b_mn=True, block_n=block_n, cta_pair=cta_pair)
    _close(out, dy.float() @ w.float(), 2e-2, 1.6e-2, "dgrad")


def test_gemm_dgrad_gelu_grad():
    from hero_b200 import ops
    m, n, k = 515, 3072, 768
    dy, w, dg = _rand((m, k), seed=10), _rand((k, n), 0.05, seed=11), _rand((m, n), 0.5, seed=12)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    ops.gemm(dy, w, out, b_mn=True, act=ops.ACT_GELU_GRAD, aux_in=dg)
    _close(out, (dy.float() @ w.float()) * dg.float(), 2e-2, 1.6e-2, "dgrad * saved gelu'")


@pytest.mark.parametrize("tokens,n_out,k_in", [(1000, 768, 768), (16000, 3072, 768),
                                               (3333, 768, 3072), (3200, 768, 4352)])
@pytest.mark.parametrize("k_splits,cta_pair", [(0, 0), (1, 1), (3, 1), (1, 2), (3, 2)])
def test_gemm_wgrad_mnmajor(tokens, n_out, k_in, k_splits, cta_pair):
    """dW[n_out, k_in] += dY[tokens, n_out]^T @ X[tokens, k_in], fp32 accumulate."""
    from hero_b200 import ops
    dy, x = _rand((tokens, n_out), 0.1, seed=13), _rand((tokens, k_in), seed=14)
    out = torch.full((n_out, k_in), 0.5, dtype=torch.float32, device=_dev())
    ops.gemm(dy, x, out, a_mn=True, b_mn=True, accumulate_f32=True, k_splits=k_splits,
             cta_pair=cta_pair)
    ref = 0.5 + dy.float().t() @ x.float()
    _close(out, ref, 5e-2, 5e-3, "wgrad")


def test_gemm_dropout_epilogue_is_deterministic_and_scaled():
    from hero_b200 import ops
    m, n, k = 512, 768, 768
    a, w = _rand((m, k), seed=15), _rand((n, k), 0.05, seed=16)
    drop = ops.drop_params(0.1, 1234)
    o1 = torch.empty(m, n, dtype=BF16, device=_dev())
    o2 = torch.empty_like(o1)
    ops.gemm(a, w, o1, drop=drop)
    ops.gemm(a, w, o2, drop=drop)
    assert torch.equal(o1, o2)
    ref = a.float() @ w.float().t()
    kept = o1.float() != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.9) < 0.01, frac
    _close(o1.float()[kept], (ref / 0.9)[kept], 3e-2, 2e-2, "kept elements scaled by 1/(1-p)")


# ------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("h,eps", [(768, 1e-12), (768, 1e-5), (4352, 1e-5), (256, 1e-5)])
def test_ln_fwd_bwd_plain(h, eps):
    from hero_b200 import ops
    n = 1037
    x = _rand((n, h), 2.0, seed=20)
    gamma = _rand((h,), 1.0, seed=21, dtype=torch.float32)
    beta = _rand((h,), 1.0, seed=22, dtype=torch.float32)
    y = torch.empty(n, h, dtype=BF16, device=_dev())
    mean = torch.empty(n, device=_dev())
    rstd = torch.empty(n, device=_dev())
    ops.ln_fwd(x, gamma, beta, eps, y, n_rows=n, mean=mean, rstd=rstd)
    xr = x.float().requires_grad_(True)
    g = gamma.clone().requires_grad_(True)
    b = beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (h,), g, b, eps)
    _close(y, ref, 2e-2, 1.6e-2, "ln fwd")
    dy = _rand((n, h), 1.0, seed=23)
    ref.backward(dy.float())
    dx = torch.empty(n, h, dtype=BF16, device=_dev())
    dgamma = torch.zeros(h, device=_dev())
    dbeta = torch.zeros(h, device=_dev())
    ops.ln_bwd(dy, x, gamma, mean, rstd, n_rows=n, dx=dx, dgamma=dgamma, dbeta=dbeta)
    _close(dx, xr.grad, 3e-2, 2e-2, "ln dx")
    _close(dgamma, g.grad, 0.5, 2e-2, "ln dgamma")
    _close(dbeta, b.grad, 0.5, 2e-2, "ln dbeta")


def test_ln_gather_add_scatter_embedding_form():
    """Text-embedding form: LN(word[ids] + pos[pid] + type) written to scattered rows."""
    from hero_b200 import ops
    h, vocab, n = 768, 1000, 777
    word = _rand((vocab, h), 0.02, seed=30, dtype=torch.float32)
    pos = _rand((64, h), 0.02, seed=31, dtype=torch.float32)
    typ = _rand((h,), 0.02, seed=32, dtype=torch.float32)
    gamma = _rand((h,), 1.0, seed=33, dtype=torch.float32)
    beta = _rand((h,), 0.1, seed=34, dtype=torch.float32)
    g = torch.Generator().manual_seed(35)
    ids = torch.randint(0, vocab, (n,), generator=g).int().to(_dev())
    pid = torch.randint(0, 64, (n,), generator=g).int().to(_dev())
    dst = torch.randperm(n + 50, generator=g)[:n].int().to(_dev())
    y = torch.zeros(n + 50, h, dtype=BF16, device=_dev())
    mean = torch.empty(n, device=_dev())
    rstd = torch.empty(n, device=_dev())
    ops.ln_fwd(word, gamma, beta, 1e-5, y, n_rows=n, x_rows=ids, add_tab=pos, add_idx=pid,
               add_vec=typ, y_rows=dst, mean=mean, rstd=rstd)
    wr = word.clone().requires_grad_(True)
    pr = pos.clone().requires_grad_(True)
    s = wr[ids.long()] + pr[pid.long()] + typ
    ref = torch.nn.functional.layer_norm(s, (h,), gamma, beta, 1e-5)
    _close(y[dst.long()], ref, 2e-2, 1.6e-2, "embedding ln fwd")
    dyfull = _rand((n + 50, h), 1.0, seed=36)
    ref.backward(dyfull[dst.long()].float())
    dword = torch.zeros_like(word)
    dpos = torch.zeros_like(pos)
    dx = torch.empty(n, h, dtype=BF16, device=_dev())
    ops.ln_bwd(dyfull, word, gamma, mean, rstd, n_rows=n, x_rows=ids, add_tab=pos, add_idx=pid,
               add_vec=typ, y_rows=dst, dx=dx, d_x_tab=dword, d_add_tab=dpos)
    _close(dword, wr.grad, 5e-2, 3e-2, "word table grad (atomic scatter)")
    _close(dpos, pr.grad, 0.3, 3e-2, "pos table grad (atomic scatter)")


def test_ln_dropout_mask_matches_between_fwd_and_bwd():
    from hero_b200 import ops
    n, h = 512, 768
    x = _rand((n, h), 1.0, seed=40)
    gamma = torch.ones(h, device=_dev())
    beta = torch.zeros(h, device=_dev())
    drop = ops.drop_params(0.1, 99)
    y = torch.empty(n, h, dtype=BF16, device=_dev())
    mean = torch.empty(n, device=_dev())
    rstd = torch.empty(n, device=_dev())
    ops.ln_fwd(x, gamma, beta, 1e-5, y, n_rows=n, mean=mean, rstd=rstd, drop=drop)
    keep = (y.float() != 0)
    assert abs(keep.float().mean().item() - 0.9) < 0.01
    # backward of sum(y): dy = 1 -> effective dy is mask/keep; compare with torch using that mask
    dy = torch.ones(n, h, dtype=BF16, device=_dev())
    dx = torch.empty(n, h, dtype=BF16, device=_dev())
    ops.ln_bwd(dy, x, gamma, mean, rstd, n_rows=n, dx=dx, drop=drop)
    xr = x.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (h,), gamma, beta, 1e-5)
    (ref * keep.float() / 0.9).sum().backward()
    _close(dx, xr.grad, 3e-2, 3e-2, "ln dx through dropout")


# ------------------------------------------------------------------------------ attention
def _attn_ref(qkv, lens, heads):
    h = heads * 64
    outs, off = [], 0
    for n in lens:
        blk = qkv[off:off + n]
        q, k, v = (blk[:, i * h:(i + 1) * h].reshape(n, heads, 64).transpose(0, 1)
                   for i in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1)
        outs.append((p @ v).transpose(0, 1).reshape(n, h))
        off += n
    return torch.cat(outs, 0)


def _att_plan(lens):
    """Device attention plan (tiles + per-token sequence ranges) for packed sequences."""
    import numpy as np
    from hero_b200.plan import DeviceIndex, SeqPlan
    mask = np.zeros((len(lens), max(max(lens), 1)), np.int64)
    for r, n in enumerate(lens):
        mask[r, :n] = 1
    sp = SeqPlan(mask)
    dev = DeviceIndex(sp.arrays("s_"), _dev())
    return sp, sp.attn(dev, "s_")


@pytest.mark.parametrize("lens", [[25] * 40, [1, 7, 33, 64, 100, 128, 2, 90], [100] * 8,
                                  [16] * 32, [3, 0, 5, 120, 9],
                                  # rows longer than one tile (max_txt_len + matched frames can
                                  # exceed 128; the position table allows 514): long-sequence path
                                  [20, 200, 31, 129, 64], [514], [300, 7, 768]])
def test_attention_fwd_bwd(lens):
    from hero_b200 import ops
    heads = 12 if max(lens) <= 200 else 3
    ntok = sum(lens)
    qkv = _rand((ntok, 3 * heads * 64), 1.0, seed=50)
    sp, att = _att_plan(lens)
    n_short = sp.n_tiles - sp.n_long
    assert all(n <= 128 for n in sp.tile_ntok[:n_short]) and int(sp.tile_ntok.sum()) == ntok
    assert sp.n_long == sum(n > 128 for n in lens)
    ctx = torch.empty(ntok, heads * 64, dtype=BF16, device=_dev())
    lse = torch.empty(ntok, heads, device=_dev())
    ops.attn_fwd(qkv, att, ctx, heads=heads, lse=lse)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, [n for n in lens if n > 0], heads)
    _close(ctx, ref, 2e-2, 1.6e-2, "attention fwd")
    dctx = _rand((ntok, heads * 64), 1.0, seed=51)
    ref.backward(dctx.float())
    dqkv = torch.empty_like(qkv)
    dbias = torch.full((3 * heads * 64,), 0.5, device=_dev())
    ops.attn_bwd(qkv, att, ctx, dctx, lse, dqkv, heads=heads, dbias=dbias)
    _close(dqkv, qr.grad, 4e-2, 3e-2, "attention bwd")
    # the bias gradient of the QKV projection is ACCUMULATED: column sums of the stored dqkv
    want = 0.5 + dqkv.float().sum(0)
    tol = 2e-3 * dqkv.float().abs().sum(0).max().item() + 1e-3
    assert (dbias - want).abs().max().item() <= tol, ((dbias - want).abs().max().item(), tol)
    # and without the bias pointer the stored gradients are the same
    dq2 = torch.empty_like(qkv)
    ops.attn_bwd(qkv, att, ctx, dctx, lse, dq2, heads=heads)
    assert torch.equal(dq2, dqkv)


def test_attention_tiles_hold_at_most_16_sequences_and_mask_out_their_neighbours():
    """The block-diagonal mask is a K = 16 membership MMA: the plan closes a tile after 16
    sequences, and scores of the neighbouring sequences must not leak even when they are far
    larger than the row's own (here by ~60 in scaled units)."""
    from hero_b200 import ops
    lens = [2] * 40 + [5, 1, 1, 7] * 6
    heads, ntok = 2, sum(lens)
    sp, att = _att_plan(lens)
    cu = np.concatenate([[0], np.cumsum(lens)])
    for a, n in zip(sp.tile_tok0.tolist(), sp.tile_ntok.tolist()):
        assert ((cu[:-1] >= a) & (cu[:-1] < a + n)).sum() <= 16
    qkv = _rand((ntok, 3 * heads * 64), 1.0, seed=58)
    seq_of = np.repeat(np.arange(len(lens)), lens)
    # keys of odd sequences are large: a leak would swamp the even sequences' rows
    big = torch.from_numpy((seq_of % 2 == 1)).to(_dev())
    H = heads * 64
    qkv[:, H:2 * H] = torch.where(big[:, None], qkv[:, H:2 * H] * 8, qkv[:, H:2 * H])
    qkv[:, :H] = qkv[:, :H] * 3
    ctx = torch.empty(ntok, H, dtype=BF16, device=_dev())
    lse = torch.empty(ntok, heads, device=_dev())
    ops.attn_fwd(qkv, att, ctx, heads=heads, lse=lse)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, lens, heads)
    _close(ctx, ref, 2e-2, 1.6e-2, "attention fwd, 16-sequence tiles")
    dctx = _rand((ntok, H), 1.0, seed=59)
    ref.backward(dctx.float())
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv, att, ctx, dctx, lse, dqkv, heads=heads)
    # (operands 3-8x larger than in the other tests: tolerance relative to the gradient scale)
    _close(dqkv, qr.grad, 6e-3 * qr.grad.abs().max().item(), 3e-2, "attention bwd, 16-sequence tiles")


def test_attention_dropout_rate_and_pair_independence():
    """V = identity block makes ctx row i = its dropped probability row: the drop rate is p and
    neighbouring columns (which share one hash through the derived pair words) are independent."""
    from hero_b200 import ops
    heads, lens = 1, [64] * 24
    ntok = sum(lens)
    qkv = torch.zeros(ntok, 3 * 64, dtype=BF16, device=_dev())      # Q = K = 0: uniform P = 1/64
    eye = torch.eye(64, dtype=BF16, device=_dev()).repeat(len(lens), 1)
    qkv[:, 128:] = eye
    _, att = _att_plan(lens)
    ctx = torch.empty(ntok, 64, dtype=BF16, device=_dev())
    ops.attn_fwd(qkv, att, ctx, heads=heads, drop=ops.drop_params(0.1, 31337))
    kept = (ctx.float() > 0)
    vals = ctx.float()[kept]
    assert torch.allclose(vals, torch.full_like(vals, 1 / 64 / 0.9), rtol=1e-2)
    rate = 1.0 - kept.float().mean().item()
    assert abs(rate - 0.1) < 6e-3, rate                    # 98k samples: sigma ~ 1e-3
    k = kept.float()
    for shift in (1, 2, 3, 4, 8):                          # pair partner, same group, next group
        a, b = k[:, :-shift].reshape(-1), k[:, shift:].reshape(-1)
        corr = ((a - a.mean()) * (b - b.mean())).mean().item() / (a.std() * b.std()).item()
        assert abs(corr) < 2e-2, (shift, corr)
    rows = torch.stack([k[:-1].reshape(-1), k[1:].reshape(-1)])      # neighbouring query rows
    assert abs(torch.corrcoef(rows)[0, 1].item()) < 2e-2


def test_attention_dropout_consistent_between_fwd_and_bwd():
    """With dropout the forward equals P_drop V for the regenerated mask: check through linearity
    (ctx is linear in V for fixed Q, K) and that backward's dV matches that same linear map."""
    from hero_b200 import ops
    heads, lens = 2, [20, 31, 64, 13]
    ntok = sum(lens)
    qkv = _rand((ntok, 3 * heads * 64), 1.0, seed=53)
    _, att = _att_plan(lens)
    drop = ops.drop_params(0.1, 4242)
    c1 = torch.empty(ntok, heads * 64, dtype=BF16, device=_dev())
    c2 = torch.empty_like(c1)
    lse = torch.empty(ntok, heads, device=_dev())
    ops.attn_fwd(qkv, att, c1, heads=heads, drop=drop, lse=lse)
    ops.attn_fwd(qkv, att, c2, heads=heads, drop=drop)
    assert torch.equal(c1, c2)
    c0 = torch.empty_like(c1)
    ops.attn_fwd(qkv, att, c0, heads=heads)
    assert (c1.float() - c0.float()).abs().mean() > 1e-3          # dropout did something
    # E[ctx_drop] = ctx: averaged over keys the deviation is zero-mean
    assert abs((c1.float() - c0.float()).mean().item()) < 5e-3
    # dV from backward must be the adjoint of V -> ctx_drop: <ctx_drop(V), dO> == <V, dV>
    dctx = _rand((ntok, heads * 64), 1.0, seed=54)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv, att, c1, dctx, lse, dqkv, heads=heads, drop=drop)
    H = heads * 64
    lhs = (c1.float() * dctx.float()).sum().item()
    rhs = (qkv[:, 2 * H:].float() * dqkv[:, 2 * H:].float()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs), 1.0), (lhs, rhs)


def test_attention_long_rows_share_the_dropout_words_of_the_tile_kernels():
    """Forward / backward of a long row regenerate the same mask (adjoint identity), and the plan
    refuses rows beyond the long-sequence limit with the offending row in the message."""
    import numpy as np
    from hero_b200 import ops
    from hero_b200.plan import SeqPlan
    with pytest.raises(ValueError, match="row 0 has 900"):
        SeqPlan(np.ones((1, 900), np.int64))
    heads, lens = 2, [150, 40]
    ntok = sum(lens)
    qkv = _rand((ntok, 3 * heads * 64), 1.0, seed=55)
    _, att = _att_plan(lens)
    drop = ops.drop_params(0.1, 777)
    c1 = torch.empty(ntok, heads * 64, dtype=BF16, device=_dev())
    lse = torch.empty(ntok, heads, device=_dev())
    ops.attn_fwd(qkv, att, c1, heads=heads, drop=drop, lse=lse)
    dctx = _rand((ntok, heads * 64), 1.0, seed=56)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv, att, c1, dctx, lse, dqkv, heads=heads, drop=drop)
    H = heads * 64
    lhs = (c1.float() * dctx.float()).sum().item()
    rhs = (qkv[:, 2 * H:].float() * dqkv[:, 2 * H:].float()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs), 1.0), (lhs, rhs)


# ------------------------------------------------------------------------------ row utilities
def test_gather_and_gather_sum_rows():
    from hero_b200 import ops
    h = 768
    src = _rand((500, h), seed=60)
    g = torch.Generator().manual_seed(61)
    idx = torch.randint(-1, 500, (700,), generator=g).int().to(_dev())
    dst = torch.empty(700, h, dtype=BF16, device=_dev())
    ops.gather_rows(src, idx, dst)
    ref = torch.where((idx >= 0)[:, None], src[idx.clamp(min=0).long()], torch.zeros_like(dst))
    assert torch.equal(dst, ref)
    counts = torch.randint(0, 4, (300,), generator=g)
    off = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).int().to(_dev())
    cidx = torch.randint(0, 500, (int(counts.sum()),), generator=g).int().to(_dev())
    out = torch.empty(300, h, dtype=BF16, device=_dev())
    ops.gather_sum_rows(src, off, cidx, out)
    ref = torch.zeros(300, h, device=_dev())
    rows = torch.repeat_interleave(torch.arange(300), counts).to(_dev())
    ref.index_add_(0, rows, src[cidx.long()].float())
    _close(out, ref, 2e-2, 1e-2, "gather_sum bf16")
    out32 = torch.ones(300, h, device=_dev())
    ops.gather_sum_rows(src, off, cidx, out32)
    _close(out32, ref + 1.0, 1e-3, 1e-3, "gather_sum f32 accumulate")


def test_colsum_relu_bwd_cast():
    from hero_b200 import ops
    x = _rand((3201, 768), seed=70)
    out = torch.ones(768, device=_dev())
    ops.colsum(x, out)
    _close(out, 1.0 + x.float().sum(0), 5e-2, 1e-3, "colsum")
    dy, pre = _rand((100, 768), seed=71), _rand((100, 768), seed=72)
    o = torch.empty_like(dy)
    ops.relu_bwd(dy, pre, o)
    assert torch.equal(o, torch.where(pre.float() > 0, dy, torch.zeros_like(dy)))
    src = _rand((1000, 77), seed=73, dtype=torch.float32).contiguous()
    dst = torch.empty(1000, 77, dtype=BF16, device=_dev())
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(BF16))


def test_adamw_matches_reference_update_rule():
    """optim/adamw.py:80-104 restated in torch on the same numbers."""
    from hero_b200 import ops
    n = 100003
    p = _rand((n,), 0.05, seed=80, dtype=torch.float32)
    g = _rand((n,), 0.01, seed=81, dtype=torch.float32)
    m = _rand((n,), 0.01, seed=82, dtype=torch.float32)
    v = _rand((n,), 0.01, seed=83, dtype=torch.float32).abs()
    lr, b1, b2, eps, wd, t = 1e-4, 0.9, 0.98, 1e-6, 0.01, 7
    rp, rm, rv = p.clone(), m.clone(), v.clone()
    rm.mul_(b1).add_(g, alpha=1 - b1)
    rv.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = rv.sqrt().add_(eps)
    step = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    rp.addcdiv_(rm, denom, value=-step)
    rp.add_(rp, alpha=-lr * wd)
    pb = torch.empty(n, dtype=BF16, device=_dev())
    ops.adamw_step(p, g, m, v, pb, step_size=step, beta1=b1, beta2=b2, eps=eps, lr_wd=lr * wd)
    _close(p, rp, 1e-7, 1e-5, "adamw p")
    _close(m, rm, 1e-8, 1e-5, "adamw m")
    _close(v, rv, 1e-9, 1e-5, "adamw v")
    assert torch.equal(pb, p.to(BF16))
    s = torch.zeros(1, device=_dev())
    ops.sumsq(g, s)
    _close(s, (g.double() ** 2).sum().float().reshape(1), 0, 1e-4, "sumsq")


# ------------------------------------------------------------------------------ fp32 residual stream
@pytest.mark.parametrize("m,n,k", [(1000, 768, 768), (16512, 768, 768), (3200, 768, 3072),
                                   (130, 768, 768)])
@pytest.mark.parametrize("block_n,cta_pair", [(128, 1), (256, 1), (256, 2)])
def test_gemm_fp32_residual_in_fp32_sum_out(m, n, k, block_n, cta_pair):
    """Forward out-projection / FFN-down form: s = x W^T + b + resid, resid and s in fp32 (the
    residual stream never passes through bf16)."""
    from hero_b200 import ops
    a, w = _rand((m, k), seed=90), _rand((n, k), 0.05, seed=91)
    bias = _rand((n,), 0.5, seed=92, dtype=torch.float32)
    resid = _rand((m, n), 3.0, seed=93, dtype=torch.float32)
    out = torch.full((m, n), float("nan"), dtype=torch.float32, device=_dev())
    ops.gemm(a, w, out, bias=bias, resid=resid, block_n=block_n, cta_pair=cta_pair)
    ref = a.float() @ w.float().t() + bias + resid
    _close(out, ref, 2e-3, 1e-4, f"fp32 residual stream {m}x{n}x{k} bn{block_n} pair{cta_pair}")
    out2 = torch.full((m, n), float("nan"), dtype=torch.float32, device=_dev())
    ops.gemm(a, w, out2, bias=bias, block_n=block_n, cta_pair=cta_pair)       # no residual
    _close(out2, ref - resid, 2e-3, 1e-4, "fp32 store without residual")


@pytest.mark.parametrize("m,n,k,act", [(16512, 3072, 768, 3), (3200, 3072, 768, 3),
                                       (1000, 768, 256, 0), (130, 192, 64, 0)])
def test_gemm_epilogue_column_sums(m, n, k, act):
    """out_colsum accumulates the column sums of the stored bf16 output (the FFN-up bias gradient
    comes out of the x gelu' dgrad this way instead of a second pass over its output)."""
    from hero_b200 import ops
    a, w = _rand((m, k), 1.0, seed=70), _rand((k, n), 0.05, seed=71)
    aux = _rand((m, n), 1.0, seed=72) if act == 3 else None
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    cs = torch.full((n,), 0.25, device=_dev())
    ops.gemm(a, w, out, b_mn=True, act=act, aux_in=aux, out_colsum=cs)
    ref = torch.empty_like(out)
    ops.gemm(a, w, ref, b_mn=True, act=act, aux_in=aux)
    assert torch.equal(out, ref)
    want = 0.25 + out.double().sum(0)
    tol = 1e-5 * out.double().abs().sum(0).max().item() + 1e-3
    assert (cs.double() - want).abs().max().item() <= tol


def test_gemm_fp32_residual_with_dropout_matches_bf16_path_mask():
    """The dropout mask depends only on (key, element index): the fp32-store kernel (32-column
    slabs) and the bf16-store kernel (64-column slabs) must drop the same elements."""
    from hero_b200 import ops
    m, n, k = 640, 768, 768
    a, w = _rand((m, k), seed=94), _rand((n, k), 0.05, seed=95)
    drop = ops.drop_params(0.1, 4321)
    o16 = torch.empty(m, n, dtype=BF16, device=_dev())
    o32 = torch.empty(m, n, dtype=torch.float32, device=_dev())
    ops.gemm(a, w, o16, drop=drop)
    ops.gemm(a, w, o32, drop=drop)
    assert torch.equal(o16.float() == 0, o32 == 0)
    _close(o32, o16, 2e-2, 1.6e-2, "same values up to the bf16 rounding of the bf16 path")


@pytest.mark.parametrize("block_n,cta_pair", [(128, 1), (256, 1), (256, 2)])
@pytest.mark.parametrize("m", [515, 16512])
def test_gemm_tma_operand_epilogues_all_tilings(m, block_n, cta_pair):
    """bf16 residual add and saved-derivative multiply, operands arriving by TMA."""
    from hero_b200 import ops
    n, k = 3072, 768
    dy, w, dg = _rand((m, k), seed=96), _rand((k, n), 0.05, seed=97), _rand((m, n), 0.5, seed=98)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    ops.gemm(dy, w, out, b_mn=True, act=ops.ACT_GELU_GRAD, aux_in=dg, block_n=block_n,
             cta_pair=cta_pair)
    _close(out, (dy.float() @ w.float()) * dg.float(), 2e-2, 1.6e-2, "dgrad * saved gelu'")
    ops.gemm(dy, w, out, b_mn=True, resid=dg, block_n=block_n, cta_pair=cta_pair)
    _close(out, dy.float() @ w.float() + dg.float(), 2e-2, 1.6e-2, "dgrad + residual")


def test_gemm_relu_residual_with_saved_preactivation():
    """frame_transform form (model/layers.py:86-93 + model/model.py:211-212)."""
    from hero_b200 import ops
    m, n, k = 3200, 768, 4352
    a, w = _rand((m, k), seed=99), _rand((n, k), 0.02, seed=100)
    bias = _rand((n,), 0.5, seed=101, dtype=torch.float32)
    resid = _rand((m, n), seed=102)
    out = torch.empty(m, n, dtype=BF16, device=_dev())
    pre = torch.empty(m, n, dtype=BF16, device=_dev())
    ops.gemm(a, w, out, bias=bias, act=ops.ACT_RELU, resid=resid, aux_out=pre)
    p = a.float() @ w.float().t() + bias
    _close(pre, p, 2e-2, 1.6e-2, "saved pre-activation")
    _close(out, torch.relu(p) + resid.float(), 2e-2, 1.6e-2, "relu + residual")


@pytest.mark.parametrize("n", [1037, 16512])
def test_ln_fp32_rows_fast_path_fwd_bwd_with_fp32_copy(n):
    """Transformer-layer LayerNorms: fp32 pre-LN sums in, bf16 + fp32 outputs, one-pass backward."""
    from hero_b200 import ops
    h, eps = 768, 1e-12
    x = _rand((n, h), 2.0, seed=110, dtype=torch.float32)
    gamma = _rand((h,), 1.0, seed=111, dtype=torch.float32)
    beta = _rand((h,), 1.0, seed=112, dtype=torch.float32)
    y = torch.empty(n, h, dtype=BF16, device=_dev())
    y32 = torch.empty(n, h, dtype=torch.float32, device=_dev())
    mean, rstd = torch.empty(n, device=_dev()), torch.empty(n, device=_dev())
    ops.ln_fwd(x, gamma, beta, eps, y, n_rows=n, mean=mean, rstd=rstd, y_f32=y32)
    xr = x.clone().requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (h,), g, b, eps)
    _close(y32, ref, 2e-5, 1e-5, "fp32 copy")
    assert torch.equal(y, y32.to(BF16))
    dy = _rand((n, h), 1.0, seed=113)
    ref.backward(dy.float())
    dx = torch.empty(n, h, dtype=BF16, device=_dev())
    dxd = torch.empty(n, h, dtype=BF16, device=_dev())
    dgamma, dbeta, dbias = (torch.zeros(h, device=_dev()) for _ in range(3))
    ops.ln_bwd(dy, x, gamma, mean, rstd, n_rows=n, dx=dx, dx_drop=dxd, dgamma=dgamma, dbeta=dbeta,
               dbias=dbias)
    _close(dx, xr.grad, 1e-2, 1.6e-2, "ln dx")
    assert torch.equal(dx, dxd)
    _close(dgamma, g.grad, 0.5, 2e-2, "ln dgamma")
    _close(dbeta, b.grad, 0.5, 2e-2, "ln dbeta")
    _close(dbias, dx.float().sum(0), 0.5, 2e-2, "bias gradient of the feeding Linear")


def test_gather_rows_fp32():
    from hero_b200 import ops
    src = _rand((500, 768), seed=120, dtype=torch.float32)
    g = torch.Generator().manual_seed(121)
    idx = torch.randint(-1, 500, (700,), generator=g).int().to(_dev())
    dst = torch.empty(700, 768, dtype=torch.float32, device=_dev())
    ops.gather_rows(src, idx, dst)
    ref = torch.where((idx >= 0)[:, None], src[idx.clamp(min=0).long()], torch.zeros_like(dst))
    assert torch.equal(dst, ref)


def test_adamw_device_side_global_norm_clip():
    """train_vcmr.py:258-259 folded into the update: scale = min(1, max_norm / (norm + 1e-6)) read
    from a device scalar (no host round trip)."""
    from hero_b200 import ops
    n = 40000
    p = _rand((n,), 0.05, seed=130, dtype=torch.float32)
    g = _rand((n,), 0.5, seed=131, dtype=torch.float32)
    m, v = torch.zeros(n, device=_dev()), torch.zeros(n, device=_dev())
    s = torch.zeros(1, device=_dev())
    ops.sumsq(g, s)
    norm = float(g.double().norm())
    scale = min(1.0, 1.0 / (norm + 1e-6))
    assert scale < 0.5
    rp, rm, rv = p.clone(), m.clone(), v.clone()
    ops.adamw_step(rp, g * scale, rm, rv, None, step_size=1e-3, beta1=0.9, beta2=0.98, eps=1e-6,
                   lr_wd=0.0)
    ops.adamw_step(p, g, m, v, None, step_size=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, lr_wd=0.0,
                   clip_sumsq=s, clip_max_norm=1.0)
    _close(p, rp, 1e-7, 1e-4, "clipped update")
    _close(m, rm, 1e-9, 1e-4, "clipped first moment")


def test_split_bf16_gemm_and_ln_low_half():
    """frame_transform form: the LayerNorm emits hi + lo bf16 halves of its fp32 output, the GEMM
    contracts hi*hi + lo*hi + hi*lo in one accumulator: the pre-activation (and therefore the ReLU
    gate) agrees with fp32 arithmetic to ~1e-4 instead of ~1e-2."""
    from hero_b200 import ops
    n, d, h = 3200, 4352, 768
    x = _rand((n, d), 1.0, seed=140, dtype=torch.float32)
    gamma = (1.0 + 0.1 * _rand((d,), 1.0, seed=141, dtype=torch.float32))
    beta = _rand((d,), 0.05, seed=142, dtype=torch.float32)
    w = _rand((h, d), 0.02, seed=143, dtype=torch.float32)
    bias = _rand((h,), 0.05, seed=144, dtype=torch.float32)
    hi = torch.empty(n, d, dtype=BF16, device=_dev())
    lo = torch.empty(n, d, dtype=BF16, device=_dev())
    ops.ln_fwd(x, gamma, beta, 1e-5, hi, n_rows=n, y_lo=lo)
    ref_xn = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    _close(hi.float() + lo.float(), ref_xn, 1e-4, 1e-4, "hi + lo reproduces the fp32 LayerNorm output")
    w_hi = w.to(BF16)
    w_lo = (w - w_hi.float()).to(BF16)
    resid = _rand((n, h), seed=145)
    out = torch.empty(n, h, dtype=BF16, device=_dev())
    pre = torch.empty(n, h, dtype=BF16, device=_dev())
    ops.gemm(hi, w_hi, out, bias=bias, act=ops.ACT_RELU, resid=resid, aux_out=pre, a_lo=lo, b_lo=w_lo)
    ref_pre = ref_xn.double() @ w.double().t() + bias.double()
    flips = ((pre.float() > 0) != (ref_pre > 0)).float().mean().item()
    assert flips < 5e-5, f"ReLU gate differs from fp64 arithmetic on {flips:.2e} of the units"
    _close(pre, ref_pre.float(), 1e-2, 1e-2, "split-bf16 pre-activation")
    _close(out, torch.relu(ref_pre.float()) + resid.float(), 2e-2, 1.6e-2, "relu + residual")
    # the plain bf16 GEMM flips far more gates: this is what the split exists for
    pre1 = torch.empty_like(pre)
    ops.gemm(hi, w_hi, out, bias=bias, act=ops.ACT_RELU, resid=resid, aux_out=pre1)
    flips1 = ((pre1.float() > 0) != (ref_pre > 0)).float().mean().item()
    assert flips1 > 5 * max(flips, 1e-6), (flips1, flips)


@pytest.mark.parametrize("block_n,cta_pair", [(128, 1), (256, 1), (256, 2)])
def test_gemm_layernorm_form_residual(block_n, cta_pair):
    """The residual handed over as a pre-LayerNorm fp32 sum + (mean, rstd, gamma, beta): the
    epilogue adds LayerNorm(resid) computed in fp32 — what the separate fp32 LayerNorm output
    would have been, without that tensor ever existing."""
    from hero_b200 import ops, _lib
    import ctypes as C
    m, n, k = 3333, 768, 3072
    a, w = _rand((m, k), seed=150), _rand((n, k), 0.03, seed=151)
    bias = _rand((n,), 0.5, seed=152, dtype=torch.float32)
    pre = _rand((m, n), 2.0, seed=153, dtype=torch.float32)
    gamma = 1.0 + 0.1 * _rand((n,), 1.0, seed=154, dtype=torch.float32)
    beta = _rand((n,), 0.2, seed=155, dtype=torch.float32)
    mean = pre.mean(-1)
    rstd = torch.rsqrt(pre.var(-1, unbiased=False) + 1e-12)
    out = torch.full((m, n), float("nan"), dtype=torch.float32, device=_dev())
    ops.gemm(a, w, out, bias=bias, resid=pre, resid_ln=(mean, rstd, gamma, beta), block_n=block_n,
             cta_pair=cta_pair)
    ref = a.float() @ w.float().t() + bias + torch.nn.functional.layer_norm(pre, (n,), gamma, beta, 1e-12)
    _close(out, ref, 2e-3, 1e-4, f"LayerNorm-form residual bn{block_n} pair{cta_pair}")
