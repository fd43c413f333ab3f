"""Synthetic HERO batches in the reference's collate layout.

`video_batch` restates the tensor layout produced by the reference's `video_collate`
(data/data.py:406-471) and `get_gather_index` (data/data.py:504-512) from per-clip items shaped
like `VideoFeatSubTokDataset.__getitem__` (data/data.py:346-397): per subtitle a token list that
starts with [SEP], the frames matched to it (or one all-zero dummy frame, masked), and the clip's
full frame-feature matrix. Everything is built on the host with a seeded generator; there is no
dataset I/O in this repo.
"""
import random

import torch

VFEAT_DIM = 4352      # utils/const.py:6
MAX_CLIP_LEN = 100    # utils/const.py:7 / config max_clip_len
SEP, CLS, PAD = 2, 0, 1


def make_clip(gen, n_frames, sub_frames, sub_lens, vfeat_dim=VFEAT_DIM, vocab=50265):
    """One clip item: (sub token ids, matched frame lists, clip features)."""
    feats = torch.randn(n_frames, vfeat_dim, generator=gen)
    subs = []
    for L in sub_lens:
        ids = torch.randint(3, vocab, (L,), generator=gen)
        ids[0] = SEP
        subs.append(ids)
    sub2frames = [(i, list(fr)) for i, fr in enumerate(sub_frames)]
    return {"feats": feats, "subs": subs, "sub2frames": sub2frames}


def video_batch(clips):
    """Collate clip items into the batch dict consumed by HierarchicalVlModel.forward."""
    sub_ids, sub_feats, sub_masks = [], [], []
    num_subs, sub_idx2frame_idx = [], []
    for c in clips:
        T, D = c["feats"].shape
        num_subs.append(len(c["subs"]))
        sub_idx2frame_idx.append(c["sub2frames"])
        for ids, (_, frames) in zip(c["subs"], c["sub2frames"]):
            frames = [f for f in frames if 0 <= f < T]
            if frames:
                sub_feats.append(c["feats"][torch.tensor(frames)])
                sub_masks.append(torch.ones(len(ids) + len(frames), dtype=torch.long))
            else:   # data/data.py:380-382: one dummy zero frame, masked out
                sub_feats.append(torch.zeros(1, D))
                sub_masks.append(torch.cat([torch.zeros(1, dtype=torch.long),
                                            torch.ones(len(ids), dtype=torch.long)]))
            sub_ids.append(ids)
    R = len(sub_ids)
    txt_lens = [len(i) for i in sub_ids]
    v_lens = [f.shape[0] for f in sub_feats]
    max_sl, max_vl = max(txt_lens), max(v_lens)
    out_size = max(m.numel() for m in sub_masks)
    D = clips[0]["feats"].shape[1]
    f_sub_input_ids = torch.full((R, max_sl), PAD, dtype=torch.long)
    f_v_feats = torch.zeros(R, max_vl, D)
    f_attn_masks = torch.zeros(R, out_size, dtype=torch.long)
    f_gather_index = torch.arange(out_size, dtype=torch.long).unsqueeze(0).repeat(R, 1)
    for r in range(R):
        tl, nf = txt_lens[r], v_lens[r]
        f_sub_input_ids[r, :tl] = sub_ids[r]
        f_v_feats[r, :nf] = sub_feats[r]
        f_attn_masks[r, :sub_masks[r].numel()] = sub_masks[r]
        f_gather_index[r, nf:nf + tl] = torch.arange(max_vl, max_vl + tl)
    f_sub_pos_ids = torch.arange(max_sl, dtype=torch.long).clamp(max=511).unsqueeze(0)
    f_v_pos_ids = torch.arange(max_vl, dtype=torch.long).unsqueeze(0)

    B = len(clips)
    n_frames = [c["feats"].shape[0] for c in clips]
    max_t = max(n_frames)
    c_v_feats = torch.zeros(B, max_t, D)
    c_attn_masks = torch.zeros(B, max_t, dtype=torch.long)
    for b, c in enumerate(clips):
        c_v_feats[b, :n_frames[b]] = c["feats"]
        c_attn_masks[b, :n_frames[b]] = 1
    c_pos_ids = torch.arange(max_t, dtype=torch.long).unsqueeze(0).repeat(B, 1)
    return {
        "f_sub_input_ids": f_sub_input_ids, "f_sub_pos_ids": f_sub_pos_ids,
        "f_v_feats": f_v_feats, "f_v_pos_ids": f_v_pos_ids,
        "f_attn_masks": f_attn_masks, "f_gather_index": f_gather_index,
        "c_v_feats": c_v_feats, "c_pos_ids": c_pos_ids, "c_attn_masks": c_attn_masks,
        "num_subs": num_subs, "sub_idx2frame_idx": sub_idx2frame_idx,
    }


def query_batch(gen, lens, vocab=50265):
    """Text-only rows for CrossModalTrm.forward(batch, 'txt') (data/data.py txt_input_collate)."""
    n, max_l = len(lens), max(lens)
    ids = torch.full((n, max_l), PAD, dtype=torch.long)
    masks = torch.zeros(n, max_l, dtype=torch.long)
    for i, L in enumerate(lens):
        row = torch.randint(3, vocab, (L,), generator=gen)
        row[0] = CLS
        ids[i, :L] = row
        masks[i, :L] = 1
    pos = torch.arange(max_l, dtype=torch.long).clamp(max=511).unsqueeze(0)
    return {"input_ids": ids, "pos_ids": pos, "attn_masks": masks}


def syn_tvr_dense(batch_size=32, seed=1234, n_frames=100, n_subs=20, frames_per_sub=5,
                  sub_len=20, query_len=16, vfeat_dim=VFEAT_DIM, vocab=50265):
    """SYN-TVR-dense: B clips x 100 frames, 20 subs of 5 frames + 20 tokens, 1 query of 16."""
    gen = torch.Generator().manual_seed(seed)
    clips = []
    for _ in range(batch_size):
        frames = [range(s * frames_per_sub, (s + 1) * frames_per_sub) for s in range(n_subs)]
        clips.append(make_clip(gen, n_frames, frames, [sub_len] * n_subs, vfeat_dim, vocab))
    return video_batch(clips), query_batch(gen, [query_len] * batch_size, vocab)


def syn_tvr_ragged(batch_size=32, seed=4321, vfeat_dim=VFEAT_DIM, vocab=50265, t_range=(40, 100),
                   s_range=(8, 30), l_range=(4, 40), q_range=(6, 24)):
    """SYN-TVR-ragged: variable T/S/L, ~10 % unmatched frames, ~5 % subs without frames."""
    gen = torch.Generator().manual_seed(seed)
    rnd = random.Random(seed)
    clips, qlens = [], []
    for _ in range(batch_size):
        T = rnd.randint(*t_range)
        S = rnd.randint(*s_range)
        cuts = sorted(rnd.sample(range(1, T), min(S - 1, T - 1)))
        bounds = [0] + cuts + [T]
        groups = []
        for s in range(len(bounds) - 1):
            fr = [f for f in range(bounds[s], bounds[s + 1]) if rnd.random() >= 0.10]
            if rnd.random() < 0.05:
                fr = []
            groups.append(fr)
        lens = [rnd.randint(*l_range) for _ in groups]
        clips.append(make_clip(gen, T, groups, lens, vfeat_dim, vocab))
        qlens.append(rnd.randint(*q_range))
    return video_batch(clips), query_batch(gen, qlens, vocab)


def syn_xm_1(seed=0, vfeat_dim=VFEAT_DIM, vocab=50265):
    """SYN-XM-1 (BASELINE config 1): 2 rows x (8 frames + 16 sub tokens), all valid."""
    gen = torch.Generator().manual_seed(seed)
    R, F, L = 2, 8, 16
    return {
        "f_v_feats": torch.randn(R, F, vfeat_dim, generator=gen),
        "f_sub_input_ids": torch.randint(3, 50000, (R, L), generator=gen),
        "f_sub_pos_ids": torch.arange(L).unsqueeze(0),
        "f_v_pos_ids": torch.arange(F).unsqueeze(0),
        "f_attn_masks": torch.ones(R, F + L, dtype=torch.long),
        "f_gather_index": torch.arange(F + L).unsqueeze(0).repeat(R, 1),
    }


def to_device(batch, device, non_blocking=True):
    """Tensors to `device`; python lists stay on the host (as data/loader.py:62-73 does)."""
    out = {}
    for k, v in batch.items():
        out[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
    return out
