"""Forward-only passes and checkpoint I/O around the encoder.

* `embed_video_corpus` — the video-embedding pass of eval_vcmr.py:161-203: every clip of a corpus
  through `v_encoder(batch, 'repr')` in eval mode, results scattered into one
  (n_videos, max_clip_len, H) tensor + mask, batch by batch. Here the batches are staged through a
  `loader.BatchStager` ring (next batch's H2D copy and plan upload overlap the current forward) and
  the packed encoder runs without an autograd graph (activations ping-pong between two workspace
  slots, `hero_bert_stack_fwd` with save=False).
* `ModelSaver` / `TrainingRestorer` — utils/save.py:112-181 with the same file layout (`vocab_padded`
  flag, `model_step_<n>.pt`, `train_state_<n>.pt`, `restore.pt` + backup), so checkpoints are
  interchangeable with the reference's; optimizer state is FusedAdamW's flat moments.
"""
import os
from os.path import exists, join

import torch

from .loader import BatchStager
from .plan import PLAN_KEY, attach_plan


@torch.no_grad()
def embed_video_corpus(model, batches, n_videos, max_clip_len, device=None, video_indices=None,
                       out_dtype=torch.float32):
    """`batches`: iterable of HOST video batch dicts (reference `video_collate` layout; plans are
    attached here when the collate has not). `video_indices`: per batch the row of each clip in the
    corpus tensors (default: consecutive). Returns (frame_embeddings (n_videos, max_clip_len', H),
    c_attn_masks (n_videos, max_clip_len')) trimmed to the longest clip seen, like the reference."""
    v_enc = model.v_encoder if hasattr(model, "v_encoder") else model
    device = device or next(v_enc.parameters()).device
    was_training = v_enc.training
    v_enc.eval()
    stager = BatchStager(device, depth=3)
    cur = torch.cuda.current_stream(device)
    total, masks, longest, row = None, None, 0, 0
    it = iter(batches)

    def stage_next():
        try:
            hb = dict(next(it))
        except StopIteration:
            return None
        if hb.get(PLAN_KEY) is None:
            attach_plan(hb)
        return stager.stage(hb)

    nxt = stage_next()
    bi = 0
    while nxt is not None:
        (batch,), ready, slot = nxt
        cur.wait_event(ready)
        emb = v_enc(batch, "repr")                       # (B, T, H), zeros at padded frames
        stager.release(slot, cur)
        nxt = stage_next()                               # next copy overlaps this forward
        B, T, Hd = emb.shape
        assert T <= max_clip_len, f"clip of {T} frames exceeds max_clip_len {max_clip_len}"
        if total is None:
            total = torch.zeros((n_videos, max_clip_len, Hd), dtype=out_dtype, device=device)
            masks = torch.zeros((n_videos, max_clip_len), dtype=batch["c_attn_masks"].dtype,
                                device=device)
        if video_indices is not None:
            idx = torch.as_tensor(video_indices[bi], device=device, dtype=torch.long)
        else:
            idx = torch.arange(row, row + B, device=device)
        total[idx, :T] = emb.to(out_dtype)
        masks[idx, :T] = batch["c_attn_masks"]
        longest = max(longest, T)
        row += B
        bi += 1
    if was_training:
        v_enc.train()
    if total is None:
        return None, None
    return total[:, :longest], masks[:, :longest]


class ModelSaver(object):
    """utils/save.py:112-130."""

    def __init__(self, output_dir, prefix="model_step", suffix="pt", half=False):
        self.output_dir, self.prefix, self.suffix, self.half = output_dir, prefix, suffix, half

    def save(self, model, step, optimizer=None):
        path = join(self.output_dir, f"{self.prefix}_{step}.{self.suffix}")
        sd = {}
        for k, v in model.state_dict().items():
            v = v.detach().cpu() if isinstance(v, torch.Tensor) else v
            if self.half and isinstance(v, torch.Tensor) and v.dtype == torch.float32:
                v = v.half()          # apex O2 checkpoints of the reference hold fp16 tensors
            sd[k] = v
        sd["vocab_padded"] = any(("word_embeddings.weight" in k or "decoder.weight" in k)
                                 and v.size(0) % 8 == 0 for k, v in sd.items()
                                 if isinstance(v, torch.Tensor))
        torch.save(sd, path)
        if optimizer is not None:
            torch.save({"step": step, "optimizer": _to_cpu(optimizer.state_dict())},
                       f"{self.output_dir}/train_state_{step}.pt")
        return path


def load_checkpoint(model, path_or_state):
    """Loads a reference-style checkpoint (fp16 or fp32 tensors, `vocab_padded` flag) into a
    hero_b200 model that may already have run (the flat bf16 mirror is refreshed)."""
    from .encoder import load_pretrained_weight
    sd = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, str) \
        else dict(path_or_state)
    sd.pop("vocab_padded", None)
    sd = {k: (v.float() if isinstance(v, torch.Tensor) and v.dtype == torch.float16 else v)
          for k, v in sd.items()}
    return load_pretrained_weight(model, sd)


def _to_cpu(state):
    if isinstance(state, torch.Tensor):
        return state.detach().cpu()
    if isinstance(state, (list, tuple)):
        return type(state)(_to_cpu(t) for t in state)
    if isinstance(state, dict):
        return {k: _to_cpu(v) for k, v in state.items()}
    return state


class TrainingRestorer(object):
    """utils/save.py:133-181 (two rotating files, resume of model + optimizer + global step)."""

    def __init__(self, output_dir, model, optimizer, save_steps=1000):
        self.save_path = f"{output_dir}/restore.pt"
        self.backup_path = f"{output_dir}/restore_backup.pt"
        self.model, self.optimizer, self.save_steps = model, optimizer, save_steps
        self.global_step = 0
        if exists(self.save_path) or exists(self.backup_path):
            self.restore()

    def step(self):
        self.global_step += 1
        if self.global_step % self.save_steps == 0:
            self.save()

    def save(self):
        ckpt = {"global_step": self.global_step,
                "model_state_dict": _to_cpu(self.model.state_dict()),
                "optim_state_dict": _to_cpu(self.optimizer.state_dict())}
        if exists(self.save_path):
            os.rename(self.save_path, self.backup_path)
        torch.save(ckpt, self.save_path)

    def restore(self):
        try:
            ckpt = torch.load(self.save_path, map_location="cpu")
        except Exception:                                   # noqa: BLE001
            ckpt = torch.load(self.backup_path, map_location="cpu")
        self.global_step = ckpt["global_step"]
        load_checkpoint(self.model, ckpt["model_state_dict"])
        dev = self.optimizer.exp_avg.device
        osd = ckpt["optim_state_dict"]
        osd = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in osd.items()}
        self.optimizer.load_state_dict(osd)
