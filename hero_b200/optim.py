"""Fused AdamW over the flat parameter buffer, with the reference's exact update rule
(optim/adamw.py:80-104) and parameter grouping (optim/misc.py:14-50): eps 1e-6 added to
sqrt(v), bias-corrected step size, decoupled weight decay applied after the Adam update with
the un-corrected lr, no decay for names containing 'bias' / 'LayerNorm.*'.

The reference launches ~10 pointwise kernels for each of ~208 tensors per step; here the flat
layout (decayed parameters first, see params.py) needs TWO launches, and the same kernel refreshes
the bf16 working copy so the next forward skips its cast pass.
"""
import math

import torch

from . import ops
from .params import FlatParams


class FusedAdamW:
    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01,
                 correct_bias=True):
        assert isinstance(flat, FlatParams) and flat.flat is not None, \
            "call flat_of(model, device) (or run one forward) before building the optimizer"
        self.flat = flat
        self._generation = flat.generation
        self.correct_bias = correct_bias
        self.eps = eps
        self.betas = betas
        split = flat.no_decay_start
        # `param_groups` keeps the reference loop working:
        #     for g in optimizer.param_groups: g['lr'] = lr_this_step   (train_vcmr.py:245-247)
        self.param_groups = [
            {"lr": lr, "weight_decay": weight_decay, "range": (0, split)},
            {"lr": lr, "weight_decay": 0.0, "range": (split, flat.total)},
        ]
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_count = 0
        flat.ensure_flat_grads()

    def zero_grad(self, set_to_none=False):
        self.flat.ensure_flat_grads().zero_()

    def grad_norm(self):
        g = self.flat.ensure_flat_grads()
        acc = torch.zeros(1, dtype=torch.float32, device=g.device)
        ops.sumsq(g, acc)
        return acc.sqrt()

    def clip_grad_norm_(self, max_norm):
        """Global-norm clip folded into the update (train_vcmr.py:258-259): returns the norm and
        remembers the scale for the next step()."""
        total = float(self.grad_norm().item())
        self._grad_scale = min(1.0, max_norm / (total + 1e-6))
        return total

    def clip_grad_norm_device_(self, max_norm):
        """Global-norm clip (train_vcmr.py:258-259) with no device->host read: the sum of squares
        stays on the device and the next step()'s AdamW kernels scale the gradients by
        min(1, max_norm / (norm + 1e-6)) themselves. Returns the device scalar (sum of squares)."""
        g = self.flat.ensure_flat_grads()
        if getattr(self, "_sumsq", None) is None:
            self._sumsq = torch.zeros(1, dtype=torch.float32, device=g.device)
        self._sumsq.zero_()
        ops.sumsq(g, self._sumsq)
        self._clip = float(max_norm)
        return self._sumsq

    def step(self):
        if self.flat.generation != self._generation or self.exp_avg.numel() != self.flat.total:
            raise RuntimeError("the parameters were re-flattened (a module was replaced or moved) "
                               "after this optimizer was built: its moment buffers and ranges no "
                               "longer describe the flat buffer; rebuild the optimizer")
        self.step_count += 1
        t = self.step_count
        b1, b2 = self.betas
        g = self.flat.ensure_flat_grads()
        scale = getattr(self, "_grad_scale", 1.0)
        self._grad_scale = 1.0
        clip = getattr(self, "_clip", None)
        self._clip = None
        for grp in self.param_groups:
            a, b = grp["range"]
            if b <= a:
                continue
            lr = grp["lr"]
            step_size = lr
            if self.correct_bias:
                step_size = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            ops.adamw_step(self.flat.flat[a:b], g[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                           self.flat.mirror[a:b], step_size=step_size, beta1=b1, beta2=b2,
                           eps=self.eps, lr_wd=lr * grp["weight_decay"], grad_scale=scale,
                           clip_sumsq=self._sumsq if clip is not None else None,
                           clip_max_norm=clip or 0.0)
        # masters changed in place through a flat view: the mirror is already fresh
        self.flat.dirty = False
        self.flat._version_sum = sum(p._version for _, p, _, _ in self.flat._probe)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items()} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g["lr"], g["weight_decay"] = s["lr"], s["weight_decay"]


def build_optimizer(model, opts, device=None):
    """optim/misc.py:14-50 for optim == 'adamw' on the flat layout. The reference builds FOUR
    groups (top x lr_mul decay / no-decay, v_encoder decay / no-decay) and some of its loops index
    them (`if i in (0, 1): lr *= lr_mul`, train_videoQA.py); the flat layout has TWO ranges
    (decay, no-decay), so anything that would make the four groups differ is refused instead of
    being silently collapsed: lr_mul != 1, optimizers other than AdamW, frozen parameters."""
    from .params import flat_of
    if getattr(opts, "optim", "adamw") != "adamw":
        raise ValueError(f"hero_b200.build_optimizer implements 'adamw' only (got {opts.optim!r}); "
                         "use the reference's torch optimizer for adam / adamax")
    if float(getattr(opts, "lr_mul", 1.0)) != 1.0:
        raise ValueError("lr_mul != 1 needs the reference's four parameter groups; FusedAdamW "
                         "keeps two flat ranges (decay / no-decay) with one learning rate")
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    if frozen:
        raise ValueError(f"frozen parameters ({frozen[:3]}...) are not supported by the flat "
                         "optimizer: every element of the flat buffer is updated")
    device = device or next(model.parameters()).device
    flat = flat_of(model, device)
    return FusedAdamW(flat, lr=opts.learning_rate, betas=tuple(opts.betas),
                      weight_decay=opts.weight_decay)
