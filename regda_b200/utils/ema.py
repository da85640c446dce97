"""ExponentialMovingAverage -- drop-in for regda/utils/ema.py:34-65 (defined and imported by the reference's
tools/train_ssl_reg.py:26 but never instantiated there; provided as the opt-in EMA teacher BASELINE.json mentions).

Same surface: EMA(model, decay); register() / update() / apply_shadow() / restore() over the parameters that require
grad.  The shadow is ONE flat float32 buffer; when the model's parameters already live in a trainer.ParamArena,
update() is a single launch of regda_ema_update over the whole arena (shadow = (1-decay)*param + decay*shadow,
ema.py:46-51), otherwise one launch per parameter."""
from __future__ import annotations

import torch

from .. import capi


class ExponentialMovingAverage:
    def __init__(self, model, decay):
        self.model = model
        self.decay = decay
        self.shadow = {}
        self.backup = {}
        self._flat = None

    def _params(self):
        return [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]

    def register(self):                                                # ema.py:41-44
        ps = self._params()
        total = sum((p.numel() + 7) // 8 * 8 for _, p in ps)
        self._flat = torch.empty(total, dtype=torch.float32, device=ps[0][1].device)
        off = 0
        for n, p in ps:
            seg = self._flat[off:off + p.numel()]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                O, I, kh, kw = p.shape                  # keep the parameter's physical (OHWI) order: the kernel works on raw memory
                view = seg.view(O, kh, kw, I).permute(0, 3, 1, 2)
            else:
                view = seg.view(p.shape)
            view.copy_(p.data)
            self.shadow[n] = view
            off += (p.numel() + 7) // 8 * 8

    def update(self):                                                  # ema.py:46-51
        for n, p in self._params():
            assert n in self.shadow
            sh = self.shadow[n]
            capi.call("regda_ema_update", capi.ptr_any(sh), capi.ptr_any(p.data), p.numel(), float(self.decay), capi.stream())

    def _written(self, params):
        """the fp32 parameters were rewritten outside the SGD kernel: refresh the bf16 copies the tcgen05 convolutions read
        (a trainer.ParamArena keeps one shadow for all of them; a free-standing parameter's cached copy is keyed on
        `_version`, which the no_grad `p.copy_()` below bumps)"""
        arenas = {id(a): a for a in (getattr(p, "_arena", None) for _, p in params) if a is not None}
        for a in arenas.values():
            a.sync_shadow()

    @torch.no_grad()
    def apply_shadow(self):                                            # ema.py:53-58
        ps = self._params()
        for n, p in ps:
            assert n in self.shadow
            self.backup[n] = p.data.clone()
            p.copy_(self.shadow[n])
        self._written(ps)

    @torch.no_grad()
    def restore(self):                                                 # ema.py:60-65
        ps = self._params()
        for n, p in ps:
            assert n in self.backup
            p.copy_(self.backup[n])
        self.backup = {}
        self._written(ps)
