"""Local Region Homogenizing -- drop-in for regda/utils/local_region_homog.py:99-152.

`Homogenizer(percent, class_num, ignore_label)(pseudo_labels, regions)` has the reference's
signature, dtype contract (int64 in, fresh int64 out, inputs never written) and error
behaviour, but runs as ONE CUDA kernel per call (regda_lrh_forward) instead of
clone -> one_hot -> torch_scatter.scatter -> max -> gather -> where.

Host synchronisation: the reference sizes its histogram with `regions.max()+1` inside
scatter (a device->host sync every step).  Pass `region_bound=` (any upper bound on
max(regions)+1, e.g. known from the region files) to run sync-free; ids outside the bound
and labels outside [0, class_num] are then reported through `check()` / `strict=True`
exactly where the reference would have raised.
"""
from __future__ import annotations

import torch

from .. import capi


class Homogenizer(torch.nn.Module):
    def __init__(self, percent=0.9, class_num=6, ignore_label=255, region_bound=None, strict=True):
        super().__init__()
        self.percent = percent
        self.class_num = class_num
        self.ignore_label = ignore_label
        self.region_bound = region_bound   # None: measured per call on the device (one sync, like the reference)
        self.strict = strict               # True: raise like the reference on out-of-domain input (one sync)
        self._flags = None

    def _flag_word(self, device):
        if self._flags is None or self._flags.device != device:
            self._flags = torch.zeros(1, dtype=torch.int32, device=device)
        return self._flags

    def check(self):
        """Raise if any call since the last check saw out-of-domain input (syncs)."""
        if self._flags is None:
            return
        f = int(self._flags.item())
        if f:
            self._flags.zero_()
        if f & capi.FLAG_LABEL_RANGE:
            raise RuntimeError("Class values must be smaller than num_classes.")   # one_hot's message (:121)
        if f & capi.FLAG_REGION_RANGE:
            raise RuntimeError("index out of range in scatter (region id < 0 or >= region_bound)")  # (:140)

    def forward(self, pseudo_labels, regions):
        assert pseudo_labels.dim() == 3                                     # (:133)
        if not pseudo_labels.is_cuda:
            raise RuntimeError("regda_b200.Homogenizer needs CUDA tensors (no CPU fallback)")
        if pseudo_labels.dtype != torch.int64 or regions.dtype != torch.int64:
            raise TypeError("Homogenizer expects int64 (LongTensor) labels and regions, as the reference does")
        b, h, w = pseudo_labels.shape
        if regions.numel() != pseudo_labels.numel():
            raise RuntimeError(f"regions with {regions.numel()} elements cannot be viewed as ({b}, {h * w}, 1)")
        labels = pseudo_labels.contiguous()
        regs = regions.contiguous()
        out = torch.empty_like(labels)
        if labels.numel() == 0:
            return out
        flags = self._flag_word(labels.device)
        bound = self.region_bound
        if bound is None:
            bound_t = torch.empty(1, dtype=torch.int64, device=labels.device)
            capi.call("regda_region_bound", capi.ptr(regs), regs.numel(), capi.ptr(bound_t), capi.ptr(flags), capi.stream())
            bound = max(int(bound_t.item()), 1)
        hw = h * w
        ws_bytes = capi.lib().regda_lrh_workspace_bytes(b, hw, self.class_num, bound)
        ws = capi.workspace.get(ws_bytes, labels.device)
        capi.call("regda_lrh_forward", capi.ptr(labels), capi.ptr(regs), capi.ptr(out), b, hw, self.class_num,
                  self.ignore_label, float(self.percent), bound, capi.ptr(flags), capi.ptr(ws), ws.numel(), capi.stream())
        if self.strict:
            self.check()
        return out
