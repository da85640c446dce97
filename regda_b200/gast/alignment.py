"""Aligner / DownscaleLabel -- drop-in for the parts of regda/gast/alignment.py that are on the
self-training path: prototypes state, update_prototype (:86-90), update_avg / init_avg
(:107-126), label_refine with label_t_sup=None (:194-265), _pearson_dist (:396-423).

`align_domain` (CORAL, alignment.py:79-84; stage 2 with --align-domain 1) is gast/coral.py on the tcgen05 kernels; the
class / instance alignment and whitening losses have no live call site and raise NotImplementedError.

Feature maps are accepted as the reference passes them ([b,k,h,w]); a channels-last tensor
(the layout the CUDA model produces) is consumed without a copy.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import capi


def _rows(feat, keep_bf16=False):
    """[b,k,h,w] -> contiguous channels-last rows [b*h*w, k] (no copy for channels_last input): float32, or -- keep_bf16 -- the
    bf16 rows as they are (the kernels that take them convert on load)"""
    b, k, h, w = feat.shape
    f = feat.detach()
    if not (keep_bf16 and f.dtype == torch.bfloat16 and k % 8 == 0):
        f = f.float()
    return f.permute(0, 2, 3, 1).contiguous().view(b * h * w, k)


class DownscaleLabel(nn.Module):
    """alignment.py:456-481."""

    def __init__(self, scale_factor=16, n_classes=7, ignore_label=-1, min_ratio=0.75):
        super().__init__()
        assert scale_factor > 1
        self.scale_factor = scale_factor
        self.n_classes = n_classes
        self.ignore_label = ignore_label
        self.min_ratio = min_ratio
        self._flags = None

    def forward(self, label):
        if len(label.shape) == 4:
            label = label.squeeze(dim=1)
        assert len(label.shape) == 3
        if not label.is_cuda:
            raise RuntimeError("regda_b200.DownscaleLabel needs a CUDA tensor (no CPU fallback)")
        bs, H, W = label.shape
        lab = label.long().contiguous()
        out = torch.empty((bs, 1, H // self.scale_factor, W // self.scale_factor), dtype=torch.int64, device=label.device)
        if self._flags is None or self._flags.device != label.device:
            self._flags = torch.zeros(1, dtype=torch.int32, device=label.device)
        capi.call("regda_downscale_label", capi.ptr(lab), capi.ptr(out), bs, H, W, self.scale_factor, self.n_classes,
                  int(self.ignore_label), float(self.min_ratio), capi.ptr(self._flags), capi.stream())
        return out


class Aligner:
    def __init__(self, logger=None, feat_channels=64, class_num=7, ignore_label=-1, decay=0.999, topk=32, resume=None,
                 device=None):
        self.feat_channels = feat_channels
        self.class_num = class_num
        self.ignore_label = ignore_label
        self.decay = decay
        self.logger = logger
        self.eps = 1e-7
        self.topk = topk
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if resume:
            self.prototypes = torch.load(resume, map_location='cpu').float().to(dev).contiguous()
            if logger is not None:
                logger.info('finish init prototypes!')
                logger.info(f'prototypes({self.prototypes.shape})={self.prototypes}')
        else:
            self.prototypes = torch.zeros([class_num, feat_channels], device=dev)
        self.downscale_gt = DownscaleLabel(scale_factor=16, n_classes=class_num, ignore_label=ignore_label, min_ratio=0.75)
        self._data_sum = torch.zeros([class_num, feat_channels], device=dev)
        self._data_cnt = torch.zeros([class_num, 1], device=dev)
        self._sums = torch.zeros([class_num, feat_channels], device=dev)
        self._cnt = torch.zeros([class_num], device=dev)

    # ---- prototypes ----------------------------------------------------------------------
    def _class_sums(self, feat, label_ds, sums, counts, accumulate):
        rows = _rows(feat, keep_bf16=True)
        n, k = rows.shape
        lab = label_ds.reshape(-1).contiguous()
        assert lab.numel() == n, "label and feature map disagree"
        ws = capi.workspace.get(capi.lib().regda_class_sums_workspace_bytes(n, self.class_num, k), rows.device)
        capi.call("regda_class_sums_bf16feat" if rows.dtype == torch.bfloat16 else "regda_class_sums", capi.ptr(rows), capi.ptr(lab), capi.ptr(sums), capi.ptr(counts), n, self.class_num, k,
                  int(self.ignore_label), int(accumulate), capi.ptr(ws), ws.numel(), capi.stream())

    def update_prototype(self, feat, label, reduce_fn=None):
        """alignment.py:86-90.  `reduce_fn(sums, counts)` (optional) is called between the
        segmented sum and the EMA -- the data-parallel trainer all-reduces there so that every
        rank keeps identical prototypes."""
        label = self.downscale_gt(label)
        self._compute_local_prototypes(feat, label, update=True, decay=self.decay, reduce_fn=reduce_fn)
        return label

    def _compute_local_prototypes(self, feat, label, update=False, decay=0.999, reduce_fn=None):
        assert 0 < decay < 1
        if not update:
            raise NotImplementedError("only the update=True form is on the self-training path")
        self._class_sums(feat, label, self._sums, self._cnt, accumulate=False)
        if reduce_fn is not None:
            reduce_fn(self._sums, self._cnt)
        if not self.prototypes.is_contiguous():
            self.prototypes = self.prototypes.contiguous()
        capi.call("regda_prototype_ema", capi.ptr(self.prototypes), capi.ptr(self._sums), capi.ptr(self._cnt),
                  self.class_num, self.prototypes.shape[1], float(decay), capi.stream())
        return None

    def update_avg(self, feat, label) -> None:
        """alignment.py:107-119 (tools/init_prototypes.py)."""
        labels = self.downscale_gt(label)
        self._class_sums(feat, labels, self._data_sum, self._data_cnt.view(-1), accumulate=True)

    def init_avg(self):
        """alignment.py:121-126."""
        self.prototypes = torch.empty_like(self._data_sum)
        capi.call("regda_prototype_init_avg", capi.ptr(self.prototypes), capi.ptr(self._data_sum), capi.ptr(self._data_cnt),
                  self.class_num, self.prototypes.shape[1], capi.stream())
        if self.logger is not None:
            self.logger.info('finish init prototypes!')
            self.logger.info(f'examples cnt({self._data_cnt.shape})={self._data_cnt}')
            self.logger.info(f'prototypes({self.prototypes.shape})={self.prototypes}')

    # ---- label refinement ------------------------------------------------------------------
    def _pearson_dist(self, feat1, feat2):
        """alignment.py:396-423: feat1 [n,k], feat2 [m,k] -> [n,m]."""
        assert feat1.shape[-1] == feat2.shape[-1]
        rows, protos = feat1.detach().float().contiguous(), feat2.detach().float().contiguous()
        n, k = rows.shape
        m = protos.shape[0]
        out = torch.empty((n, m), dtype=torch.float32, device=rows.device)
        ws = capi.workspace.get(capi.lib().regda_pearson_workspace_bytes(m, k), rows.device)
        capi.call("regda_pearson_dist", capi.ptr(rows), capi.ptr(protos), capi.ptr(out), n, m, k, capi.ptr(ws), ws.numel(), capi.stream())
        return out

    def _refine_inputs(self, feat_t, preds_t, label_t_soft, keep_bf16=False, select=False):
        b, k, h, w = feat_t.shape
        H, W = label_t_soft.shape[-2:]
        if isinstance(preds_t, (list, tuple)):
            assert len(preds_t) == 2
            p1, p2 = preds_t[0].detach().float().contiguous(), preds_t[1].detach().float().contiguous()
        else:
            p1, p2 = preds_t.detach().float().contiguous(), None
        rows = _rows(feat_t, keep_bf16)
        soft = label_t_soft.detach().float().contiguous()
        c = soft.shape[1]
        assert c == self.class_num == p1.shape[1]
        nbytes = (capi.lib().regda_refine_select_workspace_bytes(b, c, k, h, w, H, W) if select
                  else capi.lib().regda_refine_workspace_bytes(b, c, k, h, w))
        ws = capi.workspace.get(nbytes, rows.device)
        return rows, p1, p2, soft, ws, (b, c, k, h, w, H, W)

    def label_refine(self, label_t_sup, feat_t, preds_t, label_t_soft, refine=True, mode='all', temp=2.0):
        """alignment.py:194-265 for the form the self-training loop uses (label_t_sup=None, mode='all')."""
        assert mode in ['all', 's', 'p', 'n', 'l']
        if not refine:
            return label_t_soft
        if label_t_sup is not None or mode != 'all':
            raise NotImplementedError("regda_b200 implements label_refine(label_t_sup=None, mode='all') -- the form "
                                      "tools/train_ssl_reg.py:214 calls; other modes are off the hot path")
        rows, p1, p2, soft, ws, (b, c, k, h, w, H, W) = self._refine_inputs(feat_t, preds_t, label_t_soft)
        out = torch.empty_like(soft)
        capi.call("regda_label_refine", capi.ptr(rows), capi.ptr(self.prototypes), capi.ptr(p1), capi.ptr(p2), capi.ptr(soft),
                  capi.ptr(out), b, c, k, h, w, H, W, float(temp), capi.ptr(ws), ws.numel(), capi.stream())
        return out

    def refine_select(self, feat_t, preds_t, label_t_soft, temp=2.0, cutoff_top=0.8, cutoff_low=0.6):
        """Fused label_refine -> pseudo_selection (tools/train_ssl_reg.py:214-218): returns the hard
        pseudo label [b,H,W] int64 without materialising the refined probabilities."""
        rows, p1, p2, soft, ws, (b, c, k, h, w, H, W) = self._refine_inputs(feat_t, preds_t, label_t_soft, keep_bf16=True, select=True)
        out = torch.empty((b, H, W), dtype=torch.int64, device=soft.device)
        capi.call("regda_refine_select_bf16feat" if rows.dtype == torch.bfloat16 else "regda_refine_select", capi.ptr(rows), capi.ptr(self.prototypes), capi.ptr(p1), capi.ptr(p2), capi.ptr(soft),
                  capi.ptr(out), b, c, k, h, w, H, W, float(temp), float(cutoff_top), float(cutoff_low), int(self.ignore_label),
                  capi.ptr(ws), ws.numel(), capi.stream())
        return out

    # ---- stage-1/2 alignment losses --------------------------------------------------------
    def align_domain(self, feat_s, feat_t, precise=False):
        """alignment.py:79-84: Deep CORAL between the source and the target feature rows (gast/coral.py; the covariance
        contractions run on the tcgen05 kernels).  precise: float32-accuracy contractions (the float32 parity mode)."""
        assert feat_s.shape == feat_t.shape, 'tensor "feat_s" has the same shape as tensor "feat_t"'
        assert len(feat_s.shape) == 4, 'tensor "feat_s" and "feat_t" must have 4 dimensions'
        from .coral import CoralLoss
        k = self.feat_channels
        rows_s = feat_s.permute(0, 2, 3, 1).reshape([-1, k])
        rows_t = feat_t.permute(0, 2, 3, 1).reshape([-1, k])
        return CoralLoss(precise=precise)(rows_s, rows_t)

    def _out_of_scope(self, *a, **k):
        raise NotImplementedError("class / instance alignment and whitening losses have no call site on the stage-2/3 paths "
                                  "(regda/gast/alignment.py:128-170 are only reached from dead code)")

    align_class = align_instance = whiten_class_ware = _out_of_scope
