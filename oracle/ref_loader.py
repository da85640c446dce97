"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (StuLiu/RegDA).

Imports the reference's own Python modules from ``/root/reference`` on a host
that lacks seven of its third-party dependencies, by pre-seeding ``sys.modules``
with structural stubs.  No arithmetic is restated here:
``torch_scatter.scatter(reduce='sum')`` is forwarded to ``Tensor.scatter_add_``,
which is what torch_scatter's own Python wrapper does for that reduction
(reference call site: regda/utils/local_region_homog.py:140).

The reference tree does not exist on the GPU box, so this module may only be
used (a) by ``tests/golden/make_golden.py`` in the build container to produce
the committed fixtures and (b) by CPU tests that skip when the tree is absent.
Nothing under ``regda_b200/`` may import it.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("REGDA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "regda"))


class _AttrDict(dict):
    """dict with attribute access and recursive update (stand-in for ever's AttrDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, other=(), **kw):  # noqa: A003
        items = dict(other, **kw)
        for k, v in items.items():
            if isinstance(v, dict):
                cur = self.get(k)
                if not isinstance(cur, _AttrDict):
                    cur = _AttrDict()
                    dict.__setitem__(self, k, cur)
                cur.update(v)
            else:
                dict.__setitem__(self, k, v)


class _ERModule(nn.Module):
    """nn.Module + config mixin: defaults from set_default_config(), then user dict merged."""

    def __init__(self, config=None):
        super().__init__()
        self._cfg = _AttrDict()
        self.set_default_config()
        if config:
            self._cfg.update(config)

    @property
    def config(self):
        return self._cfg

    def set_default_config(self):
        pass


class _Registry(dict):
    def register(self, name=None, obj=None):
        if obj is not None:
            self[name] = obj
            return obj

        def deco(o):
            self[name or o.__name__] = o
            return o

        return deco


def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce not in ("sum", "add"):
        raise NotImplementedError("oracle loader only forwards reduce='sum' (the only one on the path)")
    index = index.expand_as(src)
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = list(src.shape)
    shape[dim] = dim_size
    return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)


# ttach==0.0.3 (requirement.txt:165) is absent from this image.  Restatement of the three classes regda/utils/tools.py:132-152
# uses, from the package's published source (ttach/base.py Compose / Transformer, ttach/transforms.py HorizontalFlip / Rotate90,
# ttach/functional.py hflip = x.flip(3), rot90 = torch.rot90(x, k, (2, 3))): Compose iterates itertools.product of the
# transforms' parameter lists in order; a view augments by applying the transforms in order and de-augments a mask by applying
# the inverses in REVERSE order.  PARITY NOTE: this stub is a restatement, not the package -- the teacher-pass goldens are
# pinned to it.
class _TtaHorizontalFlip:
    params = [False, True]

    def aug(self, x, p):
        return x.flip(3) if p else x

    deaug = aug


class _TtaRotate90:
    def __init__(self, angles):
        self.params = list(angles) if 0 in angles else [0] + list(angles)

    def aug(self, x, angle):
        import torch
        k = angle // 90 if angle >= 0 else (angle + 360) // 90
        return torch.rot90(x, k, (2, 3))

    def deaug(self, x, angle):
        return self.aug(x, -angle)


class _TtaView:
    def __init__(self, transforms, params):
        self.transforms, self.params = transforms, params

    def augment_image(self, x):
        for t, p in zip(self.transforms, self.params):
            x = t.aug(x, p)
        return x

    def deaugment_mask(self, x):
        for t, p in zip(self.transforms[::-1], self.params[::-1]):
            x = t.deaug(x, p)
        return x


class _TtaCompose:
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __iter__(self):
        import itertools
        for params in itertools.product(*[t.params for t in self.transforms]):
            yield _TtaView(self.transforms, params)

    def __len__(self):
        n = 1
        for t in self.transforms:
            n *= len(t.params)
        return n


# ever-beta==0.2.3 (requirement.txt:33) is absent from this image.  Restatement of ever/api/metric/pixel.py PixelMetric from the
# package's published source -- the base class of the reference's own PixelMetricIgnore (regda/gast/metrics.py:19), which is
# imported UNMODIFIED on top of it: a float32 scipy-sparse [C,C] confusion matrix (rows = y_true, columns = y_pred) accumulated
# by forward(); static per-class IoU / F / precision / recall on the dense matrix.  PARITY NOTE: a restatement, not the
# package -- tests/golden/miou.npz is pinned to it (like the ttach stub above).
class _PixelMetric:
    def __init__(self, num_classes, logdir=None, logger=None, class_names=None):
        import numpy as np
        from scipy import sparse
        self.num_classes = num_classes
        self._total = sparse.coo_matrix((num_classes, num_classes), dtype=np.float32)
        self._logger = logger
        self._class_names = class_names
        self.logdir = logdir

    def reset(self):
        import numpy as np
        from scipy import sparse
        self._total = sparse.coo_matrix((self.num_classes, self.num_classes), dtype=np.float32)

    def forward(self, y_true, y_pred):
        import numpy as np
        from scipy import sparse
        if isinstance(y_pred, torch.Tensor):
            y_pred = y_pred.cpu().numpy()
        if isinstance(y_true, torch.Tensor):
            y_true = y_true.cpu().numpy()
        y_pred = y_pred.reshape((-1,))
        y_true = y_true.reshape((-1,))
        v = np.ones_like(y_pred)
        cm = sparse.coo_matrix((v, (y_true, y_pred)), shape=(self.num_classes, self.num_classes), dtype=np.float32)
        self._total += cm
        return cm

    def _log_summary(self, table, dense_cm):
        pass

    @staticmethod
    def compute_iou_per_class(confusion_matrix):
        import numpy as np
        sum_over_row = np.sum(confusion_matrix, axis=0)
        sum_over_col = np.sum(confusion_matrix, axis=1)
        diag = np.diag(confusion_matrix)
        return diag / (sum_over_row + sum_over_col - diag)

    @staticmethod
    def compute_recall_per_class(confusion_matrix):
        import numpy as np
        return np.diag(confusion_matrix) / np.sum(confusion_matrix, axis=1)

    @staticmethod
    def compute_precision_per_class(confusion_matrix):
        import numpy as np
        return np.diag(confusion_matrix) / np.sum(confusion_matrix, axis=0)

    @staticmethod
    def compute_F_measure_per_class(confusion_matrix, beta=1.0):
        p = _PixelMetric.compute_precision_per_class(confusion_matrix)
        r = _PixelMetric.compute_recall_per_class(confusion_matrix)
        return (1 + beta ** 2) * p * r / ((beta ** 2) * p + r)


class _PrettyTable:
    """prettytable.PrettyTable surface used by regda/gast/metrics.py:47-61 (field_names, add_row)"""

    def __init__(self):
        self.field_names = []
        self.rows = []

    def add_row(self, row):
        self.rows.append(list(row))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    _installed = True
    _mod("torch_scatter", scatter=_scatter)
    _mod("segment_anything", sam_model_registry={}, SamAutomaticMaskGenerator=object, SamPredictor=object)
    sk = _mod("skimage")
    sk.io = _mod("skimage.io", imread=None, imsave=None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = _mod("matplotlib")
        mp.pyplot = _mod("matplotlib.pyplot")
    _mod("ttach", Compose=_TtaCompose, HorizontalFlip=_TtaHorizontalFlip, Rotate90=_TtaRotate90)
    _mod("prettytable", PrettyTable=_PrettyTable)

    import logging

    ever = _mod("ever", ERModule=_ERModule)
    registry = _mod("ever.core.registry", MODEL=_Registry(), Registry=_Registry)
    logger = _mod("ever.core.logger", get_logger=lambda *a, **k: logging.getLogger("ever"))
    core = _mod("ever.core", registry=registry, logger=logger)
    interface = _mod("ever.interface", ERModule=_ERModule)

    def freeze_params(module):
        for p in module.parameters():
            p.requires_grad = False

    def freeze_modules(module, kind):
        for m in module.modules():
            if isinstance(m, kind):
                freeze_params(m)

    param_util = _mod("ever.util.param_util", freeze_params=freeze_params, freeze_modules=freeze_modules)
    util = _mod("ever.util", param_util=param_util)
    pixel = _mod("ever.api.metric.pixel", PixelMetric=_PixelMetric)
    metric = _mod("ever.api.metric", pixel=pixel)
    api = _mod("ever.api", metric=metric)
    ever.core, ever.interface, ever.util, ever.registry, ever.api = core, interface, util, registry, api

    if not torch.cuda.is_available():
        # Aligner.__init__ calls .cuda() (alignment.py:48,56,60,76-77)
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return a namespace holding the reference's own hot-path symbols."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    import logging

    from regda.gast.alignment import Aligner, DownscaleLabel
    from regda.gast.balance import ClassBalance, CrossEntropy
    from regda.gast.pseudo_generation import pseudo_selection
    from regda.models.Encoder import Deeplabv2
    from regda.utils.local_region_homog import Homogenizer
    from regda.utils.tools import adjust_learning_rate, loss_calc
    from regda.utils.ema import ExponentialMovingAverage

    ns = types.SimpleNamespace(
        Aligner=Aligner,
        DownscaleLabel=DownscaleLabel,
        ClassBalance=ClassBalance,
        CrossEntropy=CrossEntropy,
        pseudo_selection=pseudo_selection,
        Deeplabv2=Deeplabv2,
        Homogenizer=Homogenizer,
        adjust_learning_rate=adjust_learning_rate,
        loss_calc=loss_calc,
        ExponentialMovingAverage=ExponentialMovingAverage,
        logger=logging.getLogger("regda-ref"),
    )
    return ns


def build_reference_model(ns, resnet_type="resnet101", class_num=6):
    """The model dict of tools/train_ssl_reg.py:94-111 with pretrained=False (no network)."""
    return ns.Deeplabv2(dict(
        backbone=dict(resnet_type=resnet_type, output_stride=16, pretrained=False),
        multi_layer=True, cascade=False, use_ppm=True,
        ppm=dict(num_classes=class_num, use_aux=False, fc_dim=2048),
        inchannels=2048, num_classes=class_num, is_ins_norm=True,
    ))
