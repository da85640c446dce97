"""TEST INFRASTRUCTURE ONLY: ctypes/numpy binding of oracle/regda_oracle.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libregda_oracle.so")
_lib = None

ERRORS = {1: "label out of range", 2: "region id out of range", 3: "alloc", 4: "probability outside [0,1]"}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle: {ERRORS.get(code, code)}")
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "regda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, dbl, ci, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p
        _lib.oracle_lrh.argtypes = [vp, vp, vp, i64, i64, ci, i64, dbl]
        _lib.oracle_pseudo_select.argtypes = [vp, vp, i64, ci, i64, dbl, dbl, i64]
        _lib.oracle_downscale_label.argtypes = [vp, vp, i64, i64, i64, ci, ci, i64, dbl]
        _lib.oracle_class_count.argtypes = [vp, i64, ci, i64, vp, vp]
        for f in (_lib.oracle_lrh, _lib.oracle_pseudo_select, _lib.oracle_downscale_label, _lib.oracle_class_count):
            f.restype = ci
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lrh(labels, regions, class_num, ignore_label, percent):
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    regions = np.ascontiguousarray(regions, dtype=np.int64)
    assert labels.ndim == 3 and labels.shape == regions.shape
    out = np.empty_like(labels)
    b = labels.shape[0]
    rc = lib().oracle_lrh(_p(labels), _p(regions), _p(out), b, labels[0].size if b else 0,
                          class_num, ignore_label, float(percent))
    if rc:
        raise OracleError(rc)
    return out


def pseudo_select(soft, cutoff_top=0.8, cutoff_low=0.6, ignore_label=-1):
    soft = np.ascontiguousarray(soft, dtype=np.float32)
    b, c, h, w = soft.shape
    out = np.empty((b, h, w), dtype=np.int64)
    rc = lib().oracle_pseudo_select(_p(soft), _p(out), b, c, h * w, float(cutoff_top), float(cutoff_low), ignore_label)
    if rc:
        raise OracleError(rc)
    return out


def downscale_label(label, scale=16, n_classes=6, ignore_label=-1, min_ratio=0.75):
    label = np.ascontiguousarray(label, dtype=np.int64)
    if label.ndim == 4:
        label = label[:, 0]
    b, H, W = label.shape
    out = np.empty((b, 1, H // scale, W // scale), dtype=np.int64)
    rc = lib().oracle_downscale_label(_p(label), _p(out), b, H, W, scale, n_classes, ignore_label, float(min_ratio))
    if rc:
        raise OracleError(rc)
    return out


def class_count(label, class_num, ignore_label):
    label = np.ascontiguousarray(label, dtype=np.int64)
    counts = np.zeros(class_num, dtype=np.int64)
    nv = np.zeros(1, dtype=np.int64)
    rc = lib().oracle_class_count(_p(label), label.size, class_num, ignore_label, _p(counts), _p(nv))
    if rc:
        raise OracleError(rc)
    return counts, int(nv[0])
