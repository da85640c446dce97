"""Region-ID map files -- the on-disk format on the input side of LRH.

The reference writes one map per image with `skimage.io.imsave(<img path with img_dir -> reg_dir, .png -> .tif>, int32)`
(regda/utils/local_region_homog.py:51-62: id = SAM annotation index + 1 for annotations of >= 1024 px, 0 = no region) and
reads it back with `imread(...).astype(np.int64)` -> LongTensor [1, H, W] (regda/datasets/basedata.py:60-63,74-78).  Same
paths, same dtype, same tensor here; the SAM inference that produces the maps is out of scope."""
from __future__ import annotations

import os

import numpy as np
import torch


def region_path_for(image_path: str) -> str:
    """basedata.py:61-62: <...>/img_dir/<split>/<name>.<ext> -> <...>/reg_dir/<split>/<name>.tif"""
    d, fname = os.path.split(image_path)
    return os.path.join(d.replace('img_dir', 'reg_dir'), f"{fname.split('.')[0]}.tif")


def load_region_map(path: str, device=None) -> torch.Tensor:
    """int32 TIFF -> LongTensor [1, H, W] (0 = no region), what DALoader hands to Homogenizer as `regs_t`"""
    from PIL import Image
    with Image.open(path) as im:
        a = np.array(im)
    if a.ndim != 2:
        raise ValueError(f"{path}: a region map is a single-channel image, got shape {a.shape}")
    if not np.issubdtype(a.dtype, np.integer):
        if not np.array_equal(a, np.floor(a)):
            raise ValueError(f"{path}: non-integer region ids")
    if a.min() < 0:
        raise ValueError(f"{path}: negative region id (the reference's scatter raises on it)")
    t = torch.from_numpy(a.astype(np.int64)).unsqueeze(0)
    return t.to(device) if device is not None else t


def save_region_map(path: str, mask) -> None:
    """LongTensor / ndarray [H, W] or [1, H, W] -> int32 TIFF, as local_region_homog.py:53-62 writes it"""
    from PIL import Image
    a = mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
    a = a.reshape(a.shape[-2:])
    if a.min() < 0 or a.max() > np.iinfo(np.int32).max:
        raise ValueError("region ids must fit a non-negative int32")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    Image.fromarray(a.astype(np.int32), mode="I").save(path)


def load_region_batch(image_paths, device=None) -> torch.Tensor:
    """[b, 1, H, W] int64 for a list of image paths (all tiles of one size)"""
    return torch.stack([load_region_map(region_path_for(p)) for p in image_paths]).to(device) if device is not None else \
        torch.stack([load_region_map(region_path_for(p)) for p in image_paths])
