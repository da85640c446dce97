"""CPU: region-ID map files: int32 TIFF round trip, the reference's path rule and tensor convention
(regda/utils/local_region_homog.py:51-62, regda/datasets/basedata.py:60-78), and that a map read from disk drives the LRH
oracle to the same result as the in-memory tensor."""
import numpy as np
import pytest
import torch

from oracle import cbind
from regda_b200.datasets import regions


def test_path_rule():
    assert regions.region_path_for("data/IsprsDA/Vaihingen/img_dir/train/area1_0_0_512_512.png") == \
        "data/IsprsDA/Vaihingen/reg_dir/train/area1_0_0_512_512.tif"


def test_int32_tiff_round_trip_and_lrh(tmp_path):
    g = np.random.default_rng(3)
    reg = np.kron(g.integers(0, 900, (8, 6)), np.ones((16, 16), dtype=np.int64))          # blocky ids, 0 = background
    p = tmp_path / "reg_dir" / "train" / "tile_a.tif"
    regions.save_region_map(str(p), torch.from_numpy(reg))
    back = regions.load_region_map(str(p))
    assert back.dtype == torch.int64 and tuple(back.shape) == (1, 128, 96)
    assert np.array_equal(back[0].numpy(), reg)
    from PIL import Image
    assert np.array(Image.open(p)).dtype == np.int32          # what the reference's skimage imsave(int32) leaves on disk
    lab = g.integers(-1, 6, (1, 128, 96))
    a = cbind.lrh(lab, reg[None], 6, -1, 0.5)
    b = cbind.lrh(lab, back.numpy(), 6, -1, 0.5)
    assert np.array_equal(a, b)
    batch = regions.load_region_batch([str(tmp_path / "img_dir" / "train" / "tile_a.png")])
    assert tuple(batch.shape) == (1, 1, 128, 96)


def test_rejects_bad_maps(tmp_path):
    with pytest.raises(ValueError):
        regions.save_region_map(str(tmp_path / "neg.tif"), np.full((4, 4), -1))
    from PIL import Image
    Image.fromarray(np.zeros((4, 4, 3), dtype=np.uint8)).save(tmp_path / "rgb.tif")
    with pytest.raises(ValueError):
        regions.load_region_map(str(tmp_path / "rgb.tif"))


def test_eval_tool_cli_surface():
    """tools/eval.py keeps the reference's flags (tools/eval.py:17-25) and refuses to run without a GPU"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("regda_tools_eval", os.path.join(os.path.dirname(__file__), "..", "tools", "eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = mod.parse(["--config-path", "st.regda.tiny", "--ckpt-path", "x.pth", "--tta", "1", "--test", "0"])
    assert a.tta is True and a.test is False and a.ckpt_path == "x.pth" and a.ins_norm is True
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit):
            mod.main(["--config-path", "st.regda.tiny"])
