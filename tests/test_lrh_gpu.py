"""GPU: Local Region Homogenizing (regda_lrh_forward through the Homogenizer mirror) against
the golden vectors of the unmodified reference and the C oracle.  Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import lrh_golden_cases
from oracle import cbind

pytestmark = pytest.mark.gpu

CASES = lrh_golden_cases()
PATHS = {"auto": 0, "generic": 1}


def _hom(**kw):
    from regda_b200.utils.local_region_homog import Homogenizer
    return Homogenizer(**kw)


def _set_path(mode):
    from regda_b200 import capi
    capi.check(capi.lib().regda_set_lrh_path(mode))


def _last_path():
    import ctypes
    from regda_b200 import capi
    cl = ctypes.c_int(0)
    return capi.lib().regda_lrh_last_path(ctypes.byref(cl)), cl.value


@pytest.fixture(autouse=True)
def _reset_path():
    yield
    _set_path(0)


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("case", CASES, ids=[str(c["name"]) for c in CASES])
def test_golden_bit_exact(case, path):
    _set_path(PATHS[path])
    lab = torch.from_numpy(case["labels"]).cuda()
    reg = torch.from_numpy(case["regions"]).cuda()
    lab0, reg0 = lab.clone(), reg.clone()
    h = _hom(percent=float(case["percent"]), class_num=int(case["class_num"]), ignore_label=int(case["ignore"]))
    out = h(lab, reg)
    assert out.dtype == torch.int64 and out.shape == lab.shape
    assert out.data_ptr() != lab.data_ptr()
    assert torch.equal(lab, lab0) and torch.equal(reg, reg0)          # inputs are never mutated
    assert np.array_equal(out.cpu().numpy(), case["out"])


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("n_regions", [50, 500, 5000])
@pytest.mark.parametrize("shape", [(2, 512, 512), (1, 1024, 1024), (3, 96, 200), (2, 37, 53)])
def test_random_tiles_vs_oracle(shape, n_regions, path):
    from regda_b200 import synth
    _set_path(PATHS[path])
    b, hh, ww = shape
    reg = synth.region_maps(b, hh, ww, n_regions, device="cuda", seed=7 + n_regions)
    lab = synth.lrh_labels(reg, 6, -1, seed=11)
    for pct in (0.5, 0.9):
        out = _hom(percent=pct, class_num=6, ignore_label=-1)(lab, reg)
        want = cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), 6, -1, pct)
        assert np.array_equal(out.cpu().numpy(), want)
    if path == "auto" and shape == (2, 512, 512) and n_regions <= 500:
        assert _last_path()[0] == 2, "the aligned case must take the cluster (fast) path"


def test_exact_percent_regions_straddling_256_at_scale():
    """regions built at exactly `percent` with valid counts on both sides of 256 inside a big tile."""
    b, hh, ww = 2, 512, 512
    lab = torch.full((b, hh * ww), -1, dtype=torch.int64)
    reg = torch.zeros((b, hh * ww), dtype=torch.int64)
    pos = 0
    rid = 1
    for n in (4, 100, 254, 256, 258, 512, 1000, 1024, 4096, 30000):
        k = n // 2
        reg[:, pos:pos + n + 5] = rid          # 5 ignored px inside the region
        lab[:, pos:pos + k] = 2
        lab[:, pos + k:pos + n] = 4
        pos += n + 5
        rid += 3
    lab, reg = lab.view(b, hh, ww).cuda(), reg.view(b, hh, ww).cuda()
    want = cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), 6, -1, 0.5)
    for mode in (0, 1):
        _set_path(mode)
        out = _hom(percent=0.5, class_num=6, ignore_label=-1)(lab, reg)
        assert np.array_equal(out.cpu().numpy(), want)
    flat = want.reshape(b, -1)
    assert (flat[0, :2] == np.array([2, 2])).all() and flat[0, 3] == 4      # n=4: unchanged
    start256 = 4 + 5 + 100 + 5 + 254 + 5
    assert (flat[0, start256:start256 + 256 + 5] == 2).all()                # n=256: homogenised, ignored px filled


@pytest.mark.parametrize("class_num,ignore", [(7, -1), (16, 255), (2, -1), (1, -1), (17, -1), (40, 255)])
def test_class_counts_and_ignore_values(class_num, ignore):
    from regda_b200 import synth
    reg = synth.region_maps(2, 64, 128, 30, device="cuda", seed=5)
    lab = synth.lrh_labels(reg, class_num, ignore, seed=6)
    for mode in (0, 1):
        _set_path(mode)
        out = _hom(percent=0.4, class_num=class_num, ignore_label=ignore)(lab, reg)
        want = cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), class_num, ignore, 0.4)
        assert np.array_equal(out.cpu().numpy(), want)


def test_region_bound_hint_is_sync_free_and_equivalent():
    from regda_b200 import synth
    reg = synth.region_maps(4, 256, 256, 100, device="cuda", seed=1)
    lab = synth.lrh_labels(reg, 6, -1, seed=2)
    a = _hom(percent=0.5, class_num=6, ignore_label=-1)(lab, reg)
    for bound in (int(reg.max()) + 1, 4096, 60000, 200000):
        bth = _hom(percent=0.5, class_num=6, ignore_label=-1, region_bound=bound, strict=False)
        assert torch.equal(bth(lab, reg), a)
        bth.check()


def test_relabelling_region_ids_is_unobservable():
    from regda_b200 import synth
    reg = synth.region_maps(2, 128, 128, 40, device="cuda", seed=3)
    lab = synth.lrh_labels(reg, 6, -1, seed=4)
    h = _hom(percent=0.5, class_num=6, ignore_label=-1)
    a = h(lab, reg)
    remap = torch.where(reg == 0, reg, reg * 37 + 11)
    assert torch.equal(h(lab, remap), a)


def test_domain_errors_raise_like_the_reference():
    lab = torch.zeros(1, 8, 8, dtype=torch.int64, device="cuda")
    reg = torch.ones(1, 8, 8, dtype=torch.int64, device="cuda")
    h = _hom(percent=0.5, class_num=6, ignore_label=-1)
    bad = lab.clone(); bad[0, 0, 0] = 9
    with pytest.raises(RuntimeError):
        h(bad, reg)
    bad = lab.clone(); bad[0, 0, 0] = -2
    with pytest.raises(RuntimeError):
        h(bad, reg)
    badr = reg.clone(); badr[0, 3, 3] = -1
    with pytest.raises(RuntimeError):
        h(lab, badr)
    ok = lab.clone(); ok[0, 0, 0] = 6            # label == class_num is legal (counts nowhere)
    out = h(ok, reg)
    assert int(out[0, 0, 0]) == 0 and int(out.sum()) == 0
    with pytest.raises(AssertionError):
        h(lab[0], reg[0])
    hb = _hom(percent=0.5, class_num=6, ignore_label=-1, region_bound=1)
    with pytest.raises(RuntimeError):
        hb(lab, reg)                              # id 1 is outside bound 1


def test_empty_and_ragged():
    h = _hom(percent=0.5, class_num=6, ignore_label=-1)
    e = torch.zeros(0, 4, 4, dtype=torch.int64, device="cuda")
    assert h(e, e).shape == (0, 4, 4)
    e = torch.zeros(2, 0, 4, dtype=torch.int64, device="cuda")
    assert h(e, e).shape == (2, 0, 4)
    lab = torch.randint(-1, 6, (3, 1, 1), device="cuda")
    reg = torch.randint(0, 3, (3, 1, 1), device="cuda")
    want = cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), 6, -1, 0.5)
    assert np.array_equal(h(lab, reg).cpu().numpy(), want)
    # non-contiguous views are accepted like any tensor the reference would reshape
    big = torch.randint(-1, 6, (2, 40, 80), device="cuda")
    regs = torch.randint(0, 9, (2, 40, 80), device="cuda")
    lv, rv = big[:, ::2, ::2], regs[:, ::2, ::2]
    want = cbind.lrh(lv.cpu().numpy(), rv.cpu().numpy(), 6, -1, 0.5)
    assert np.array_equal(h(lv, rv).cpu().numpy(), want)


def test_full_microbench_size_properties():
    """128x512x512 (BASELINE config 5): two independent kernels agree, per-image oracle on a
    sample, region-0 pixels untouched, every homogenised region is constant."""
    from regda_b200 import synth
    b = 128
    reg = synth.region_maps(b, 512, 512, 500, device="cuda", seed=21)
    lab = synth.lrh_labels(reg, 6, -1, seed=22)
    bound = int(reg.max()) + 1
    h = _hom(percent=0.5, class_num=6, ignore_label=-1, region_bound=bound)
    _set_path(0)
    fast = h(lab, reg)
    assert _last_path()[0] == 2
    _set_path(1)
    slow = h(lab, reg)
    assert torch.equal(fast, slow)
    assert torch.equal(fast[reg == 0], lab[reg == 0])
    for i in (0, 63, 127):
        want = cbind.lrh(lab[i:i + 1].cpu().numpy(), reg[i:i + 1].cpu().numpy(), 6, -1, 0.5)
        assert np.array_equal(fast[i:i + 1].cpu().numpy(), want)
    changed = fast != lab
    assert bool(changed.any())
    # a pixel that changed took its region's winner: all changed pixels of one (image, region) agree
    key = (torch.arange(b, device="cuda").view(b, 1, 1) * bound + reg)[changed]
    val = fast[changed]
    mx = torch.full((b * bound,), -1, dtype=torch.int64, device="cuda").scatter_reduce(0, key, val, "amax")
    mn = torch.full((b * bound,), 99, dtype=torch.int64, device="cuda").scatter_reduce(0, key, val, "amin")
    seen = mx >= 0
    assert torch.equal(mx[seen], mn[seen])


@pytest.mark.parametrize("n_regions", [1000, 2000, 5000, 9000])
def test_large_region_tables(n_regions):
    """1000 .. 9000 regions per 512x512 tile (region ids up to 2x that; BASELINE.json configs[4] spans 50 .. 5000): bit-exact on
    whichever path the planner picks (the single-pass cluster kernel while its per-CTA bin table leaves room for two CTAs per
    SM, the global-bin path beyond), plus an adversarial map with thousands of distinct ids in every image row block."""
    from regda_b200 import synth
    reg = synth.region_maps(4, 512, 512, n_regions, device="cuda", seed=3 + n_regions)
    lab = synth.lrh_labels(reg, 6, -1, seed=5)
    for pct in (0.5, 0.75):
        out = _hom(percent=pct, class_num=6, ignore_label=-1)(lab, reg)
        want = cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), 6, -1, pct)
        assert np.array_equal(out.cpu().numpy(), want)
    reg2 = (torch.arange(512 * 512, device="cuda").view(1, 512, 512) % 16000 + 1).long()
    lab2 = synth.lrh_labels(reg2, 6, -1, seed=9)
    out2 = _hom(percent=0.3, class_num=6, ignore_label=-1)(lab2, reg2)
    assert np.array_equal(out2.cpu().numpy(), cbind.lrh(lab2.cpu().numpy(), reg2.cpu().numpy(), 6, -1, 0.3))
