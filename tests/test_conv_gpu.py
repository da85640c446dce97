"""GPU: the tcgen05 implicit-GEMM convolution (regda_conv_fprop_bf16 through regda_b200.ops.conv)
against a plain PyTorch float32 convolution of the same bf16-rounded operands.
Tolerance: bf16 output rounding (2^-9 relative) + fp32 accumulation-order noise -> 1e-2 of the
output scale, and a much tighter bound on the mean error."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (n, cin, h, w, cout, k, pad, dil): every conv family of ResNet-101 OS16 + the PPM fuse conv
SHAPES = [
    (2, 64, 32, 32, 64, 1, 0, 1),
    (2, 256, 32, 32, 128, 1, 0, 1),
    (2, 64, 16, 16, 256, 1, 0, 1),
    (1, 128, 64, 64, 128, 3, 1, 1),
    (2, 256, 32, 32, 256, 3, 1, 1),
    (2, 512, 32, 32, 512, 3, 2, 2),        # layer4 dilated
    (2, 64, 128, 128, 64, 3, 1, 1),        # layer1: one image row per tile
    (1, 64, 256, 256, 64, 1, 0, 1),        # two tiles per image row
    (2, 128, 24, 40, 192, 3, 1, 1),        # ragged: patches overhang the image, cout = 3 x 64
    (1, 1024, 32, 32, 256, 1, 0, 1),
    (1, 4096, 32, 32, 512, 3, 1, 1),       # PPM fuse conv, K = 36864
    (3, 64, 12, 12, 64, 3, 1, 1),          # small map, patch 8 x 16 overhangs
]


def _ref(x, w, pad, dil):
    return F.conv2d(x.float(), w.float(), None, 1, pad, dil)


@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
def test_fprop_matches_float32_reference(shape):
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    assert tc.supports_fprop(x.shape, wt.shape, 1, pad, dil, x.dtype)
    y = tc.fprop(x, wt, 1, pad, dil)
    ref = _ref(x, wt, pad, dil)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    err = (y.float() - ref).abs()
    scale = ref.abs().max()
    assert float(err.max()) <= 1e-2 * float(scale), (float(err.max()), float(scale))
    assert float(err.mean()) <= 2e-3 * float(ref.abs().mean())


@pytest.mark.parametrize("shape", [SHAPES[1], SHAPES[4], SHAPES[5], SHAPES[8]], ids=str)
def test_conv2d_module_forward_backward(shape):
    """the nn.Module front end: forward, input gradient (forward weights read in place as an MN-major operand)
    and weight gradient (accumulated into weight.grad) all through the tcgen05 kernels."""
    from regda_b200.ops import conv as C
    n, cin, h, w, cout, k, pad, dil = shape
    torch.manual_seed(0)
    m = C.Conv2d(cin, cout, k, padding=pad, dilation=dil, bias=False).cuda()
    x = torch.randn(n, cin, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    before = dict(C.stats)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    assert C.stats["tcgen05_fprop"] == before["tcgen05_fprop"] + 1
    assert C.stats["tcgen05_dgrad"] == before["tcgen05_dgrad"] + 1
    assert C.stats["tcgen05_wgrad"] == before["tcgen05_wgrad"] + 1
    xr = x.detach().float().requires_grad_(True)
    wr = m.weight.detach().bfloat16().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, pad, dil)
    yr.backward(gy.float())
    for got, want in ((y, yr), (x.grad, xr.grad), (m.weight.grad, wr.grad)):
        err = (got.float() - want).abs()
        assert float(err.max()) <= 1.5e-2 * float(want.abs().max())


def test_unsupported_shapes_are_refused():
    from regda_b200 import capi
    L = capi.lib()
    assert L.regda_conv_fprop_supported(2, 32, 32, 64, 64, 1, 1, 1, 0, 1) == 1
    assert L.regda_conv_fprop_supported(2, 32, 32, 3, 64, 7, 7, 2, 3, 1) == 0      # stem
    assert L.regda_conv_fprop_supported(2, 32, 32, 64, 6, 1, 1, 1, 0, 1) == 0      # classifier
    assert L.regda_conv_fprop_supported(2, 6, 6, 2048, 512, 1, 1, 1, 0, 1) == 0    # PPM branch on a pooled map
    assert L.regda_conv_fprop_supported(2, 64, 64, 128, 128, 3, 3, 2, 1, 1) == 1   # stride 2: TMA element strides
    assert L.regda_conv_fprop_supported(2, 64, 64, 128, 128, 3, 3, 3, 1, 1) == 0   # stride 3
    assert L.regda_conv_dgrad_supported(2, 64, 64, 128, 128, 3, 3, 2, 1, 1) == 0   # strided dgrad stays on the library


# (n, cin, h, w, cout, k, stride, pad, dil)
STRIDED = [
    (2, 128, 64, 64, 128, 3, 2, 1, 1),     # layer2.0.conv2
    (2, 256, 64, 64, 512, 1, 2, 0, 1),     # layer2.0.downsample
    (1, 256, 32, 48, 256, 3, 2, 1, 1),     # ragged
    (2, 64, 30, 30, 64, 3, 2, 1, 1),       # odd output size 15x15
]


@pytest.mark.parametrize("shape", STRIDED, ids=str)
def test_strided_fprop_and_wgrad(shape):
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, stride, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    assert tc.supports_fprop(x.shape, wt.shape, stride, pad, dil, x.dtype)
    y = tc.fprop(x, wt, stride, pad, dil)
    xr, wr = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride, pad, dil)
    assert y.shape == ref.shape
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    gy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    ref.backward(gy.float())
    assert tc.supports_wgrad(x.shape, wt.shape, stride, pad, dil, x.dtype)
    gw = torch.zeros(cout, cin, k, k, device="cuda").contiguous(memory_format=torch.channels_last)
    tc.wgrad_accumulate(gy, x, gw, stride, pad, dil)
    tc.wgrad_accumulate(gy, x, gw, stride, pad, dil)          # accumulates: second call doubles it
    assert float((gw / 2 - wr.grad).abs().max()) <= 5e-3 * float(wr.grad.abs().max())


WGRAD = [SHAPES[0], SHAPES[1], SHAPES[3], SHAPES[5], SHAPES[6], SHAPES[8], SHAPES[9], SHAPES[11], (2, 4096, 32, 32, 512, 3, 1, 1)]


@pytest.mark.parametrize("epilogue", ["tma", "direct"])
@pytest.mark.parametrize("shape", WGRAD, ids=str)
def test_wgrad_and_dgrad_match_float32_reference(shape, epilogue, monkeypatch):
    from regda_b200.ops import tc
    monkeypatch.setenv("REGDA_CONV_EPILOGUE", epilogue)
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    xr, wr = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, 1, pad, dil)
    gy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    ref.backward(gy.float())
    assert tc.supports_wgrad(x.shape, wt.shape, 1, pad, dil, x.dtype)
    gw = torch.zeros(cout, cin, k, k, device="cuda").contiguous(memory_format=torch.channels_last)
    tc.wgrad_accumulate(gy, x, gw, 1, pad, dil)
    assert float((gw - wr.grad).abs().max()) <= 5e-3 * float(wr.grad.abs().max())
    assert tc.supports_dgrad(x.shape, wt.shape, 1, pad, dil, x.dtype)
    gx = tc.dgrad(gy, wt, x.shape, 1, pad, dil)
    assert float((gx.float() - xr.grad).abs().max()) <= 1e-2 * float(xr.grad.abs().max())


def test_conv_tap_adds_the_residual_gradient_in_the_dgrad_epilogue():
    """y, stats, x_tap = conv(x, tap=True): the gradient arriving through x_tap (the residual branch) is added inside the
    dgrad kernel; statistics from the epilogue equal the sums of the bf16 outputs."""
    from regda_b200.ops import conv as C
    torch.manual_seed(2)
    m = C.Conv2d(256, 128, 3, padding=1, bias=False).cuda()
    x = torch.randn(4, 256, 32, 32, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y, st, xt = m.forward_with_bn_stats(x, 2, tap=True)
    assert st is not None and st.shape == (2, 2, 128)
    yf = y.float()
    want = torch.stack([torch.stack([yf[:2].sum((0, 2, 3)), yf[:2].square().sum((0, 2, 3))]), torch.stack([yf[2:].sum((0, 2, 3)), yf[2:].square().sum((0, 2, 3))])])
    assert float((st - want).abs().max()) <= 2e-3 * float(want.abs().max())
    gy = torch.randn_like(y)
    gt = torch.randn_like(x)
    torch.autograd.backward([y, xt], [gy, gt])
    xr = x.detach().float().requires_grad_(True)
    wr = m.weight.detach().bfloat16().float()
    yr = F.conv2d(xr, wr, None, 1, 1, 1)
    torch.autograd.backward([yr, xr * 1.0], [gy.float(), gt.float()])
    assert float((x.grad.float() - xr.grad).abs().max()) <= 1.5e-2 * float(xr.grad.abs().max())


@pytest.mark.parametrize("epilogue", ["tma", "direct"])
@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[2], SHAPES[4], SHAPES[8], SHAPES[11], (4, 256, 32, 32, 1024, 1, 0, 1),
                                   (2, 128, 5, 40, 128, 3, 1, 1)], ids=str)
def test_fused_bn_statistics_and_both_epilogues(shape, epilogue, monkeypatch):
    """The conv epilogue's per-group per-channel (sum, sum of squares) equal those of the bf16 output it wrote, for the
    TMA-store epilogue and the per-lane-store epilogue, and both epilogues write the same output bits."""
    from regda_b200.ops import tc
    monkeypatch.setenv("REGDA_CONV_EPILOGUE", epilogue)
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    groups = 2 if n % 2 == 0 else 1
    y, st = tc.fprop(x, wt, 1, pad, dil, groups)
    y_plain = tc.fprop(x, wt, 1, pad, dil)
    assert torch.equal(y, y_plain)
    monkeypatch.setenv("REGDA_CONV_EPILOGUE", "direct" if epilogue == "tma" else "tma")
    assert torch.equal(y, tc.fprop(x, wt, 1, pad, dil))
    ref = _ref(x, wt, pad, dil)
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    yf = y.float().reshape(groups, n // groups, cout, -1)
    want_sum = yf.sum(dim=(1, 3))
    want_sq = (yf * yf).sum(dim=(1, 3))
    assert st.shape == (groups, 2, cout)
    cnt = (n // groups) * y.shape[2] * y.shape[3]
    assert torch.allclose(st[:, 0], want_sum, rtol=1e-4, atol=1e-4 * cnt ** 0.5 * float(yf.abs().max()))
    assert torch.allclose(st[:, 1], want_sq, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("with_addend", [False, True])
@pytest.mark.parametrize("shape", [SHAPES[1], SHAPES[2], SHAPES[4], SHAPES[5], SHAPES[8], SHAPES[11], (4, 1024, 32, 32, 256, 1, 0, 1)], ids=str)
def test_dgrad_bnred_matches_masked_dgrad_and_reductions(shape, with_addend):
    """regda_conv_dgrad_bnred_bf16: dz == bf16((dgrad + addend) * mask) bit for bit against the plain dgrad kernel's unrounded
    sum (checked at bf16 resolution), and red == (sum dz, sum dz * bn_y) per statistics group and channel in float32."""
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    cl = torch.channels_last
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=cl)
    oh, ow = tc.out_hw(h, w, k, k, 1, pad, dil)
    gy = torch.randn(n, cout, oh, ow, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    bn_y = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    keep = torch.rand(n, cin, h, w, device="cuda", generator=g) > 0.4
    add = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl) if with_addend else None
    # mask bytes: NHWC element e -> bit e % 8 of byte e / 8
    bits = keep.permute(0, 2, 3, 1).reshape(-1, 8).to(torch.uint8)
    mask = (bits << torch.arange(8, device="cuda", dtype=torch.uint8)).sum(dim=1).to(torch.uint8).contiguous()
    groups = 2 if n % 2 == 0 else 1
    red = torch.zeros(groups, 2, cin, device="cuda")
    dz = tc.dgrad_bnred(gy, wt, (n, cin, h, w), 1, pad, dil, add, bn_y, mask, red, groups)
    plain = tc.dgrad(gy, wt, (n, cin, h, w), 1, pad, dil, addend=add)        # same kernel family, unmasked
    want = torch.where(keep, plain, torch.zeros_like(plain))
    assert torch.equal(dz, want)
    dzf = dz.float().reshape(groups, n // groups, cin, -1)
    yf = bn_y.float().reshape(groups, n // groups, cin, -1)
    s1, s2 = dzf.sum(dim=(1, 3)), (dzf * yf).sum(dim=(1, 3))
    tol = 1e-4 * float(dzf.abs().max()) * (dzf.shape[1] * dzf.shape[3]) ** 0.5 * 4
    assert torch.allclose(red[:, 0], s1, rtol=1e-4, atol=tol), float((red[:, 0] - s1).abs().max())
    assert torch.allclose(red[:, 1], s2, rtol=1e-4, atol=4 * tol), float((red[:, 1] - s2).abs().max())


@pytest.mark.parametrize("shape", [(2, 64, 64), (4, 96, 160), (2, 37, 53)], ids=str)
def test_stem_conv_patch_matrix_path_matches_float32_reference(shape):
    """Conv2d(3, 64, 7, stride 2, padding 3) through regda_stem_im2col_bf16 + the 1x1 tcgen05 kernels: output, fused BatchNorm
    statistics and weight gradient against a float32 convolution of the same bf16 operands."""
    from regda_b200.ops import stem
    from regda_b200.ops.conv import Conv2d
    n, h, w = shape
    torch.manual_seed(3)
    conv = Conv2d(3, 64, 7, stride=2, padding=3, bias=False).cuda()
    x = torch.randn(n, 3, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    assert stem.supported(conv, x)
    groups = 2 if n % 2 == 0 else 1
    y, st = stem.stem_conv(x, conv.weight, groups)
    wr = conv.weight.detach().bfloat16().float().requires_grad_(True)
    ref = F.conv2d(x.float(), wr, None, 2, 3, 1)
    assert y.shape == ref.shape
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    yf = y.float().reshape(groups, n // groups, 64, -1)
    assert torch.allclose(st[:, 0], yf.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[:, 1], (yf * yf).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    gy = torch.randn(ref.shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    ref.backward(gy.float())
    assert float((conv.weight.grad - wr.grad).abs().max()) <= 5e-3 * float(wr.grad.abs().max())
