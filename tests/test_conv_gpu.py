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
    # maps smaller than 128 pixels: several images per M tile (pyramid-pooling branches, deep layers of small tiles)
    (16, 2048, 1, 1, 512, 1, 0, 1),        # PPM scale 1: 16 rows of a 128-row tile
    (16, 2048, 2, 2, 512, 1, 0, 1),
    (16, 2048, 3, 3, 512, 1, 0, 1),        # 3x3: 2x2 patches overhang
    (16, 2048, 6, 6, 512, 1, 0, 1),        # 6x6: four 4x4 patches per 8 images
    (2, 256, 6, 6, 256, 3, 1, 1),          # 96x96 tile at 1/16: 3x3 conv, halo zero-fill per image, image block overhang
    (5, 128, 4, 4, 64, 3, 2, 2),           # dilated, 5 images of a block of 8
    (3, 512, 8, 8, 256, 3, 1, 1),          # 64 pixels: 2 images per tile, odd image count
]
SMALL = SHAPES[12:]


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
    assert L.regda_conv_fprop_supported(2, 6, 6, 2048, 512, 1, 1, 1, 0, 1) == 1    # PPM branch on a pooled map: multi-image tiles
    assert L.regda_conv_fprop_supported(2, 64, 64, 128, 128, 3, 3, 2, 1, 1) == 1   # stride 2: TMA element strides
    assert L.regda_conv_fprop_supported(2, 64, 64, 128, 128, 3, 3, 3, 1, 1) == 0   # stride 3
    assert L.regda_conv_dgrad_supported(2, 64, 64, 128, 128, 3, 3, 2, 1, 1) == 0   # C ABI: stride-1 dgrad only (stride 2 = zero insertion)
    # fused statistics need every tile inside one statistics group
    assert L.regda_conv_fprop_stats_supported(16, 1, 1, 2048, 512, 1, 1, 1, 0, 1, 2) == 0     # 128 images per tile, 8 per group
    assert L.regda_conv_fprop_stats_supported(16, 6, 6, 2048, 512, 1, 1, 1, 0, 1, 2) == 1     # 8 images per tile
    assert L.regda_conv_fprop_stats_supported(16, 32, 32, 256, 256, 3, 3, 1, 1, 1, 2) == 1


def test_engine_has_no_fallback():
    """a shape outside the kernels' coverage raises instead of reaching a library kernel"""
    from regda_b200.ops import conv as C
    assert C.ENGINE == "tcgen05"
    m = C.Conv2d(48, 64, 1, bias=False).cuda()
    before = dict(C.stats)
    with pytest.raises(RuntimeError, match="no tcgen05 kernel"):
        m(torch.randn(2, 48, 16, 16, device="cuda").bfloat16())
    with pytest.raises(RuntimeError, match="CUDA"):
        m.cpu()(torch.randn(2, 48, 16, 16))
    assert C.stats == before


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


WGRAD = [SHAPES[0], SHAPES[1], SHAPES[3], SHAPES[5], SHAPES[6], SHAPES[8], SHAPES[9], SHAPES[11], (2, 4096, 32, 32, 512, 3, 1, 1)] + SMALL


@pytest.mark.parametrize("shape", WGRAD, ids=str)
def test_wgrad_and_dgrad_match_float32_reference(shape):
    from regda_b200.ops import tc
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


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[2], SHAPES[4], SHAPES[8], SHAPES[11], (4, 256, 32, 32, 1024, 1, 0, 1),
                                   (2, 128, 5, 40, 128, 3, 1, 1), (16, 2048, 6, 6, 512, 1, 0, 1), (4, 512, 8, 8, 256, 3, 1, 1),
                                   (2, 256, 6, 6, 256, 3, 1, 1),
                                   # 8 channel blocks of 256 do not divide the 148-CTA grid: contiguous tile ranges, several
                                   # statistics flushes per CTA (layer4 conv3 / downsample, ResNet width 2048)
                                   (16, 512, 32, 32, 2048, 1, 0, 1), (6, 256, 24, 40, 2048, 1, 0, 1)], ids=str)
def test_fused_bn_statistics_and_float32_epilogue(shape):
    """The conv epilogue's per-group per-channel (sum, sum of squares) equal those of the bf16 output it wrote; the
    float32-output epilogue returns the same accumulators un-rounded (their bf16 rounding is the bf16 output, bit for bit)."""
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    groups = 2 if n % 2 == 0 else 1
    if not tc.supports_fprop_stats(x.shape, wt.shape, 1, pad, dil, x.dtype, groups):
        groups = 1
    assert tc.supports_fprop_stats(x.shape, wt.shape, 1, pad, dil, x.dtype, groups)
    y, st = tc.fprop(x, wt, 1, pad, dil, groups)
    y_plain = tc.fprop(x, wt, 1, pad, dil)
    assert torch.equal(y, y_plain)
    y32 = tc.fprop(x, wt, 1, pad, dil, out_f32=True)
    assert y32.dtype == torch.float32 and torch.equal(y32.bfloat16(), y)
    ref = _ref(x, wt, pad, dil)
    assert float((y32 - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) * max(1.0, (cin * k * k / 4096) ** 0.5)
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    yf = y.float().reshape(groups, n // groups, cout, -1)
    want_sum = yf.sum(dim=(1, 3))
    want_sq = (yf * yf).sum(dim=(1, 3))
    assert st.shape == (groups, 2, cout)
    cnt = (n // groups) * y.shape[2] * y.shape[3]
    assert torch.allclose(st[:, 0], want_sum, rtol=1e-4, atol=1e-4 * cnt ** 0.5 * float(yf.abs().max()))
    assert torch.allclose(st[:, 1], want_sq, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("shape", [(16, 256, 32, 32, 1024, 1, 0, 1), (4, 128, 24, 24, 128, 3, 1, 1), (16, 2048, 6, 6, 512, 1, 0, 1)], ids=str)
def test_static_weight_hint_changes_nothing_but_the_order_of_the_first_loads(shape):
    """regda_conv_hint_static_weights: the next convolution requests the weight halves of its first pipeline stages before its
    programmatic-dependent-launch wait -- same bits out, forward and data gradient, and the hint is consumed by one launch"""
    from regda_b200 import capi
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    cl = torch.channels_last
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=cl)
    y0, st0 = tc.fprop(x, wt, 1, pad, dil, 2)
    for _ in range(3):
        capi.check(capi.lib().regda_conv_hint_static_weights())
        y1, st1 = tc.fprop(x, wt, 1, pad, dil, 2)
        assert torch.equal(y0, y1) and torch.allclose(st0, st1, rtol=1e-5, atol=1e-3)
    gy = torch.randn_like(y0)
    d0 = tc.dgrad(gy, wt, x.shape, 1, pad, dil)
    capi.check(capi.lib().regda_conv_hint_static_weights())
    d1 = tc.dgrad(gy, wt, x.shape, 1, pad, dil)
    assert torch.equal(d0, d1)
    assert torch.equal(tc.dgrad(gy, wt, x.shape, 1, pad, dil), d0)          # (hint consumed: a plain launch again)


@pytest.mark.parametrize("with_addend", [False, True])
@pytest.mark.parametrize("shape", [SHAPES[1], SHAPES[2], SHAPES[4], SHAPES[5], SHAPES[8], SHAPES[11], (4, 1024, 32, 32, 256, 1, 0, 1),
                                   (16, 256, 6, 6, 256, 3, 1, 1), (4, 512, 8, 8, 128, 1, 0, 1),
                                   (16, 2048, 32, 32, 512, 1, 0, 1)], ids=str)        # 8 channel blocks: contiguous tile ranges
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
    assert tc.supports_dgrad_bnred((n, cin, h, w), wt.shape, 1, pad, dil, torch.bfloat16, groups)
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
    # the loader's float32 NCHW image goes in directly (rounded to bf16 inside the patch kernel): same bits out
    xf = torch.randn(n, 3, h, w, device="cuda")
    y16, st16 = stem.stem_conv(xf.bfloat16().contiguous(memory_format=torch.channels_last), conv.weight, groups)
    y32, st32 = stem.stem_conv(xf, conv.weight, groups, torch.bfloat16)
    assert torch.equal(y16, y32)


@pytest.mark.parametrize("with_bnred", [False, True])
@pytest.mark.parametrize("shape", STRIDED, ids=str)
def test_strided_dgrad_through_zero_insertion(shape, with_bnred):
    """data gradient of the stride-2 convolutions (layer2.0 / layer3.0 conv2 and downsample): the stride-1 kernel on the
    zero-inserted dY, plain and fused with the BatchNorm reductions, against a float32 convolution's input gradient"""
    from regda_b200.ops import tc
    n, cin, h, w, cout, k, stride, pad, dil = shape
    g = torch.Generator(device="cuda").manual_seed(9)
    cl = torch.channels_last
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=cl)
    xr = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    ref = F.conv2d(xr, wt.float(), None, stride, pad, dil)
    gy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    ref.backward(gy.float())
    assert tc.supports_dgrad(xr.shape, wt.shape, stride, pad, dil, torch.bfloat16)
    if not with_bnred:
        gx = tc.dgrad(gy, wt, xr.shape, stride, pad, dil)
        assert float((gx.float() - xr.grad).abs().max()) <= 1e-2 * float(xr.grad.abs().max())
        return
    assert tc.supports_dgrad_bnred(xr.shape, wt.shape, stride, pad, dil, torch.bfloat16, 1)
    bn_y = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    keep = torch.rand(n, cin, h, w, device="cuda", generator=g) > 0.3
    bits = keep.permute(0, 2, 3, 1).reshape(-1, 8).to(torch.uint8)
    mask = (bits << torch.arange(8, device="cuda", dtype=torch.uint8)).sum(dim=1).to(torch.uint8).contiguous()
    red = torch.zeros(1, 2, cin, device="cuda")
    dz = tc.dgrad_bnred(gy, wt, xr.shape, stride, pad, dil, None, bn_y, mask, red, 1)
    want = torch.where(keep, xr.grad, torch.zeros_like(xr.grad))
    assert float((dz.float() - want).abs().max()) <= 1e-2 * float(want.abs().max())
    dzf = dz.float()
    assert torch.allclose(red[0, 0], dzf.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(red[0, 1], (dzf * bn_y.float()).sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)


F32 = [(2, 64, 16, 16, 64, 3, 1, 1, 1), (2, 256, 6, 6, 512, 1, 1, 0, 1), (2, 128, 32, 32, 128, 3, 2, 1, 1), (2, 256, 16, 16, 512, 1, 2, 0, 1),
       (2, 512, 6, 6, 512, 3, 1, 2, 2), (2, 2048, 1, 1, 512, 1, 1, 0, 1), (1, 4096, 8, 8, 512, 3, 1, 1, 1)]


@pytest.mark.parametrize("shape", F32, ids=str)
def test_float32_parity_convolution_on_the_tcgen05_kernels(shape):
    """compute_dtype=float32: Conv2d runs on the bf16 tensor-core kernels through the three-way operand split of ops/tc.py (hi*hi
    in one accumulator, the five cross terms in another, K cut into short chunks because the tensor cores truncate when they
    align an update with a long accumulation) -- float32-class accuracy (3e-5 of the output scale; a plain bf16 product is at
    1e-2) for the output, the input gradient and the weight gradient, with no library convolution."""
    from regda_b200.ops import conv as C
    n, cin, h, w, cout, k, stride, pad, dil = shape
    torch.manual_seed(4)
    torch.backends.cudnn.allow_tf32 = False
    m = C.Conv2d(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=False).cuda()
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    before = dict(C.stats)
    y = m(x)
    assert y.dtype == torch.float32
    gy = torch.randn_like(y)
    y.backward(gy)
    assert C.stats["cudnn"] == before["cudnn"] and C.stats["tcgen05_fprop"] == before["tcgen05_fprop"] + 1
    xr = x.detach().double().requires_grad_(True)
    wr = m.weight.detach().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride, pad, dil)
    yr.backward(gy.double())
    for name, got, want in (("y", y, yr), ("dx", x.grad, xr.grad), ("dw", m.weight.grad, wr.grad)):
        err = float((got.double() - want).abs().max()) / float(want.abs().max())
        assert err <= 3e-5, (name, err)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_dropout_classifier_forward_backward(dtype, p):
    """Dropout2d -> Conv2d(512, C, 1) + bias as one kernel each way (regda/models/Encoder.py:39-40) against torch on the mask
    the kernel drew (recovered from the output of an all-ones probe), float32 logits in NCHW like the reference's"""
    from regda_b200.ops import head
    from regda_b200.ops.conv import Conv2d
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False
    b, cin, h, w, ncls = 4, 512, 16, 24, 6
    conv = Conv2d(cin, ncls, 1, bias=True).cuda()
    y = torch.randn(b, cin, h, w, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert head.supported(y, conv.weight)
    head.seed(1234)
    out = head.dropout_classifier(y, conv.weight, conv.bias, p, True)
    assert out.shape == (b, ncls, h, w) and out.dtype == torch.float32 and out.is_contiguous()
    gout = torch.randn_like(out)
    out.backward(gout)
    # the same draw again -> the same mask (counter-based generator, state re-seeded)
    head.seed(1234)
    keep = torch.empty(b, cin, device="cuda")
    from regda_b200 import capi
    if p > 0:
        capi.call("regda_dropout2d_mask", capi.ptr(head._state(y.device)), float(p), capi.ptr(keep), b * cin, capi.stream())
        frac = float((keep == 0).float().mean())
        assert abs(frac - p) < 0.05 and torch.all((keep == 0) | ((keep - 1 / (1 - p)).abs() < 1e-6))
        head.seed(1235)
        keep2 = torch.empty_like(keep)
        capi.call("regda_dropout2d_mask", capi.ptr(head._state(y.device)), float(p), capi.ptr(keep2), b * cin, capi.stream())
        assert not torch.equal(keep, keep2)
    else:
        keep.fill_(1.0)
    yr = y.detach().float().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    ref = F.conv2d(yr * keep.view(b, cin, 1, 1), wr, br)
    ref.backward(gout)
    tol = 2e-5 if dtype == torch.float32 else 2e-3
    assert float((out - ref).abs().max()) <= tol * float(ref.abs().max())
    assert float((y.grad.float() - yr.grad).abs().max()) <= (1e-5 if dtype == torch.float32 else 1e-2) * float(yr.grad.abs().max())
    assert float((conv.weight.grad - wr.grad).abs().max()) <= 1e-4 * float(wr.grad.abs().max())
    assert float((conv.bias.grad - br.grad).abs().max()) <= 1e-4 * float(br.grad.abs().max())
    # eval mode: no dropout
    out_eval = head.dropout_classifier(y.detach(), conv.weight, conv.bias, p, False)
    ref_eval = F.conv2d(y.detach().float(), conv.weight.detach(), conv.bias.detach())
    assert float((out_eval - ref_eval).abs().max()) <= tol * float(ref_eval.abs().max())
