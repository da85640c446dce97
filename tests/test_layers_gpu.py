"""GPU: the fused BatchNorm(+residual+ReLU) and pyramid-pooling kernels (csrc/norm.cu, csrc/ppm.cu) against plain
PyTorch float32 references of the same ops on the same bf16-rounded inputs, forward and backward; and the bf16
fused model against the float32 eager model.
Tolerances: outputs are bf16 (2^-9 relative rounding) of fp32 arithmetic -> 1e-2 of the output scale on the max
error; fp32 outputs (statistics, parameter gradients, pooled maps) 2e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _close_most(got, want, tol, frac=2e-3):
    """like _close, but a fraction `frac` of the elements may be off (a ReLU mask bit that flips where the bf16
    pre-activation rounds to +-0 changes that element's gradient by the full incoming value)"""
    err = (got.float() - want.float()).abs()
    scale = float(want.float().abs().max())
    bad = float((err > tol * scale + 1e-6).float().mean())
    assert bad <= frac, (bad, float(err.max()), scale)


def _close(got, want, tol):
    err = float((got.float() - want.float()).abs().max())
    scale = float(want.float().abs().max())
    assert err <= tol * scale + 1e-6, (err, scale)


@pytest.mark.parametrize("shape", [(2, 64, 33, 20), (8, 256, 16, 16), (2, 2048, 8, 8), (3, 512, 7, 9), (4, 128, 64, 64)], ids=str)
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False)])
def test_bn_act_forward_backward(shape, relu, res):
    from regda_b200.ops import norm
    n, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(c + h)
    y = _cl((torch.randn(shape, device="cuda", generator=g) * 1.7 + 0.4).bfloat16()).requires_grad_(True)
    r = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16()).requires_grad_(True) if res else None
    bn = torch.nn.BatchNorm2d(c).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
    ref_bn = torch.nn.BatchNorm2d(c).cuda().train()
    ref_bn.load_state_dict(bn.state_dict())
    assert norm.supported(y, bn)
    out = norm.bn_act(y, bn, residual=r, relu=relu)
    gout = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16())
    out.backward(gout)

    yr = y.detach().float().requires_grad_(True)
    rr = r.detach().float().requires_grad_(True) if res else None
    o = ref_bn(yr)
    if res:
        o = o + rr
    if relu:
        o = F.relu(o)
    o.backward(gout.float())
    _close(out, o, 1e-2)
    _close(bn.running_mean, ref_bn.running_mean, 2e-3)
    _close(bn.running_var, ref_bn.running_var, 2e-3)
    assert int(bn.num_batches_tracked) == 1
    # the reference backward sees the un-rounded fp32 output for the ReLU mask; elements within bf16 rounding of 0 may differ
    _close(y.grad, yr.grad, 2e-2)
    if res:
        _close(r.grad, rr.grad, 2e-2)
    _close(bn.weight.grad, ref_bn.weight.grad, 1e-2)
    _close(bn.bias.grad, ref_bn.bias.grad, 1e-2)
    # gradients accumulate into existing .grad buffers
    before = bn.weight.grad.clone()
    norm.bn_act(y.detach().requires_grad_(True), bn, residual=None, relu=relu).backward(gout)
    assert not torch.equal(before, bn.weight.grad)


@pytest.mark.parametrize("shape", [(2, 2048, 32, 32), (2, 256, 8, 8), (1, 64, 20, 13), (2, 128, 64, 64)], ids=str)
def test_ppm_pool_forward_backward(shape):
    from regda_b200.ops import ppm
    scales = (1, 2, 3, 6)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16()).requires_grad_(True)
    pooled = ppm.pool(x, scales)
    xr = x.detach().float().requires_grad_(True)
    refs = [F.adaptive_avg_pool2d(xr, s).flatten(2).transpose(1, 2) for s in scales]     # [b, s*s, c]
    ref = torch.cat(refs, 1)
    assert pooled.shape == ref.shape and pooled.dtype == torch.float32
    _close(pooled, ref, 2e-3)
    gp = torch.randn(ref.shape, device="cuda", generator=g)
    pooled.backward(gp)
    ref.backward(gp)
    _close(x.grad, xr.grad, 1e-2)


@pytest.mark.parametrize("shape", [(2, 2048, 32, 32), (1, 64, 20, 13), (2, 640, 16, 16)], ids=str)
def test_ppm_pool_tap_adds_the_other_readers_gradient(shape):
    """the feature map's second reader goes through the tap: its gradient is added inside the pooling backward kernel"""
    from regda_b200.ops import ppm
    scales = (1, 2, 3, 6)
    g = torch.Generator(device="cuda").manual_seed(9)
    x = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16()).requires_grad_(True)
    pooled, x_tap = ppm.pool(x, scales, tap=True)
    assert x_tap.data_ptr() == x.data_ptr()
    xr = x.detach().float().requires_grad_(True)
    ref = torch.cat([F.adaptive_avg_pool2d(xr, s).flatten(2).transpose(1, 2) for s in scales], 1)
    gp = torch.randn(ref.shape, device="cuda", generator=g)
    gt = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16())
    torch.autograd.backward([pooled, x_tap], [gp, gt])
    torch.autograd.backward([ref, xr * 1.0], [gp, gt.float()])
    _close(x.grad, xr.grad, 1e-2)
    # only the tap carries a gradient
    x2 = x.detach().clone().requires_grad_(True)
    _, t2 = ppm.pool(x2, scales, tap=True)
    t2.backward(gt)
    assert torch.equal(x2.grad, gt)


def test_folded_fuse_conv_matches_concat_conv_and_adds_tap_gradient():
    """ops/ppm_fold.py against the reference formulation (Encoder.py:43-52: conv3x3 over cat(feat, up(branch_k))), float32 torch; with
    tap=True the next reader's gradient is added inside the data-gradient kernel"""
    from regda_b200.ops import ppm_fold
    scales = (1, 2, 3, 6)
    b, cf, cb, h, w, O = 4, 256, 64, 16, 16, 128
    g = torch.Generator(device="cuda").manual_seed(12)
    fin = _cl((torch.randn(b, cf, h, w, device="cuda", generator=g)).bfloat16()).requires_grad_(True)
    brs = [_cl(torch.randn(b, cb, s, s, device="cuda", generator=g).bfloat16()).requires_grad_(True) for s in scales]
    wt = (torch.randn(O, cf + 4 * cb, 3, 3, device="cuda", generator=g) / (9 * (cf + 4 * cb)) ** 0.5).contiguous(memory_format=torch.channels_last)
    wt = torch.nn.Parameter(wt)
    y, st, fin_tap = ppm_fold.fuse_conv(fin, brs, wt, scales, 2, tap=True)
    finr = fin.detach().float().requires_grad_(True)
    brr = [t.detach().float().requires_grad_(True) for t in brs]
    wr = wt.detach().bfloat16().float().requires_grad_(True)
    cat = torch.cat([finr] + [F.interpolate(t, (h, w), mode="bilinear", align_corners=False) for t in brr], 1)
    ref = F.conv2d(cat, wr, padding=1)
    _close(y, ref, 1.5e-2)
    yf = y.detach().float().view(2, b // 2, O, h * w)
    _close(st[:, 0], yf.sum((1, 3)), 1e-2)
    _close(st[:, 1], (yf * yf).sum((1, 3)), 1e-2)
    gy = _cl(torch.randn(ref.shape, device="cuda", generator=g).bfloat16())
    gt = _cl(torch.randn(fin.shape, device="cuda", generator=g).bfloat16())
    torch.autograd.backward([y, fin_tap], [gy, gt])
    torch.cuda.synchronize()
    from regda_b200.ops import conv as C
    C.join_wgrad_stream()
    torch.autograd.backward([ref, finr * 1.0], [gy.float(), gt.float()])
    _close(fin.grad, finr.grad, 1.5e-2)
    for t, tr in zip(brs, brr):
        _close(t.grad, tr.grad, 2e-2)
    _close(wt.grad, wr.grad, 1.5e-2)


def test_ppm_cells_split_and_gradient():
    from regda_b200.ops import ppm
    scales = (1, 2, 3, 6)
    b, c = 3, 256
    g = torch.Generator(device="cuda").manual_seed(10)
    pooled = torch.randn(b, 50, c, device="cuda", generator=g).requires_grad_(True)
    ps = ppm.cells(pooled, scales)
    off = 0
    gs = []
    for p, s in zip(ps, scales):
        want = pooled.detach()[:, off:off + s * s, :].reshape(b, s, s, c).permute(0, 3, 1, 2).bfloat16()
        assert p.shape == want.shape and p.is_contiguous(memory_format=torch.channels_last) and torch.equal(p, want)
        gs.append(torch.randn(want.shape, device="cuda", generator=g).bfloat16())
        off += s * s
    torch.autograd.backward(list(ps), gs)
    want = torch.cat([t.permute(0, 2, 3, 1).reshape(b, -1, c) for t in gs], 1).float()
    assert torch.equal(pooled.grad, want)


@pytest.mark.parametrize("shape", [(2, 2048, 32, 32, 512), (2, 64, 8, 8, 32), (1, 128, 20, 13, 64)], ids=str)
def test_ppm_upsample_concat_forward_backward(shape):
    from regda_b200.ops import ppm
    b, c, h, w, cb = shape
    scales = (1, 2, 3, 6)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = _cl(torch.randn(b, c, h, w, device="cuda", generator=g).bfloat16()).requires_grad_(True)
    brs = [_cl(torch.randn(b, cb, s, s, device="cuda", generator=g).bfloat16()).requires_grad_(True) for s in scales]
    cat = ppm.upsample_concat(x, brs, scales)
    xr = x.detach().float().requires_grad_(True)
    brr = [t.detach().float().requires_grad_(True) for t in brs]
    ref = torch.cat([xr] + [F.interpolate(t, (h, w), mode="bilinear", align_corners=False) for t in brr], 1)
    assert cat.shape == ref.shape and cat.is_contiguous(memory_format=torch.channels_last)
    _close(cat, ref, 1e-2)
    gc = _cl(torch.randn(ref.shape, device="cuda", generator=g).bfloat16())
    cat.backward(gc)
    ref.backward(gc.float())
    _close(x.grad, xr.grad, 1e-2)
    for t, tr in zip(brs, brr):
        _close(t.grad, tr.grad, 1e-2)


def _run_model(fused, engine, dtype, x, label):
    from oracle import step_oracle as so
    from regda_b200.gast.balance import CrossEntropy
    from regda_b200.models import Encoder as E
    from regda_b200.ops import conv as C
    from regda_b200.utils.tools import loss_calc
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    old = (E.FUSED, C.ENGINE)
    E.set_fused(fused)
    C.set_engine(engine)
    try:
        m = E.Deeplabv2(cfg, compute_dtype=dtype)
        m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        m = m.cuda().train()
        rec = {}
        for name in ("encoder.resnet.layer1", "encoder.resnet.layer2"):
            m.get_submodule(name).register_forward_hook(lambda mod, i, o, name=name: rec.__setitem__(name, o.detach().float()))
        before = dict(C.stats)
        x1, x2, feat = m(x)
        loss = loss_calc([x1, x2], label, CrossEntropy(-1), multi=True)
        loss.backward()
        rec.update(x1=x1.detach(), x2=x2.detach(), feat=feat.detach(), loss=float(loss),
                   gnorm={n: p.grad.float().norm().item() for n, p in m.named_parameters()},
                   used={k: C.stats[k] - before[k] for k in C.stats},
                   bn1_rm=m.encoder.resnet.bn1.running_mean.clone(), nbt=int(m.encoder.resnet.layer3[0].bn2.num_batches_tracked))
        return rec
    finally:
        E.set_fused(old[0])
        C.set_engine(old[1])


def test_fused_bf16_model_is_as_accurate_as_the_library_bf16_model():
    """Whole model, train mode, forward + backward, same weights / input three ways: float32 eager (the parity
    mode), bf16 eager with library kernels, bf16 with every hand-written kernel (tcgen05 fprop/dgrad/wgrad, fused BN,
    PPM).  The seeded random weights make the net chaotic (bf16 library compute drifts ~50 % from float32 at the
    logits), so the bar is: early layers within 3e-2 / 9e-2 of float32, and everywhere the hand-written path is no
    further from float32 than 1.5x the library bf16 path."""
    torch.manual_seed(0)
    x = torch.randn(4, 3, 256, 256, device="cuda").clamp(max=1.0)
    label = torch.randint(0, 6, (4, 256, 256), device="cuda")
    ref = _run_model(False, "cudnn", torch.float32, x, label)
    lib = _run_model(False, "cudnn", torch.bfloat16, x, label)
    our = _run_model(True, "auto", torch.bfloat16, x, label)
    assert our["used"]["tcgen05_fprop"] > 40 and our["used"]["tcgen05_dgrad"] > 40 and our["used"]["tcgen05_wgrad"] > 40
    assert lib["used"]["tcgen05_fprop"] == 0

    def err(r, k):
        return float((r[k].float() - ref[k].float()).abs().max() / ref[k].float().abs().max())

    assert err(our, "encoder.resnet.layer1") < 3e-2 and err(our, "encoder.resnet.layer2") < 9e-2
    for k in ("encoder.resnet.layer1", "encoder.resnet.layer2", "feat", "x1", "x2"):
        assert err(our, k) <= 1.5 * err(lib, k) + 1e-3, (k, err(our, k), err(lib, k))
    assert abs(our["loss"] - ref["loss"]) <= 1.5 * abs(lib["loss"] - ref["loss"]) + 2e-2 * abs(ref["loss"])
    # parameter-gradient norms: same distribution of deviations as the library path
    def gdev(r):
        return sorted(abs(r["gnorm"][n] - ref["gnorm"][n]) / (ref["gnorm"][n] + 1e-12) for n in ref["gnorm"])
    ours, libs = gdev(our), gdev(lib)
    med = len(ours) // 2
    # (two bf16 evaluations of this network differ run to run -- fp32 atomics re-amplified by every re-quantisation, see
    #  tests/test_bf16_parity_gpu.py -- so the ratio of two such noise samples is itself noisy: observed 0.8x .. 3x)
    assert ours[med] <= 3.0 * libs[med] + 1e-2, (ours[med], libs[med])
    assert ours[int(0.9 * len(ours))] <= 3.0 * libs[int(0.9 * len(libs))] + 5e-2
    # BatchNorm bookkeeping is part of the checkpoint ABI
    _close(our["bn1_rm"], ref["bn1_rm"], 1e-2)
    assert our["nbt"] == 1


def test_bn_statistics_groups_equal_separate_calls():
    """groups=2 over a concatenated batch == two consecutive calls (outputs, gradients, running statistics)."""
    from regda_b200.ops import norm
    g = torch.Generator(device="cuda").manual_seed(21)
    shape = (4, 256, 16, 16)
    ya = _cl((torch.randn(shape, device="cuda", generator=g) * 2 + 1).bfloat16())
    yb = _cl((torch.randn(shape, device="cuda", generator=g) * 0.5 - 2).bfloat16())
    ra, rb = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16()), _cl(torch.randn(shape, device="cuda", generator=g).bfloat16())
    ga, gb = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16()), _cl(torch.randn(shape, device="cuda", generator=g).bfloat16())
    bn1, bn2 = torch.nn.BatchNorm2d(256).cuda().train(), torch.nn.BatchNorm2d(256).cuda().train()
    ins = [t.clone().requires_grad_(True) for t in (ya, yb, ra, rb)]
    oa = norm.bn_act(ins[0], bn1, residual=ins[2], relu=True)
    ob = norm.bn_act(ins[1], bn1, residual=ins[3], relu=True)
    torch.autograd.backward([oa, ob], [ga, gb])
    ycat = _cl(torch.cat([ya, yb])).requires_grad_(True)
    rcat = _cl(torch.cat([ra, rb])).requires_grad_(True)
    oc = norm.bn_act(ycat, bn2, residual=rcat, relu=True, groups=2)
    oc.backward(_cl(torch.cat([ga, gb])))
    # the fp32 statistics are accumulated with atomics (order differs between launches): outputs may differ by one
    # bf16 ulp, and a ReLU mask bit may flip where the pre-activation rounds to +-0
    _close(oc, torch.cat([oa, ob]), 8e-3)
    _close_most(ycat.grad, torch.cat([ins[0].grad, ins[1].grad]), 1e-2)
    _close_most(rcat.grad, torch.cat([ins[2].grad, ins[3].grad]), 1e-2)
    _close(bn2.running_mean, bn1.running_mean, 1e-5)
    _close(bn2.running_var, bn1.running_var, 1e-5)
    _close(bn2.weight.grad, bn1.weight.grad, 3e-2)      # sums over dz: a handful of flipped ReLU mask bits
    _close(bn2.bias.grad, bn1.bias.grad, 3e-2)
    assert int(bn2.num_batches_tracked) == 2 == int(bn1.num_batches_tracked)


def _pair_run(dtype, pair, xs, xt):
    from oracle import step_oracle as so
    from regda_b200.models import Encoder as E
    from regda_b200.ops import conv as C
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    m = E.Deeplabv2(cfg, compute_dtype=dtype)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m = m.cuda().train()
    if dtype == torch.float64:
        m = m.double()
        C.set_engine("cudnn")         # float64 exists only on the library engine: this run pins host LOGIC, not kernels
    try:
        if pair:
            (a1, a2, fa), (b1, b2, fb) = m.forward_pair(xs, xt)
        else:
            a1, a2, fa = m(xs)
            b1, b2, fb = m(xt)
        (a1.square().mean() + a2.mean() + b1.square().mean() - 0.5 * b2.mean()).backward()   # no exactly-cancelling terms
    finally:
        C.set_engine("tcgen05")
    return dict(out=[t.detach().double() for t in (a1, a2, fa, b1, b2, fb)],
                g={n: p.grad.double().clone() for n, p in m.named_parameters()},
                rm=m.encoder.resnet.layer2[0].bn1.running_mean.clone(), rv=m.layer5.conv_last[1].running_var.clone())


def test_forward_pair_equals_two_forward_calls_float64():
    """Deeplabv2.forward_pair(xs, xt) == (model(xs), model(xt)) -- logits, features, parameter gradients, BN buffers --
    in float64, where the library has exact algorithms for every batch size: pins the LOGIC (statistics groups,
    slicing, running-stat order) at 1e-6."""
    torch.manual_seed(1)
    xs = torch.randn(2, 3, 64, 64, device="cuda").clamp(max=1.0)
    xt = (torch.randn(2, 3, 64, 64, device="cuda") * 0.7 + 0.3).clamp(max=1.0)
    sep, par = _pair_run(torch.float64, False, xs, xt), _pair_run(torch.float64, True, xs, xt)
    for a, b in zip(sep["out"], par["out"]):
        _close(b, a, 1e-5)
    _close(par["rm"], sep["rm"], 1e-6)
    _close(par["rv"], sep["rv"], 1e-6)
    gmax = max(float(v.norm()) for v in sep["g"].values())
    worst = max(float((par["g"][n] - sep["g"][n]).norm() / (sep["g"][n].norm() + 1e-6 * gmax)) for n in sep["g"])
    assert worst < 1e-4, worst


def test_forward_pair_bf16_is_as_accurate_as_two_calls():
    """The bf16 hand-written path, paired vs two calls, both measured against the float32 two-call model (the random
    50-layer network amplifies one-ulp differences, so pair-vs-separate is compared through their distance to the
    float32 anchor): the paired forward may not be further from float32 than 2x the two-call forward (the exact
    statement about the pairing logic is the float64 test above; the tight bf16 statement is tests/test_bf16_parity_gpu.py)."""
    torch.manual_seed(1)
    xs = torch.randn(2, 3, 256, 256, device="cuda").clamp(max=1.0)
    xt = (torch.randn(2, 3, 256, 256, device="cuda") * 0.7 + 0.3).clamp(max=1.0)
    ref, sep, par = _pair_run(torch.float32, False, xs, xt), _pair_run(torch.bfloat16, False, xs, xt), _pair_run(torch.bfloat16, True, xs, xt)

    def oerr(r):
        return [float((a - b).abs().max() / b.abs().max()) for a, b in zip(r["out"], ref["out"])]

    for ep, es in zip(oerr(par), oerr(sep)):
        assert ep <= 2.0 * es + 5e-3, (ep, es)            # (a ratio of two bf16 noise samples: observed 0.7x .. 1.5x)
    gmax = max(float(v.norm()) for v in ref["g"].values())

    def gdev(r):
        return sorted(float((r["g"][n] - ref["g"][n]).norm() / (ref["g"][n].norm() + 1e-6 * gmax)) for n in ref["g"])

    dp, ds = gdev(par), gdev(sep)
    for q in (0.5, 0.9):
        i = int(q * len(dp))
        assert dp[i] <= 2.0 * ds[i] + 1e-2, (q, dp[i], ds[i])
    _close(par["rm"], sep["rm"], 6e-2)
    _close(par["rv"], sep["rv"], 6e-2)


@pytest.mark.parametrize("shape", [(2, 64, 32, 32), (1, 64, 17, 23), (2, 128, 9, 8)], ids=str)
def test_maxpool3s2_matches_torch(shape):
    from regda_b200.ops import norm
    g = torch.Generator(device="cuda").manual_seed(5)
    # post-ReLU-like input: many exact zeros, so window ties (first maximum wins) are exercised
    x = _cl(torch.relu(torch.randn(shape, device="cuda", generator=g)).bfloat16()).requires_grad_(True)
    y = norm.max_pool3s2(x)
    xr = x.detach().clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y, yr)
    gy = _cl(torch.randn(yr.shape, device="cuda", generator=g).bfloat16())
    y.backward(gy)
    yr.backward(gy)
    _close(x.grad, xr.grad, 1e-2)          # sums of up to four bf16 gradients: rounding order only


@pytest.mark.parametrize("shape", [(4, 2048, 16, 16), (2, 256, 8, 8), (3, 512, 7, 9)], ids=str)
def test_instance_norm_matches_torch(shape):
    from regda_b200.ops import norm
    g = torch.Generator(device="cuda").manual_seed(9)
    x = _cl((torch.randn(shape, device="cuda", generator=g) * 1.5 + 0.5).bfloat16()).requires_grad_(True)
    assert norm.instance_norm_supported(x)
    y = norm.instance_norm(x, 1e-5)
    xr = x.detach().float().requires_grad_(True)
    yr = F.instance_norm(xr, eps=1e-5)
    _close(y, yr, 1e-2)
    gy = _cl(torch.randn(shape, device="cuda", generator=g).bfloat16())
    y.backward(gy)
    yr.backward(gy.float())
    _close(x.grad, xr.grad, 1.5e-2)


@pytest.mark.parametrize("case", ["plain", "downsample_s1", "downsample_s2"])
def test_bn_backward_fused_into_dgrad_epilogue_matches_standalone_reduce(case):
    """Three bottleneck blocks, bf16, ONE forward and two backward passes over the retained graph: first with the BatchNorm
    backward reductions done in the consumer convolutions' data-gradient epilogue (regda_conv_dgrad_bnred_bf16: masked dz +
    sum dz, sum dz*y), then with every handle broken, i.e. the standalone reduce + apply kernels.  Same activations, same ReLU
    masks; what differs is summation order and where the masked gradient is rounded to bf16, hence 1e-2 of the gradient
    scale and 1e-2 of its mean (one bf16 ulp is 3.9e-3; measured 3e-3 .. 6e-3 after nine layers; a wrong mask or sum is O(1)).  (Two separate forwards are NOT comparable at this tolerance: the statistics are fp32 atomics
    and nine bf16 layers amplify their last bit.)  The exact op-level check is
    tests/test_conv_gpu.py::test_dgrad_bnred_matches_masked_dgrad_and_reductions."""
    from regda_b200.models import Encoder as E
    from regda_b200.ops import norm as fnorm

    handles = []
    orig_init = fnorm.BnHandle.__init__

    def tracking_init(self):
        orig_init(self)
        handles.append(self)

    torch.manual_seed(7)
    if case == "plain":
        blocks = [E.Bottleneck(256, 64), E.Bottleneck(256, 64), E.Bottleneck(256, 64)]
    elif case == "downsample_s1":
        blocks = [E.Bottleneck(256, 64), E.Bottleneck(256, 128, 1, 2, downsample=True), E.Bottleneck(512, 128, 1, 2)]
    else:
        blocks = [E.Bottleneck(256, 64), E.Bottleneck(256, 128, 2, 1, downsample=True), E.Bottleneck(512, 128)]
    net = torch.nn.Sequential(*blocks).cuda().train().to(memory_format=torch.channels_last)
    x = torch.randn(4, 256, 24, 40, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    old = fnorm.FUSE_BN_BWD
    fnorm.FUSE_BN_BWD = True
    fnorm.BnHandle.__init__ = tracking_init
    E._GROUPS = 2
    try:
        out = net(x)
    finally:
        E._GROUPS = 1
        fnorm.BnHandle.__init__ = orig_init
        fnorm.FUSE_BN_BWD = old
    n_fused = sum(1 for h in handles if h.fused)
    # every BatchNorm+ReLU output inside the stack is consumed by convolutions only, and every one of them (the stride-2 ones
    # through the zero-inserted dY) carries the reductions in its data-gradient epilogue
    assert n_fused == {"plain": 8, "downsample_s1": 8, "downsample_s2": 8}[case], (n_fused, len(handles))
    g = torch.randn(out.shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)

    def backward():
        x.grad = None
        for p in net.parameters():
            p.grad = None
        out.backward(g, retain_graph=True)
        return x.grad.float().clone(), {n: p.grad.float().clone() for n, p in net.named_parameters()}

    gx1, g1 = backward()
    for h in handles:
        h.broken = True
    gx0, g0 = backward()

    def close(a, b, name):
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 1e-2 * scale, (name, float((a - b).abs().max()), scale)
        assert float((a - b).abs().mean()) <= 1e-2 * (float(b.abs().mean()) + 1e-12), (name, float((a - b).abs().mean()), float(b.abs().mean()))
    close(gx1, gx0, "dx")
    for n in g0:
        close(g1[n], g0[n], n)
