"""CPU: the host-side layout algebra behind two kernels, checked against plain torch -- no GPU, no CUDA call.

* the folded PPM fuse convolution (regda_b200/ops/ppm_fold.py): conv3x3(cat(fin, up(p_k))) == conv3x3(fin, W[:, :cf]) + A . G with the
  constant basis A of `basis_weight` and G_k = p_k . W_k^T laid out as the kernels lay it out ((o, tap) columns -> GT[o][(cell, tap)]);
  reference formulation: regda/models/Encoder.py:43-52 + the 3x3 convolution of :33-34.
* the stem patch matrix (regda_b200/ops/stem.py, csrc/stem.cu): k = r*24 + s*3 + c, 8 groups of 24, against F.conv2d(7x7 / 2)
  (regda/_resnets.py:150).
"""
import torch
import torch.nn.functional as F


def test_folded_fuse_convolution_algebra_matches_concat_then_conv():
    from regda_b200.ops import ppm_fold
    torch.manual_seed(0)
    scales = (1, 2, 3, 6)
    b, cf, cb, h, w, O = 2, 16, 8, 12, 12, 5
    fin = torch.randn(b, cf, h, w, dtype=torch.float64)
    brs = [torch.randn(b, cb, s, s, dtype=torch.float64) for s in scales]
    wt = torch.randn(O, cf + 4 * cb, 3, 3, dtype=torch.float64)
    cat = torch.cat([fin] + [F.interpolate(t, (h, w), mode="bilinear", align_corners=False) for t in brs], 1)
    want = F.conv2d(cat, wt, padding=1)
    # the folded form, with the layouts of the CUDA path
    kp = ppm_fold._kp(scales)
    a = ppm_fold.basis_weight(h, w, scales, "cpu").double().view(h * w, kp)                   # A[px][cell*9 + tap] (bf16-rounded weights)
    gt = torch.zeros(b, O, kp, dtype=torch.float64)                                            # GT[img][o][(cell, tap)]
    cell0 = 0
    for k, (p, s) in enumerate(zip(brs, scales)):
        wk = wt[:, cf + k * cb:cf + (k + 1) * cb].permute(0, 2, 3, 1).reshape(O * 9, cb)       # rows (o, tap) of the OHWI weight block
        g = p.permute(0, 2, 3, 1).reshape(b, s * s, cb) @ wk.t()                               # G_k[(img, cell)][o*9 + tap]
        gt[:, :, cell0 * 9:(cell0 + s * s) * 9] = g.view(b, s * s, O, 9).permute(0, 2, 1, 3).reshape(b, O, s * s * 9)
        cell0 += s * s
    yppm = torch.einsum("pk,bok->bop", a, gt).view(b, O, h, w)
    got = F.conv2d(fin, wt[:, :cf], padding=1) + yppm
    # A's entries are bilinear weights rounded to bf16 (the GEMM operand): 3 significant digits
    assert float((got - want).abs().max()) <= 4e-3 * float(want.abs().max())
    # with the exact basis the identity is exact
    cols = []
    for s in scales:
        eye = torch.eye(s * s, dtype=torch.float64).view(s * s, 1, s, s)
        up = F.pad(F.interpolate(eye, (h, w), mode="bilinear", align_corners=False)[:, 0], (1, 1, 1, 1))
        cols.append(torch.stack([up[:, r:r + h, c:c + w].reshape(s * s, h * w) for r in range(3) for c in range(3)], 1).reshape(s * s * 9, h * w))
    a64 = torch.cat(cols, 0).t()
    got64 = F.conv2d(fin, wt[:, :cf], padding=1) + torch.einsum("pk,bok->bop", a64, gt[:, :, :a64.shape[1]]).view(b, O, h, w)
    assert float((got64 - want).abs().max()) <= 1e-10 * float(want.abs().max())


def test_stem_patch_rows_are_eight_groups_of_24():
    from regda_b200.ops import stem
    torch.manual_seed(1)
    n, h, w, cout = 2, 20, 26, 64
    x = torch.randn(n, 3, h, w, dtype=torch.float64)
    wt = torch.randn(cout, 3, 7, 7, dtype=torch.float64)
    want = F.conv2d(x, wt, stride=2, padding=3)
    oh, ow = want.shape[-2:]
    # patch matrix in the kernel's order: k = r*24 + s*3 + c for the 7 filter rows, zero elsewhere
    cols = F.unfold(x, 7, padding=3, stride=2).view(n, 3, 7, 7, oh * ow)                        # [n][c][r][s][px]
    a = torch.zeros(n, oh * ow, stem.K_PAD, dtype=torch.float64)
    a.view(n, oh * ow, 8, 24)[:, :, :7, :21] = cols.permute(0, 4, 2, 3, 1).reshape(n, oh * ow, 7, 21)
    wp = torch.zeros(cout, stem.K_PAD, dtype=torch.float64)
    stem._rows(wp, cout).copy_(wt.permute(0, 2, 3, 1).reshape(cout, 7, 21))                     # what _StemConvFn.forward packs
    got = (a @ wp.t()).permute(0, 2, 1).reshape(n, cout, oh, ow)
    assert float((got - want).abs().max()) <= 1e-10 * float(want.abs().max())
    # the gradient's way back: rows of a patch-ordered weight gradient land in the OHWI gradient
    gw = torch.randn(cout, stem.K_PAD, dtype=torch.float64)
    back = torch.zeros(cout, 3, 7, 7, dtype=torch.float64).contiguous(memory_format=torch.channels_last)
    back.permute(0, 2, 3, 1).reshape(cout, 7, 21).add_(stem._rows(gw, cout))
    assert torch.equal(back.permute(0, 2, 3, 1).reshape(cout, 7, 21), stem._rows(gw, cout))
