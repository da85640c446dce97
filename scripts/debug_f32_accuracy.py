"""Accuracy of the float32 (split-operand) convolution path per shape, against float64 (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from regda_b200.ops import conv as C
torch.manual_seed(0)
for (n, cin, h, w, cout, k, stride, pad, dil) in [(2, 64, 16, 16, 64, 3, 1, 1, 1), (2, 256, 16, 16, 1024, 1, 1, 0, 1), (2, 512, 6, 6, 512, 3, 1, 2, 2),
                                                  (2, 4096, 8, 8, 512, 3, 1, 1, 1), (2, 128, 32, 32, 128, 3, 2, 1, 1)]:
    m = C.Conv2d(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=False).cuda()
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(x); gy = torch.randn_like(y); y.backward(gy)
    xr = x.detach().double().requires_grad_(True); wr = m.weight.detach().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride, pad, dil); yr.backward(gy.double())
    x32 = x.detach().clone().requires_grad_(True); w32 = m.weight.detach().clone().requires_grad_(True)
    torch.backends.cudnn.allow_tf32 = False
    y32 = F.conv2d(x32, w32, None, stride, pad, dil); y32.backward(gy)
    e = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
    print(f"K={cin*k*k:6d} ours y {e(y, yr):.1e} dx {e(x.grad, xr.grad):.1e} dw {e(m.weight.grad, wr.grad):.1e} | cudnn-f32 y {e(y32, yr):.1e} dx {e(x32.grad, xr.grad):.1e} dw {e(w32.grad, wr.grad):.1e}")
