"""Launch loops of the step's two largest backward kernels on their layer3 shapes (16 images of 32x32), for ncu captures:
   python scripts/ncu_l3_bwd.py wgrad   -> conv_wgrad_persistent_kernel<256,4>: dW of the 3x3 256->256 convolution
   python scripts/ncu_l3_bwd.py bnred   -> conv_persistent_kernel<256,3,true,..,true>: data gradient of the 1x1 256->1024 convolution
                                            + shortcut addend + ReLU mask + the BatchNorm-backward sums, as the step runs it"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regda_b200.ops import tc

torch.manual_seed(0)
n, hw = 16, 32


def cl(*shape):
    return torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if sys.argv[1] == "wgrad":
    gy, x = cl(n, 256, hw, hw), cl(n, 256, hw, hw)
    gw = torch.zeros(256, 256, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
    launch = lambda: tc.wgrad_accumulate(gy, x, gw, 1, 1, 1)
else:
    cin, cout = 1024, 256                       # the next block's conv1 (1024 -> 256): its data gradient has N = 1024, K = 256
    gy, w = cl(n, cout, hw, hw), cl(cout, cin, 1, 1)
    addend, bn_y = cl(n, cin, hw, hw), cl(n, cin, hw, hw)
    mask = torch.randint(0, 256, (n * hw * hw * cin // 8,), dtype=torch.uint8, device="cuda")
    red = torch.zeros(2, 2, cin, device="cuda")
    launch = lambda: tc.dgrad_bnred(gy, w, (n, cin, hw, hw), 1, 0, 1, addend, bn_y, mask, red, 2)
for _ in range(8):
    flush.zero_()
    launch()
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == "time":          # CUDA events around single launches, L2 flushed before each
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{sys.argv[1]}: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us per launch (L2 flushed)")
