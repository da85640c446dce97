"""Per-shape timing of the tcgen05 convolution kernels against the library (cuDNN through
torch) on the ResNet-101 OS16 + PPM shapes at the benchmark batch (8 images per forward).
Usage: python scripts/bench_conv.py [--wgrad]   (prints one line per shape; CUDA events, L2 flushed)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from regda_b200.ops import tc

# (name, count, cin, hw, cout, k, pad, dil)
SHAPES = [
    ("l1.conv1a", 1, 64, 128, 64, 1, 0, 1), ("l1.conv2", 3, 64, 128, 64, 3, 1, 1), ("l1.conv3", 4, 64, 128, 256, 1, 0, 1),
    ("l1.conv1", 2, 256, 128, 64, 1, 0, 1), ("l2.0.conv1", 1, 256, 128, 128, 1, 0, 1), ("l2.conv3", 4, 128, 64, 512, 1, 0, 1),
    ("l2.conv1", 3, 512, 64, 128, 1, 0, 1), ("l2.conv2", 3, 128, 64, 128, 3, 1, 1), ("l3.0.conv1", 1, 512, 64, 256, 1, 0, 1),
    ("l3.conv3", 23, 256, 32, 1024, 1, 0, 1), ("l3.conv1", 22, 1024, 32, 256, 1, 0, 1), ("l3.conv2", 22, 256, 32, 256, 3, 1, 1),
    ("l4.0.conv1", 1, 1024, 32, 512, 1, 0, 1), ("l4.0.conv2", 1, 512, 32, 512, 3, 1, 1), ("l4.conv3", 3, 512, 32, 2048, 1, 0, 1),
    ("l4.0.down", 1, 1024, 32, 2048, 1, 0, 1), ("l4.conv1", 2, 2048, 32, 512, 1, 0, 1), ("l4.conv2d", 2, 512, 32, 512, 3, 2, 2),
    ("head.fuse", 2, 4096, 32, 512, 3, 1, 1),
]


GRAPH_INNER = 0      # --graph N: time N back-to-back launches replayed from a CUDA graph (no host launch floor, warm L2)


def timeit(fn, flush, reps=10):
    for _ in range(3):
        fn()
    if GRAPH_INNER > 0:
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(GRAPH_INNER):
                fn()
        gr.replay()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / GRAPH_INNER)
        ts.sort()
        return ts[len(ts) // 2]
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--wgrad", action="store_true", help="time the weight-gradient kernel instead of fprop")
    ap.add_argument("--dgrad", action="store_true", help="time the data-gradient kernel instead of fprop")
    ap.add_argument("--stats", action="store_true", help="fprop with the fused BatchNorm statistics epilogue (2 groups), as the step runs it")
    ap.add_argument("--graph", type=int, default=0, help="time N launches replayed from one CUDA graph")
    ap.add_argument("--min-tiles-256", type=int, default=0, help="tile policy: minimum number of 128x256 tiles for the 256-wide tile (0 = default)")
    a = ap.parse_args()
    if a.min_tiles_256:
        from regda_b200 import capi
        capi.check(capi.lib().regda_conv_tune(a.min_tiles_256))
    global GRAPH_INNER
    GRAPH_INNER = a.graph
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = {"tc": 0.0, "lib": 0.0, "flop": 0.0}
    for name, cnt, cin, hw, cout, k, pad, dil in SHAPES:
        if a.only and a.only not in name:
            continue
        x = torch.randn(a.n, cin, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
        flop = 2.0 * a.n * hw * hw * cout * cin * k * k
        if a.wgrad or a.dgrad:
            gy = torch.randn(a.n, cout, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
            gw = torch.zeros(cout, cin, k, k, device="cuda").contiguous(memory_format=torch.channels_last)
            mask = [a.dgrad, a.wgrad, False]
            if a.wgrad:
                t_tc = timeit(lambda: tc.wgrad_accumulate(gy, x, gw, 1, pad, dil), flush)
            else:
                t_tc = timeit(lambda: tc.dgrad(gy, w, x.shape, 1, pad, dil), flush)
            t_lib = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [pad, pad], [dil, dil], False, [0, 0], 1, mask), flush)
        else:
            t_tc = timeit(lambda: tc.fprop(x, w, 1, pad, dil, 2 if a.stats else None), flush)
            t_lib = timeit(lambda: F.conv2d(x, w, None, 1, pad, dil), flush)
        tot["tc"] += cnt * t_tc; tot["lib"] += cnt * t_lib; tot["flop"] += cnt * flop
        print(f"{name:12s} x{cnt:2d} M={a.n*hw*hw:6d} N={cout:4d} K={cin*k*k:5d}  tcgen05 {t_tc*1e3:8.1f} us {flop/t_tc/1e9:7.1f} TF/s | "
              f"library {t_lib*1e3:8.1f} us {flop/t_lib/1e9:7.1f} TF/s | speedup {t_lib/t_tc:5.2f}x", flush=True)
    print(f"{'wgrad' if a.wgrad else 'dgrad' if a.dgrad else 'fprop'} total (weighted by layer count): tcgen05 {tot['tc']:.3f} ms ({tot['flop']/tot['tc']/1e9:.1f} TF/s), "
          f"library {tot['lib']:.3f} ms ({tot['flop']/tot['lib']/1e9:.1f} TF/s)")


if __name__ == "__main__":
    main()
