"""Benchmark harness of the self-training step (bench.py's default workload):
BASELINE.json configs[1] -- st.regda.2potsdam, ResNet-101 DeepLab, bf16, 8 source + 8 target
512x512 tiles per GPU, 6 classes.  Called by bench.py; see there for the JSON contract."""
from __future__ import annotations

import json
import os
import time

import torch

from regda_b200 import capi, synth

CLASS_NUM, H, W, B = 6, 512, 512, 8
FLOP_PER_IMG_FWD_BWD = 543.5e9       # 2 * 90.59 GMAC * 3 (R101, 512^2, C=6)
# BASELINE.json configs: P = configs[1] (the headline: st.regda.2potsdam, ResNet-101, 6 classes, 8 + 8 tiles of 512x512 per GPU);
# L = configs[3] (ResNet-50, 7 classes, 1024x1024 tiles, global batch 16 over 2 GPUs = 8 + 8 tiles per GPU: the large-tile stress)
CONFIGS = {
    "P": dict(resnet="resnet101", classes=6, hw=(512, 512), batch=8, flop_per_img=543.5e9,
              name="st.regda.2potsdam (ResNet-101 DeepLab, 6 classes, 512x512 tiles)"),
    "L": dict(resnet="resnet50", classes=7, hw=(1024, 1024), batch=8, flop_per_img=1706.8e9,
              name="large-tile stress (ResNet-50 DeepLab, 7 classes, 1024x1024 tiles; BASELINE.json configs[3])"),
}


def model_cfg(resnet, classes):
    return dict(backbone=dict(resnet_type=resnet, output_stride=16, pretrained=False), multi_layer=True, cascade=False,
                use_ppm=True, ppm=dict(num_classes=classes, use_aux=False, fc_dim=2048), inchannels=2048,
                num_classes=classes, is_ins_norm=True)


MODEL_CFG = model_cfg("resnet101", CLASS_NUM)


def build(device, world, resnet="resnet101", use_graph=True, n_regions=200, seed=2333, batch=B, hw=(H, W), stage=3, classes=CLASS_NUM):
    """stage 3: the self-training step (tools/train_ssl_reg.py, the headline workload); stage 2: the alignment step
    (tools/train_align_reg.py) -- same model and shapes, inputs without the offline soft labels"""
    from regda_b200.gast.alignment import Aligner
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.trainer import AlignStep, GraphedStep, SelfTrainingStep
    from regda_b200.utils.local_region_homog import Homogenizer
    torch.manual_seed(seed)
    model = Deeplabv2(model_cfg(resnet, classes), compute_dtype=torch.bfloat16).to(device).train()
    inputs = synth.step_inputs(batch, hw[0], hw[1], classes, n_regions, device=device, seed=seed)
    images_s, label_s, images_t, soft_t, regs_t, proto = inputs
    aligner = Aligner(None, 2048, classes, -1, 0.996, device=device)
    aligner.prototypes = proto.clone()
    bound = int(regs_t.max()) + 1
    hom = Homogenizer(percent=0.5, class_num=classes, ignore_label=-1, region_bound=bound, strict=False)
    if stage == 2:
        step = AlignStep(model, aligner, hom, class_num=classes, ignore_label=-1, world_size=world)
        tensors = [images_s, label_s, images_t, regs_t]
    else:
        step = SelfTrainingStep(model, aligner, hom, class_num=classes, ignore_label=-1, world_size=world)
        tensors = [images_s, label_s, images_t, soft_t, regs_t]
    runner = None
    if use_graph:
        runner = GraphedStep(step, tensors, lr=1e-2)
    return model, step, runner, tensors


def ncu_evidence(key):
    """dram bytes / tensor-pipe figures of one `ncu --set full` capture per kernel, committed under profiles/ (a run under the
    profiler is never a bench value, so they are READ here, not measured): profiles/ncu_metrics.json {key: {...}}"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_metrics.json")
    try:
        return json.load(open(path)).get(key)
    except (OSError, ValueError):
        return None


def head_conv_launcher():
    """The dominant kernel of the step by algorithmic FLOPs, launched exactly as the product path launches it
    (regda_b200/ops/ppm_fold.py): the 3x3 PPM fuse convolution over the 2048 feature channels of the 4096-channel OHWI weight,
    read in place (channel stride 4096), on the 32x32 feature maps of the 16 images of a step -- implicit GEMM M=16384, N=512,
    K=18432 -- with the pyramid branches' contribution as the bf16 epilogue addend and the next BatchNorm's statistics (2 groups)
    fused.  Returns (launch, flop)."""
    n, cf, ct, hw, cout = 2 * B, 2048, 4096, H // 16, 512
    cl = torch.channels_last
    x = torch.randn(n, cf, hw, hw, device="cuda").bfloat16().contiguous(memory_format=cl)
    w = (torch.randn(cout, ct, 3, 3, device="cuda") / (ct * 9) ** 0.5).bfloat16().contiguous(memory_format=cl)
    add = torch.randn(n, cout, hw, hw, device="cuda").bfloat16().contiguous(memory_format=cl)
    y = torch.empty((n, cout, hw, hw), dtype=torch.bfloat16, device="cuda", memory_format=cl)
    stats = torch.zeros(2, 2, cout, device="cuda")

    def launch():
        capi.call("regda_conv_fprop_addend_bf16", capi.ptr_any(x), capi.ptr_any(w), ct, capi.ptr_any(y), n, hw, hw, cf, cout, 3, 3, 1, 1, 1,
                  capi.ptr_any(add), capi.ptr_any(stats), 2, 0, capi.stream())

    launch.keep = (x, w, add, y, stats)
    return launch, 2.0 * n * hw * hw * cout * cf * 9


def top_kernel_roofline(pk, reps=10):
    """head_conv_launcher() timed alone with CUDA events on the launch stream, L2 flushed between launches (burst peak is the
    denominator); DRAM traffic / tensor-pipe figures are read from the committed ncu capture of the same launch."""
    launch, flop = head_conv_launcher()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        launch()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    ach = flop / (ms * 1e-3) / 1e12
    ev = ncu_evidence("conv_head_fprop") or {}
    return {"bound": "tensor", "achieved": round(ach, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_burst"], 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed ncu --set full capture
            "traffic": ev.get("dram_bytes"), "traffic_source": ev.get("source"),
            "ncu_tensor_pipe_active_pct": ev.get("tensor_pipe_active_pct"),
            "peak_source": pk["source"] + " (burst: kernel timed alone)",
            "kernel": "conv_persistent_kernel<256,4> (tcgen05 128x256 tiles, TMA-store epilogue with addend + BatchNorm statistics) on the folded PPM "
                      "fuse conv (3x3 over the 2048 feature channels of the 4096-channel weight, 16x32x32 px: M=16384 N=512 K=18432)",
            "algorithmic_flop_per_launch": flop, "us_per_launch": round(ms * 1e3, 1)}


def lrh_subrecords(pk, device, steps=10):
    """BASELINE.json metric part (ii): Homogenizer.forward on 128 x 512 x 512 int64 tiles at 50 / 500 / 5000 regions per tile --
    Gpix/s, fraction of the HBM roofline at 24 algorithmic bytes per pixel, and a bit-exact parity gate against the C oracle
    on sampled tiles (the full-size bit-exact checks are tests/test_lrh_gpu.py)."""
    import numpy as np
    from oracle import cbind
    from regda_b200.utils.local_region_homog import Homogenizer
    out = []
    for n_regions in (50, 500, 5000):
        reg = synth.region_maps(128, 512, 512, n_regions, device=device, seed=2333)
        lab = synth.lrh_labels(reg, 6, -1, seed=2334)
        hom = Homogenizer(percent=0.5, class_num=6, ignore_label=-1, region_bound=int(reg.max()) + 1, strict=False)
        o = hom(lab, reg)
        parity = all(np.array_equal(o[i:i + 1].cpu().numpy(), cbind.lrh(lab[i:i + 1].cpu().numpy(), reg[i:i + 1].cpu().numpy(), 6, -1, 0.5))
                     for i in (0, 127))
        for _ in range(3):
            hom(lab, reg)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            hom(lab, reg)
            ev[i + 1].record()
        torch.cuda.synchronize()
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
        ms = per[len(per) // 2]
        npx = lab.numel()
        gbs = 24 * npx / (ms * 1e-3) / 1e9
        out.append({"regions_per_tile": n_regions, "gpix_per_s": round(npx / (ms * 1e-3) / 1e9, 2), "us_per_launch": round(ms * 1e3, 1),
                    "achieved_gbs": round(gbs, 1), "frac": round(gbs / pk["hbm"], 4), "bit_exact_vs_oracle": bool(parity)})
        del reg, lab, o
    return {"workload": "128x512x512 int64 tiles, 6 classes, percent 0.5 (BASELINE.json configs[4])", "bound": "hbm", "peak": pk["hbm"],
            "unit": "GB/s", "algorithmic_bytes_per_pixel": 24, "points": out}


def library_baseline(device, steps=5):
    """the same step with every convolution on the library (cuDNN through torch, bf16) and torch's own BatchNorm / pooling /
    upsampling kernels (ops.conv.set_engine('cudnn'), Encoder.set_fused(False)), CUDA-graph replayed like the product path:
    the informative GPU baseline (what PyTorch + cuDNN gives on this B200), next to the CPU reference arm"""
    from regda_b200.models import Encoder as E
    from regda_b200.ops import conv as C
    old = (E.FUSED, C.ENGINE)
    E.set_fused(False)
    C.set_engine("cudnn")
    try:
        model, step, runner, tensors = build(device, 1, use_graph=True, seed=2333)
        for _ in range(3):
            runner(*tensors, lr=1e-2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            runner(*tensors, lr=1e-2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del model, step, runner, tensors
        torch.cuda.empty_cache()
        return {"value": round(2 * B / (ms * 1e-3), 1), "unit": "images/s", "ms_per_step": round(ms, 2), "steps": steps,
                "what": "same step, torch bf16 + cuDNN convolutions + ATen BatchNorm/pool/upsample, CUDA graph, 1 GPU"}
    finally:
        E.set_fused(old[0])
        C.set_engine(old[1])


def workload_config(stage, imgs, world, use_graph, engine, config="P"):
    """the `config` object of the bench line; `stage` is the integer the run was BUILT with (3: the headline workload)"""
    assert stage in (2, 3)
    if stage == 3 and config == "L":
        what = ("self-training step (tools/train_ssl_reg.py:198-241) at BASELINE.json configs[3]: ResNet-50 DeepLab (2 PPM heads), 7 classes, "
                "8 source + 8 target 1024x1024 tiles per GPU, refine+select+LRH+prototype EMA+4 CE+backward+clip+SGD")
    elif stage == 3:
        what = ("st.regda.2potsdam self-training step (tools/train_ssl_reg.py:198-241): ResNet-101 DeepLab (2 PPM heads), 8 source + 8 target "
                "512x512 tiles per GPU, refine+select+LRH+prototype EMA+4 CE+backward+clip+SGD")
    else:
        what = ("st.regda.2potsdam stage-2 alignment step (tools/train_align_reg.py:144-196): ResNet-101 DeepLab, 8 source + 8 target "
                "512x512 tiles per GPU, prototype EMA+own-prediction soft labels+refine+select+LRH+2 CE+2 prototype-contrastive "
                "losses+backward+clip+SGD")
    return {"workload": what, "global_batch": imgs, "parallelism": f"dp{world}", "cuda_graph": use_graph, "conv_engine": engine,
            "l2": "per-step working set (activations ~6 GB) far larger than L2"}


def run(args, rank, world, local, pk, ClockSampler, barrier, max_over_ranks):
    dev = torch.device("cuda", local)
    if os.environ.get("REGDA_TUNE_MIN256"):                 # tile-policy sweeps only (profiles/): never set by the driver
        capi.check(capi.lib().regda_conv_tune(int(os.environ["REGDA_TUNE_MIN256"])))
    if os.environ.get("REGDA_TUNE_BN_TRIGGER"):             # PDL sweeps only (profiles/pdl_sweep_*): never set by the driver
        torch.cuda.set_device(dev)
        torch.zeros(1, device=dev)
        capi.check(capi.lib().regda_bn_tune(int(os.environ["REGDA_TUNE_BN_TRIGGER"])))
    use_graph = os.environ.get("REGDA_GRAPH", "1") != "0"
    calls0 = capi.launch_count
    stage = 2 if getattr(args, "workload", None) == "align" else 3
    cname = getattr(args, "config", "P") or "P"
    conf = CONFIGS[cname]
    model, step, runner, tensors = build(dev, world, resnet=conf["resnet"], use_graph=use_graph, seed=2333 + rank, stage=stage,
                                         batch=conf["batch"], hw=conf["hw"], classes=conf["classes"])
    lr = 1e-2

    def one_step(inp):
        if runner is not None:
            return runner(*inp, lr=lr)
        return step(*inp, lr)

    calls_per_step = None
    for i in range(max(args.warmup, 3)):
        c0 = capi.launch_count
        out = one_step(tensors)
        calls_per_step = capi.launch_count - c0
    if runner is not None:
        # kernels are replayed from the graph: the per-step count is what one eager pass issued during capture
        c0 = capi.launch_count
        step._step_impl(*tensors)
        calls_per_step = capi.launch_count - c0
    torch.cuda.synchronize()
    loss0 = float(out["loss"])
    assert loss0 == loss0 and abs(loss0) < 1e4, f"non-finite / diverged loss {loss0}"

    sampler = ClockSampler(local)
    barrier(world)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = one_step(tensors)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    imgs = 2 * conf["batch"] * world
    value = imgs / (ms * 1e-3)

    # ---- e2e: pinned host inputs -> H2D every step (double-buffered on a copy stream) -> step -> loss D2H
    host = [t.cpu().pin_memory() for t in tensors]
    bufs = [[torch.empty_like(t) for t in tensors] for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, h in zip(bufs[slot], host):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    for s in range(2):
        consumed[s].record()
    h2d = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    n_e2e = max(3, min(args.steps, 10))

    def e2e_loop(n):
        stage_inputs(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                stage_inputs(slot ^ 1)
            torch.cuda.current_stream().wait_event(ready[slot])
            o = one_step(bufs[slot])
            consumed[slot].record()
            loss_host.copy_(o["loss"].view(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()        # the trainer reads the loss every step
        return float(loss_host[0])

    e2e_loop(2)
    barrier(world)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(n_e2e)
    t1.record()
    barrier(world)
    e2e_ms = max_over_ranks(t0.elapsed_time(t1), world) / n_e2e

    if rank != 0:
        return
    from regda_b200.ops import conv as convmod
    flops = conf["flop_per_img"] * 2 * conf["batch"]
    ach = flops / (ms * 1e-3) / 1e12
    line = {
        "metric": "train images/sec (512x512, 6-class)" if cname == "P" else "train images/sec (1024x1024, 7-class)",
        "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(stage, imgs, world, use_graph, convmod.ENGINE, cname),
        "e2e": {"value": round(imgs / (e2e_ms * 1e-3), 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(calls_per_step) * args.steps,
        "roofline": top_kernel_roofline(pk),
        "clocks": clocks,
        "conv_dispatch": dict(convmod.stats),
        "loss_first": round(loss0, 4),
    }
    # the dominant kernel explains the ceiling; the whole step is what the metric is made of: both live in `roofline`
    line["roofline"]["whole_step"] = {"achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_sustained"], 4),
                                      "note": f"{flops / 1e12:.2f} TFLOP algorithmic conv work ({conf['flop_per_img'] / 1e9:.1f} GFLOP/img fwd+bwd, SURVEY.md 8d) "
                                              "per 16-image step / step time, sustained bf16 peak"}
    if world == 1 and stage == 3 and cname == "P" and not getattr(args, "no_extras", False):
        del model, step, runner, tensors, bufs, host
        torch.cuda.empty_cache()
        line["lrh"] = lrh_subrecords(pk, dev)
        line["gpu_library_baseline"] = library_baseline(dev)
    if world == 1 and not args.no_cpu and cname == "P":
        line["cpu_baseline"] = cpu_step_baseline(reps=2)
    print(json.dumps(line), flush=True)


def cpu_step_baseline(reps=2, batch=2):
    """The CPU arm: the torch fp32 port of the whole inner step (oracle/step_oracle.py, pinned to the
    reference's golden vectors) with every host thread, on a bounded sample: `batch` source +
    `batch` target 512x512 tiles (train-mode BatchNorm needs >= 2 images)."""
    from oracle import step_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = so.DeeplabOracle("resnet101", CLASS_NUM)
    images_s, label_s, images_t, soft_t, regs_t, proto = synth.step_inputs(batch, H, W, CLASS_NUM, 200, device="cpu")
    st = so.StepState(model, proto)
    so.inner_step(st, images_s, label_s, images_t, soft_t, regs_t, class_num=CLASS_NUM, lr=1e-2)
    t0 = time.perf_counter()
    for _ in range(reps):
        so.inner_step(st, images_s, label_s, images_t, soft_t, regs_t, class_num=CLASS_NUM, lr=1e-2)
    dt = (time.perf_counter() - t0) / reps
    return dict(value=round(2 * batch / dt, 3), unit="images/s", cores=cores, kind="port",
                sample=f"{batch}+{batch} tiles of 512x512 per step (of the 8+8), {reps} timed steps, torch fp32 port of the inner step",
                ms_per_step=round(dt * 1e3, 1))


def reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    r = cpu_step_baseline(reps=steps)
    print(json.dumps({
        "impl": "reference", "metric": "train images/sec (512x512, 6-class)", "value": r["value"], "unit": "images/s",
        "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as our arm (same label), timed per step on a bounded sample of it: see cpu_baseline.sample
        "config": {"workload": workload_config(3, 16, 1, False, "cpu")["workload"], "global_batch": 16, "parallelism": "host cpu",
                   "sample": r["sample"], "engine": "torch fp32 CPU port of the reference's inner step (oracle/step_oracle.py)"},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)
