"""The self-training inner step (reference tools/train_ssl_reg.py:198-241) as one object.

    step = SelfTrainingStep(model, aligner, homogenizer, ...)
    out = step(images_s, label_s, images_t, soft_t, regs_t, lr)

Per iteration, exactly what the reference loop does: two train-mode forwards (source, target:
separate calls, so BatchNorm statistics are per domain batch), label_refine -> pseudo_selection
(fused, regda_refine_select), Local Region Homogenizing, prototype EMA update, the four
cross-entropy terms (fused bilinear-upsample CE), backward, clip_grad_norm_(32), SGD(0.9, 5e-4).

B200-first choices: every parameter / gradient / momentum tensor is a view into one flat fp32
arena, so gradient all-reduce is one NCCL call and clip + SGD are three kernel launches for the
whole model; no host synchronisation anywhere inside the step (the reference has >= 3), which
also makes the step capturable in a CUDA graph (`use_cuda_graph=True`).
Data-parallel: one process per GPU, images sharded by rank, gradient arena all-reduced (mean)
and prototype sums all-reduced (sum) so every rank keeps identical weights and prototypes.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import capi, parallel
from .ops import conv as conv_ops
from .gast.balance import CrossEntropy
from .utils.tools import loss_calc


class _Range:
    """NVTX range around one phase of the step (the reference has no profiler hooks).  Host-side markers:
    visible in an nsys / ncu timeline of the eager step; inside a replayed CUDA graph the whole step is one launch."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push("regda/" + self.name)
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


def _is_channels_last_4d(p):
    return p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous()


class ParamArena:
    """All parameters of a module re-homed into one flat float32 buffer (plus gradient and
    momentum buffers of the same layout).  Physical layouts are kept (conv weights stay
    channels-last), so the flat order is the kernels' order."""

    def __init__(self, module):
        params = [p for p in module.parameters() if p.requires_grad]
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8          # segments 16-byte aligned in the bf16 shadow too (TMA base alignment)
        dev = params[0].device
        self.numel = total
        self.param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.momentum = torch.zeros(total, dtype=torch.float32, device=dev)
        self.param_bf16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)   # rewritten by the SGD kernel every step
        self.params = params
        for p, o in zip(params, offs):
            n = p.numel()

            def view(buf):
                seg = buf[o:o + n]
                if _is_channels_last_4d(p):
                    O, I, kh, kw = p.shape
                    return seg.view(O, kh, kw, I).permute(0, 3, 1, 2)
                return seg.view(p.shape)

            v = view(self.param)
            v.copy_(p.data)
            p.data = v
            p.grad = view(self.grad)
            p._bf16 = view(self.param_bf16)            # what the tcgen05 convolutions read (ops/tc.py weight_shadow)
            p._arena = self                            # writers of p.data outside the SGD kernel call p._arena.sync_shadow()
            p._arena_off = o                           # the backward ops report it to the gradient buckets (parallel.GradBuckets)
        self.lr_device = torch.zeros(1, dtype=torch.float32, device=dev)
        self.sync_shadow()
        # load_state_dict() (resume, reload-best, evaluate(ckpt)) writes the fp32 arena behind the SGD kernel's back:
        # refresh the bf16 shadow the convolutions read
        module.register_load_state_dict_post_hook(lambda _m, _incompatible: self.sync_shadow())
        self.first_step = True
        self.offset_of = {n: p._arena_off for n, p in module.named_parameters() if hasattr(p, "_arena_off")}
        self.buckets = None                            # parallel.GradBuckets when data-parallel (set by the trainer)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self):
        self.grad.zero_()

    def bucket_boundaries(self, prefixes=("layer5.", "encoder.resnet.layer4.", "encoder.resnet.layer3.")):
        """element offsets of the first parameter of each named sub-module: where the backward pass of a sequential network
        has finished everything behind it (the two PPM heads, layer5 / layer6, are parallel branches and form ONE bucket)"""
        cuts = []
        for pre in prefixes:
            offs = [o for n, o in self.offset_of.items() if n.startswith(pre)]
            if offs:
                cuts.append(min(offs))
        return cuts

    def backward_reached(self, p):
        """called by the backward ops (ops/conv.py, ops/norm.py, ops/head.py, ops/stem.py) when they enter parameter p"""
        if self.buckets is not None:
            self.buckets.reached(p._arena_off)

    def sync_shadow(self):
        """refresh the bf16 shadow after the fp32 parameters were written from outside the SGD kernel
        (load_state_dict, broadcast)"""
        self.param_bf16.copy_(self.param)
        # a plainly launched kernel after the copy: convolutions prefetch these weights ahead of their programmatic-dependent-launch
        # wait (regda_conv_hint_static_weights), which orders them only after the kernels BEFORE their immediate predecessor
        self.lr_device.add_(0.0)

    def set_lr(self, lr):
        """host float -> the device scalar the SGD kernel reads (graph-replay safe)."""
        self.lr_device.fill_(float(lr))

    def clip_and_sgd(self, max_norm=32.0, momentum=0.9, weight_decay=5e-4, grad_scale=1.0):
        """clip_grad_norm_(max_norm) + SGD step over the whole arena (train_ssl_reg.py:239-241);
        the learning rate is whatever set_lr() last wrote."""
        ws = capi.workspace.get(capi.lib().regda_sumsq_workspace_bytes(self.numel), self.param.device)
        capi.call("regda_sumsq", capi.ptr(self.grad), self.numel, capi.ptr(self._sumsq), 0, capi.ptr(ws), ws.numel(), capi.stream())
        capi.call("regda_sgd_step", capi.ptr(self.param), capi.ptr(self.grad), capi.ptr(self.momentum), capi.ptr(self.param_bf16), self.numel,
                  capi.ptr(self._sumsq), float(max_norm), float(grad_scale), 0.0, capi.ptr(self.lr_device), float(momentum),
                  float(weight_decay), int(self.first_step), capi.stream())
        self.first_step = False

    def grad_norm(self):
        return self._sumsq.sqrt()


class SelfTrainingStep:
    def __init__(self, model, aligner, homogenizer, class_num=6, ignore_label=-1, cutoff_top=0.8, cutoff_low=0.6,
                 refine_temp=2.0, sam_refine=True, refine_label=True, max_norm=32.0, momentum=0.9, weight_decay=5e-4,
                 loss_fn_s=None, loss_fn_t=None, world_size=1, use_cuda_graph=False, pair_forward=None):
        self.model, self.aligner, self.homogenizer = model, aligner, homogenizer
        self.class_num, self.ignore_label = class_num, ignore_label
        self.cutoff_top, self.cutoff_low, self.refine_temp = cutoff_top, cutoff_low, refine_temp
        self.sam_refine, self.refine_label = sam_refine, refine_label
        self.max_norm, self.momentum, self.weight_decay = max_norm, momentum, weight_decay
        self.loss_fn_s = loss_fn_s or CrossEntropy(ignore_label=ignore_label)
        self.loss_fn_t = loss_fn_t or CrossEntropy(ignore_label=ignore_label)
        self.world_size = world_size
        # paired forward (both domain batches as one tensor, per-domain BatchNorm groups) is the bf16 performance mode;
        # the float32 parity mode mirrors the reference's two model calls so that the library picks the same
        # convolution algorithms the golden vectors were produced with
        if pair_forward is None:
            pair_forward = getattr(model, "compute_dtype", None) == torch.bfloat16
        if os.environ.get("REGDA_PAIR_FORWARD") == "0":      # A/B measurements only
            pair_forward = False
        self.pair_forward = bool(pair_forward) and hasattr(model, "forward_pair")
        self.arena = ParamArena(model)
        if world_size > 1:
            # every rank trains the same weights: start from rank 0's (model construction may have been seeded per rank)
            parallel.broadcast_parameters(self.arena.param)
            self.arena.sync_shadow()
            for buf in model.buffers():
                parallel.broadcast_parameters(buf)
            # gradient all-reduce in buckets launched from inside the backward pass, ordered after the weight-gradient
            # side stream (ops/conv.py) so that the collective sees those kernels' results
            dev = self.arena.grad.device

            def order_after():
                key, side = conv_ops._wgrad_stream(dev)
                side.wait_stream(torch.cuda.current_stream())      # BatchNorm / bias gradients are produced on the main stream
                conv_ops._side_used.add(key)
                return side

            self.arena.buckets = parallel.GradBuckets(self.arena.grad, self.arena.bucket_boundaries(),
                                                      order_after if dev.type == "cuda" else None)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None
        self._lr = torch.zeros((), dtype=torch.float64)

    # ---- state that a step advances (besides the weights): what a CUDA-graph warm-up must not leave changed ----
    def persistent_state(self):
        """every tensor one step mutates in place and the next step reads: parameters (+ bf16 shadow), momentum, the
        aligner's prototypes, the BatchNorm running statistics / counters, ClassBalance frequencies"""
        ts = [self.arena.param, self.arena.param_bf16, self.arena.momentum, self.aligner.prototypes]
        ts += [b for b in self.model.buffers()]
        for fn in (self.loss_fn_s, self.loss_fn_t):
            cb = getattr(fn, "class_balancer", None)
            if cb is not None and isinstance(getattr(cb, "freq", None), torch.Tensor):
                ts.append(cb.freq)
        return ts

    def _reduce_proto(self, sums, counts):
        if self.world_size > 1:
            parallel.allreduce_sum_(sums, counts)

    def _backward(self, loss):
        """loss.backward() with the weight gradients on a side stream (joined at the end) and, data-parallel, the gradient
        arena all-reduced bucket by bucket while the backward pass is still running (sum; the mean is the SGD kernel's
        grad_scale = 1 / world)"""
        bk = self.arena.buckets
        if bk is not None:
            bk.begin()
        with conv_ops.wgrad_side_stream():
            loss.backward()
            if bk is not None:
                bk.finish()                    # remaining bucket(s) + the current stream waits for every collective

    def _step_impl(self, images_s, label_s, images_t, soft_t, regs_t):
        m = self.model
        self.arena.zero_grad()
        capi.zero_pool.reset(self.arena.param.device)      # one memset for every small accumulator of the step
        with _Range("forward"):
            if self.pair_forward and images_s.shape == images_t.shape:
                # both domain batches through the network as one tensor, BatchNorm statistics per domain (models/Encoder.py)
                # (features stay bf16: the Aligner kernels of this step read bf16 rows and nothing differentiates through them)
                (pred_s1, pred_s2, feat_s), (pred_t1, pred_t2, feat_t) = m.forward_pair(images_s, images_t, feat_dtype=torch.bfloat16)   # :210-212
            else:
                pred_s1, pred_s2, feat_s = m(images_s)                             # :210
                pred_t1, pred_t2, feat_t = m(images_t)                             # :212
        with torch.no_grad(), _Range("pseudo_labels"):
            if self.refine_label:
                hard = self.aligner.refine_select(feat_t, [pred_t1, pred_t2], soft_t, self.refine_temp,
                                                  self.cutoff_top, self.cutoff_low)   # :214-218
            else:
                from .gast.pseudo_generation import pseudo_selection
                hard = pseudo_selection(soft_t, self.cutoff_top, self.cutoff_low, 'tensor', self.ignore_label, check=False)
            if self.sam_refine:
                hard = self.homogenizer(hard, regs_t.squeeze(1))                   # :223
            self.aligner.update_prototype(feat_s, label_s, reduce_fn=self._reduce_proto)   # :225
        with _Range("losses"):
            loss_source = loss_calc([pred_s1, pred_s2], label_s, loss_fn=self.loss_fn_s, multi=True)   # :228
            loss_target = loss_calc([pred_t1, pred_t2], hard, loss_fn=self.loss_fn_t, multi=True)      # :233
            loss = loss_source + loss_target
        with _Range("backward+allreduce"):
            self._backward(loss)                                                   # :238 (+ the bucketed gradient all-reduce)
        with _Range("clip+sgd"):
            self.arena.clip_and_sgd(self.max_norm, self.momentum, self.weight_decay, 1.0 / self.world_size)   # :239-241
        capi.zero_pool.disarm()
        return loss.detach(), loss_source.detach(), loss_target.detach(), hard

    def __call__(self, images_s, label_s, images_t, soft_t, regs_t, lr):
        if not self.use_cuda_graph:
            self.arena.set_lr(lr)
            loss, ls, lt, hard = self._step_impl(images_s, label_s, images_t, soft_t, regs_t)
            return dict(loss=loss, loss_source=ls, loss_target=lt, hard=hard, grad_norm=self.arena.grad_norm())
        raise NotImplementedError("use GraphedStep for CUDA-graph replay")


class AlignStep(SelfTrainingStep):
    """Stage-2 step (tools/train_align_reg.py:144-196): the same paired forward, prototype EMA, label
    refinement, selection and LRH as stage 3, but the target soft labels come from the model's own predictions of this step
    (:165-167), there is no target segmentation loss, and both domains' features are pulled towards the prototypes by the
    prototype-contrastive loss (:186-189).  `_step_impl(images_s, label_s, images_t, regs_t)` returns
    (loss, loss_seg, loss_align, hard) -- the slots GraphedStep reports as loss / loss_source / loss_target / hard.
    align_domain=True adds the CORAL loss between the two domains' features (:187, --align-domain 1 as in the shipped
    recipe runs/regda/run_2potsdam.sh:15); its value is kept in `self.loss_domain`."""

    def __init__(self, *args, pcl_temp=8.0, align_domain=False, **kwargs):
        super().__init__(*args, **kwargs)
        from .loss import PrototypeContrastiveLoss
        self.loss_fn_pcl = PrototypeContrastiveLoss(temperature=pcl_temp, ignore_label=self.ignore_label)
        self.align_domain = bool(align_domain)
        self.loss_domain = None

    def _step_impl(self, images_s, label_s, images_t, regs_t):
        from .ops import ppm as fppm
        m = self.model
        self.arena.zero_grad()
        capi.zero_pool.reset(self.arena.param.device)
        if self.pair_forward and images_s.shape == images_t.shape:
            (pred_s1, pred_s2, feat_s), (pred_t1, pred_t2, feat_t) = m.forward_pair(images_s, images_t)   # :155, :163
        else:
            pred_s1, pred_s2, feat_s = m(images_s)
            pred_t1, pred_t2, feat_t = m(images_t)
        with torch.no_grad():
            label_s_down = self.aligner.update_prototype(feat_s, label_s, reduce_fn=self._reduce_proto)   # :158 (before the refinement)
            soft_t = fppm.upsample_softmax_mean(pred_t1.detach(), pred_t2.detach(), images_t.shape[-2:])   # :165-167
            if self.refine_label:
                hard = self.aligner.refine_select(feat_t, [pred_t1, pred_t2], soft_t, self.refine_temp, self.cutoff_top, self.cutoff_low)  # :168-171
            else:
                from .gast.pseudo_generation import pseudo_selection
                hard = pseudo_selection(soft_t, self.cutoff_top, self.cutoff_low, 'tensor', self.ignore_label, check=False)
            if self.sam_refine:
                hard = self.homogenizer(hard, regs_t.squeeze(1))                   # :176-178
            label_t = self.aligner.downscale_gt(hard)                              # :182
        loss_seg = loss_calc([pred_s1, pred_s2], label_s, loss_fn=self.loss_fn_s, multi=True)          # :186
        loss_align = (self.loss_fn_pcl(self.aligner.prototypes, feat_s, label_s_down) +
                      self.loss_fn_pcl(self.aligner.prototypes, feat_t, label_t)) * 0.5                # :188-189
        loss = loss_seg + loss_align
        if self.align_domain:                                                      # :187
            precise = getattr(m, "compute_dtype", None) == torch.float32
            loss_domain = self.aligner.align_domain(feat_s, feat_t, precise=precise)
            self.loss_domain = loss_domain.detach()
            loss = loss + loss_domain
        self._backward(loss)                                                       # :193
        self.arena.clip_and_sgd(self.max_norm, self.momentum, self.weight_decay, 1.0 / self.world_size)   # :194-196
        capi.zero_pool.disarm()
        return loss.detach(), loss_seg.detach(), loss_align.detach(), hard

    def __call__(self, images_s, label_s, images_t, regs_t, lr):
        self.arena.set_lr(lr)
        loss, lseg, lal, hard = self._step_impl(images_s, label_s, images_t, regs_t)
        return dict(loss=loss, loss_seg=lseg, loss_align=lal, loss_domain=self.loss_domain, hard=hard, grad_norm=self.arena.grad_norm())


class GraphedStep:
    """Captures SelfTrainingStep into a CUDA graph (fixed shapes).  The learning rate lives in a
    device scalar, so the schedule is followed without re-capture."""

    def __init__(self, step: SelfTrainingStep, example_inputs, lr=0.0, warmup=3):
        self.step = step
        self.static_in = [t.clone() for t in example_inputs]
        step.arena.set_lr(lr)
        # The warm-up runs real steps (allocator / autotune warm-up before capture).  Training must start from the state
        # it was given -- the reference's loop has no warm-up -- so everything a step advances is snapshotted here and
        # put back after the capture: weights, momentum, prototypes, BatchNorm running statistics, ClassBalance state.
        state = step.persistent_state()
        assert all(t.data_ptr() != 0 for t in state)
        saved = [t.clone() for t in state]
        first_step = step.arena.first_step
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                step._step_impl(*self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step._step_impl(*self.static_in)
        with torch.no_grad():
            for t, v in zip(step.persistent_state(), saved):
                t.copy_(v)
        if first_step:
            # the captured SGD launch has first_step = 0 baked in (momentum * buf + g); with a zero buffer that IS the
            # first-step rule buf = g, so a fresh run replays correctly from the very first step
            step.arena.momentum.zero_()
        torch.cuda.synchronize()

    def __call__(self, *inputs, lr=None):
        if lr is not None:
            self.step.arena.set_lr(lr)
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        loss, ls, lt, hard = self.out
        return dict(loss=loss, loss_source=ls, loss_target=lt, hard=hard, grad_norm=self.step.arena.grad_norm())
