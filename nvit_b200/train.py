"""The reference's train-step contract (/root/reference/nvit/train.py:885-993) on the sm_100a engine.

    forward (bf16 GEMMs, fp32 residual) -> cross-entropy -> backward -> [data-parallel gradient all-reduce, overlapped]
    -> clip_grad_norm_(1.0) -> AdamW(betas 0.9/0.95, wd 0.1) -> zero_grad -> normalize_matrices

``Trainer`` drives the engine directly (no autograd graph): gradients land in the engine's flat fp32 buffer, the
gradient norm, clip, AdamW update and the weight normalisation are four launches over contiguous memory, and in
data-parallel runs each transformer block's gradient slice is all-reduced over NCCL on a side stream as soon as that
block's backward has been issued (train.py:434-446 intends DDP; SURVEY.md 2.3 #3 explains why the reference never
actually reduces — the intended semantics, mean of gradients over ranks, are what is built here).

Deviation kept on purpose: the reference wraps bf16 training in a GradScaler (train.py:135-136); its power-of-two
scale/unscale is the identity for finite gradients, so it is omitted.
"""
from __future__ import annotations

import torch

from . import ops

F32 = torch.float32


class GradReducer:
    """Bucketed all-reduce (sum) of ranges of one flat gradient buffer, issued as ranges become final.

    Device-agnostic host logic (tested with gloo on CPU); on CUDA the collectives run on `comm_stream` so they overlap
    the rest of the backward pass.  Ranges smaller than `min_bucket` elements are coalesced with their neighbours and
    flushed at `finish()`.
    """

    def __init__(self, flat: torch.Tensor, group=None, min_bucket: int = 1 << 20, bucket_elems: int = 0,
                 reserve_sms: int = 0, reserve_calls: int = 0):
        """`bucket_elems` > 0 merges adjacent large ranges (consecutive transformer blocks arrive back to front) until a
        bucket holds at least that many elements: fewer, longer collectives beside the backward pass."""
        import torch.distributed as dist
        self.dist = dist
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.min_bucket = min_bucket
        self.bucket_elems = bucket_elems
        self.acc: list[int] | None = None          # [lo, hi) of the bucket being filled
        self.pending = []
        self.small: list[tuple[int, int]] = []
        self.comm_stream = torch.cuda.Stream() if flat.is_cuda else None
        self.reduced_elems = 0
        # `reserve_sms` > 0: the `reserve_calls` kernel launches that follow a bucket's all-reduce leave that many SMs free
        # (nvit_set_sm_budget for a window of launches): NCCL's CTAs cannot share an SM with a 227 KB persistent CTA, and a
        # collective that has to wait for SMs costs the kernel beside it a wave.  MEASURED (N = 2): slower than
        # plain overlap, which is slower than one all-reduce after backward; off by default.
        self.reserve_sms, self.reserve_calls = reserve_sms, reserve_calls
        self.total_sms = torch.cuda.get_device_properties(flat.device).multi_processor_count if flat.is_cuda else 0

    def _issue(self, lo: int, hi: int):
        if hi <= lo or self.world == 1:
            return
        view = self.flat[lo:hi]
        self.reduced_elems += hi - lo
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                self.pending.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
            if self.reserve_sms > 0 and self.reserve_calls > 0:
                from . import _lib
                _lib.sm_budget_window(self.reserve_calls, self.total_sms - self.reserve_sms)
        else:
            self.pending.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def ready(self, lo: int, hi: int):
        if hi - lo >= self.min_bucket:
            if self.bucket_elems <= 0:
                self._issue(lo, hi)
                return
            if self.acc is not None and (hi == self.acc[0] or lo == self.acc[1]):
                self.acc = [min(lo, self.acc[0]), max(hi, self.acc[1])]
            else:
                self._flush_acc()
                self.acc = [lo, hi]
            if self.acc[1] - self.acc[0] >= self.bucket_elems:
                self._flush_acc()
        elif hi > lo:
            self.small.append((lo, hi))

    def _flush_acc(self):
        if self.acc is not None:
            self._issue(*self.acc)
            self.acc = None

    def finish(self):
        """Flush coalesced small ranges and make the current stream wait for every collective."""
        self._flush_acc()
        self.small.sort()
        merged: list[list[int]] = []
        for lo, hi in self.small:
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        for lo, hi in merged:
            self._issue(lo, hi)
        self.small = []
        for w in self.pending:
            w.wait()
        self.pending = []


class DeviceLoader:
    """Host -> device double buffering for the training loop's batches (the reference does
    ``X.pin_memory().to(device, non_blocking=True)`` on the compute stream for every batch, train.py:886-890).

    Wraps any iterable of ``(X, y)`` CPU tensors — float images ``[B, C, S, S]`` as the reference's DataLoader yields them,
    or raw uint8 ``[B, S, S, C]`` images, which ``ViT.forward`` / ``Trainer.step`` normalise on the fly (4x fewer PCIe
    bytes).  Batch i+1 is staged in pinned memory and copied on a side stream while step i computes; a slot is reused
    only after the step that read it has been enqueued.
    """

    def __init__(self, batches, device, depth: int = 2, transform=None):
        """transform: optional callable applied ON THE DEVICE to each uint8 batch as it is handed out (nvit_b200.augment's
        AutoAugment: the reference augments in its DataLoader workers instead, train.py:1081-1092)."""
        self.transform = transform
        self.batches, self.device, self.depth = batches, torch.device(device), max(2, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * self.depth

    def _stage(self, i, X, y):
        slot = self.slots[i % self.depth]
        if slot is None or slot["hx"].shape != X.shape or slot["hx"].dtype != X.dtype or slot["hy"].shape != y.shape:
            slot = {"hx": torch.empty(X.shape, dtype=X.dtype).pin_memory(), "hy": torch.empty(y.shape, dtype=y.dtype).pin_memory(),
                    "dx": torch.empty(X.shape, dtype=X.dtype, device=self.device), "dy": torch.empty(y.shape, dtype=y.dtype, device=self.device),
                    "ready": torch.cuda.Event(), "free": None}
            self.slots[i % self.depth] = slot
        slot["ready"].synchronize()               # the previous copy out of the pinned staging buffers has finished
        slot["hx"].copy_(X)
        slot["hy"].copy_(y)
        with torch.cuda.stream(self.stream):
            if slot["free"] is not None:
                self.stream.wait_event(slot["free"])      # the step that consumed this slot's device buffers is enqueued
            slot["dx"].copy_(slot["hx"], non_blocking=True)
            slot["dy"].copy_(slot["hy"], non_blocking=True)
            slot["ready"].record(self.stream)
        return slot

    def __iter__(self):
        pending = None
        for i, (X, y) in enumerate(self.batches):
            slot = self._stage(i, X, y)          # batch i goes out while the consumer still works on batch i-1
            if pending is not None:
                yield self._hand_out(pending)
                self._release(pending)           # back from the consumer: its work on `pending` is enqueued
            pending = slot
        if pending is not None:
            yield self._hand_out(pending)

    def _hand_out(self, slot):
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        if self.transform is not None:
            return self.transform(slot["dx"]), slot["dy"]      # a new tensor on the compute stream; the slot is reusable as before
        return slot["dx"], slot["dy"]

    def _release(self, slot):
        slot["free"] = torch.cuda.Event()
        slot["free"].record(torch.cuda.current_stream(self.device))


class Trainer:
    """One rank of the (optionally data-parallel) nViT training loop; mirrors Trainer.train's inner step."""

    def __init__(self, model, learning_rate: float = 1e-3, betas=(0.9, 0.95), weight_decay: float = 0.1, grad_clip: float = 1.0,
                 eps: float = 1e-8, gradient_accumulation_steps: int = 1, process_group=None, data_parallel: bool | None = None,
                 cuda_graph: bool = False, graph_warmup_steps: int = 2, overlap_allreduce: bool = False, sm_budget: int = 0,
                 overlap_reserve_sms: int = 0, overlap_reserve_calls: int = 0,
                 bucket_blocks: int = 1, fused_tail: bool = True,
                 consistency_weight: float = 0.1, smoothness_weight: float = 0.1):
        import torch.distributed as dist
        self.model = model
        self.engine = model.engine
        self.lr, self.betas, self.wd, self.clip, self.eps = learning_rate, betas, weight_decay, grad_clip, eps
        self.grad_accum = gradient_accumulation_steps
        # weights of the Kohonen losses (settings.training.*, train.py:909-926); the other three come from the model config
        self.consistency_weight, self.smoothness_weight = consistency_weight, smoothness_weight
        self.iter_num = 0
        self.opt_step = 0
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.dp = data_parallel
        self.group = process_group
        self.world = dist.get_world_size(process_group) if self.dp else 1
        self._state_for = None
        self.reducer = None
        self.launches = 0
        # Data parallel: ONE all-reduce of the flat gradient buffer after the backward pass by default.  MEASURED at N = 8
        # (profiles/r01_summary.md): bucketed collectives overlapped with backward are SLOWER (53.0 vs 51.7 ms/step),
        # because NCCL's CTAs cannot share an SM with the persistent 227 KB GEMM CTAs and every bucket costs the GEMM
        # running beside it a wave; overlap_allreduce=True keeps the bucketed form available.
        self.overlap = overlap_allreduce
        self.overlap_reserve = (int(overlap_reserve_sms), int(overlap_reserve_calls))
        self.bucket_blocks = max(1, bucket_blocks)   # transformer blocks per overlapped bucket
        # clip + AdamW + normalize_matrices + bf16 operand emit + zero_grad as one launch (nvit_adamw_norm_fused);
        # False keeps the separate kernels (sumsq, adamw_flat, weight_norm_multi, cast, memset)
        self.fused_tail = fused_tail
        if self.dp and sm_budget:
            from . import _lib
            _lib.call("nvit_set_sm_budget", int(sm_budget))
        # CUDA-graph replay of the whole step - forward, backward, the NCCL all-reduce (data parallel) and the optimizer
        # tail: the ~270 launches, their tensor-map encodes and the Python between them are captured once; learning rate
        # and step count then live in device memory (self.hyper).  The Kohonen learning-rate schedule is host-side state,
        # so that mode runs eagerly.
        self.use_graph = bool(cuda_graph) and not model.config.use_kohonen
        self.graph_warmup_steps = max(1, graph_warmup_steps)
        self._graph = None
        self._graph_inputs = None
        self._graph_launches = 0
        self._graph_B = None
        self.replays = 0

    # ---- optimizer state lives beside the flat parameter buffer
    def _ensure_state(self):
        eng = self.engine
        eng.param_list()
        if self._state_for is not eng.P32:
            old_m, old_v = getattr(self, "m", None), getattr(self, "v", None)
            self.m = torch.zeros_like(eng.P32)
            self.v = torch.zeros_like(eng.P32)
            if old_m is not None and self.opt_step > 0:
                # the engine re-materialised its flat buffers (model.to(...) with a real move / cast): the moments follow
                # when the layout is the same, otherwise say so instead of silently restarting AdamW at step opt_step
                if old_m.numel() == self.m.numel():
                    self.m.copy_(old_m)
                    self.v.copy_(old_v)
                else:
                    import warnings
                    warnings.warn("nvit_b200.Trainer: the parameter layout changed under a running optimizer; AdamW moments "
                                  "were reset (load_optimizer_state_dict restores them)")
            # [0] sum of squared gradients (clip), [1] unit counter of the fused tail (uint32 bits); zeroed together
            self.tail_scratch = torch.zeros(4, device=eng.P32.device, dtype=F32)
            self.gnorm = self.tail_scratch[0:1]
            self.unit_counter = self.tail_scratch[1:2].view(torch.int32)
            self.sumsq_ws = torch.zeros(1024, device=eng.P32.device, dtype=F32)     # nvit_sumsq_f32_det scratch ([0] = ticket)
            self.loss_buf = torch.zeros(1, device=eng.P32.device, dtype=F32)
            self.hyper = torch.tensor([self.lr, float(self.opt_step)], device=eng.P32.device, dtype=F32)   # {lr, step}
            self._state_for = eng.P32
            self._drop_graph()
            cfg = self.model.config
            if cfg.use_kohonen:
                k = 1.0 / (self.grad_accum * self.world)
                eng.set_aux_weights(reconstruction=cfg.reconstruction_weight * k, consistency=self.consistency_weight * k,
                                    local_quantization=cfg.local_quantization_weight * k,
                                    global_quantization=cfg.global_quantization_weight * k, smoothness=self.smoothness_weight * k)
            if self.dp:
                lo, hi = eng.block_grad_range(0)
                self.reducer = GradReducer(eng.G32, self.group,
                                           bucket_elems=(hi - lo) * self.bucket_blocks if self.bucket_blocks > 1 else 0,
                                           reserve_sms=self.overlap_reserve[0] if self.overlap else 0,
                                           reserve_calls=self.overlap_reserve[1] if self.overlap else 0)
                self._broadcast_params()

    def _drop_graph(self):
        if self._graph is not None and self._graph_B is not None:
            self.engine._pinned.discard(self._graph_B)
        self._graph = None
        self._graph_B = None

    def _broadcast_params(self):
        import torch.distributed as dist
        dist.broadcast(self.engine.P32, src=0, group=self.group)
        self.engine.invalidate_operands()

    def normalize_matrices(self):
        """Trainer.normalize_matrices (train.py:461-480): one multi-tensor launch."""
        if self.model.config.use_nvit:
            self.engine.normalize_matrices()

    def micro_step(self, X: torch.Tensor, y: torch.Tensor, last: bool = True):
        """forward + loss + backward of one micro-batch; gradients accumulate in the flat buffer (train.py:898-933)."""
        eng = self.engine
        self._ensure_state()
        if y.dim() != 1 or y.shape[0] != X.shape[0]:
            raise ValueError(f"labels must have shape ({X.shape[0]},), got {tuple(y.shape)}")
        if y.dtype != torch.int64 or not y.is_contiguous():
            y = y.long().contiguous()             # the kernel reads int64 class indices (F.cross_entropy's target type)
        if self.model.training:
            self.model.step += 1               # ViT.forward's own counter (model.py:404-405): drives the Kohonen schedule
        logits, recon = eng.forward(X, save=True)
        B, N = logits.shape
        dlogits = eng._acts[B].setdefault("dlogits", torch.empty(B, N, device=logits.device, dtype=F32))
        # mean CE over the batch, scaled for accumulation and for the mean over data-parallel ranks
        ops.cross_entropy(logits, y, self.loss_buf, dlogits, 1.0 / (self.grad_accum * self.world))
        self.launches += 1
        eng.grad_ready_hook = self.reducer.ready if (self.dp and last and self.overlap) else None
        eng.backward(dlogits)
        self.last_recon = recon
        self.last_aux = eng.last_aux
        return logits

    def _all_reduce_grads(self):
        """Mean of the gradients over the data-parallel ranks (the 1/N is folded into the loss scale): what DDP is meant
        to do at train.py:434-446, 899-902."""
        eng = self.engine
        if self.overlap:
            self.reducer.finish()             # the buckets went out behind the backward pass; flush the rest and wait
            return
        import torch.distributed as dist
        # one collective over the whole flat buffer on the compute stream's order (capturable in a CUDA graph)
        dist.all_reduce(eng.G32[:eng.n_active], op=dist.ReduceOp.SUM, group=self.group)
        self.reducer.reduced_elems += eng.n_active

    def optimizer_step(self):
        """clip -> AdamW -> zero_grad -> normalize_matrices (train.py:935-946, 989-990)."""
        eng = self.engine
        if self.dp:
            self._all_reduce_grads()
        self.opt_step += 1
        self.hyper[1:2].add_(1.0)          # device-side step count (what a replayed graph advances)
        na = eng.n_active
        gn = None
        self.tail_scratch.zero_()
        if self.clip and self.clip > 0:
            # reproducible sum: replicas must derive the SAME clip coefficient from their identical all-reduced gradients
            ops.sumsq_det(eng.G32[:na], self.gnorm, self.sumsq_ws)
            gn = self.gnorm
            self.launches += 1
        if self.fused_tail:
            eng.fused_tail(self.m, self.v, self.lr, self.betas, self.eps, self.wd, self.opt_step, self.unit_counter, gnorm_sq=gn,
                           max_norm=self.clip or 0.0, dev_lr_step=self.hyper)
        else:
            ops.adamw_flat(eng.P32[:na], eng.G32[:na], self.m[:na], self.v[:na], eng.n_decay, self.lr, self.betas[0], self.betas[1],
                           self.eps, self.wd, self.opt_step, gn, self.clip or 0.0, dev_lr_step=self.hyper)
            self.launches += 1
            eng.invalidate_operands()
            eng.zero_grad()
            self.normalize_matrices()
        if self.dp and self.model.config.use_kohonen:
            # the in-forward map update (kohonen.py:121-165) saw different images on every rank: average the node tables so
            # that the replicas stay identical (the reference's DDP would leave them diverged)
            import torch.distributed as dist
            for tag in ("local", "global"):
                nodes = eng.p(tag + "_kohonen.nodes")
                dist.all_reduce(nodes, op=dist.ReduceOp.SUM, group=self.group)
                nodes.mul_(1.0 / self.world)

    def set_lr(self, lr: float):
        """Learning-rate schedule hook (train.py:874-876 sets it per epoch); safe with a captured graph."""
        self.lr = lr
        if self._state_for is not None:
            self.hyper[0:1].fill_(lr)

    def _step_eager(self, X, y):
        self.loss_buf.zero_()
        for k in range(self.grad_accum):
            self.micro_step(X, y, last=(k == self.grad_accum - 1))
        if self.grad_accum > 1:
            self.loss_buf.mul_(1.0 / self.grad_accum)      # mean over the micro-steps (each adds its mean CE)
        self.optimizer_step()
        return self.loss_buf

    def input_buffers(self, like_X: torch.Tensor, like_y: torch.Tensor):
        """Static device buffers the captured graph reads; fill them in place (e.g. H2D copies) to avoid a staging copy."""
        if (self._graph_inputs is None or self._graph_inputs[0].shape != like_X.shape or self._graph_inputs[0].dtype != like_X.dtype
                or self._graph_inputs[1].dtype != like_y.dtype):
            self._graph_inputs = (torch.empty_like(like_X, device=self.engine.P32.device),
                                  torch.empty_like(like_y, device=self.engine.P32.device))
            self._drop_graph()
        return self._graph_inputs

    def step(self, X: torch.Tensor, y: torch.Tensor):
        """One full training iteration on one (micro-)batch; returns the device scalar of the mean CE loss."""
        self._ensure_state()
        self.iter_num += 1
        if not self.use_graph or self.iter_num <= self.graph_warmup_steps:
            return self._step_eager(X, y)
        Xs, ys = self.input_buffers(X, y)
        if X.data_ptr() != Xs.data_ptr():
            Xs.copy_(X, non_blocking=True)
            ys.copy_(y, non_blocking=True)
        if self._graph is None:
            eng = self.engine
            before = self.total_launches
            host_step = self.opt_step
            # The graph holds raw pointers into this batch size's activation set: pin it, so that forwards at other batch
            # sizes between replays (an eval loader, a partial last batch) can never hand its memory back to the allocator.
            B = X.shape[0]
            eng._buffers(B)
            eng._pinned.add(B)
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._step_eager(Xs, ys)
            except Exception as e:           # e.g. a collective that cannot be captured: keep training, eagerly, and say so
                import warnings
                warnings.warn(f"nvit_b200.Trainer: CUDA-graph capture of the step failed ({type(e).__name__}: {e}); running eagerly")
                eng._pinned.discard(B)
                self.use_graph = False
                self.opt_step = host_step
                self.hyper[1:2].fill_(float(host_step))
                self.launches -= self.total_launches - before
                eng.zero_grad()
                eng.invalidate_operands()
                return self._step_eager(X, y)
            self.opt_step = host_step           # capture does not execute; the replay below is the real step
            self.hyper[1:2].fill_(float(host_step))
            self._graph_launches = self.total_launches - before
            self.launches -= self._graph_launches   # count launches when they run, i.e. per replay
            self._graph = g
            self._graph_B = B
            # the capture recorded the operand-refresh decision of ITS first forward; replays always find the operands the
            # previous replay's fused tail wrote, but the first replay may not: refresh them by hand once
            eng.refresh_operands(force=True)
        self._graph.replay()
        self.opt_step += 1
        self.replays += 1
        self.launches += self._graph_launches
        # what the replay left behind: operands written by the fused tail of the captured step (when it is in use)
        self.engine._p16_version = (tuple(s.param._version for s in self.engine.slots.values()) if self.fused_tail else None)
        return self.loss_buf

    # ---- checkpoint interchange (train.py:629-655 saves {"model", "optimizer", "model_args", "iter_num", ...})
    def _torch_optimizer(self):
        dev = self.engine.P32.device.type
        return self.model.configure_optimizers(self.wd, self.lr, self.betas, dev)

    def optimizer_state_dict(self) -> dict:
        """The flat AdamW moments in torch.optim.AdamW's state_dict layout over ViT.configure_optimizers' groups, i.e. what
        the reference stores under checkpoint["optimizer"] (train.py:641).  Parameters that never receive a gradient
        have no entry, as in torch."""
        self._ensure_state()
        eng, opt = self.engine, self._torch_optimizer()
        by_param = {id(s.param): s for s in eng.slots.values()}
        if self.opt_step > 0:
            for group in opt.param_groups:
                for prm in group["params"]:
                    s = by_param[id(prm)]
                    if s.off >= eng.n_active:
                        continue
                    opt.state[prm] = {"step": torch.tensor(float(self.opt_step), device=prm.device),
                                      "exp_avg": self.m[s.off:s.off + s.numel].view(s.shape).clone(),
                                      "exp_avg_sq": self.v[s.off:s.off + s.numel].view(s.shape).clone()}
        return opt.state_dict()

    def load_optimizer_state_dict(self, state: dict) -> None:
        """Inverse of optimizer_state_dict: accepts the reference's checkpoint["optimizer"] (train.py:380)."""
        self._ensure_state()
        eng, opt = self.engine, self._torch_optimizer()
        opt.load_state_dict(state)
        by_param = {id(s.param): s for s in eng.slots.values()}
        self.m.zero_()
        self.v.zero_()
        step = 0
        for prm, st in opt.state.items():
            s = by_param[id(prm)]
            self.m[s.off:s.off + s.numel].copy_(st["exp_avg"].reshape(-1))
            self.v[s.off:s.off + s.numel].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        group = opt.param_groups[0]
        self.lr, self.betas, self.eps = group["lr"], tuple(group["betas"]), group["eps"]
        self.opt_step = step
        self.hyper.copy_(torch.tensor([self.lr, float(step)], dtype=F32))
        self._drop_graph()

    @property
    def total_launches(self) -> int:
        return self.launches + self.engine.launches
