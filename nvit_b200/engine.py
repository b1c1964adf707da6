"""Hand-scheduled forward and backward pass of the nViT training hot path over the C-ABI kernels.

Restates, kernel by kernel, ``ViT.forward`` (/root/reference/nvit/model.py:403-470), ``CrossAttentionBlock.forward``
(:219-275), ``Block.forward`` (:92-169) and ``Block.norm_skip`` (:84-87), plus the backward that autograd derives
from them.  PyTorch provides device memory, streams and the parameter objects; all arithmetic is in
``libnvit_b200.so``.  Precision follows the reference under ``torch.autocast(bf16)``:
GEMM and attention operands bf16 with fp32 accumulation, norms / residual stream / parameters / gradients fp32.

Memory layout (HBM):
  * ``P32``  one flat fp32 buffer holding every parameter; ``model.parameters()`` are views into it.  Order:
    [GEMM weights in GEMM-operand order | other weight-decayed params | no-decay params || params that never get a
    gradient].  AdamW, the gradient norm and the data-parallel all-reduce stream over contiguous ranges of it.
  * ``G32``  the gradients in the same layout (the autograd path hands out clones of its slices as ``p.grad``).
  * ``W16``  bf16 GEMM operands, one cast launch from the head of ``P32`` (q/k/v concatenated to [3C, C], etc).
  * per-layer activations saved for backward (bf16 GEMM operands, fp32 residual stream), allocated once per batch size.
"""
from __future__ import annotations

import math
import os

import torch

from . import ops

F32 = torch.float32
BF16 = torch.bfloat16


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class _Slot:
    __slots__ = ("name", "param", "off", "numel", "shape")

    def __init__(self, name, param, off):
        self.name, self.param, self.off = name, param, off
        self.numel, self.shape = param.numel(), tuple(param.shape)


class Engine:
    def __init__(self, model):
        self.model = model
        self.cfg = model.config
        self.P32 = None
        self._acts = {}                 # batch size -> activation set (insertion order = least recently used first)
        self._pinned = set()            # batch sizes a captured CUDA graph points into: never evicted
        self.max_resident_batches = 2   # activation sets kept resident (train + eval / last partial batch)
        self.acts_epoch = 0             # bumped whenever an activation set is dropped or rebuilt
        self._saved = None
        self.grad_ready_hook = None     # callable(lo, hi): flat-gradient range [lo, hi) is final (data-parallel overlap)
        self.launches = 0
        self._p16_version = None
        self.last_aux = {}              # Kohonen-map losses / units of the last forward (empty without the maps)
        self.probe = None               # list to receive (start, end) CUDA event pairs around the c_fc GEMM (bench.py)
        # SiLU-gate backward inside the dgrad GEMM epilogue (nvit_gemm_gate_bwd); NVIT_FUSE_GATE_BWD=0 keeps the two-kernel path
        self.fuse_gate_bwd = os.environ.get("NVIT_FUSE_GATE_BWD", "1") != "0"
        # q/k normalised and sqk-scaled in the projection GEMM's epilogue (nvit_gemm_qknorm) instead of inside the attention
        # kernels; NVIT_FUSE_QKNORM=0 keeps the in-kernel normalisation
        self.fuse_qknorm = os.environ.get("NVIT_FUSE_QKNORM", "1") != "0"

    # ------------------------------------------------------------------------------------------ parameter layout
    def invalidate(self):
        self.P32 = None

    def _named(self):
        return dict(self.model.named_parameters())

    def _build_layout(self):
        m, cfg = self.model, self.cfg
        named = self._named()
        C, L = cfg.n_embd, cfg.n_layer
        order: list[str] = []
        # region A: GEMM weights, in the order/concatenation the bf16 operand buffer uses
        order += ["local_patch_embed.weight", "global_patch_embed.1.weight"]
        order += ["cross_attention.q_local.weight", "cross_attention.k_global.weight", "cross_attention.v_global.weight",
                  "cross_attention.proj.weight", "cross_attention.out_proj.weight"]
        for i in range(L):
            b = f"transformer.h.{i}."
            order += [b + "query.weight", b + "key.weight", b + "value.weight", b + "att_c_proj.weight", b + "c_fc.weight",
                      b + "mlp_c_proj.weight"]
        order += ["mlp_head.1.weight"]
        # the reconstruction loss joins the objective only with the Kohonen maps (train.py:909-926); then its head trains
        self.rec_active = bool(cfg.use_kohonen)
        if self.rec_active:
            order += ["reconstruction_head.0.weight"]
        n_gemm = len(order)
        # region A': remaining weight-decayed parameters (dim >= 2)
        inactive = {n for n in named if (cfg.use_nvit and ".rmsnorm_" in n) or n == "map_balance"
                    or (n.startswith("reconstruction_head.") and not self.rec_active)}
        rest = [n for n in named if n not in order and n not in inactive]
        order += [n for n in rest if named[n].dim() >= 2 and "sz" not in n]
        n_decay_names = len(order)
        # region B: no-decay parameters; biases of fused GEMMs kept adjacent (q,k,v) so one column-sum serves them
        nodecay = [n for n in rest if not (named[n].dim() >= 2 and "sz" not in n)]

        def key(n):
            for i, suffix in enumerate(("query.bias", "key.bias", "value.bias")):
                if n.endswith(suffix):
                    return (n.rsplit(".", 2)[0], 0, i)
            for i, suffix in enumerate(("k_global.bias", "v_global.bias")):
                if n.endswith(suffix):
                    return ("cross_attention", 0, i)
            return (n, 1, 0)
        order += sorted(nodecay, key=key)
        n_active_names = len(order)
        # region C: never receive a gradient: the reconstruction head first (it is a GEMM operand)
        if not self.rec_active:
            order += ["reconstruction_head.0.weight", "reconstruction_head.0.bias"]
        order += sorted(n for n in inactive if not n.startswith("reconstruction_head."))
        assert len(order) == len(named) and set(order) == set(named), "parameter layout does not cover the model"

        slots, off = {}, 0
        marks = {}
        for idx, n in enumerate(order):
            if idx == n_gemm:
                marks["gemm_end"] = off
            if idx == n_decay_names:
                marks["decay_end"] = off
            if idx == n_active_names:
                marks["active_end"] = off
            s = _Slot(n, named[n], off)
            slots[n] = s
            off = _align(off + s.numel, 8)
        marks.setdefault("gemm_end", off)
        marks.setdefault("decay_end", off)
        marks.setdefault("active_end", off)
        self.slots, self.order, self.n_total = slots, order, off
        self.n_gemm, self.n_decay, self.n_active = marks["gemm_end"], marks["decay_end"], marks["active_end"]

    def _materialize(self, device):
        """(Re)build the flat buffers on `device` and alias every parameter (and its grad) into them."""
        self._build_layout()
        self.device = device
        P32 = torch.zeros(self.n_total, device=device, dtype=F32)
        self.G32 = torch.zeros(self.n_total, device=device, dtype=F32)
        with torch.no_grad():
            for s in self.slots.values():
                view = P32[s.off:s.off + s.numel].view(s.shape)
                view.copy_(s.param.data.to(device=device, dtype=F32))
                s.param.data = view
        self.P32 = P32
        # bf16 GEMM operands: region A + the reconstruction weight
        rec = self.slots["reconstruction_head.0.weight"]
        self.W16 = torch.empty(self.n_gemm + (0 if self.rec_active else _align(rec.numel, 8)), device=device, dtype=BF16)
        self._rec16_off = rec.off if self.rec_active else self.n_gemm
        self._p16_version = None
        self.tail_table = None
        self._acts = {}
        self._pinned = set()
        self.acts_epoch += 1
        self._build_norm_table()
        self.scratch = torch.zeros(8, device=device, dtype=F32)   # [0] recon loss, [1:4] map pair-loss sums, [4] smoothness
        # weights of the auxiliary losses times the incoming gradient: [0] reconstruction, [1] consistency,
        # [2] local quantization, [3] global quantization, [4] smoothness; [5] Kohonen lr * alpha (device-side for graphs)
        self.aux_w = torch.zeros(8, device=device, dtype=F32)
        self._ptrs = {n: s.param.data_ptr() for n, s in self.slots.items()}

    def _check_alias(self, device):
        if self.P32 is None or self.P32.device != device:
            self._materialize(device)
            return
        for n, s in self.slots.items():
            if s.param.data_ptr() != self._ptrs[n]:
                self._materialize(device)
                return

    def param_list(self):
        if self.P32 is None:
            dev = next(self.model.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("nvit_b200.ViT must live on a CUDA device (model.cuda()) before forward")
            self._materialize(dev)
        return [self.slots[n].param for n in self.order]

    def p(self, name):       # fp32 parameter view
        s = self.slots[name]
        return self.P32[s.off:s.off + s.numel].view(s.shape)

    def g(self, name):       # fp32 gradient view
        s = self.slots[name]
        return self.G32[s.off:s.off + s.numel].view(s.shape)

    def w16(self, name, rows=None):
        """bf16 operand of a GEMM weight as [rows, cols]; `rows` > own rows spans the following concatenated weights."""
        s = self.slots[name]
        off = s.off if name != "reconstruction_head.0.weight" else self._rec16_off
        r = s.shape[0] if rows is None else rows
        cols = s.numel // s.shape[0]
        return self.W16[off:off + r * cols].view(r, cols)

    def g2d(self, name, rows=None):
        s = self.slots[name]
        r = s.shape[0] if rows is None else rows
        cols = s.numel // s.shape[0]
        return self.G32[s.off:s.off + r * cols].view(r, cols)

    def block_grad_range(self, i):
        b = f"transformer.h.{i}."
        lo = self.slots[b + "query.weight"].off
        last = self.slots[b + "mlp_c_proj.weight"]
        return lo, last.off + last.numel

    def _build_norm_table(self, columns_first: bool = True):
        """Device table for nvit_weight_norm_multi (Trainer.normalize_matrices, train.py:461-480)."""
        rows, first = [], 0
        # Units are handed to CTAs in table order.  A column unit (axis 0: 128 columns x all rows, two passes) runs ~15x
        # longer than a row unit (8 rows), so the column-normalised matrices come first: their units start at t = 0 and
        # the short row units fill in behind them, instead of a handful of long units forming the tail of the launch.
        for want in ((0, 1) if columns_first else (None,)):      # None: plain block order (scripts/wnorm_bench.py A/B)
            for i in range(self.cfg.n_layer if self.cfg.use_nvit else 0):
                b = f"transformer.h.{i}."
                for nm, axis in (("query", 1), ("key", 1), ("value", 1), ("att_c_proj", 0), ("c_fc", 1), ("mlp_c_proj", 0)):
                    if want is not None and axis != want:
                        continue
                    s = self.slots[b + nm + ".weight"]
                    r, c = s.shape
                    rows.append([self.P32.data_ptr() + 4 * s.off, 0, r, c, axis, first])
                    first += (r + 7) // 8 if axis == 1 else (c + 127) // 128
        self.norm_table = torch.tensor(rows if rows else [[0] * 6], dtype=torch.int64, device=self.device)
        self.norm_units = first

    def normalize_matrices(self):
        if not self.cfg.use_nvit:        # train.py:463-464: only in nViT mode
            return
        ops.weight_norm_multi(self.norm_table, self.norm_table.shape[0], self.norm_units)
        self.launches += 1
        self._p16_version = None

    def zero_grad(self):
        self.G32.zero_()

    def refresh_operands(self, force=False):
        """autocast's weight casts: fp32 master -> bf16 GEMM operands, one launch for region A, one for the recon head.

        The bf16 operands are rebuilt on EVERY forward except the first one after the engine itself produced them
        together with the current fp32 values: only `fused_tail` (the Trainer's optimizer step, which writes fp32 and
        bf16 in the same pass) marks them current, and the mark is spent by the next forward.  Writes the engine cannot see - an external optimizer, the reference's normalize_matrices going
        through `W.data.copy_` (train.py:474-480; `.data` writes do not bump the tensor version) - are therefore always
        picked up by the next forward; the cost is two streaming launches (~0.2 ms at B/16)."""
        ver = tuple(s.param._version for s in self.slots.values())
        if not force and self._p16_version is not None and ver == self._p16_version:
            self._p16_version = None      # the mark covers exactly one forward: the one right after the fused tail
            return
        ops.cast_bf16(self.P32[:self.n_gemm], self.W16[:self.n_gemm])
        self.launches += 1
        if not self.rec_active:
            rec = self.slots["reconstruction_head.0.weight"]
            ops.cast_bf16(self.P32[rec.off:rec.off + rec.numel], self.W16[self._rec16_off:self._rec16_off + rec.numel])
            self.launches += 1
        self._p16_version = None          # valid for this forward only

    def invalidate_operands(self):
        """Force the next forward to rebuild the bf16 operands (call after writing weights behind the Trainer's back)."""
        self._p16_version = None

    # ------------------------------------------------------------------------------------------ fused optimizer tail
    def _build_tail_table(self):
        """Segment table of nvit_adamw_norm_fused: one entry per trained tensor, the column-normalised matrices first
        (their units move ~20x the bytes of a row unit), then the row-normalised ones, then everything else."""
        cfg = self.cfg
        normed = {}
        if cfg.use_nvit:
            for i in range(cfg.n_layer):
                b = f"transformer.h.{i}."
                for nm, axis in (("query", 1), ("key", 1), ("value", 1), ("att_c_proj", 0), ("c_fc", 1), ("mlp_c_proj", 0)):
                    normed[b + nm + ".weight"] = axis
        segs = {2: [], 1: [], 0: []}
        for n in self.order:
            s = self.slots[n]
            if s.off >= self.n_active:
                continue
            decay = 1 if s.off < self.n_decay else 0
            w16_off = s.off if s.off < self.n_gemm else -1
            if n in normed:
                r, c = s.shape
                kind = 1 if normed[n] == 1 else 2
                units = (r + 31) // 32 if kind == 1 else (c + 127) // 128
                segs[kind].append([s.off, r, c, kind, w16_off, decay, units])
            else:
                segs[0].append([s.off, 1, s.numel, 0, w16_off, decay, (s.numel + 8191) // 8192])
        rows, first = [], 0
        for kind in (2, 1, 0):
            for off, r, c, k, w16, decay, units in segs[kind]:
                rows.append([off, r, c, k, w16, decay, first, 0])
                first += units
        self.tail_table = torch.tensor(rows, dtype=torch.int64, device=self.device)
        self.tail_units = first

    def fused_tail(self, m, v, lr, betas, eps, wd, step, counter, gnorm_sq=None, max_norm=0.0, dev_lr_step=None):
        """clip scale + AdamW + normalize_matrices + bf16 operand emit + zero_grad, one launch (train.py:935-946, 461-480)."""
        if getattr(self, "tail_table", None) is None:
            self._build_tail_table()
        ops.adamw_norm_fused(self.P32, self.G32, m, v, self.W16, self.tail_table, self.tail_table.shape[0], self.tail_units,
                             lr, betas[0], betas[1], eps, wd, step, counter, gnorm_sq=gnorm_sq, max_norm=max_norm,
                             dev_lr_step=dev_lr_step, zero_grad=True)
        self.launches += 1
        # fp32 and bf16 were written together: the operands are current until somebody bumps a parameter version
        self._p16_version = tuple(s.param._version for s in self.slots.values())

    # ------------------------------------------------------------------------------------------ activations
    def _buffers(self, B):
        if B in self._acts:
            self._acts[B] = self._acts.pop(B)       # most recently used last
            return self._acts[B]
        # make room BEFORE allocating (an activation set is tens of GB): drop the least recently used unpinned sets
        while len(self._acts) >= max(1, self.max_resident_batches):
            victim = next((b for b in self._acts if b not in self._pinned), None)
            if victim is None:
                break
            del self._acts[victim]
            self.acts_epoch += 1
        cfg, dev = self.cfg, self.device
        C, L, P, G = cfg.n_embd, cfg.n_layer, cfg.local_patch_size, cfg.global_patch_size
        T = (cfg.image_size // P) ** 2
        M, H = B * T, cfg.n_head
        Kl, Kg = cfg.channels * P * P, cfg.channels * G * G
        ncls = cfg.num_classes

        def e(*shape, dtype=BF16):
            return torch.empty(*shape, device=dev, dtype=dtype)
        a = {
            "A_l": e(M, Kl), "A_g": e(M, Kg), "local32": e(M, C, dtype=F32), "local16": e(M, C), "global16": e(M, C),
            "ca": [{"q": e(M, C), "kv": e(M, 2 * C), "att": e(M, C), "lse": e(B, H, T, dtype=F32), "uv": e(M, 2 * C),
                    "x": e(M, C), "o": e(M, C), "inv": e(M, 2 * H, dtype=F32)} for _ in range(3 if cfg.use_kohonen else 1)],
            "qk_inv": [e(M, 2 * H, dtype=F32) for _ in range(L)],
            "h32": [e(M, C, dtype=F32) for _ in range(L + 1)], "h16": [e(M, C) for _ in range(L + 1)],
            "qkv": [e(M, 3 * C) for _ in range(L)], "att": [e(M, C) for _ in range(L)],
            "lse": [e(B, H, T, dtype=F32) for _ in range(L)], "h_att": [e(M, C) for _ in range(L)],
            "h1_32": [e(M, C, dtype=F32) for _ in range(L)], "h1_16": [e(M, C) for _ in range(L)],
            "uv": [e(M, 8 * C) for _ in range(L)], "x": [e(M, 4 * C) for _ in range(L)], "h_mlp": [e(M, C) for _ in range(L)],
            "y16": e(B, C), "xhat": e(B, C, dtype=F32), "rstd": e(B, dtype=F32),
            "raw": e(B, ncls, dtype=F32), "logits": e(B, ncls, dtype=F32), "pred": e(M, Kl),
            # backward workspace
            "G": e(M, C, dtype=F32), "dHin": e(M, C, dtype=F32), "dH1": e(M, C, dtype=F32),
            "d_c16": e(M, C), "d_c16b": e(M, C), "d_4c": e(M, 4 * C), "d_8c": e(M, 8 * C), "d_3c": e(M, 3 * C),
            "draw16": torch.zeros(B, _align(ncls, 8), device=dev, dtype=BF16), "dy16": e(B, C),
        }
        if not cfg.use_nvit:
            a.update({"global32": e(M, C, dtype=F32), "y1_32": [e(M, C, dtype=F32) for _ in range(L)],
                      "y1_16": [e(M, C) for _ in range(L)], "dY1": e(M, C, dtype=F32)})
        if cfg.use_kohonen:
            Gn = self.slots["local_kohonen.nodes"].shape[0]
            a["global32"] = e(M, C, dtype=F32)
            a["som"] = {tag: {"node_sq": e(Gn, dtype=F32), "hi": e(Gn, C), "lo": e(Gn, C), "snap": e(Gn, C, dtype=F32),
                              "xlo": e(M, C), "dots": e(M, Gn, dtype=F32), "idx": e(M, dtype=torch.int32),
                              "idx64": e(B, T, dtype=torch.int64), "onehot": e(M, Gn), "counts": e(Gn, dtype=F32),
                              "repr32": e(M, C, dtype=F32), "repr16": e(M, C), "pooled": e(B, C, dtype=F32)}
                        for tag in ("local", "global")}
            a.update({"ln32": e(M, C, dtype=F32), "ln16": e(M, C), "gn32": e(M, C, dtype=F32), "gn16": e(M, C),
                      "k32a": e(M, C, dtype=F32), "k32b": e(M, C, dtype=F32), "dpred": e(M, Kl)})
        for k, v in a["ca"][0].items():      # the original-ViT branch addresses its single cross-attention call by these names
            a["ca_" + k] = v
        self._acts[B] = a
        return a

    def _wgrad(self, dy, x, gw):
        """gw[N,K] += dy[M,N]^T x[M,K]"""
        ops.linear_wgrad(dy, x, gw, splits=0, accumulate=True)     # splits=0: the library picks the split-K factor
        self.launches += 1

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, img: torch.Tensor, save: bool = True):
        cfg = self.cfg
        # uint8 [B, S, S, channels] (HWC, as the data loader holds images before ToTensor): normalised on the fly by the
        # im2col kernels with input_mean / input_std (Normalize(0.5, 0.5) of train.py:1081-1092 by default)
        self._u8 = img.dtype == torch.uint8
        S = cfg.image_size
        if self._u8:
            if img.dim() != 4 or img.shape[-1] != cfg.channels:
                raise ValueError("uint8 input must be [B, S, S, channels] (HWC)")
            if tuple(img.shape[1:3]) != (S, S):
                raise ValueError("uint8 input must be [B, S, S, channels] with S = image_size")
        else:
            # the activation buffers are sized from the config: any other image shape would write past them (the
            # reference fails at the position-embedding add, model.py:411-412)
            if img.dim() != 4 or tuple(img.shape[1:]) != (cfg.channels, S, S):
                raise ValueError(f"input must be [B, {cfg.channels}, {S}, {S}] (config channels / image_size), got {tuple(img.shape)}")
            if img.dtype != F32:
                img = img.float()
        if img.shape[0] < 1:
            raise ValueError("empty batch")
        img = img.contiguous()
        self._check_alias(img.device)
        self.refresh_operands()
        B = img.shape[0]
        if not cfg.use_nvit:
            return self._forward_orig(img, save)
        C, L, H, P, G = cfg.n_embd, cfg.n_layer, cfg.n_head, cfg.local_patch_size, cfg.global_patch_size
        T = (cfg.image_size // P) ** 2
        a = self._buffers(B)
        self._cur_B = B
        bias = cfg.bias
        p, w16 = self.p, self.w16
        amul = 0.05 / cfg.base_scale
        smul = 1.0 / cfg.base_scale
        att_scale = float(C // H) ** 0.5
        n0 = self.launches

        # ---- dual patch embedding as im2col + GEMM, bias and position embedding in the epilogue (model.py:407-415)
        self._im2col(img, a, P, G)
        ops.linear_fwd(a["A_l"], w16("local_patch_embed.weight"), a["local32"], bias=p("local_patch_embed.bias"),
                       rowadd=p("local_pos_embed").view(T, C), rowadd_period=T, c2=a["local16"])
        if cfg.use_kohonen:      # the maps and the quantization loss read the fp32 global embedding as well
            ops.linear_fwd(a["A_g"], w16("global_patch_embed.1.weight"), a["global32"], bias=p("global_patch_embed.1.bias"),
                           rowadd=p("global_pos_embed").view(T, C), rowadd_period=T, c2=a["global16"])
        else:
            ops.linear_fwd(a["A_g"], w16("global_patch_embed.1.weight"), a["global16"], bias=p("global_patch_embed.1.bias"),
                           rowadd=p("global_pos_embed").view(T, C), rowadd_period=T)
        self.launches += 4

        # ---- cross attention (model.py:219-275): q <- local, k,v <- global; three shared-weight calls around the Kohonen
        # maps (model.py:424-445), one without them (model.py:448)
        if cfg.use_kohonen:
            self._kohonen_fwd(a, B, T)
            som = a["som"]
            self._ca_fwd(a["ca"][0], B, T, som["local"]["repr32"], som["local"]["repr16"], a["local16"], a["ln32"], a["ln16"])
            self._ca_fwd(a["ca"][1], B, T, som["global"]["repr32"], som["global"]["repr16"], a["global16"], a["gn32"], a["gn16"])
            self._ca_fwd(a["ca"][2], B, T, a["ln32"], a["ln16"], a["gn16"], a["h32"][0], a["h16"][0])
        else:
            self._ca_fwd(a["ca"][0], B, T, a["local32"], a["local16"], a["global16"], a["h32"][0], a["h16"][0])

        # ---- transformer blocks + norm_skip (model.py:92-169, 84-87, 450-452)
        for i in range(L):
            b = f"transformer.h.{i}."
            qkv, inv = a["qkv"][i], (a["qk_inv"][i] if self.fuse_qknorm else None)
            qkv_bias = self._cat_bias(b + "query.bias", 3 * C) if bias else None
            if inv is not None:
                ops.gemm_qknorm(a["h16"][i], w16(b + "query.weight", rows=3 * C), qkv, p(b + "sqk"), smul, C, 2 * C, inv, bias=qkv_bias)
            else:
                ops.linear_fwd(a["h16"][i], w16(b + "query.weight", rows=3 * C), qkv, bias=qkv_bias)
            ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], p(b + "sqk"), smul, att_scale, a["att"][i], a["lse"][i], B, H, T,
                              inv_q=None if inv is None else inv[:, :H], inv_k=None if inv is None else inv[:, H:])
            ops.linear_fwd(a["att"][i], w16(b + "att_c_proj.weight"), a["h_att"][i], bias=p(b + "att_c_proj.bias") if bias else None)
            ops.residual_fwd(a["h32"][i], a["h_att"][i], p(b + "attn_alpha"), amul, a["h1_32"][i], a["h1_16"][i])
            self._gated_fwd(a["h1_16"][i], w16(b + "c_fc.weight"), p(b + "c_fc.bias") if bias else None, p(b + "suv"), math.sqrt(C),
                            a["uv"][i], a["x"][i], 4 * C)
            ops.linear_fwd(a["x"][i], w16(b + "mlp_c_proj.weight"), a["h_mlp"][i], bias=p(b + "mlp_c_proj.bias") if bias else None)
            ops.residual_fwd(a["h1_32"][i], a["h_mlp"][i], p(b + "mlp_alpha"), amul, a["h32"][i + 1], a["h16"][i + 1],
                             h0=a["h32"][i], skip=p(b + "skip_param"))
            self.launches += 6

        # ---- classifier head (model.py:455-456, 466-468) and reconstruction loss (model.py:459-464)
        ops.pool_ln_fwd(a["h32"][L], p("mlp_head.0.weight"), p("mlp_head.0.bias"), 1e-5, a["y16"], a["xhat"], a["rstd"], B, T, C)
        ops.linear_fwd(a["y16"], w16("mlp_head.1.weight"), a["raw"], bias=p("mlp_head.1.bias"))
        ops.head_scale_fwd(a["raw"], p("sz"), cfg.sz_init_value / cfg.sz_init_scaling, a["logits"])
        ops.linear_fwd(a["h16"][L], w16("reconstruction_head.0.weight"), a["pred"], bias=p("reconstruction_head.0.bias"))
        self.scratch[0:1].zero_()
        ops.tanh_mse(a["pred"], a["A_l"], self.scratch[0:1])
        self.launches += 5
        self.last_forward_launches = self.launches - n0
        self._saved = (B, T) if save else None
        self.last_aux = self._kohonen_aux(B * T) if cfg.use_kohonen else {}
        return a["logits"].clone(), self.scratch[0].clone()

    input_mean, input_std = 0.5, 0.5      # Normalize(mean, std) applied to uint8 inputs (train.py:1081-1092)

    def _im2col(self, img, a, P, G):
        """Patch operands of the two embeddings (model.py:286-304): local P x P, global G x G reflect-padded, stride P."""
        if self._u8:
            scale, shift = 1.0 / (255.0 * self.input_std), -self.input_mean / self.input_std
            ops.im2col_u8(img, a["A_l"], P, P, 0, scale, shift)
            ops.im2col_u8(img, a["A_g"], G, P, (G - P) // 2, scale, shift)
        else:
            ops.im2col(img, a["A_l"], P, P, 0)
            ops.im2col(img, a["A_g"], G, P, (G - P) // 2)

    def _ca_fwd(self, s, B, T, loc32, loc16, glob16, out32, out16):
        """CrossAttentionBlock.forward, nViT mode (model.py:219-275); `s` holds this call's saved activations."""
        cfg = self.cfg
        C, H = cfg.n_embd, cfg.n_head
        p, w16, bias = self.p, self.w16, cfg.bias
        ca = "cross_attention."
        smul = 1.0 / cfg.base_scale
        inv = s["inv"] if self.fuse_qknorm else None
        kv_bias = self._cat_bias(ca + "k_global.bias", 2 * C) if bias else None
        if inv is not None:
            ops.gemm_qknorm(loc16, w16(ca + "q_local.weight"), s["q"], p(ca + "sqk"), smul, C, C, inv[:, :H],
                            bias=p(ca + "q_local.bias") if bias else None)
            ops.gemm_qknorm(glob16, w16(ca + "k_global.weight", rows=2 * C), s["kv"], p(ca + "sqk"), smul, C, C, inv[:, H:], bias=kv_bias)
        else:
            ops.linear_fwd(loc16, w16(ca + "q_local.weight"), s["q"], bias=p(ca + "q_local.bias") if bias else None)
            ops.linear_fwd(glob16, w16(ca + "k_global.weight", rows=2 * C), s["kv"], bias=kv_bias)
        ops.attention_fwd(s["q"], s["kv"][:, :C], s["kv"][:, C:], p(ca + "sqk"), smul, float(C // H) ** 0.5, s["att"],
                          s["lse"], B, H, T, inv_q=None if inv is None else inv[:, :H], inv_k=None if inv is None else inv[:, H:])
        self._gated_fwd(s["att"], w16(ca + "proj.weight"), p(ca + "proj.bias") if bias else None, None, 1.0, s["uv"], s["x"], C)
        ops.linear_fwd(s["x"], w16(ca + "out_proj.weight"), s["o"], bias=p(ca + "out_proj.bias") if bias else None)
        ops.residual_fwd(loc32, s["o"], p(ca + "attn_alpha"), 0.05 / cfg.base_scale, out32, out16)
        self.launches += 5

    # ------------------------------------------------------------------------------------------ Kohonen maps (config 5)
    def som_geometry(self):
        """Grid rows/columns and neighbourhood width of one map (kohonen.py:52-69)."""
        per_map = self.cfg.kohonen_nodes // 2
        m = int(per_map ** 0.5)
        n = per_map // m
        return m, n, (m * n) ** 0.5 / 2.0

    def _kohonen_fwd(self, a, B, T):
        """Best-matching units, representations and (training mode) the in-forward map update (model.py:417-431).

        Distances on tcgen05: dots = x n^T from bf16 hi/lo splits of both fp32 operands (three GEMMs, fp32 accumulate)."""
        cfg, model = self.cfg, self.model
        C = cfg.n_embd
        M = B * T
        gm, gn, sigma = self.som_geometry()
        alpha = cfg.kohonen_scheduler_min_lr if cfg.kohonen_scheduler_enabled else cfg.kohonen_alpha     # model.py:316, 321
        if model.training:
            self.aux_w[5:6].fill_(float(model.get_kohonen_lr(model.step)) * alpha)
        for tag, x32, x16 in (("local", a["local32"], a["local16"]), ("global", a["global32"], a["global16"])):
            s = a["som"][tag]
            nodes = self.p(tag + "_kohonen.nodes")
            Gn = nodes.shape[0]
            ops.som_prepare(nodes, s["node_sq"], s["hi"], s["lo"], s["snap"])
            ops.split_bf16(x32, None, s["xlo"])
            for k, (xa, nb) in enumerate(((x16, s["hi"]), (x16, s["lo"]), (s["xlo"], s["hi"]))):
                ops.gemm(xa, nb, s["dots"], M=M, N=Gn, K=C, lda=C, ldb=C, ldc=Gn, accumulate=(k > 0))
            s["counts"].zero_()
            ops.som_select(s["dots"], s["node_sq"], s["snap"], s["idx"], s["idx64"], s["onehot"], s["counts"], s["repr32"], s["repr16"])
            self.launches += 6
            if model.training:
                # kohonen.py:138-165: update i pairs image i with the unit of flattened token i, for i < min(B*T, B)
                ops.som_pool(x32, B * C, T, s["pooled"])
                ops.som_update(nodes, s["pooled"], s["idx"], min(M, B), gm, gn, self.aux_w[5:6], sigma)
                self.launches += 2

    def _kohonen_aux(self, M):
        """The four Kohonen losses (model.py:437-442) as device scalars; smoothness sees the UPDATED nodes."""
        cfg = self.cfg
        C = cfg.n_embd
        a = self._acts[self._cur_B]
        som = a["som"]
        side = int(math.sqrt(cfg.kohonen_nodes // 2))
        if side * side != cfg.kohonen_nodes // 2:
            raise ValueError(f"Number of nodes per map ({cfg.kohonen_nodes // 2}) must be a perfect square. "
                             f"Got {cfg.kohonen_nodes} total nodes.")
        self.scratch[1:5].zero_()
        ops.som_pair_losses(som["local"]["repr32"], som["global"]["repr32"], a["local32"], a["global32"], self.scratch[1:4])
        for tag in ("local", "global"):
            ops.som_smoothness(self.p(tag + "_kohonen.nodes"), som[tag]["counts"], side, M, self.scratch[4:5])
        self.launches += 3
        sums = self.scratch.clone()
        return {"kohonen_consistency": 1.0 - sums[1] / M, "kohonen_smoothness": sums[4],
                "local_quantization": sums[2] / (M * C), "global_quantization": sums[3] / (M * C),
                "local_indices": som["local"]["idx64"], "global_indices": som["global"]["idx64"]}

    def set_aux_weights(self, reconstruction=0.0, consistency=0.0, local_quantization=0.0, global_quantization=0.0, smoothness=0.0):
        """Weights (times the incoming gradient scale) of the auxiliary losses for the next backward; device-resident."""
        self.param_list()
        self.aux_w[:5].copy_(torch.tensor([reconstruction, consistency, local_quantization, global_quantization, smoothness],
                                          dtype=F32), non_blocking=True)

    def _cat_bias(self, first_name, n):
        s = self.slots[first_name]
        return self.P32[s.off:s.off + n]

    def _gated_fwd(self, x, w, bias, suv, mul, uv, out, Fh):
        """uv = x W^T (+b); out = (u*su) * silu(v*sv) — fused into the GEMM epilogue unless a bias is present."""
        M, K = x.shape
        if bias is None:
            probe = self.probe is not None and Fh > K          # the block MLP (Fh = 4C), not the cross-attention gate
            if probe:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ops.gemm(x, w, out, M=M, N=Fh, K=K, lda=x.stride(0), ldb=w.stride(0), ldc=Fh, colscale=suv, colscale_mul=mul,
                     c2=uv, ldc2=2 * Fh, swiglu_half=Fh)
            if probe:
                e1.record()
                self.probe.append((e0, e1))
            self.launches += 1
        else:
            ops.linear_fwd(x, w, uv, bias=bias)
            ops.swiglu_fwd(uv, suv, mul, out)
            self.launches += 2

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, dlogits: torch.Tensor):
        """Accumulate dL/dparams into G32 given dL/dlogits (fp32 [B, classes]).  Forward must have run with save=True."""
        if self._saved is None:
            raise RuntimeError("Engine.backward called without a saved forward pass")
        cfg = self.cfg
        if not cfg.use_nvit:
            return self._backward_orig(dlogits)
        B, T = self._saved
        C, L, H = cfg.n_embd, cfg.n_layer, cfg.n_head
        M = B * T
        ncls = cfg.num_classes
        a = self._acts[B]
        p, g, w16, g2d = self.p, self.g, self.w16, self.g2d
        bias = cfg.bias
        amul = 0.05 / cfg.base_scale
        smul = 1.0 / cfg.base_scale
        att_scale = float(C // H) ** 0.5
        n0 = self.launches
        dlogits = dlogits.to(F32).contiguous()

        # ---- head
        draw = a["draw16"][:, :ncls]
        ops.head_scale_bwd(dlogits, a["raw"], p("sz"), cfg.sz_init_value / cfg.sz_init_scaling, a["draw16"], g("sz"))
        ops.colsum(draw, g("mlp_head.1.bias"))
        self._wgrad(draw, a["y16"], g2d("mlp_head.1.weight"))
        ops.linear_dgrad(draw, w16("mlp_head.1.weight"), a["dy16"])
        ops.pool_ln_bwd(a["dy16"], p("mlp_head.0.weight"), a["xhat"], a["rstd"], a["G"], g("mlp_head.0.weight"), g("mlp_head.0.bias"), B, T, C)
        self.launches += 4
        G, dHin, dH1 = a["G"], a["dHin"], a["dH1"]
        if self.rec_active:
            # reconstruction head backward (model.py:459-464 in the objective, train.py:925-926)
            ops.tanh_mse_bwd(a["pred"], a["A_l"], self.aux_w[0:1], a["dpred"])
            self._wgrad(a["dpred"], a["h16"][L], g2d("reconstruction_head.0.weight"))
            ops.colsum(a["dpred"], g("reconstruction_head.0.bias"))
            ops.linear_dgrad(a["dpred"], w16("reconstruction_head.0.weight"), G, accumulate=True)
            self.launches += 3

        # ---- blocks, last to first
        for i in reversed(range(L)):
            b = f"transformer.h.{i}."
            dHmlp, dHatt, dAtt = a["d_c16"], a["d_c16"], a["d_c16b"]
            ops.residual_bwd(G, a["h1_32"][i], a["h_mlp"][i], p(b + "mlp_alpha"), amul, dH1, dHmlp, g(b + "mlp_alpha"),
                             h0=a["h32"][i], skip=p(b + "skip_param"), dh0=dHin, dskip=g(b + "skip_param"))
            self._wgrad(dHmlp, a["x"][i], g2d(b + "mlp_c_proj.weight"))
            if bias:
                ops.colsum(dHmlp, g(b + "mlp_c_proj.bias"))
            if self.fuse_gate_bwd and not bias:
                # dx = dHmlp W stays on chip: the dgrad GEMM's epilogue applies the gate backward against the saved raw u|v;
                # dL/dsuv[c] = W_fc[c,:] . dW_fc[c,:] / suv[c] follows from the weight gradient (18 MB instead of [M, 8C])
                ops.gemm_gate_bwd(dHmlp, w16(b + "mlp_c_proj.weight"), a["uv"][i], p(b + "suv"), math.sqrt(C), a["d_8c"])
                self._wgrad(a["d_8c"], a["h1_16"][i], g2d(b + "c_fc.weight"))
                ops.rowdot_div(p(b + "c_fc.weight"), g(b + "c_fc.weight"), p(b + "suv"), g(b + "suv"))
            else:
                ops.linear_dgrad(dHmlp, w16(b + "mlp_c_proj.weight"), a["d_4c"])
                ops.swiglu_bwd(a["d_4c"], a["uv"][i], p(b + "suv"), math.sqrt(C), a["d_8c"], g(b + "suv"))
                self._wgrad(a["d_8c"], a["h1_16"][i], g2d(b + "c_fc.weight"))
            if bias:
                ops.colsum(a["d_8c"], g(b + "c_fc.bias"))
            ops.linear_dgrad(a["d_8c"], w16(b + "c_fc.weight"), dH1, accumulate=True)
            ops.residual_bwd(dH1, a["h32"][i], a["h_att"][i], p(b + "attn_alpha"), amul, dHin, dHatt, g(b + "attn_alpha"),
                             dh_accumulate=True)
            self._wgrad(dHatt, a["att"][i], g2d(b + "att_c_proj.weight"))
            if bias:
                ops.colsum(dHatt, g(b + "att_c_proj.bias"))
            ops.linear_dgrad(dHatt, w16(b + "att_c_proj.weight"), dAtt)
            qkv, dqkv = a["qkv"][i], a["d_3c"]
            inv = a["qk_inv"][i] if self.fuse_qknorm else None
            ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], p(b + "sqk"), smul, att_scale, a["att"][i], dAtt, a["lse"][i],
                              dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], g(b + "sqk"), B, H, T,
                              inv_q=None if inv is None else inv[:, :H], inv_k=None if inv is None else inv[:, H:])
            self._wgrad(dqkv, a["h16"][i], g2d(b + "query.weight", rows=3 * C))
            if bias:
                ops.colsum(dqkv, self._cat_grad(b + "query.bias", 3 * C))
            ops.linear_dgrad(dqkv, w16(b + "query.weight", rows=3 * C), dHin, accumulate=True)
            self.launches += 8 + (4 if bias else 0)
            G, dHin = dHin, G
            if self.grad_ready_hook is not None:
                self.grad_ready_hook(*self.block_grad_range(i))

        # ---- cross attention: G = dL/d(h0)
        if cfg.use_kohonen:
            dLocal, dGlobal = self._kohonen_bwd(a, B, T, G, dHin, dH1)
        else:
            dLocal, dGlobal = dHin, dH1
            self._ca_bwd(a["ca"][0], B, T, G, a["local32"], a["local16"], a["global16"], dLocal, dGlobal)

        # ---- patch embeddings: position / bias gradients and the two conv weight gradients (no dX: images need none)
        for dX32, A, wname, bname, pos in ((dLocal, a["A_l"], "local_patch_embed.weight", "local_patch_embed.bias", "local_pos_embed"),
                                           (dGlobal, a["A_g"], "global_patch_embed.1.weight", "global_patch_embed.1.bias", "global_pos_embed")):
            ops.pos_bias_grad(dX32, B, T, C, g(pos).view(T, C), g(bname))
            ops.cast_bf16(dX32, a["d_c16"])
            self._wgrad(a["d_c16"], A, g2d(wname))
            self.launches += 2
        self.last_backward_launches = self.launches - n0
        self._saved = None
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(0, self.block_grad_range(0)[0])
            self.grad_ready_hook(self.block_grad_range(cfg.n_layer - 1)[1], self.n_active)


    def _ca_bwd(self, s, B, T, G, loc32, loc16, glob16, dLoc, dGlob):
        """Backward of one cross-attention call: G = dL/d(out) -> dLoc = dL/d(local) and dGlob = dL/d(global), both fp32
        [M, C] and WRITTEN; parameter gradients accumulate (the three Kohonen-mode calls share weights)."""
        cfg = self.cfg
        C, H = cfg.n_embd, cfg.n_head
        M = B * T
        a = self._acts[B]
        p, g, w16, g2d, bias = self.p, self.g, self.w16, self.g2d, cfg.bias
        ca = "cross_attention."
        dO, dX = a["d_c16"], a["d_c16b"]
        ops.residual_bwd(G, loc32, s["o"], p(ca + "attn_alpha"), 0.05 / cfg.base_scale, dLoc, dO, g(ca + "attn_alpha"))
        self._wgrad(dO, s["x"], g2d(ca + "out_proj.weight"))
        if bias:
            ops.colsum(dO, g(ca + "out_proj.bias"))
        duv = a["d_8c"].view(-1)[:M * 2 * C].view(M, 2 * C)     # contiguous [M, 2C] scratch
        if self.fuse_gate_bwd and C % 64 == 0:
            ops.gemm_gate_bwd(dO, w16(ca + "out_proj.weight"), s["uv"], None, 1.0, duv)
        else:
            ops.linear_dgrad(dO, w16(ca + "out_proj.weight"), dX)
            ops.swiglu_bwd(dX, s["uv"], None, 1.0, duv, None)
        self._wgrad(duv, s["att"], g2d(ca + "proj.weight"))
        if bias:
            ops.colsum(duv, g(ca + "proj.bias"))
        dAtt = a["d_c16"]
        ops.linear_dgrad(duv, w16(ca + "proj.weight"), dAtt)
        dq = a["d_c16b"]
        dkv = a["d_3c"].view(-1)[:M * 2 * C].view(M, 2 * C)
        inv = s["inv"] if self.fuse_qknorm else None
        ops.attention_bwd(s["q"], s["kv"][:, :C], s["kv"][:, C:], p(ca + "sqk"), 1.0 / cfg.base_scale, float(C // H) ** 0.5, s["att"], dAtt,
                          s["lse"], dq, dkv[:, :C], dkv[:, C:], g(ca + "sqk"), B, H, T,
                          inv_q=None if inv is None else inv[:, :H], inv_k=None if inv is None else inv[:, H:])
        self._wgrad(dq, loc16, g2d(ca + "q_local.weight"))
        self._wgrad(dkv, glob16, g2d(ca + "k_global.weight", rows=2 * C))
        if bias:
            ops.colsum(dq, g(ca + "q_local.bias"))
            ops.colsum(dkv, self._cat_grad(ca + "k_global.bias", 2 * C))
            self.launches += 4
        ops.linear_dgrad(dq, w16(ca + "q_local.weight"), dLoc, accumulate=True)
        ops.linear_dgrad(dkv, w16(ca + "k_global.weight", rows=2 * C), dGlob)
        self.launches += 6 if (self.fuse_gate_bwd and C % 64 == 0) else 7

    def _kohonen_bwd(self, a, B, T, G, f1, f2):
        """Backward of model.py:417-445: the three shared-weight cross-attention calls, the map losses and the scatter
        of the representation gradients into the node tables.  Returns (dL/d local patches, dL/d global patches)."""
        som = a["som"]
        M = B * T
        free = [f1, f2, a["k32a"], a["k32b"]]                 # fp32 [M, C] scratch besides G
        d_ln, d_gn = free.pop(), free.pop()
        self._ca_bwd(a["ca"][2], B, T, G, a["ln32"], a["ln16"], a["gn16"], d_ln, d_gn)
        free.append(G)
        d_rg, d_xg = free.pop(), free.pop()
        self._ca_bwd(a["ca"][1], B, T, d_gn, som["global"]["repr32"], som["global"]["repr16"], a["global16"], d_rg, d_xg)
        free.append(d_gn)
        d_rl, d_xl = free.pop(), free.pop()
        self._ca_bwd(a["ca"][0], B, T, d_ln, som["local"]["repr32"], som["local"]["repr16"], a["local16"], d_rl, d_xl)
        # consistency and quantization gradients join the representation / patch gradients (model.py:437, 441-442)
        ops.som_pair_losses(som["local"]["repr32"], som["global"]["repr32"], a["local32"], a["global32"], None, self.aux_w[1:4],
                            d_rl, d_rg, d_xl, d_xg)
        side = int(math.sqrt(self.cfg.kohonen_nodes // 2))
        hi, lo = a["d_c16"], a["d_c16b"]
        for tag, d_r in (("local", d_rl), ("global", d_rg)):
            # nodes[idx] gather backward = onehot^T dRepr on the tensor cores, dRepr as bf16 hi + lo
            gn = self.g2d(tag + "_kohonen.nodes")
            ops.split_bf16(d_r, hi, lo)
            self._wgrad(som[tag]["onehot"], hi, gn)
            self._wgrad(som[tag]["onehot"], lo, gn)
            ops.som_smoothness(self.p(tag + "_kohonen.nodes"), som[tag]["counts"], side, M, self.scratch[5:6], self.aux_w[4:5], gn)
            self.launches += 2
        self.launches += 1
        return d_xl, d_xg

    # ------------------------------------------------------------------------------------------ original-ViT branch
    # config.use_nvit = False (BASELINE config 4): RMSNorm pre-norm that OVERWRITES h, plain residual adds, 1/sqrt(D)
    # attention without q/k normalisation, no suv / sz, cross-attention without residual - and norm_skip still applied
    # (model.py:95-96, 132-133, 145-146, 157-158, 221-223, 450-452; SURVEY.md appendix B "original ViT branch as written").
    def _forward_orig(self, img, save=True):
        cfg = self.cfg
        B = img.shape[0]
        C, L, H, P, G = cfg.n_embd, cfg.n_layer, cfg.n_head, cfg.local_patch_size, cfg.global_patch_size
        T = (cfg.image_size // P) ** 2
        a = self._buffers(B)
        self._cur_B = B
        bias = cfg.bias
        p, w16 = self.p, self.w16
        att_scale = 1.0 / float(C // H) ** 0.5
        eps = 1e-6
        n0 = self.launches
        self._im2col(img, a, P, G)
        ops.linear_fwd(a["A_l"], w16("local_patch_embed.weight"), a["local32"], bias=p("local_patch_embed.bias"),
                       rowadd=p("local_pos_embed").view(T, C), rowadd_period=T)
        ops.linear_fwd(a["A_g"], w16("global_patch_embed.1.weight"), a["global32"], bias=p("global_patch_embed.1.bias"),
                       rowadd=p("global_pos_embed").view(T, C), rowadd_period=T)
        ca = "cross_attention."
        ops.add_rmsnorm_fwd(a["local32"], None, p(ca + "local_norm.weight"), eps, None, a["local16"])
        ops.add_rmsnorm_fwd(a["global32"], None, p(ca + "global_norm.weight"), eps, None, a["global16"])
        ops.linear_fwd(a["local16"], w16(ca + "q_local.weight"), a["ca_q"], bias=p(ca + "q_local.bias") if bias else None)
        ops.linear_fwd(a["global16"], w16(ca + "k_global.weight", rows=2 * C), a["ca_kv"],
                       bias=self._cat_bias(ca + "k_global.bias", 2 * C) if bias else None)
        ops.attention_fwd(a["ca_q"], a["ca_kv"][:, :C], a["ca_kv"][:, C:], None, 1.0, att_scale, a["ca_att"], a["ca_lse"], B, H, T)
        self._gated_fwd(a["ca_att"], w16(ca + "proj.weight"), p(ca + "proj.bias") if bias else None, None, 1.0, a["ca_uv"], a["ca_x"], C)
        ops.linear_fwd(a["ca_x"], w16(ca + "out_proj.weight"), a["h32"][0], bias=p(ca + "out_proj.bias") if bias else None, c2=a["h16"][0])
        self.launches += 11
        for i in range(L):
            b = f"transformer.h.{i}."
            ops.add_rmsnorm_fwd(a["h32"][i], None, p(b + "rmsnorm_att.weight"), eps, a["y1_32"][i], a["y1_16"][i])
            ops.linear_fwd(a["y1_16"][i], w16(b + "query.weight", rows=3 * C), a["qkv"][i],
                           bias=self._cat_bias(b + "query.bias", 3 * C) if bias else None)
            qkv = a["qkv"][i]
            ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], None, 1.0, att_scale, a["att"][i], a["lse"][i], B, H, T)
            ops.linear_fwd(a["att"][i], w16(b + "att_c_proj.weight"), a["h_att"][i], bias=p(b + "att_c_proj.bias") if bias else None)
            ops.add_rmsnorm_fwd(a["y1_32"][i], a["h_att"][i], p(b + "rmsnorm_mlp.weight"), eps, a["h1_32"][i], a["h1_16"][i])
            self._gated_fwd(a["h1_16"][i], w16(b + "c_fc.weight"), p(b + "c_fc.bias") if bias else None, None, 1.0, a["uv"][i], a["x"][i], 4 * C)
            ops.linear_fwd(a["x"][i], w16(b + "mlp_c_proj.weight"), a["h_mlp"][i], bias=p(b + "mlp_c_proj.bias") if bias else None)
            ops.add_skipnorm_fwd(a["h1_32"][i], a["h_mlp"][i], a["h32"][i], p(b + "skip_param"), a["h32"][i + 1], a["h16"][i + 1])
            self.launches += 7
        ops.pool_ln_fwd(a["h32"][L], p("mlp_head.0.weight"), p("mlp_head.0.bias"), 1e-5, a["y16"], a["xhat"], a["rstd"], B, T, C)
        ops.linear_fwd(a["y16"], w16("mlp_head.1.weight"), a["logits"], bias=p("mlp_head.1.bias"))
        ops.linear_fwd(a["h16"][L], w16("reconstruction_head.0.weight"), a["pred"], bias=p("reconstruction_head.0.bias"))
        self.scratch[0:1].zero_()
        ops.tanh_mse(a["pred"], a["A_l"], self.scratch[0:1])
        self.launches += 4
        self.last_forward_launches = self.launches - n0
        self._saved = (B, T) if save else None
        return a["logits"].clone(), self.scratch[0].clone()

    def _backward_orig(self, dlogits):
        cfg = self.cfg
        B, T = self._saved
        C, L, H = cfg.n_embd, cfg.n_layer, cfg.n_head
        M = B * T
        ncls = cfg.num_classes
        a = self._acts[B]
        p, g, w16, g2d = self.p, self.g, self.w16, self.g2d
        bias = cfg.bias
        att_scale = 1.0 / float(C // H) ** 0.5
        eps = 1e-6
        n0 = self.launches
        dlogits = dlogits.to(F32).contiguous()
        draw = a["draw16"][:, :ncls]
        ops.head_scale_bwd(dlogits, a["logits"], None, 1.0, a["draw16"], None)     # no sz: draw = bf16(dlogits)
        ops.colsum(draw, g("mlp_head.1.bias"))
        self._wgrad(draw, a["y16"], g2d("mlp_head.1.weight"))
        ops.linear_dgrad(draw, w16("mlp_head.1.weight"), a["dy16"])
        ops.pool_ln_bwd(a["dy16"], p("mlp_head.0.weight"), a["xhat"], a["rstd"], a["G"], g("mlp_head.0.weight"), g("mlp_head.0.bias"), B, T, C)
        self.launches += 4
        G, dHin, dY2, dY1 = a["G"], a["dHin"], a["dH1"], a["dY1"]
        for i in reversed(range(L)):
            b = f"transformer.h.{i}."
            dHmlp, dHatt, dAtt = a["d_c16"], a["d_c16"], a["d_c16b"]
            ops.add_skipnorm_bwd(G, a["h1_32"][i], a["h_mlp"][i], a["h32"][i], p(b + "skip_param"), dY2, dHmlp, dHin, g(b + "skip_param"))
            self._wgrad(dHmlp, a["x"][i], g2d(b + "mlp_c_proj.weight"))
            if bias:
                ops.colsum(dHmlp, g(b + "mlp_c_proj.bias"))
            if self.fuse_gate_bwd:
                ops.gemm_gate_bwd(dHmlp, w16(b + "mlp_c_proj.weight"), a["uv"][i], None, 1.0, a["d_8c"])
            else:
                ops.linear_dgrad(dHmlp, w16(b + "mlp_c_proj.weight"), a["d_4c"])
                ops.swiglu_bwd(a["d_4c"], a["uv"][i], None, 1.0, a["d_8c"], None)
            self._wgrad(a["d_8c"], a["h1_16"][i], g2d(b + "c_fc.weight"))
            if bias:
                ops.colsum(a["d_8c"], g(b + "c_fc.bias"))
            ops.linear_dgrad(a["d_8c"], w16(b + "c_fc.weight"), dY2, accumulate=True)
            ops.add_rmsnorm_bwd(dY2, a["y1_32"][i], a["h_att"][i], p(b + "rmsnorm_mlp.weight"), eps, dY1, dHatt, g(b + "rmsnorm_mlp.weight"))
            self._wgrad(dHatt, a["att"][i], g2d(b + "att_c_proj.weight"))
            if bias:
                ops.colsum(dHatt, g(b + "att_c_proj.bias"))
            ops.linear_dgrad(dHatt, w16(b + "att_c_proj.weight"), dAtt)
            qkv, dqkv = a["qkv"][i], a["d_3c"]
            ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], None, 1.0, att_scale, a["att"][i], dAtt, a["lse"][i],
                              dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], None, B, H, T)
            self._wgrad(dqkv, a["y1_16"][i], g2d(b + "query.weight", rows=3 * C))
            if bias:
                ops.colsum(dqkv, self._cat_grad(b + "query.bias", 3 * C))
            ops.linear_dgrad(dqkv, w16(b + "query.weight", rows=3 * C), dY1, accumulate=True)
            ops.add_rmsnorm_bwd(dY1, a["h32"][i], None, p(b + "rmsnorm_att.weight"), eps, dHin, None, g(b + "rmsnorm_att.weight"),
                                dh_accumulate=True)
            self.launches += 9 + (4 if bias else 0)
            G, dHin = dHin, G
            if self.grad_ready_hook is not None:
                self.grad_ready_hook(*self.block_grad_range(i))
        # cross attention (no residual in this mode): G = dL/d(out_proj output)
        ca = "cross_attention."
        dO, dX = a["d_c16"], a["d_c16b"]
        ops.cast_bf16(G, dO)
        self._wgrad(dO, a["ca_x"], g2d(ca + "out_proj.weight"))
        if bias:
            ops.colsum(dO, g(ca + "out_proj.bias"))
        duv = a["d_8c"].view(-1)[:M * 2 * C].view(M, 2 * C)
        if self.fuse_gate_bwd and C % 64 == 0:
            ops.gemm_gate_bwd(dO, w16(ca + "out_proj.weight"), a["ca_uv"], None, 1.0, duv)
        else:
            ops.linear_dgrad(dO, w16(ca + "out_proj.weight"), dX)
            ops.swiglu_bwd(dX, a["ca_uv"], None, 1.0, duv, None)
        self._wgrad(duv, a["ca_att"], g2d(ca + "proj.weight"))
        if bias:
            ops.colsum(duv, g(ca + "proj.bias"))
        dAtt = a["d_c16"]
        ops.linear_dgrad(duv, w16(ca + "proj.weight"), dAtt)
        dq = a["d_c16b"]
        dkv = a["d_3c"].view(-1)[:M * 2 * C].view(M, 2 * C)
        ops.attention_bwd(a["ca_q"], a["ca_kv"][:, :C], a["ca_kv"][:, C:], None, 1.0, att_scale, a["ca_att"], dAtt, a["ca_lse"],
                          dq, dkv[:, :C], dkv[:, C:], None, B, H, T)
        self._wgrad(dq, a["local16"], g2d(ca + "q_local.weight"))
        self._wgrad(dkv, a["global16"], g2d(ca + "k_global.weight", rows=2 * C))
        if bias:
            ops.colsum(dq, g(ca + "q_local.bias"))
            ops.colsum(dkv, self._cat_grad(ca + "k_global.bias", 2 * C))
        dLn, dGn = dY2, dY1
        ops.linear_dgrad(dq, w16(ca + "q_local.weight"), dLn)
        ops.linear_dgrad(dkv, w16(ca + "k_global.weight", rows=2 * C), dGn)
        dLocal, dGlobal = dHin, G
        ops.add_rmsnorm_bwd(dLn, a["local32"], None, p(ca + "local_norm.weight"), eps, dLocal, None, g(ca + "local_norm.weight"))
        ops.add_rmsnorm_bwd(dGn, a["global32"], None, p(ca + "global_norm.weight"), eps, dGlobal, None, g(ca + "global_norm.weight"))
        self.launches += 12
        for dX32, A, wname, bname, pos in ((dLocal, a["A_l"], "local_patch_embed.weight", "local_patch_embed.bias", "local_pos_embed"),
                                           (dGlobal, a["A_g"], "global_patch_embed.1.weight", "global_patch_embed.1.bias", "global_pos_embed")):
            ops.pos_bias_grad(dX32, B, T, C, g(pos).view(T, C), g(bname))
            ops.cast_bf16(dX32, a["d_c16"])
            self._wgrad(a["d_c16"], A, g2d(wname))
            self.launches += 2
        self.last_backward_launches = self.launches - n0
        self._saved = None
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(0, self.block_grad_range(0)[0])
            self.grad_ready_hook(self.block_grad_range(cfg.n_layer - 1)[1], self.n_active)

    def _cat_grad(self, first_name, n):
        s = self.slots[first_name]
        return self.G32[s.off:s.off + n]


class NViTFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward/backward are the engine's hand-scheduled passes.

    Outputs (logits, reconstruction) or, with Kohonen maps, (logits, reconstruction, consistency, smoothness, local
    quantization, global quantization); the gradients arriving for the scalar losses become the device-side weights of
    the fused loss-gradient kernels."""

    AUX = ("kohonen_consistency", "kohonen_smoothness", "local_quantization", "global_quantization")

    @staticmethod
    def forward(ctx, engine, img, *params):
        logits, recon = engine.forward(img, save=True)
        ctx.engine = engine
        ctx.n_params = len(params)
        if not engine.cfg.use_kohonen:
            # Not in the reference's objective without the maps (train.py:909-926), so its backward is not scheduled
            # (reconstruction_head.* stay grad-less, as in the reference's loop).  The output stays differentiable on
            # purpose: a loop that DOES put it into its loss gets an error from backward() instead of silent zeros.
            return logits, recon
        return (logits, recon, *(engine.last_aux[k] for k in NViTFunction.AUX))

    @staticmethod
    def backward(ctx, dlogits, *daux):
        eng = ctx.engine
        if not eng.cfg.use_kohonen and daux and daux[0] is not None and bool((daux[0] != 0).any()):
            raise RuntimeError("nvit_b200: a gradient arrived for aux_losses['reconstruction'], but without Kohonen maps the "
                               "reconstruction loss is not part of the reference's objective (train.py:909-926) and its "
                               "backward pass is not built in this mode; use use_kohonen=True or keep it out of the loss")
        eng.zero_grad()
        if eng.cfg.use_kohonen:
            drec, dcons, dsmooth, dlq, dgq = ((d if d is not None else eng.aux_w.new_zeros(())) for d in daux)
            eng.aux_w[:5].copy_(torch.stack([drec, dcons, dlq, dgq, dsmooth]).to(F32))
        if dlogits is None:
            dlogits = torch.zeros_like(eng._acts[eng._cur_B]["logits"])
        eng.backward(dlogits)
        grads = []
        for n in eng.order:
            s = eng.slots[n]
            if s.off >= eng.n_active or not s.param.requires_grad:
                grads.append(None)
            else:
                grads.append(eng.g(n).clone())
        return (None, None, *grads)
