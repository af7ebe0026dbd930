"""Host-side driver: packs a GPT's parameters into `lp_weight` records and issues the C-ABI calls of one
forward pass (reference call order: GPT.forward model.py:63-111 -> Block.forward 158-180 ->
CausalSelfAttention.forward 194-254 -> MLP 284-301).  PyTorch is used for device memory, streams and CUDA-graph
capture only; every arithmetic operation on the path is a kernel of liblitparrot_b200.so.
"""
import ctypes
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from lit_parrot_b200 import _lib
from lit_parrot_b200._lib import LpWeight

_FMT_OF_DTYPE = {torch.float32: _lib.LP_W_F32, torch.bfloat16: _lib.LP_W_BF16}
_KV_OF_DTYPE = {torch.float32: _lib.LP_F32, torch.bfloat16: _lib.LP_BF16}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class PackedLinear:
    """One linear layer as the kernels see it.  Keeps the tensors alive that the raw pointers refer to."""

    def __init__(self, w: torch.Tensor, fmt: int, N: int, K: int, bias: Optional[torch.Tensor] = None,
                 aux0: Optional[torch.Tensor] = None, aux1: Optional[torch.Tensor] = None, group: int = 0,
                 aux2: Optional[torch.Tensor] = None, flags: int = 0, out_bias: Optional[torch.Tensor] = None,
                 out_scale: Optional[torch.Tensor] = None) -> None:
        assert w.is_contiguous()
        self.keep = (w, bias, aux0, aux1, aux2)
        self.affine = (out_bias, out_scale)  # adapter v2: scale * (linear(x) + bias), fp32 [N] each (adapter_v2.py:34-35)
        self.N, self.K, self.fmt = N, K, fmt
        self.rec = LpWeight(_ptr(w), _ptr(aux0), _ptr(aux1), _ptr(aux2), _ptr(bias), fmt, N, K, group, flags, 0, _ptr(out_bias),
                            _ptr(out_scale))
        self.ref = ctypes.byref(self.rec)

    @property
    def stored_bytes(self) -> int:
        """Bytes a decode step must stream for this layer (weights + scales/zeros/absmax).  When the tile-major
        scale/zero buffer (aux2) exists it is what the streaming kernel reads instead of aux0/aux1."""
        w, _bias, aux0, aux1, aux2 = self.keep
        # int4: aux2 (tile-major scale / zero words) replaces aux0 / aux1; NF4: aux2 is aux0 (absmax) re-tiled with zero padding,
        # the canonical count is aux0
        aux = (aux2,) if (aux2 is not None and self.fmt != _lib.LP_W_NF4) else (aux0, aux1)
        return sum(t.numel() * t.element_size() for t in (w, *aux) if t is not None)


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to(torch.float32).contiguous()


def pack_linear(mod: torch.nn.Module) -> PackedLinear:
    """nn.Linear (fp32 / bf16) is used in place; quantised modules provide their own packing."""
    if hasattr(mod, "lp_pack"):
        return mod.lp_pack()
    w = mod.weight.data
    if w.dtype not in _FMT_OF_DTYPE:
        raise RuntimeError(f"unsupported parameter dtype {w.dtype}: use float32 or bfloat16")
    if not w.is_contiguous():
        w = w.contiguous()
        mod.weight.data = w
    return PackedLinear(w, _FMT_OF_DTYPE[w.dtype], w.shape[0], w.shape[1], bias=_f32(getattr(mod, "bias", None)),
                        out_bias=_f32(getattr(mod, "adapter_bias", None)), out_scale=_f32(getattr(mod, "adapter_scale", None)))


def pack_swiglu(fc_1: torch.nn.Module, fc_2: torch.nn.Module) -> PackedLinear:
    """fc_1 / fc_2 (model.py:293-300) as ONE weight with interleaved rows (2i = fc_1 row i, 2i+1 = fc_2 row i) so a warp
    produces silu(a)*b without a round trip.  The parameters are re-pointed at strided views of the interleaved
    storage, so nothing is duplicated and `state_dict()` / `load_state_dict()` keep working."""
    if hasattr(fc_1, "lp_pack_pair"):
        return fc_1.lp_pack_pair(fc_2)
    w1, w2 = fc_1.weight.data, fc_2.weight.data
    if w1.dtype not in _FMT_OF_DTYPE:
        raise RuntimeError(f"unsupported parameter dtype {w1.dtype}: use float32 or bfloat16")
    I, E = w1.shape
    already = (w1.stride() == (2 * E, 1) and w2.stride() == (2 * E, 1)
               and w2.data_ptr() == w1.data_ptr() + E * w1.element_size())
    if already:
        inter = torch.as_strided(w1, (2 * I, E), (E, 1))
    else:
        inter = torch.empty((2 * I, E), dtype=w1.dtype, device=w1.device)
        inter[0::2].copy_(w1)
        inter[1::2].copy_(w2)
        fc_1.weight.data = inter[0::2]
        fc_2.weight.data = inter[1::2]
    def pair(name):  # per-row vectors of the two layers, interleaved like the rows
        if getattr(fc_1, name, None) is None:
            return None
        return torch.stack((getattr(fc_1, name).data.float(), getattr(fc_2, name).data.float()), dim=1).reshape(-1).contiguous()

    return PackedLinear(inter, _FMT_OF_DTYPE[inter.dtype], 2 * I, E, bias=pair("bias"), out_bias=pair("adapter_bias"),
                        out_scale=pair("adapter_scale"))


class _Layer:
    pass


class Engine:
    def __init__(self, model, device: torch.device) -> None:
        self.device = device
        self.precision = model.precision
        self.use_graph = bool(model.use_cuda_graph)
        self.round = 1 if model.precision == "bf16" else 0
        self.cfg = cfg = model.config
        self.tp = model.tp_context if cfg.tp_size > 1 else None
        if cfg.tp_size > 1 and self.tp is None:
            raise RuntimeError("config.tp_size > 1: set model.tp_context = lit_parrot_b200.tp.TPContext(...) before the first forward")
        self.lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
        self.param_dtype = model.transformer.wte.weight.dtype
        if self.param_dtype not in _FMT_OF_DTYPE:
            raise RuntimeError(f"unsupported parameter dtype {self.param_dtype}: use float32 or bfloat16")
        for p in model.parameters():
            if p.device != device:
                raise RuntimeError(f"all parameters must live on {device}; found one on {p.device}")
        self.norm_kind = _lib.LP_NORM_RMS if cfg._norm_class == "RMSNorm" else _lib.LP_NORM_LAYERNORM
        self.act = _lib.LP_EPI_SWIGLU if cfg._mlp_class == "LLaMAMLP" else _lib.LP_EPI_GELU
        self.wte = model.transformer.wte.weight.data
        self.layers: List[_Layer] = []
        for blk in model.transformer.h:
            L = _Layer()
            L.n1_w, L.n1_b = _f32(blk.norm_1.weight), _f32(getattr(blk.norm_1, "bias", None))
            if cfg.shared_attention_norm:
                L.n2_w = L.n2_b = None
            else:
                L.n2_w, L.n2_b = _f32(blk.norm_2.weight), _f32(getattr(blk.norm_2, "bias", None))
            L.qkv = pack_linear(blk.attn.attn)
            L.proj = pack_linear(blk.attn.proj)
            if cfg._mlp_class == "LLaMAMLP":
                L.fc = pack_swiglu(blk.mlp.fc_1, blk.mlp.fc_2)
            else:
                L.fc = pack_linear(blk.mlp.fc)
            L.mlp_proj = pack_linear(blk.mlp.proj)
            # LLaMA-Adapter (lit_gpt/adapter.py:171-177): adaption prompt embedding + per-head gate of the adapted layers
            L.adapter = None
            if hasattr(blk.attn, "adapter_wte"):
                L.adapter = dict(wte=_f32(blk.attn.adapter_wte.weight), gate=_f32(blk.attn.gating_factor).reshape(-1).contiguous(), kv=None)
            self.layers.append(L)
        self.lnf_w, self.lnf_b = _f32(model.transformer.ln_f.weight), _f32(getattr(model.transformer.ln_f, "bias", None))
        self.lm_head = pack_linear(model.lm_head)
        self.cos = self.sin = None
        self._bufs: Dict[int, Dict[str, torch.Tensor]] = {}
        self._graphs: Dict[Tuple, Tuple] = {}
        self._scratch: Optional[Tuple] = None
        self._consecutive = False
        # batch-1 decode: the whole step as one persistent kernel (lp_decode_step) where the library covers the model
        self.has_adapter = any(L.adapter is not None for L in self.layers)
        if self.has_adapter and self.tp is not None:
            raise NotImplementedError("adapter models are not sharded tensor-parallel")
        # adapter models take the per-op path (the prefix attention is its own launch; the v2 output affine is refused by the plan)
        self.use_step_kernel = (os.environ.get("LP_DECODE_STEP", "1") != "0" and bool(getattr(model, "use_step_kernel", True))
                                and not self.has_adapter)
        self._steps: Dict[Tuple, Optional[Tuple]] = {}
        self._flag = torch.zeros(1, dtype=torch.int32).pin_memory()  # mapped host word written by lp_validate_inputs
        self._slabs = None  # fused column->row pairs of the step kernel: None = not decided yet, False = not applicable

    # ------------------------------------------------------------------ bookkeeping
    def set_rope(self, rope) -> None:
        cos, sin = (t.to(self.device, torch.float32) for t in rope)
        n_elem = self.cfg.rope_n_elem
        if n_elem % 2 and cos.size(-1) != n_elem:
            # n_elem == 1 (hs 4 at rotary_percentage 0.25: the reference's own unit-test config).  The reference's table is
            # then 2 wide and `x[..., :1] * cos` broadcasts (model.py:229, 336): the head grows to hs + 1 dims with the rotated
            # value r = x0 * (cos - sin) duplicated, so q.k gains 2 * r_q * r_k.  Same scores within hs dims: rotate by
            # sqrt(2) * (cos, -sin) (the kernels' partner for d >= n_elem / 2 is +x).  Other odd sizes fail in the reference too.
            if n_elem != 1:
                raise RuntimeError(f"rotary dimension {n_elem} is odd: the reference cannot apply RoPE to it either")
            cos, sin = cos[:, :1] * math.sqrt(2.0), -sin[:, :1] * math.sqrt(2.0)
        self.cos, self.sin = cos.contiguous(), sin.contiguous()

    def drop_graphs(self) -> None:
        self._graphs.clear()
        self._steps.clear()

    # ------------------------------------------------------------------ LLaMA-Adapter prefix keys / values (load time)
    def _adapter_kv(self, L, stream: int):
        """(ak, av) fp32 [G, aT, hs] of an adapted layer: the k / v parts of attn.attn(adapter_wte.weight) (adapter.py:238-249, incl.
        the v2 output affine of attn.attn), computed once — the reference caches them as `adapter_kv_cache` as well."""
        ad = L.adapter
        if ad["kv"] is None:
            cfg = self.cfg
            aT, G, qpk, hs = ad["wte"].shape[0], cfg.n_query_groups, cfg.q_per_kv, cfg.head_size
            pre = torch.empty((aT, cfg.qkv_rows), device=self.device, dtype=torch.float32)
            for r0 in range(0, aT, 8):  # lp_linear rows per call
                m = min(8, aT - r0)
                _lib.check(self.lib.lp_linear(ad["wte"][r0:].data_ptr(), m, L.qkv.ref, _lib.LP_EPI_NONE, None, pre[r0:].data_ptr(),
                                              self.round, stream), "lp_linear(adapter prefix)")
            pre = pre.view(aT, G, qpk + 2, hs)
            ad["kv"] = (pre[:, :, qpk].permute(1, 0, 2).contiguous(), pre[:, :, qpk + 1].permute(1, 0, 2).contiguous())
        return ad["kv"]

    def drop_adapter_kv(self) -> None:
        for L in self.layers:
            if L.adapter is not None:
                L.adapter["kv"] = None
        self.drop_graphs()

    def adapter_kv_views(self):
        """What the reference keeps in `GPT.adapter_kv_caches` (adapter.py:103-107): per layer (ak, av) as (1, G, aT, hs), or None."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        out = []
        for L in self.layers:
            if L.adapter is None:
                out.append(None)
            else:
                ak, av = self._adapter_kv(L, stream)
                out.append((ak.unsqueeze(0), av.unsqueeze(0)))
        return out

    def _adapter_attn(self, L, qkv: int, pos_ptr: int, att: int, B: int, T: int, scale: float, stream: int) -> None:
        """att += gating * softmax(q . ak^T * scale) . av over the adaption prompt (adapter.py:251-254)."""
        cfg = self.cfg
        ak, av = self._adapter_kv(L, stream)
        _lib.check(self.lib.lp_adapter_attn(qkv, _ptr(self.cos), _ptr(self.sin), pos_ptr, ak.data_ptr(), av.data_ptr(),
                                            L.adapter["gate"].data_ptr(), att, B, T, cfg.n_head, cfg.n_query_groups, cfg.head_size,
                                            cfg.rope_n_elem, ak.shape[1], scale, self.round, stream), "lp_adapter_attn")

    # ------------------------------------------------------------------ slab images of the fused pairs (load time)
    def _build_slabs(self):
        """Sequential-residual SwiGLU models with GPTQ int4 g128 weights (bf16-exact scales) and 128-wide heads — Llama-2 —
        run attention -> attn.proj and fc -> mlp.proj as column->row pairs INSIDE each CTA of the step kernel (SLAB ops): the
        projections are re-laid out once into per-CTA K-slab images (lp_decode_step_slab_build).  LP_DS_FUSE=0 keeps the five-op
        layer (A/B measurements).  Returns (meta_mlp, meta_att) device tensors, or False."""
        if self._slabs is not None:
            return self._slabs
        self._slabs = False
        cfg, lib = self.cfg, self.lib
        if (os.environ.get("LP_DS_FUSE", "1") == "0" or self.tp is not None or cfg.parallel_residual
                or self.act != _lib.LP_EPI_SWIGLU or cfg.head_size != 128):
            return False
        for L in self.layers:
            for m in (L.proj, L.mlp_proj):
                if (m.fmt != _lib.LP_W_INT4 or m.rec.group != 128 or not (m.rec.flags & _lib.LP_WF_AUX_PACKED) or not m.rec.aux2
                        or m.rec.bias):
                    return False
            if L.fc.fmt != _lib.LP_W_INT4 or L.fc.N != 2 * L.mlp_proj.K:
                return False
        E, I, H = cfg.n_embd, cfg.intermediate_size, cfg.n_head
        metas = []
        for src, (N, K, arg, hs) in enumerate(((E, I, (2 * I) // 16, 0), (E, E, H, cfg.head_size))):
            rec = (_lib.LpSlabMeta * 256)()
            n, nbytes = ctypes.c_int(0), ctypes.c_size_t(0)
            rc = lib.lp_decode_step_slab_layout(src, N, K, arg, hs, rec, 256, ctypes.byref(n), ctypes.byref(nbytes))
            if rc == -2:
                return False
            _lib.check(rc, "lp_decode_step_slab_layout")
            raw = bytes(rec)[: n.value * ctypes.sizeof(_lib.LpSlabMeta)]
            metas.append((torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device), n.value, nbytes.value))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for L in self.layers:
            for name, W, (meta, n, nbytes) in (("mlp_slab", L.mlp_proj, metas[0]), ("att_slab", L.proj, metas[1])):
                img = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
                _lib.check(lib.lp_decode_step_slab_build(W.ref, meta.data_ptr(), n, img.data_ptr(), stream), "lp_decode_step_slab_build")
                setattr(L, name, img)
        self._slabs = (metas[0][0], metas[1][0])
        return self._slabs

    # ------------------------------------------------------------------ batch-1 decode step as one persistent kernel
    def _step_plan(self, b: Dict[str, torch.Tensor], idx_ptr: int, idx64: int, idx_off: Optional[int], pos_ptr: int, caches):
        """Op table of one decode step (same op order as `_run`) -> lp_decode_step_plan.  Returns the handle, or None when the
        library does not cover this model (GQA / MQA, fp32 cache, NF4 / int8 / fp32 weights, ...): the per-op path runs then."""
        cfg, lib = self.cfg, self.lib
        max_seq = caches[0][0].size(2)
        key = (idx_ptr, idx64, idx_off, pos_ptr, caches[0][0].data_ptr(), max_seq, b["x"].data_ptr())
        if key in self._steps:
            ent = self._steps[key]
            return None if ent is None else ent[0]
        if len(self._steps) > 16:
            self._steps.clear()
        E, H, G, hs = cfg.n_embd, cfg.n_head_local, cfg.n_query_groups_local, cfg.head_size
        x, qkv, xmid, u, logits = (b[k].data_ptr() for k in ("x", "qkv", "xmid", "u", "logits"))
        ops: List[_lib.LpStepOp] = []

        def linear(W, src, norm, epi, res, dst, dep, from_attn=False, keep_local=False):
            op = _lib.LpStepOp()
            op.kind, op.dep, op.W = _lib.LP_STEP_LINEAR, dep, ctypes.pointer(W.rec)
            op.keep_local = int(keep_local)
            op.x, op.x_is_attention = (None if from_attn else src), int(from_attn)
            op.norm_kind = -1
            if norm is not None:
                op.norm_kind, op.norm_w, op.norm_b, op.eps = self.norm_kind, _ptr(norm[0]), _ptr(norm[1]), cfg.norm_eps
            op.epilogue, op.residual, op.out = epi, res, dst
            ops.append(op)
            return len(ops) - 1

        def attention(li, dep):
            op = _lib.LpStepOp()
            op.kind, op.dep, op.norm_kind = _lib.LP_STEP_ATTENTION, dep, -1
            op.qkv, op.k_cache, op.v_cache = qkv, caches[li][0].data_ptr(), caches[li][1].data_ptr()
            ops.append(op)
            return len(ops) - 1

        def slab(W, src_kind, image, meta, dep):
            """x += W[:, columns this CTA produced itself] . v: the second half of a column->row pair (SLAB op)."""
            op = _lib.LpStepOp()
            op.kind, op.dep, op.W, op.norm_kind = _lib.LP_STEP_SLAB, dep, ctypes.pointer(W.rec), -1
            op.slab_src, op.slab_image, op.slab_meta, op.out = src_kind, image.data_ptr(), meta.data_ptr(), x
            ops.append(op)
            return len(ops) - 1

        slabs = self._build_slabs()
        tp = self.tp
        n_exch = [0]
        # QKV rows accumulate in place (stream-K split: every SM streams the same number of weight bytes whatever the tile count)
        # into a zeroed buffer of their own per layer; the step's prologue clears them (lp_step_geom.zero_ptr).  LP_DS_QKV_SK=0:
        # whole tiles per CTA into the shared qkv buffer, for A/B.
        qkv_rows = cfg.qkv_rows_local
        qkv_sk = os.environ.get("LP_DS_QKV_SK", "1") != "0" and qkv_rows % 4 == 0
        qkv_all = torch.zeros((cfg.n_layer, qkv_rows), dtype=torch.float32, device=self.device) if qkv_sk else None

        def row_parallel(W, src, dep, from_attn=False):
            """x += src . W^T.  Single GPU: in place (stage-granular split, atomic accumulation).  Tensor parallel: W is a column
            shard, the product a partial sum.  PUSH exchange: the linear op stores its partial into slot [s][rank] of EVERY rank's
            symmetric buffer as {value, epoch} pairs while its tiles finish; the EXCHANGE op polls its slice of those pairs in local
            memory, adds the tp values and the residual (one-shot NVLink all-reduce inside the kernel, one traversal, no fence)."""
            if tp is None:
                return linear(W, src, None, _lib.LP_EPI_RESIDUAL, x, x, dep, from_attn=from_attn)
            slot = n_exch[0] % 2
            n_exch[0] += 1
            slot_base = tp.push_off + slot * tp.size * E * 8  # {value, epoch} pairs: 8 bytes per element
            pad_base = tp.push_pad + slot * tp.size
            state = tp.push_state[slot].data_ptr()
            i_part = linear(W, src, None, _lib.LP_EPI_NONE, None, tp.buf.data_ptr() + slot_base + tp.rank * E * 8, dep, from_attn=from_attn)
            snd = ops[i_part]
            snd.tp_buf_ptrs, snd.tp_pad_ptrs, snd.tp_state = tp.buf_ptrs, tp.pad_ptrs, state
            snd.tp_buf_offset, snd.tp_pad_base, snd.tp_rank, snd.tp_size = slot_base + tp.rank * E * 8, pad_base, tp.rank, tp.size
            op = _lib.LpStepOp()
            op.kind, op.dep, op.norm_kind = _lib.LP_STEP_EXCHANGE, i_part, -1
            op.tp_buf_ptrs, op.tp_pad_ptrs, op.tp_state = tp.buf_ptrs, tp.pad_ptrs, state
            op.tp_buf_offset, op.tp_pad_base, op.tp_rank, op.tp_size = slot_base, pad_base, tp.rank, tp.size
            op.residual, op.out = x, x
            ops.append(op)
            return len(ops) - 1

        last = -1
        qkv_shared = qkv
        for li, L in enumerate(self.layers):
            if qkv_sk:
                qkv = qkv_all[li].data_ptr()
                i_qkv = linear(L.qkv, x, (L.n1_w, L.n1_b), _lib.LP_EPI_RESIDUAL, qkv, qkv, last)
            else:
                qkv = qkv_shared
                i_qkv = linear(L.qkv, x, (L.n1_w, L.n1_b), _lib.LP_EPI_NONE, None, qkv, last)
            if cfg.parallel_residual:
                n2 = (L.n1_w, L.n1_b) if cfg.shared_attention_norm else (L.n2_w, L.n2_b)
                i_fc = linear(L.fc, x, n2, self.act, None, u, last)   # reads the old x: streams right behind the QKV weights
                i_att = attention(li, i_qkv)
                # x += attn.proj(att); x += mlp.proj(u): both in place (atomic accumulation, stage-granular split), no barrier
                if tp is None and os.environ.get("LP_DS_PROJ_LAST", "1") != "0":
                    # mlp.proj FIRST: its input (u) has been complete since fc, so a CTA streams its 4x larger weights the moment it
                    # leaves the attention op, while the slower heads finish; attn.proj, which needs every head's partials (a
                    # grid-wide dependency), runs last, when they have long arrived.  (LP_DS_PROJ_LAST=0: reference order, for A/B.)
                    row_parallel(L.mlp_proj, u, i_fc)
                    last = row_parallel(L.proj, None, i_att, from_attn=True)
                    continue
                i_proj = row_parallel(L.proj, None, i_att, from_attn=True)
                last = row_parallel(L.mlp_proj, u, i_fc if tp is None else max(i_fc, i_proj))
            else:
                if cfg.shared_attention_norm:
                    return self._steps.setdefault(key, None)
                i_att = attention(li, i_qkv)
                if slabs:
                    # three grid-wide dependencies per layer instead of five: the head's CTAs apply attn.proj to their own head,
                    # every CTA applies mlp.proj to its own SwiGLU outputs; both add partial rows into x
                    i_proj = slab(L.proj, 1, L.att_slab, slabs[1], i_att)
                    linear(L.fc, x, (L.n2_w, L.n2_b), self.act, None, u, i_proj, keep_local=True)
                    last = slab(L.mlp_proj, 0, L.mlp_slab, slabs[0], -1)
                    continue
                i_proj = row_parallel(L.proj, None, i_att, from_attn=True)
                i_fc = linear(L.fc, x, (L.n2_w, L.n2_b), self.act, None, u, i_proj)
                last = row_parallel(L.mlp_proj, u, i_fc)
        linear(self.lm_head, x, (self.lnf_w, self.lnf_b), _lib.LP_EPI_NONE, None, logits, last)

        n = len(ops)
        arr = (_lib.LpStepOp * n)(*ops)
        plan = torch.zeros(lib.lp_decode_step_plan_bytes(n), dtype=torch.uint8, device=self.device)
        ws = torch.zeros(max(lib.lp_decode_step_workspace_bytes(H, hs), 16), dtype=torch.uint8, device=self.device)
        gm = _lib.LpStepGeom()
        gm.pos, gm.idx, gm.idx_offset, gm.wte, gm.x0 = pos_ptr, idx_ptr, idx_off, self.wte.data_ptr(), x
        gm.cos, gm.sin = _ptr(self.cos), _ptr(self.sin)
        gm.workspace, gm.workspace_bytes = ws.data_ptr(), ws.numel()
        gm.idx_is_int64, gm.wte_dtype = idx64, _KV_OF_DTYPE[self.wte.dtype]
        gm.E, gm.H, gm.G, gm.hs, gm.n_elem, gm.max_seq = E, H, G, hs, cfg.rope_n_elem, max_seq
        gm.kv_dtype, gm.scale = _KV_OF_DTYPE[caches[0][0].dtype], 1.0 / math.sqrt(hs)
        if qkv_all is not None:
            gm.zero_ptr, gm.zero_bytes = qkv_all.data_ptr(), qkv_all.numel() * 4
        handle = _lib.LpStepHandle()
        rc = lib.lp_decode_step_plan(arr, n, ctypes.byref(gm), plan.data_ptr(), plan.numel(), ctypes.byref(handle))
        if rc == -2:
            self._steps[key] = None
            return None
        _lib.check(rc, "lp_decode_step_plan")
        self._steps[key] = (handle, plan, ws, arr, n, qkv_all)
        return handle

    def _raise_if_flagged(self) -> None:
        f = int(self._flag[0])
        if f:
            self._flag.zero_()
            what = " and ".join(w for bit, w in ((1, "a token id outside the embedding table"), (2, "a position outside the RoPE table"))
                                if f & bit)
            raise IndexError(f"an earlier decode step was given {what} (the value was clamped on the device; the reference raises)")

    def check_step_health(self) -> None:
        """Synchronous: raise if the watchdog of the persistent step kernel fired in any step launched so far (a dependency
        counter or a tensor-parallel peer flag that never arrived, lp_decode_step_status).  generate() calls it once per call."""
        torch.cuda.synchronize(self.device)
        self._raise_if_flagged()
        info = (ctypes.c_int32 * 8)()
        first = None
        for ent in self._steps.values():  # every plan is read (and its record cleared) before anything is raised
            if ent is None:
                continue
            rc = self.lib.lp_decode_step_status(ctypes.byref(ent[0]), info)
            if rc != 0 and first is None:
                first = (rc, list(info))
        if first is not None:
            rc, info = first
            what = {1: "dependency counter", 2: "tensor-parallel peer flag"}.get(info[0], f"code {info[0]}")
            raise RuntimeError(f"liblitparrot_b200: lp_decode_step timed out ({self.lib.lp_status_str(rc).decode()}): {what} of op "
                               f"{info[1]} (dep / peer {info[2]}) never arrived in CTA {info[3]} (value seen {info[4]}); the "
                               "results of that step are invalid")

    def scratch_cache(self, B: int, T: int):
        key = (B, T)
        if self._scratch is None or self._scratch[0] != key:
            cfg = self.cfg
            store = torch.empty((cfg.n_layer, 2, B, cfg.n_query_groups_local, T, cfg.head_size), device=self.device,
                                dtype=self.param_dtype)
            self._scratch = (key, [(store[l, 0], store[l, 1]) for l in range(cfg.n_layer)])
        return self._scratch[1]

    def weight_bytes_per_token(self) -> int:
        """Algorithmic bytes of weights one decode step streams (all linears incl. lm_head + one embedding row)."""
        n = self.lm_head.stored_bytes + self.cfg.n_embd * self.wte.element_size()
        for L in self.layers:
            n += L.qkv.stored_bytes + L.proj.stored_bytes + L.fc.stored_bytes + L.mlp_proj.stored_bytes
        return n

    # ------------------------------------------------------------------ prefill on the tcgen05 GEMM
    def tc_eligible(self, rows: int) -> bool:
        """T > 1 goes through lp_gemm_bf16_tc when every linear is bf16 (fp32-activation accuracy via bf16 term splitting) or,
        in bf16-faithful mode, quantised (weights expanded to bf16 like the reference's bf16 dequantisation)."""
        if rows < 9 or getattr(self, "disable_tc", False) or self.tp is not None:
            return False
        mats = [self.lm_head] + [m for L in self.layers for m in (L.qkv, L.proj, L.fc, L.mlp_proj)]
        if any(m.N % 8 or m.K % 8 for m in mats):
            return False
        if all(m.fmt == _lib.LP_W_BF16 for m in mats):
            return True
        return self.round == 1 and all(m.fmt in (_lib.LP_W_BF16, _lib.LP_W_INT4, _lib.LP_W_NF4, _lib.LP_W_INT8) for m in mats)

    def _bf16_weight(self, W: "PackedLinear", stream: int) -> int:
        """Device pointer of W as dense bf16 [N, K]; quantised formats are expanded into a scratch buffer (stream ordered)."""
        if W.fmt == _lib.LP_W_BF16:
            return W.rec.w
        need = W.N * W.K
        if getattr(self, "_wscratch", None) is None or self._wscratch.numel() < need:
            biggest = max(m.N * m.K for L in self.layers for m in (L.qkv, L.proj, L.fc, L.mlp_proj))
            self._wscratch = torch.empty(max(need, biggest), dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.lp_dequant_bf16(W.ref, self._wscratch.data_ptr(), stream), "lp_dequant_bf16")
        return self._wscratch.data_ptr()

    def buffers(self, rows: int, B: int, T: int, max_seq: int) -> Dict[str, torch.Tensor]:
        key = (B, T, max_seq)  # the attention workspace is sized from (B, T), not from rows alone
        b = self._bufs.get(key)
        if b is None:
            cfg, dev = self.cfg, self.device
            f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)  # noqa: E731
            E, I, V = cfg.n_embd, cfg.intermediate_size_local, cfg.padded_vocab_size
            Hl, Gl = cfg.n_head_local, cfg.n_query_groups_local
            ws_bytes = max(self.lib.lp_attn_workspace_bytes(B, T, Hl, cfg.head_size, max_seq),
                           self.lib.lp_attn_fused_workspace_bytes(B, Hl, Gl, cfg.head_size, max_seq))
            b = dict(x=f(rows, E), n1=f(rows, E), n2=f(rows, E), qkv=f(rows, cfg.qkv_rows_local), q=f(rows, Hl * cfg.head_size),
                     att=f(rows, Hl * cfg.head_size), xmid=f(rows, E), u=f(rows, I), xf=f(rows, E), logits=f(rows, V),
                     ws=torch.zeros(max(ws_bytes, 16), device=dev, dtype=torch.uint8))  # zeroed: split-merge tickets
            if self.tc_eligible(rows):
                # bf16 split terms of the GEMM operands: [3][rows, width]
                bf = lambda w: torch.empty((3, rows, w), device=dev, dtype=torch.bfloat16)  # noqa: E731
                b.update(t_n=bf(E), t_att=bf(Hl * cfg.head_size), t_u=bf(I))
            if len(self._bufs) > 8:
                self._bufs.clear()
            self._bufs[key] = b
        return b

    # ------------------------------------------------------------------ one forward pass = a fixed list of C-ABI calls
    def _run(self, b: Dict[str, torch.Tensor], idx_ptr: int, idx64: int, idx_off: Optional[int], pos_ptr: int,
             caches, B: int, T: int, stream: int, last_only: bool = False) -> None:
        lib, cfg, r, chk = self.lib, self.cfg, self.round, _lib.check
        rows = B * T
        if idx_off is None and self.tc_eligible(rows):
            # wide decode batches (B >= 9) and prefill: projections on the tcgen05 GEMM, weights still streamed once
            return self._run_tc(b, idx_ptr, idx64, pos_ptr, caches, B, T, stream, last_only)
        if (rows == 1 and self.use_step_kernel and r == 0 and getattr(self, "trace", None) is None
                and self.cos is not None and caches[0][0].dtype == torch.bfloat16):
            handle = self._step_plan(b, idx_ptr, idx64, idx_off, pos_ptr, caches)
            if handle is not None:
                chk(lib.lp_decode_step(ctypes.byref(handle), stream), "lp_decode_step")
                return
        E, H, G, hs = cfg.n_embd, cfg.n_head_local, cfg.n_query_groups_local, cfg.head_size
        max_seq = caches[0][0].size(2)
        kvd = _KV_OF_DTYPE[caches[0][0].dtype]
        x, n1, n2, qkv, q, att, xmid, u = (b[k].data_ptr() for k in ("x", "n1", "n2", "qkv", "q", "att", "xmid", "u"))
        tp = self.tp
        if tp is not None:
            if rows * E > tp.slot_floats:
                raise RuntimeError(f"tensor-parallel exchange buffer holds {tp.slot_floats // E} rows; got {rows}")
            tp.begin_forward()

        def row_parallel(src, W, res, dst, what):
            """dst = res + src . W^T.  Tensor parallel: W is a column shard, so the product is a partial sum: it goes to this
            rank's slot of the symmetric buffer and lp_tp_allreduce_residual adds all ranks' partials and the residual."""
            mark()
            if tp is None:
                chk(lib.lp_linear(src, rows, W.ref, _lib.LP_EPI_RESIDUAL, res, dst, r, stream), what)
                return
            slot, part = tp.next_slot()
            chk(lib.lp_linear(src, rows, W.ref, _lib.LP_EPI_NONE, None, part, 0, stream), what)
            chk(lib.lp_tp_allreduce_residual(tp.buf_ptrs, tp.pad_ptrs, tp.rank, tp.size, slot * tp.slot_floats * 4, slot * tp.size,
                                             tp.state[slot].data_ptr(), rows * E, res, dst, r, stream), "lp_tp_allreduce_residual")
        ws, ws_bytes = b["ws"].data_ptr(), b["ws"].numel()
        scale = 1.0 / math.sqrt(hs)
        wte_dt = _KV_OF_DTYPE[self.wte.dtype]
        chk(lib.lp_embed(idx_ptr, idx64, idx_off, self.wte.data_ptr(), wte_dt, x, rows, E, r, stream), "lp_embed")
        trace = getattr(self, "trace", None)  # debug: (ncalls, 148 * 8) int64, one row of stamps per streaming GEMV
        tcount = [0]

        def mark():
            if trace is not None:
                lib.lp_debug_stream_trace(trace[tcount[0] % trace.size(0)].data_ptr())
                tcount[0] += 1

        def norm_linear(src, nw, nb, W, epi, dst, scratch, what):
            mark()
            """norm fused into the GEMV prologue where the kernel supports it, else lp_norm + lp_linear."""
            rc = lib.lp_norm_linear(self.norm_kind, _ptr(nw), _ptr(nb), cfg.norm_eps, src, rows, W.ref, epi, None, dst, r, stream)
            if rc == -2:
                chk(lib.lp_norm(self.norm_kind, src, _ptr(nw), _ptr(nb), cfg.norm_eps, scratch, rows, E, r, stream), "lp_norm")
                rc = lib.lp_linear(scratch, rows, W.ref, epi, None, dst, r, stream)
            chk(rc, what)

        for li, L in enumerate(self.layers):
            kc, vc = caches[li][0].data_ptr(), caches[li][1].data_ptr()
            norm_linear(x, L.n1_w, L.n1_b, L.qkv, _lib.LP_EPI_NONE, qkv, n1, "lp_linear(qkv)")
            rc = -2
            if T == 1:  # one launch: RoPE + cache append + split-K tensor-core attention + merge
                rc = lib.lp_attn_decode_fused(qkv, _ptr(self.cos), _ptr(self.sin), pos_ptr, att, kc, vc, kvd, ws, ws_bytes, B, H, G,
                                              hs, cfg.rope_n_elem, max_seq, scale, r, stream)
                if rc != -2:
                    chk(rc, "lp_attn_decode_fused")
            if rc == -2:
                chk(lib.lp_rope_kv_append(qkv, _ptr(self.cos), _ptr(self.sin), pos_ptr, q, kc, vc, kvd, B, T, H, G, hs,
                                          cfg.rope_n_elem, max_seq, r, stream), "lp_rope_kv_append")
                self._attention_rows(q, kc, vc, kvd, pos_ptr, att, ws, ws_bytes, B, T, max_seq, scale, stream)
            if L.adapter is not None:
                self._adapter_attn(L, qkv, pos_ptr, att, B, T, scale, stream)
            if cfg.parallel_residual:
                # x + attn(n1) + mlp(n2), n2 = n1 when the norm is shared (model.py:169-171); both GEMVs read the old x
                n2w, n2b = (L.n1_w, L.n1_b) if cfg.shared_attention_norm else (L.n2_w, L.n2_b)
                norm_linear(x, n2w, n2b, L.fc, self.act, u, n2, "lp_linear(fc)")
                row_parallel(att, L.proj, x, xmid, "lp_linear(proj)")
                row_parallel(u, L.mlp_proj, xmid, x, "lp_linear(mlp.proj)")
            else:
                if cfg.shared_attention_norm:
                    raise NotImplementedError("No checkpoint amongst the ones we support uses this configuration"
                                              " (non-parallel residual and shared attention norm).")
                row_parallel(att, L.proj, x, x, "lp_linear(proj)")
                norm_linear(x, L.n2_w, L.n2_b, L.fc, self.act, u, n2, "lp_linear(fc)")
                row_parallel(u, L.mlp_proj, x, x, "lp_linear(mlp.proj)")
        xf, logits = b["xf"].data_ptr(), b["logits"].data_ptr()
        if tp is not None and tp.count % 2:  # keep the slot sequence of a (replayed) forward even: see TPContext
            slot, part = tp.next_slot()
            chk(lib.lp_tp_allreduce_residual(tp.buf_ptrs, tp.pad_ptrs, tp.rank, tp.size, slot * tp.slot_floats * 4, slot * tp.size,
                                             tp.state[slot].data_ptr(), 4, None, b["xmid"].data_ptr(), 0, stream), "lp_tp_allreduce_residual")
        if trace is not None:
            lib.lp_debug_stream_trace(None)
        if last_only and T > 1:
            # only the last position of each sequence feeds the sampler (generate/base.py:136): gather those rows
            for bi in range(B):
                src = x + ((bi + 1) * T - 1) * E * 4
                chk(lib.lp_norm(self.norm_kind, src, _ptr(self.lnf_w), _ptr(self.lnf_b), cfg.norm_eps, xf + bi * E * 4, 1, E, r,
                                stream), "lp_norm")
            chk(lib.lp_linear(xf, B, self.lm_head.ref, _lib.LP_EPI_NONE, None, logits, r, stream), "lp_linear(lm_head)")
        else:
            norm_linear(x, self.lnf_w, self.lnf_b, self.lm_head, _lib.LP_EPI_NONE, logits, xf, "lp_linear(lm_head)")

    def _attention_rows(self, q, kc, vc, kvd, pos_ptr, att, ws, ws_bytes, B, T, max_seq, scale, stream) -> None:
        """Attention of T query rows against the cache: the tensor-core causal kernel when the positions are consecutive
        and do not wrap (`self._consecutive`, set by forward()), else the generic exact kernel."""
        lib, cfg, r = self.lib, self.cfg, self.round
        rc = -2
        if T > 1 and self._consecutive:
            rc = lib.lp_attn_prefill(q, kc, vc, kvd, pos_ptr, att, B, T, cfg.n_head_local, cfg.n_query_groups_local, cfg.head_size,
                                     max_seq, scale, r, stream)
            if rc != -2:
                _lib.check(rc, "lp_attn_prefill")
        if rc == -2:
            _lib.check(lib.lp_attn_decode(q, kc, vc, kvd, pos_ptr, att, ws, ws_bytes, B, T, cfg.n_head_local, cfg.n_query_groups_local,
                                          cfg.head_size, max_seq, scale, r, stream), "lp_attn_decode")

    def _run_tc(self, b: Dict[str, torch.Tensor], idx_ptr: int, idx64: int, pos_ptr: int, caches, B: int, T: int, stream: int,
                last_only: bool = False) -> None:
        """T > 1: the same op sequence as `_run`, with every projection on the tcgen05 GEMM (lp_gemm_bf16_tc)."""
        lib, cfg, r, chk = self.lib, self.cfg, self.round, _lib.check
        rows = B * T
        E, H, G, hs, I = cfg.n_embd, cfg.n_head, cfg.n_query_groups, cfg.head_size, cfg.intermediate_size
        max_seq = caches[0][0].size(2)
        kvd = _KV_OF_DTYPE[caches[0][0].dtype]
        x, qkv, q, att, xmid, u = (b[k].data_ptr() for k in ("x", "qkv", "q", "att", "xmid", "u"))
        t_n, t_att, t_u = b["t_n"].data_ptr(), b["t_att"].data_ptr(), b["t_u"].data_ptr()
        ws, ws_bytes = b["ws"].data_ptr(), b["ws"].numel()
        scale = 1.0 / math.sqrt(hs)
        nt = 1 if r else (3 if rows <= 16 else 2)  # bf16 terms per activation (3 terms = all 24 bits; 2 = 16 bits, error 2^-17)
        nk = self.norm_kind
        chk(lib.lp_embed(idx_ptr, idx64, None, self.wte.data_ptr(), _KV_OF_DTYPE[self.wte.dtype], x, rows, E, r, stream), "lp_embed")

        def gemm(terms, W, epi, res, out_f32, out_bf=None):
            chk(lib.lp_gemm_bf16_tc_affine(terms, nt, rows, self._bf16_weight(W, stream), W.N, W.K, W.rec.bias, W.rec.out_bias,
                                           W.rec.out_scale, epi, res, out_f32, out_bf, nt, r, stream), "lp_gemm_bf16_tc")

        def norm_split(src, nw, nb, dst):
            chk(lib.lp_split_bf16(src, dst, rows, E, nt, nk, _ptr(nw), _ptr(nb), cfg.norm_eps, r, stream), "lp_split_bf16")

        for li, L in enumerate(self.layers):
            kc, vc = caches[li][0].data_ptr(), caches[li][1].data_ptr()
            norm_split(x, L.n1_w, L.n1_b, t_n)
            gemm(t_n, L.qkv, _lib.LP_EPI_NONE, None, qkv)
            rc = -2
            if T == 1:
                rc = lib.lp_attn_decode_fused(qkv, _ptr(self.cos), _ptr(self.sin), pos_ptr, att, kc, vc, kvd, ws, ws_bytes, B, H, G, hs,
                                              cfg.rope_n_elem, max_seq, scale, r, stream)
                if rc != -2:
                    chk(rc, "lp_attn_decode_fused")
            if rc == -2:
                chk(lib.lp_rope_kv_append(qkv, _ptr(self.cos), _ptr(self.sin), pos_ptr, q, kc, vc, kvd, B, T, H, G, hs, cfg.rope_n_elem,
                                          max_seq, r, stream), "lp_rope_kv_append")
                self._attention_rows(q, kc, vc, kvd, pos_ptr, att, ws, ws_bytes, B, T, max_seq, scale, stream)
            if L.adapter is not None:
                self._adapter_attn(L, qkv, pos_ptr, att, B, T, scale, stream)
            chk(lib.lp_split_bf16(att, t_att, rows, E, nt, -1, None, None, 0.0, 0, stream), "lp_split_bf16")
            if cfg.parallel_residual:
                if not cfg.shared_attention_norm:
                    norm_split(x, L.n2_w, L.n2_b, t_n)  # shared norm: t_n already holds norm_1(x)
                gemm(t_n, L.fc, self.act, None, None, t_u)
                if r == 0:  # x += attn.proj(att); x += mlp.proj(u), in place (decode batches: split-K with atomic accumulation)
                    gemm(t_att, L.proj, _lib.LP_EPI_RESIDUAL, x, x)
                    gemm(t_u, L.mlp_proj, _lib.LP_EPI_RESIDUAL, x, x)
                else:  # bf16-faithful: the reference rounds (x + h) before adding the MLP branch (model.py:171)
                    gemm(t_att, L.proj, _lib.LP_EPI_RESIDUAL, x, xmid)
                    gemm(t_u, L.mlp_proj, _lib.LP_EPI_RESIDUAL, xmid, x)
            else:
                if cfg.shared_attention_norm:
                    raise NotImplementedError("No checkpoint amongst the ones we support uses this configuration"
                                              " (non-parallel residual and shared attention norm).")
                gemm(t_att, L.proj, _lib.LP_EPI_RESIDUAL, x, x)
                norm_split(x, L.n2_w, L.n2_b, t_n)
                gemm(t_n, L.fc, self.act, None, None, t_u)
                gemm(t_u, L.mlp_proj, _lib.LP_EPI_RESIDUAL, x, x)
        xf, logits = b["xf"].data_ptr(), b["logits"].data_ptr()
        if last_only and T > 1:
            # only the last position of each sequence feeds the sampler (generate/base.py:136)
            for bi in range(B):
                src = x + ((bi + 1) * T - 1) * E * 4
                chk(lib.lp_norm(nk, src, _ptr(self.lnf_w), _ptr(self.lnf_b), cfg.norm_eps, xf + bi * E * 4, 1, E, r, stream), "lp_norm")
            chk(lib.lp_linear(xf, B, self.lm_head.ref, _lib.LP_EPI_NONE, None, logits, r, stream), "lp_linear(lm_head)")
        elif True:
            norm_split(x, self.lnf_w, self.lnf_b, t_n)
            gemm(t_n, self.lm_head, _lib.LP_EPI_NONE, None, logits)
        else:
            chk(lib.lp_norm(nk, x, _ptr(self.lnf_w), _ptr(self.lnf_b), cfg.norm_eps, xf, rows, E, r, stream), "lp_norm")
            chk(lib.lp_linear(xf, rows, self.lm_head.ref, _lib.LP_EPI_NONE, None, logits, r, stream), "lp_linear(lm_head)")

    # ------------------------------------------------------------------ public entry used by GPT.forward
    def _check_caches(self, caches, B: int) -> None:
        k = caches[0][0]
        cfg = self.cfg
        if k.dim() != 4 or k.size(0) != B or k.size(1) != cfg.n_query_groups_local or k.size(3) != cfg.head_size:
            raise RuntimeError(f"kv cache shape {tuple(k.shape)} does not match (B={B}, G={cfg.n_query_groups_local}, max_seq, "
                               f"hs={cfg.head_size}); call model.reset_cache() when the batch size changes")
        if k.dtype not in _KV_OF_DTYPE or not k.is_contiguous():
            raise RuntimeError("kv cache must be a contiguous float32 / bfloat16 tensor")

    def forward(self, idx: torch.Tensor, pos: torch.Tensor, caches, allow_graph: bool = True,
                last_only: bool = False, raw_logits: bool = False) -> torch.Tensor:
        B, T = idx.shape
        cfg = self.cfg
        self._check_caches(caches, B)
        if pos.numel() != T:
            raise ValueError(f"input_pos has {pos.numel()} entries for a sequence of length {T}")
        max_seq = caches[0][0].size(2)
        V = cfg.padded_vocab_size
        if T == 1 and allow_graph:
            return self._forward_graph(idx, pos, caches, B, max_seq)
        b = self.buffers(B * T, B, T, max_seq)
        if idx.dtype not in (torch.int32, torch.int64):
            idx = idx.long()
        idx = idx.contiguous()
        pos32 = pos.to(device=self.device, dtype=torch.int32).contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._consecutive = False
        # one host read per call on this (not replayed) path: the reference raises on token ids outside the embedding table
        # (nn.Embedding, model.py:99) and on positions outside the RoPE table (index_select, model.py:88-92); the kernels index
        # with both.  Are the positions p, p+1, ... without wrapping the cache?
        lim = torch.stack((idx.min(), idx.max())).tolist()
        ph = pos32.cpu()
        if lim[0] < 0 or lim[1] >= V:
            raise IndexError(f"token id out of range: ids span [{int(lim[0])}, {int(lim[1])}], the embedding has {V} rows")
        if int(ph.min()) < 0 or int(ph.max()) >= cfg.block_size:
            raise IndexError(f"input_pos out of range: positions span [{int(ph.min())}, {int(ph.max())}], block_size is {cfg.block_size}")
        if T > 1:
            self._consecutive = bool((ph[1:] - ph[:-1] == 1).all()) and int(ph[-1]) < max_seq
        self._run(b, idx.data_ptr(), int(idx.dtype == torch.int64), None, pos32.data_ptr(), caches, B, T, stream, last_only)
        out = b["logits"][:B].view(B, 1, V) if (last_only and T > 1) else b["logits"].view(B, T, V)
        return out if raw_logits else out.to(self.param_dtype, copy=True)

    # ------------------------------------------------------------------ device-resident generate loop
    def gen_state(self, capacity: int) -> Dict[str, torch.Tensor]:
        st = getattr(self, "_gen", None)
        if st is None or st["seq"].numel() < capacity:
            i32 = lambda n: torch.zeros(n, dtype=torch.int32, device=self.device)  # noqa: E731
            st = dict(seq=i32(capacity + 1), pos=i32(1), step=i32(1), tok=i32(1))
            self._gen = st
            self._graphs = {k: v for k, v in self._graphs.items() if k[0] != "gen"}
        return st

    def decode_step(self, caches, temperature: float, top_k: int, seed: int, B: int = 1):
        """Returns a callable that runs ONE decode step: embed(seq[pos]) .. lm_head -> sample -> seq[pos+1], pos += 1.
        Captured once per (cache, sampling parameters); position, tokens and the Philox step live in device memory.
        B > 1 (lock-step batch, model.py:66): the B sampled ids in `tok` are the next step's input; no history is kept."""
        self._check_caches(caches, B)
        max_seq = caches[0][0].size(2)
        st = self._gen
        if st["tok"].numel() < B:
            st["tok"] = torch.zeros(B, dtype=torch.int32, device=self.device)
        key = ("gen", B, max_seq, caches[0][0].data_ptr(), temperature, top_k, seed, st["seq"].data_ptr(), st["tok"].data_ptr())
        g = self._graphs.get(key)
        if g is None:
            b = self.buffers(B, B, 1, max_seq)
            V = self.cfg.padded_vocab_size

            def step_fn(stream: int) -> None:
                if B == 1:
                    self._run(b, st["seq"].data_ptr(), 0, st["pos"].data_ptr(), st["pos"].data_ptr(), caches, 1, 1, stream)
                    seq = st["seq"].data_ptr()
                else:
                    self._run(b, st["tok"].data_ptr(), 0, None, st["pos"].data_ptr(), caches, B, 1, stream)
                    seq = None
                _lib.check(self.lib.lp_sample(b["logits"].data_ptr(), B, V, temperature, top_k, seed, st["step"].data_ptr(),
                                              st["tok"].data_ptr(), seq, st["pos"].data_ptr(), stream), "lp_sample")

            if self.use_graph:
                # warm-up = one real step (loads the M=1 kernels before capture); it is then undone by restoring the
                # device-side position / Philox step, so the first replay recomputes exactly the same token and KV slot
                saved = (st["pos"].clone(), st["step"].clone(), st["tok"].clone())
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    step_fn(side.cuda_stream)
                    st["pos"].copy_(saved[0])
                    st["step"].copy_(saved[1])
                    st["tok"].copy_(saved[2])
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    step_fn(torch.cuda.current_stream(self.device).cuda_stream)
                g = graph.replay
            else:
                g = lambda: step_fn(torch.cuda.current_stream(self.device).cuda_stream)  # noqa: E731
            self._graphs[key] = g
        return g

    def _forward_graph(self, idx, pos, caches, B: int, max_seq: int) -> torch.Tensor:
        """T == 1: the whole step (embed .. lm_head) is captured once per (B, cache) and replayed."""
        V = self.cfg.padded_vocab_size
        key = ("fwd", B, max_seq, caches[0][0].data_ptr())
        g = self._graphs.get(key)
        if g is None:
            b = self.buffers(B, B, 1, max_seq)
            s_idx = torch.zeros(B, dtype=torch.int64, device=self.device)
            s_pos = torch.zeros(1, dtype=torch.int32, device=self.device)
            # the warm-up run inside _capture executes the real step (it writes this position's KV slot)
            s_idx.copy_(idx.reshape(-1))
            s_pos.copy_(pos.reshape(-1))
            graph = self._capture(lambda st: self._run(b, s_idx.data_ptr(), 1, None, s_pos.data_ptr(), caches, B, 1, st))
            g = (graph, b, s_idx, s_pos)
            self._graphs[key] = g
        graph, b, s_idx, s_pos = g
        self._raise_if_flagged()  # an out-of-range id / position of an EARLIER replay (seen without a sync: mapped host flag)
        # replayed path, inputs stay on the device: ONE launch copies them into the graph's static buffers and range-checks /
        # clamps them there (lp_stage_inputs); a violation is reported at the next call
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if (idx.dtype in (torch.int32, torch.int64) and pos.dtype in (torch.int32, torch.int64) and idx.is_contiguous()
                and pos.is_contiguous() and idx.device == self.device and pos.device == self.device):
            _lib.check(self.lib.lp_stage_inputs(idx.data_ptr(), int(idx.dtype == torch.int64), B, pos.data_ptr(),
                                                int(pos.dtype == torch.int64), 1, s_idx.data_ptr(), s_pos.data_ptr(), V,
                                                self.cfg.block_size, self._flag.data_ptr(), stream), "lp_stage_inputs")
        else:
            s_idx.copy_(idx.reshape(-1))
            s_pos.copy_(pos.reshape(-1))
            _lib.check(self.lib.lp_validate_inputs(s_idx.data_ptr(), 1, B, V, s_pos.data_ptr(), 1, self.cfg.block_size,
                                                   self._flag.data_ptr(), stream), "lp_validate_inputs")
        if graph is None:
            self._run(b, s_idx.data_ptr(), 1, None, s_pos.data_ptr(), caches, B, 1, torch.cuda.current_stream(self.device).cuda_stream)
        else:
            graph.replay()
        return b["logits"].view(B, 1, V).to(self.param_dtype, copy=True)

    def _capture(self, fn):
        """Warm up on a side stream, then capture `fn(stream)` into a CUDA graph (None if graphs are disabled)."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            fn(side.cuda_stream)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if not self.use_graph:
            return None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn(torch.cuda.current_stream(self.device).cuda_stream)
        return graph
