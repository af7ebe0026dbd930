"""GPTQ quantiser on the device — drop-in surface of the reference's ``quantize/gptq.py:267-609``:
``GPTQQuantizer`` (same constructor keywords, ``collect_input_stats`` / ``find_params_weight`` / ``quantize_weight`` /
``quantize``), ``blockwise_quantization(model, sample_inputs, working_device, bits=4, groupsize=-1)`` and ``main``.

Underneath (B200-first, csrc/gptq.cu):
  * the Hessian ``H += 2/n X^T X`` of every layer is one tcgen05 GEMM per calibration batch with the tokens as the reduction
    dimension (``lp_gptq_hessian_update``: transpose + bf16 term split, the pairwise products side by side along the reduction);
  * the column sweep runs one warp per weight row with the 128-column block in registers (``lp_gptq_block_sweep``), the error
    feedback into the columns right of the block is ``lp_gptq_trailing_update``, the grids come from ``lp_gptq_find_params``;
  * the three Cholesky factorisations (gptq.py:387-391) are cuSOLVER calls through ``torch.linalg`` — a one-off O(K^3) step;
  * the reference re-runs a Block's *module* forward with hooks to collect the inputs of each linear layer (gptq.py:494-507).  The
    modules here own parameters only, so ``_BlockRunner`` is the per-block entry: it issues the block's kernels (lp_norm, GEMV /
    tcgen05 GEMM, RoPE + causal attention, SwiGLU / GELU, residuals) over a batch of calibration sequences and hands the input of
    the requested layer to the quantiser — same observation order as the reference: a layer sees the outputs of the layers
    quantised before it.
"""
import gc
import json
import math
import sys
import time
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch

from lit_parrot_b200 import _lib
from lit_parrot_b200.engine import _KV_OF_DTYPE, _f32, _ptr, pack_linear
from lit_parrot_b200.quantize import ColBlockQuantizedLinear


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class GPTQQuantizer:
    """quantize/gptq.py:267-431.  The algorithm is Frantar et al., GPTQ (arXiv:2210.17323)."""

    def __init__(self, linear_module, *, bits, perchannel=True, sym=False, blocksize=128, percdamp=0.01, groupsize=-1,
                 actorder=False, hessian_terms: int = 3):
        assert isinstance(linear_module, torch.nn.Linear)
        self.linear_module = linear_module
        self.dev = linear_module.weight.device
        if self.dev.type != "cuda":
            raise RuntimeError("lit_parrot_b200.GPTQQuantizer runs on the GPU (sm_100a) only; move the layer to CUDA first")
        if not perchannel:
            raise NotImplementedError("perchannel=False (one grid for the whole matrix) is not used by quantize/gptq.py's callers")
        if blocksize > 128 or blocksize <= 0:
            raise NotImplementedError("the sweep kernel holds a block of at most 128 columns per warp (the reference's default)")
        self.lib = _lib.init(self.dev.index if self.dev.index is not None else torch.cuda.current_device())
        self.rows, self.columns = linear_module.weight.shape
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0
        self.bits = bits
        self.maxq = 2 ** bits - 1
        self.perchannel = perchannel
        self.sym = sym
        self.blocksize = blocksize
        self.percdamp = percdamp
        self.groupsize = groupsize
        self.actorder = actorder
        self.hessian_terms = hessian_terms  # bf16 terms per activation in the Hessian GEMM: 3 = all 24 bits, 2 = 16 bits
        self.tile_cols = self.columns if groupsize == -1 else groupsize
        n_groups = (self.columns + self.tile_cols - 1) // self.tile_cols
        self.scales = torch.zeros((self.rows, n_groups), dtype=linear_module.weight.dtype, device=self.dev)
        self.zeros = torch.zeros_like(self.scales)
        assert not (self.actorder and self.groupsize != -1), "The permutation trick does not work for grouped quantization"
        self._ws: Optional[torch.Tensor] = None

    # ---- gptq.py:313-316 -----------------------------------------------------------------------------------------
    @staticmethod
    def quantize_weight(x, scale, zero, maxq):
        q = torch.clamp(torch.round(x / scale) + zero, 0, maxq)
        return scale * (q - zero)

    # ---- gptq.py:318-347 (per-channel): one launch of lp_gptq_find_params ------------------------------------------
    def find_params_weight(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        shape = x.shape
        x2 = x.flatten(1).float().contiguous()
        scale = torch.empty((x2.shape[0], 1), device=x2.device)
        zero = torch.empty_like(scale)
        _lib.check(self.lib.lp_gptq_find_params(x2.data_ptr(), x2.shape[0], x2.shape[1], 0, 1, x2.shape[1], self.maxq, int(self.sym),
                                                scale.data_ptr(), zero.data_ptr(), 1, _stream(x2.device)), "lp_gptq_find_params")
        shape = [-1] + [1] * (len(shape) - 1)
        return scale.reshape(shape), zero.reshape(shape)

    # ---- gptq.py:349-363 -----------------------------------------------------------------------------------------
    def collect_input_stats(self, _1, inp, _2):
        inp = inp[0].detach()
        self.last_inp = inp
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).float().contiguous()
        keep = self.nsamples / (self.nsamples + tmp)
        self.nsamples += tmp
        rows = x.shape[0]
        need = self.lib.lp_gptq_hessian_workspace_bytes(rows, self.columns, self.hessian_terms)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.lp_gptq_hessian_update(self.H.data_ptr(), self.columns, x.data_ptr(), rows, keep,
                                                   math.sqrt(2 / self.nsamples), self.hessian_terms, self._ws.data_ptr(),
                                                   self._ws.numel(), _stream(self.dev)), "lp_gptq_hessian_update")

    # ---- gptq.py:365-431 -----------------------------------------------------------------------------------------
    def quantize(self):
        lib, dev, st = self.lib, self.dev, _stream(self.dev)
        N, K = self.rows, self.columns
        W = self.linear_module.weight.detach().to(dtype=torch.float, copy=True).contiguous()
        g = self.tile_cols
        n_groups = self.scales.shape[1]
        sc = torch.empty((N, n_groups), device=dev)  # fp32 grids the sweep uses (the stored buffers are in the weight's dtype)
        ze = torch.empty_like(sc)
        # scale, zero = self.find_params_weight(W): the per-row grid over the WHOLE row (kept for groupsize == -1)
        _lib.check(lib.lp_gptq_find_params(W.data_ptr(), N, K, 0, 1, K, self.maxq, int(self.sym), sc.data_ptr(), ze.data_ptr(),
                                           n_groups, st), "lp_gptq_find_params")
        if self.groupsize != -1:
            row_sc, row_ze = sc[:, :1].clone(), ze[:, :1].clone()
            self.scales[:] = row_sc
            self.zeros[:] = row_ze
        H = self.H
        del self.H
        self._ws = None
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        W[:, dead] = 0
        if self.actorder:
            perm = torch.argsort(torch.diag(H), descending=True)
            W = W[:, perm].contiguous()
            H = H[perm][:, perm]
        damp = self.percdamp * torch.mean(torch.diag(H))
        diag = torch.arange(K, device=dev)
        H[diag, diag] += damp
        H = torch.linalg.cholesky(H)
        H = torch.cholesky_inverse(H)
        Hinv = torch.linalg.cholesky(H, upper=True).contiguous()
        del H
        Q = torch.zeros_like(W)
        Err = torch.zeros((N, 128), device=dev)
        loss = torch.zeros(N, device=dev)
        for i1 in range(0, K, self.blocksize):
            count = min(self.blocksize, K - i1)
            if self.groupsize != -1:
                # grids of the groups that START inside this block, from the current W (gptq.py:409-412)
                g0 = (i1 + g - 1) // g
                g1 = min(n_groups, (i1 + count - 1) // g + 1)
                if g1 > g0:
                    _lib.check(lib.lp_gptq_find_params(W.data_ptr(), N, K, g0, g1 - g0, g, self.maxq, int(self.sym), sc.data_ptr(),
                                                       ze.data_ptr(), n_groups, st), "lp_gptq_find_params")
            sweep_group = g if self.groupsize != -1 else K
            _lib.check(lib.lp_gptq_block_sweep(W.data_ptr(), N, K, i1, count, Hinv.data_ptr(), sc.data_ptr(), ze.data_ptr(), n_groups,
                                               sweep_group, self.maxq, Q.data_ptr(), Err.data_ptr(), loss.data_ptr(), st),
                       "lp_gptq_block_sweep")
            _lib.check(lib.lp_gptq_trailing_update(W.data_ptr(), N, K, i1, count, Hinv.data_ptr(), Err.data_ptr(), st),
                       "lp_gptq_trailing_update")
        if self.groupsize != -1:
            self.scales[:] = sc
            self.zeros[:] = ze
        else:
            self.scales[:] = sc[:, :1]
            self.zeros[:] = ze[:, :1]
        if self.actorder:
            invperm = torch.argsort(perm)
            Q = Q[:, invperm]
        weight = Q.reshape(self.linear_module.weight.shape).to(self.linear_module.weight.data.dtype)
        error = torch.sum(loss).item()
        q_module = ColBlockQuantizedLinear(self.linear_module.in_features, self.linear_module.out_features,
                                           self.linear_module.bias is not None, bits=self.bits, tile_cols=self.groupsize,
                                           device=self.dev, dtype=self.scales.dtype)
        q_module.scales = self.scales
        q_module.zeros = self.zeros
        # pack_weight (gptq.py:233-241) recovers the integer codes as trunc(weight / scales + zeros).  With fp32 grids that is
        # exact (|code - zero| <= 15: the quotient rounds back to the integer), so the codes are recovered from the fp32 values
        # and the fp32 grids the sweep used.  (Under bf16-true the reference divides the bf16-rounded values by the bf16-rounded
        # scales and truncates, which lowers a code by one whenever the quotient lands just below the integer — an artefact of
        # its float round trip that is not reproduced; for fp32 models the codes are identical to the reference's.)
        cols = torch.arange(K, device=dev) // (g if self.groupsize != -1 else K)
        codes = (Q / sc[:, cols] + ze[:, cols]).clamp_(0, self.maxq).round_().to(torch.uint8)
        q_module.quant_weight.copy_(codes[:, 0::2] | (codes[:, 1::2] << 4))
        del weight
        q_module.bias = None if self.linear_module.bias is None else self.linear_module.bias.detach()
        return q_module, error


# ----------------------------------------------------------------------------------------------------------------------
# per-block entry for the calibration passes
# ----------------------------------------------------------------------------------------------------------------------
class _BlockRunner:
    """Block.forward (model.py:158-180) without a KV cache over a batch of calibration sequences, issued op by op through the C
    ABI, with the input of one named linear layer handed out (`capture`)."""

    def __init__(self, model, device: torch.device, T: int) -> None:
        self.cfg = cfg = model.config
        self.dev = device
        self.lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
        cos, sin = model.build_rope_cache(torch.zeros(1, 1, device=device))
        n_elem = cfg.rope_n_elem
        if n_elem % 2:
            raise NotImplementedError("odd rotary dimension")
        self.cos, self.sin = cos.float().contiguous(), sin.float().contiguous()
        self.norm_kind = _lib.LP_NORM_RMS if cfg._norm_class == "RMSNorm" else _lib.LP_NORM_LAYERNORM
        self.kv_dtype = torch.float32 if model.transformer.wte.weight.dtype == torch.float32 else torch.bfloat16
        self.pos = torch.arange(T, device=device, dtype=torch.int32)
        self.T = T
        self._packed: Dict[int, Tuple] = {}
        self._wscratch: Optional[torch.Tensor] = None
        self._bufs: Dict[Tuple, Dict[str, torch.Tensor]] = {}

    def _pack(self, mod):
        ent = self._packed.get(id(mod))
        if ent is None or ent[0] is not mod:
            ent = (mod, pack_linear(mod))
            self._packed[id(mod)] = ent
        return ent[1]

    def _linear(self, src: torch.Tensor, mod, epi: int, res: Optional[torch.Tensor], dst: torch.Tensor, terms: torch.Tensor) -> None:
        """dst = epilogue(src . W^T + b [, res]) for `rows` calibration tokens: tcgen05 GEMM for bf16 / int4 weights (the int4 ones
        expanded to bf16 like the reference's own bf16 dequantisation), streaming GEMV in row chunks for fp32 weights."""
        lib, st = self.lib, _stream(self.dev)
        W = self._pack(mod)
        rows = src.shape[0]
        # bf16 checkpoints (kv_dtype bf16): tensor cores; fp32 checkpoints keep the exact fp32 arithmetic of the GEMV kernels
        if self.kv_dtype == torch.bfloat16 and rows >= 16 and W.fmt in (_lib.LP_W_BF16, _lib.LP_W_INT4) and W.N % 8 == 0 and W.K % 8 == 0:
            nt = 2
            wp, rc = W.rec.w, 0
            if W.fmt != _lib.LP_W_BF16:
                if self._wscratch is None or self._wscratch.numel() < W.N * W.K:
                    self._wscratch = torch.empty(W.N * W.K, dtype=torch.bfloat16, device=self.dev)
                rc = lib.lp_dequant_bf16(W.ref, self._wscratch.data_ptr(), st)
                wp = self._wscratch.data_ptr()
            if rc == 0:
                _lib.check(lib.lp_split_bf16(src.data_ptr(), terms.data_ptr(), rows, W.K, nt, -1, None, None, 0.0, 0, st), "lp_split_bf16")
                rc = lib.lp_gemm_bf16_tc(terms.data_ptr(), nt, rows, wp, W.N, W.K, W.rec.bias, epi, _ptr(res), dst.data_ptr(), None, 0, 0, st)
            if rc != -2:
                _lib.check(rc, "lp_gemm_bf16_tc")
                return
        _lib.check(lib.lp_linear(src.data_ptr(), rows, W.ref, epi, _ptr(res), dst.data_ptr(), 0, st), "lp_linear")

    def _buffers(self, B: int) -> Dict[str, torch.Tensor]:
        cfg, T = self.cfg, self.T
        key = (B,)
        b = self._bufs.get(key)
        if b is None:
            rows = B * T
            f = lambda *s: torch.empty(s, device=self.dev, dtype=torch.float32)  # noqa: E731
            E, I, H, G, hs = cfg.n_embd, cfg.intermediate_size, cfg.n_head, cfg.n_query_groups, cfg.head_size
            kv_dtype = self.kv_dtype  # fp32 models: exact fp32 k / v like the reference's cache-less forward; bf16 models: bf16
            ws = max(self.lib.lp_attn_workspace_bytes(B, T, H, hs, T), 16)
            b = dict(n1=f(rows, E), n2=f(rows, E), qkv=f(rows, cfg.qkv_rows), q=f(rows, H * hs), att=f(rows, H * hs), x1=f(rows, E),
                     u=f(rows, I), u2=f(rows, I), out=f(rows, E), kc=torch.zeros((B, G, T, hs), device=self.dev, dtype=kv_dtype),
                     vc=torch.zeros((B, G, T, hs), device=self.dev, dtype=kv_dtype), ws=torch.zeros(ws, dtype=torch.uint8, device=self.dev),
                     terms=torch.empty((3, rows, max(E, I, H * hs)), device=self.dev, dtype=torch.bfloat16))
            self._bufs = {key: b}
        return b

    def run(self, blk, x: torch.Tensor, capture: Optional[str] = None):
        """x fp32 [B, T, E] -> (block output fp32 [B, T, E], input of the layer named `capture` as [B * T, in_features] or None)."""
        cfg, lib, st = self.cfg, self.lib, _stream(self.dev)
        B, T, E = x.shape
        assert T == self.T
        rows = B * T
        b = self._buffers(B)
        H, G, hs = cfg.n_head, cfg.n_query_groups, cfg.head_size
        xr = x.reshape(rows, E)
        got = None

        def norm(mod, src, dst):
            _lib.check(lib.lp_norm(self.norm_kind, src.data_ptr(), _f32(mod.weight).data_ptr(), _ptr(_f32(getattr(mod, "bias", None))),
                                   cfg.norm_eps, dst.data_ptr(), rows, E, 0, st), "lp_norm")

        norm(blk.norm_1, xr, b["n1"])
        if capture == "attn.attn":
            got = b["n1"]
        self._linear(b["n1"], blk.attn.attn, _lib.LP_EPI_NONE, None, b["qkv"], b["terms"])
        kvd = _KV_OF_DTYPE[b["kc"].dtype]
        scale = 1.0 / math.sqrt(hs)
        _lib.check(lib.lp_rope_kv_append(b["qkv"].data_ptr(), self.cos.data_ptr(), self.sin.data_ptr(), self.pos.data_ptr(),
                                         b["q"].data_ptr(), b["kc"].data_ptr(), b["vc"].data_ptr(), kvd, B, T, H, G, hs, cfg.rope_n_elem,
                                         T, 0, st), "lp_rope_kv_append")
        rc = -2
        if T > 1:
            rc = lib.lp_attn_prefill(b["q"].data_ptr(), b["kc"].data_ptr(), b["vc"].data_ptr(), kvd, self.pos.data_ptr(), b["att"].data_ptr(),
                                     B, T, H, G, hs, T, scale, 0, st)
        if rc == -2:
            rc = lib.lp_attn_decode(b["q"].data_ptr(), b["kc"].data_ptr(), b["vc"].data_ptr(), kvd, self.pos.data_ptr(), b["att"].data_ptr(),
                                    b["ws"].data_ptr(), b["ws"].numel(), B, T, H, G, hs, T, scale, 0, st)
        _lib.check(rc, "attention")
        if capture == "attn.proj":
            got = b["att"]
        # x1 = x + attn.proj(att)
        self._linear(b["att"], blk.attn.proj, _lib.LP_EPI_RESIDUAL, xr, b["x1"], b["terms"])
        if cfg.parallel_residual:
            if cfg.shared_attention_norm:
                n2 = b["n1"]
            else:
                norm(blk.norm_2, xr, b["n2"])
                n2 = b["n2"]
        else:
            if cfg.shared_attention_norm:
                raise NotImplementedError("No checkpoint amongst the ones we support uses this configuration"
                                          " (non-parallel residual and shared attention norm).")
            norm(blk.norm_2, b["x1"], b["n2"])
            n2 = b["n2"]
        if capture in ("mlp.fc", "mlp.fc_1", "mlp.fc_2"):
            got = n2
        if cfg._mlp_class == "LLaMAMLP":
            self._linear(n2, blk.mlp.fc_1, _lib.LP_EPI_NONE, None, b["u"], b["terms"])
            self._linear(n2, blk.mlp.fc_2, _lib.LP_EPI_NONE, None, b["u2"], b["terms"])
            _lib.check(lib.lp_swiglu(b["u"].data_ptr(), b["u2"].data_ptr(), b["u"].data_ptr(), b["u"].numel(), 0, st), "lp_swiglu")
        else:
            self._linear(n2, blk.mlp.fc, _lib.LP_EPI_GELU, None, b["u"], b["terms"])
        if capture == "mlp.proj":
            got = b["u"]
        self._linear(b["u"], blk.mlp.proj, _lib.LP_EPI_RESIDUAL, b["x1"], b["out"], b["terms"])  # (x + h) + mlp, both residual forms
        return b["out"].view(B, T, E), got


def get_sample_data():
    """gptq.py:434-439: 2000 random C4 documents.  Needs the `datasets` package and network access."""
    from datasets import load_dataset

    traindata = load_dataset("allenai/c4", "allenai--c4", data_files={"train": "en/c4-train.00000-of-01024.json.gz"}, split="train")
    return "\n".join(traindata[i]["text"] for i in torch.randperm(len(traindata))[:2000].tolist())


@torch.no_grad()
def blockwise_quantization(model, sample_inputs, working_device, *, bits=4, groupsize=-1, batch: int = 8, verbose: bool = True,
                           _on_layer=None):
    """gptq.py:442-548: quantise every linear layer of the model in order; a layer's statistics are collected from the outputs of
    the layers quantised before it.  The whole model stays on `working_device` (the reference shuttles blocks between the CPU and
    the GPU to fit 40 GB; a B200 holds the models this path is used for).  `_on_layer(key, q_module)` (test hook) sees every
    quantised layer and may return the module to install instead (teacher forcing with another quantiser's result)."""
    say = print if verbose else (lambda *a, **k: None)
    dev = torch.device(working_device)
    if dev.type != "cuda":
        raise RuntimeError("blockwise_quantization runs on a CUDA (sm_100a) device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    model.to(dev)
    if hasattr(model, "invalidate"):
        model.invalidate()
    cfg = model.config
    lib = _lib.init(dev.index)
    sample_inputs = sample_inputs.to(dev)
    n, T = sample_inputs.shape
    E = cfg.n_embd
    wte = model.transformer.wte.weight.data
    idx = sample_inputs.reshape(-1).contiguous()
    if idx.dtype not in (torch.int32, torch.int64):
        idx = idx.long()
    inps = torch.empty((n, T, E), device=dev, dtype=torch.float32)
    _lib.check(lib.lp_embed(idx.data_ptr(), int(idx.dtype == torch.int64), None, wte.data_ptr(), _KV_OF_DTYPE[wte.dtype], inps.data_ptr(),
                            n * T, E, 0, _stream(dev)), "lp_embed")
    outs = torch.zeros_like(inps)
    runner = _BlockRunner(model, dev, T)
    names = ["attn.attn", "attn.proj", "mlp.proj"] + (["mlp.fc"] if cfg._mlp_class == "GptNeoxMLP" else ["mlp.fc_1", "mlp.fc_2"])
    say("Starting to quantize blocks")
    for i, block in enumerate(model.transformer.h):
        for name in names:
            t0 = time.perf_counter()
            module = block.get_submodule(name)
            gptq = GPTQQuantizer(module, bits=bits, groupsize=groupsize, actorder=(groupsize == -1))
            for j in range(0, n, batch):
                _, x_in = runner.run(block, inps[j:j + batch], capture=name)
                K = x_in.shape[-1]
                gptq.collect_input_stats(None, (x_in.view(-1, T, K),), None)
            q_module, error = gptq.quantize()
            if _on_layer is not None:
                q_module = _on_layer(f"transformer.h.{i}.{name}", q_module) or q_module
            pname, dname = name.rsplit(".", 1)
            setattr(block.get_submodule(pname), dname, q_module)
            del gptq
            say(f"{i} {name} collecting stats quantizing time {int(time.perf_counter() - t0 + 0.5)}s quantization error {error:.1f}")
        for j in range(0, n, batch):
            out, _ = runner.run(block, inps[j:j + batch])
            outs[j:j + batch].copy_(out)
        runner._packed.clear()
        gc.collect()
        inps, outs = outs, inps  # the outputs are the next block's inputs
    # ln_f, then the lm_head sees its outputs
    lnf = model.transformer.ln_f
    for j in range(0, n, batch):
        rows = inps[j:j + batch].numel() // E
        _lib.check(lib.lp_norm(runner.norm_kind, inps[j:j + batch].data_ptr(), _f32(lnf.weight).data_ptr(),
                               _ptr(_f32(getattr(lnf, "bias", None))), cfg.norm_eps, outs[j:j + batch].data_ptr(), rows, E, 0, _stream(dev)),
                   "lp_norm")
    inps, outs = outs, inps
    gptq = GPTQQuantizer(model.lm_head, bits=bits, groupsize=groupsize, actorder=(groupsize == -1))
    for j in range(0, n, batch):
        gptq.collect_input_stats(None, (inps[j:j + batch],), None)
    q_module, error = gptq.quantize()
    if _on_layer is not None:
        q_module = _on_layer("lm_head", q_module) or q_module
    model.lm_head = q_module
    if hasattr(model, "invalidate"):
        model.invalidate()


def main(*, checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"), output_path: Optional[Path] = None,
         n_samples: int = 128, precision: str = "bf16-true", sample_text: Optional[str] = None) -> None:
    """Quantises a checkpoint to 4 bits with GPTQ and writes `lit_model_gptq.4bit.pth` (gptq.py:551-602).

    `sample_text` (extension): calibration text instead of the C4 shard the reference downloads (there may be no network)."""
    from lit_parrot_b200.checkpoint import check_valid_checkpoint_dir, lazy_load
    from lit_parrot_b200.cli import _param_dtype
    from lit_parrot_b200.config import Config
    from lit_parrot_b200.model import GPT
    from lit_parrot_b200.tokenizer import Tokenizer

    checkpoint_dir = Path(checkpoint_dir)
    if output_path is None:
        output_path = checkpoint_dir / "lit_model_gptq.4bit.pth"
    check_valid_checkpoint_dir(checkpoint_dir)
    with open(checkpoint_dir / "lit_config.json") as fp:
        config = Config(**json.load(fp))
    device = torch.device("cuda", torch.cuda.current_device())
    checkpoint_path = checkpoint_dir / "lit_model.pth"
    print(f"Loading model {str(checkpoint_path)!r} with {config.__dict__}", file=sys.stderr)
    t0 = time.time()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(_param_dtype(precision))
    try:
        model = GPT(config)
    finally:
        torch.set_default_dtype(prev)
    with lazy_load(checkpoint_path) as checkpoint:
        model.load_state_dict(checkpoint.get("model", checkpoint))
    print(f"Time to load model: {time.time() - t0:.02f} seconds.", file=sys.stderr)
    model.eval()
    tokenizer = Tokenizer(checkpoint_dir)
    test_string = sample_text if sample_text is not None else get_sample_data()
    encoded_text = tokenizer.encode(test_string, eos=True)
    block_size = config.block_size
    n_samples = min(n_samples, encoded_text.numel() // block_size)
    if n_samples < 1:
        raise ValueError(f"the calibration text has {encoded_text.numel()} tokens; one sample needs block_size = {block_size}")
    encoded_text = encoded_text[: n_samples * block_size].reshape(n_samples, block_size)
    t0 = time.perf_counter()
    blockwise_quantization(model, encoded_text, device, bits=4)
    t = time.perf_counter() - t0
    print(f"\n\nTime for quantization: {t:.02f} sec total", file=sys.stderr)
    print(f"Memory used: {torch.cuda.max_memory_allocated() / 1e9:.02f} GB", file=sys.stderr)
    torch.save(model.state_dict(), output_path)
