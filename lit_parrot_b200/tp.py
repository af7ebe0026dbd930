"""Tensor parallelism for the decode hot path (BASELINE configs[4]: Llama-2-70b over 2/4/8 B200).

The reference has no tensor parallelism (its multi-GPU inference is FSDP, generate/base.py:187-205, which re-gathers
every layer's weights for every token); this module is the B200-native replacement: one process per GPU,
``torch.distributed`` (NCCL) for bootstrap, weights sharded by query group / MLP column, and ONE exchange per
sub-block — a one-shot all-reduce over NVLink peer memory fused with the residual add (``lp_tp_allreduce_residual``),
fed straight from the projection GEMV's output buffer.  No NCCL call on the per-token path.

    cfg  = Config.from_name("Llama-2-70b-hf").with_tp(world, rank)
    model = GPT(cfg); model.load_state_dict(shard_state_dict(full_sd, cfg)); model.cuda()
    model.tp_context = TPContext(dist.group.WORLD, device, max_rows=..., n_embd=cfg.n_embd)
"""
from typing import Dict

import torch

from lit_parrot_b200.config import Config


def shard_state_dict(sd: Dict[str, torch.Tensor], cfg: Config) -> Dict[str, torch.Tensor]:
    """Rank ``cfg.tp_rank``'s shard of a full (reference-layout) state dict.

    * ``attn.attn`` rows are group-major ``[q x q_per_kv, k, v] x hs`` per query group (model.py:210-214): a rank takes the
      contiguous rows of its groups; ``attn.proj`` takes the matching input columns (heads are group-major after the
      transpose at model.py:249);
    * ``mlp.fc / fc_1 / fc_2`` rows and ``mlp.proj`` columns are split evenly;
    * biases of the row-parallel outputs (``attn.proj``, ``mlp.proj``) live on rank 0 only (they are added once, after the
      exchange); everything else (norms, ``wte``, ``lm_head``) is replicated.
    """
    tp, r = cfg.tp_size, cfg.tp_rank
    if tp == 1:
        return dict(sd)
    out = {}
    for k, v in sd.items():
        if ".attn.attn." in k or ".mlp.fc" in k:  # column-parallel: split output rows (weight dim 0, bias dim 0)
            out[k] = v.chunk(tp, dim=0)[r].contiguous()
        elif k.endswith(".attn.proj.weight") or k.endswith(".mlp.proj.weight"):  # row-parallel: split input columns
            out[k] = v.chunk(tp, dim=1)[r].contiguous()
        elif k.endswith(".attn.proj.bias") or k.endswith(".mlp.proj.bias"):
            out[k] = v if r == 0 else torch.zeros_like(v)
        else:
            out[k] = v
    return out


class TPContext:
    """Symmetric (peer-mapped) exchange buffers of one tensor-parallel group + the slot bookkeeping of the exchanges."""

    def __init__(self, group, device: torch.device, max_rows: int, n_embd: int) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.slot_floats = max_rows * n_embd
        self.n_embd = n_embd
        # [2 slots][max_rows * E]: pull protocol of the per-op kernel (lp_tp_allreduce_residual: prefill, batches > 1), then
        # [2 slots][size ranks][E] {value, epoch} pairs: push protocol of the step kernel (batch-1 decode; every rank's partial
        # lands in every buffer, 8 bytes per element)
        self.push_off = 2 * self.slot_floats * 4
        self.buf = symm.empty(2 * self.slot_floats + 2 * self.size * n_embd * 2, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        if self.hdl.signal_pad_size < 4 * self.size * 4:
            raise RuntimeError("symmetric-memory signal pad too small")
        self.buf_ptrs = self.hdl.buffer_ptrs_dev      # device arrays of `size` peer-mapped addresses
        self.pad_ptrs = self.hdl.signal_pad_ptrs_dev
        self.state = torch.zeros(2, 2, dtype=torch.int32, device=device)  # pull protocol, per slot: epoch, CTA ticket
        self.push_state = torch.zeros(2, 2, dtype=torch.int32, device=device)  # push protocol, per slot: epoch, -
        self.push_pad = 2 * self.size  # first signal-pad word of the push protocol (the pull protocol owns [0, 2 * size))
        self.slot = 0
        self.count = 0
        dist.barrier(group)

    def begin_forward(self) -> None:
        self.slot, self.count = 0, 0

    def next_slot(self):
        """(slot index, device pointer of this rank's slot) for the next exchange; slots alternate."""
        s = self.slot
        self.slot ^= 1
        self.count += 1
        return s, self.buf.data_ptr() + s * self.slot_floats * 4


class LoopbackTPContext:
    """One-rank stand-in for ``TPContext`` (single process, no ``torch.distributed``): the only 'peer' is this GPU's own
    buffer, so an exchange adds ONE partial to the residual.  It lets a tensor-parallel SHARD (``Config.with_tp``) run alone —
    the parity tests of the tp-local kernel shapes on one GPU use it; it is not a way to run a sharded model."""

    def __init__(self, device: torch.device, max_rows: int, n_embd: int) -> None:
        self.group = None
        self.rank, self.size = 0, 1
        self.slot_floats = max_rows * n_embd
        self.n_embd = n_embd
        self.push_off = 2 * self.slot_floats * 4
        self.buf = torch.zeros(2 * self.slot_floats + 2 * n_embd * 2, dtype=torch.float32, device=device)
        self.pad = torch.zeros(64, dtype=torch.int32, device=device)
        self.push_state = torch.zeros(2, 2, dtype=torch.int32, device=device)
        self.push_pad = 2
        self._ptrs = torch.tensor([self.buf.data_ptr(), self.pad.data_ptr()], dtype=torch.int64, device=device)
        self.buf_ptrs = self._ptrs.data_ptr()
        self.pad_ptrs = self._ptrs.data_ptr() + 8
        self.state = torch.zeros(2, 2, dtype=torch.int32, device=device)
        self.slot = 0
        self.count = 0

    begin_forward = TPContext.begin_forward
    next_slot = TPContext.next_slot
