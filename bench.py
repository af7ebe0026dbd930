#!/usr/bin/env python
"""bench.py — decode throughput of the hot path on B200 (contract: see the task prompt / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--no-extras]

A "step" is ONE decode step of the hot path (embed -> n_layer blocks -> ln_f -> lm_head -> sampling) over one
batch; `value` is decoded tokens per second with everything resident in HBM (captured-graph replay, CUDA events);
`e2e` is the same metric through the public drop-in API (`model(idx, max_seq_length, input_pos)` + `sample`) driven
from the host with pinned-host token buffers: H2D copy of the token ids and D2H read of the sampled ids every step.

Default workload (N = 1): BASELINE.json configs[1] — stablelm-base-alpha-3b, bf16, batch 1, 2k context.
N > 1: N independent replicas of the same workload (weak scaling, no collective; the models of configs 1-4 do not
shard).  Weights are random-init (`_init_weights`: N(0, 0.02)), prompts/KV contents synthetic.

`--impl reference` times the CPU oracle port of the reference (`oracle/lit_oracle.py`, torch CPU ops, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (preset, quantization mode, gptq tile, batch, context)
    "stablelm-3b-bf16-b1": ("stablelm-base-alpha-3b", None, 0, 1, 2048),
    "stablelm-3b-bf16-b32": ("stablelm-base-alpha-3b", None, 0, 32, 2048),
    "llama2-7b-int4g128-b1": ("Llama-2-7b-hf", "gptq.int4", 128, 1, 2048),
    "llama2-7b-nf4-b1": ("Llama-2-7b-hf", "bnb.nf4", 0, 1, 2048),
    "llama2-7b-bf16-b1": ("Llama-2-7b-hf", None, 0, 1, 2048),
    "falcon-7b-bf16-b1": ("falcon-7b", None, 0, 1, 2048),
    "pythia-70m-bf16-b1": ("pythia-70m", None, 0, 1, 2048),
    # tensor parallel over all launched ranks (BASELINE configs[4]); at 1 GPU it is the unsharded 70B model (137 GB of 180)
    "llama2-70b-bf16-b1-tp": ("Llama-2-70b-hf", None, 0, 1, 2048),
}
DEFAULT = "stablelm-3b-bf16-b1"
EXTRAS = ["llama2-7b-int4g128-b1", "llama2-7b-nf4-b1", "falcon-7b-bf16-b1", "stablelm-3b-bf16-b32"]  # BASELINE configs[1..3]


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same workload
    (profiles/step_kernel_ncu.json, written by tools/ncu_step_summary.py), or None."""
    path = os.path.join(REPO, "profiles", "step_kernel_ncu.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(workload)
        return None if ent is None else int(ent["dram_bytes_read"] + ent["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        return None


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows, self.proc, self.index = [], None, index
        self.nvml = None

    def _nvml_handle(self):
        """NVML handle of CUDA device `index` (by UUID: the two enumerations differ under CUDA_VISIBLE_DEVICES), or None."""
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            return pynvml, h
        except Exception:
            return None

    def sample_until(self, event, max_samples: int = 200) -> None:
        """In-process NVML samples WHILE the GPU works through the queued timed steps (the nvidia-smi poller below needs ~100 ms
        to deliver its first line, longer than a short timed region): one sample at once, then until `event` has completed."""
        if self.nvml is None:
            self.nvml = self._nvml_handle() or False
        if not self.nvml:
            return
        nv, h = self.nvml
        n = 0
        while True:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                flag = lambda bit: "Active" if rs & bit else "Not Active"  # noqa: E731
                self.rows.append([str(sm), str(mx), "", flag(0x8), flag(0x40), flag(0x20), flag(0x4)])  # hw, hw-thermal, sw-thermal, sw power cap
            except Exception:
                return
            n += 1
            if event.query() or n >= max_samples:
                return
            time.sleep(0.002)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def build_model(workload: str, device):
    import torch

    import lit_parrot_b200 as lp

    preset, quant, tile, B, ctx = WORKLOADS[workload]
    cfg = lp.Config.from_name(preset)
    torch.manual_seed(1234)
    with torch.device(device):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            dense = lp.GPT(cfg)
        finally:
            torch.set_default_dtype(prev)
    dense.apply(dense._init_weights)
    if quant is None:
        return dense.eval(), cfg, B, ctx
    # quantise the same random weights layer by layer (round-to-nearest on the reference grid / NF4 code book)
    with torch.device(device):
        with lp.quantization(quant, **({"gptq_tile_cols": tile} if quant == "gptq.int4" else {})):
            q = lp.GPT(cfg)
    qmods, dmods = dict(q.named_modules()), dict(dense.named_modules())
    for name, mod in qmods.items():
        src = dmods[name]
        if hasattr(mod, "quantize_rtn_"):
            # GPTQ: scales live in the model dtype like the reference's buffers under bf16-true (quantize/gptq.py:223-226); the
            # weights are quantised against the rounded scales.  0.5 + 4/128 bytes per weight.
            mod.quantize_rtn_(src.weight.data.float(), **({"scale_dtype": torch.bfloat16} if quant == "gptq.int4" else {}))
            src.weight.data = torch.empty(0, device=device)
        elif isinstance(mod, (torch.nn.Linear, torch.nn.Embedding)) or type(mod).__name__ in ("RMSNorm", "LayerNorm"):
            for pn, p in mod.named_parameters(recurse=False):
                p.data = getattr(src, pn).data.to(torch.bfloat16)
    q = q.to(torch.bfloat16) if quant != "gptq.int4" else q
    del dense
    torch.cuda.empty_cache()
    return q.eval(), cfg, B, ctx


def build_tp_model(preset, device, rank, world):
    """Llama-2-70b (or any GQA preset) sharded over `world` GPUs: every rank materialises ITS shard of the same random
    model (per-tensor seeds, so all ranks agree on the global weights without ever holding them all)."""
    import torch
    import torch.distributed as dist

    import lit_parrot_b200 as lp
    from lit_parrot_b200.tp import TPContext

    cfg = lp.Config.from_name(preset).with_tp(world, rank)
    with torch.device(device):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            model = lp.GPT(cfg)
        finally:
            torch.set_default_dtype(prev)
    gen = torch.Generator(device=device)
    for i, (name, p) in enumerate(model.named_parameters()):
        if p.dim() == 2:
            # one seed per tensor: replicated tensors agree on every rank, sharded ones differ through the rank offset
            gen.manual_seed(1234 + i * 64 + (rank if (".attn." in name or ".mlp." in name) else 0))
            p.data.normal_(0.0, 0.02, generator=gen)
        elif name.endswith("weight"):
            p.data.fill_(1.0)
        else:
            p.data.zero_()
    model.eval()
    if world > 1:
        model.tp_context = TPContext(dist.group.WORLD, device, max_rows=8, n_embd=cfg.n_embd)
    return model, cfg


def algorithmic_bytes_per_step(eng, cfg, B, kv_len_avg, kv_elem_bytes):
    """weights of every linear at stored width (+scales/zeros/absmax) + one embedding row per sequence
    + compact KV read B*2*L*G*hs*bytes*(pos+1) + KV write."""
    w = eng.weight_bytes_per_token() + (B - 1) * cfg.n_embd * eng.wte.element_size()
    kv_row = 2 * cfg.n_layer * cfg.n_query_groups_local * cfg.head_size * kv_elem_bytes
    return w, B * kv_row * kv_len_avg + B * kv_row


def run_workload(workload, steps, warmup, device, rank=0, world=1, with_e2e=True, with_kernel=True):
    import torch
    import torch.distributed as dist

    import lit_parrot_b200 as lp
    from lit_parrot_b200 import _lib

    tp = workload.endswith("-tp")
    if tp:
        preset, _q, _t, B, ctx = WORKLOADS[workload]
        model, cfg = build_tp_model(preset, device, rank, world)
    else:
        model, cfg, B, ctx = build_model(workload, device)
    lib = _lib.init(device.index)
    V = cfg.padded_vocab_size
    start = ctx - steps - warmup - 1
    assert start > 0, "context too short for steps + warmup"
    # synthetic KV cache filled up to `start` positions; decoding continues from there
    model.kv_caches = model.build_kv_caches(torch.zeros(B, 1, device=device), ctx)
    for k, v in model.kv_caches:
        k[:, :, :start].normal_(0, 1)
        v[:, :, :start].normal_(0, 1)
    gen = torch.Generator(device="cpu").manual_seed(1 + (0 if tp else rank))  # tensor parallel: ONE sequence over all ranks
    tok0 = torch.randint(0, cfg.vocab_size, (B, 1), generator=gen)

    def sync_barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident loop: ONE captured graph per step (forward + on-device greedy sampling for B == 1) -------
    count0 = lib.lp_launch_count()
    pos = torch.tensor([start], device=device)
    model(tok0.to(device), ctx, pos)  # builds engine + rope, warm-up run + graph capture of one forward step
    eng = model._get_engine(device)
    launches_per_step = (lib.lp_launch_count() - count0) // 2 + 1  # (warm-up + capture) / 2, + the sampling kernel
    st = eng.gen_state(cfg.block_size)
    st["seq"].zero_()
    st["seq"][start + 1] = int(tok0[0, 0])
    st["pos"].fill_(start + 1)
    replay = eng.decode_step(model.kv_caches, 1.0, 1, 1234, B)
    if B > 1:
        st["tok"][:B].copy_(tok0[:, 0])
    for _ in range(warmup):
        replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_barrier()
    with ClockSampler(device.index) as clk:
        e0.record()
        for _ in range(steps):
            replay()
        e1.record()
        clk.sample_until(e1)  # clocks / throttle reasons while the queued steps execute
        sync_barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    eng.check_step_health()  # the step kernel's watchdog must not have fired in any timed step
    tokens_agree = None
    if tp and world > 1:
        # every rank decoded the SAME greedy tokens (replicated lm_head over bit-identical all-reduced rows)
        mine = st["seq"][start + 1:start + 1 + warmup + steps].clone()
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        tokens_agree = all(bool(torch.equal(allt[0], t)) for t in allt)
        assert tokens_agree, "tensor-parallel ranks decoded different tokens"
    kv_b = model.kv_caches[0][0].element_size()
    wbytes, kvbytes = algorithmic_bytes_per_step(eng, cfg, B, start + warmup + steps / 2 + 1, kv_b)
    replicas = 1 if tp else world  # tensor parallel: ONE model over all ranks (strong scaling); else independent replicas
    res = {"workload": workload, "ms_per_step": ms / steps, "tok_s": replicas * B * steps / (ms / 1e3),
           "launches_per_step": int(launches_per_step), "bytes_per_step": {"weights": int(wbytes), "kv": int(kvbytes)},
           "step_gbs": (wbytes + kvbytes) / (ms / steps) / 1e6, "clocks": clk.summary(), "B": B, "ctx": ctx}
    if tokens_agree is not None:
        res["tokens_agree"] = tokens_agree

    # ---- e2e: public API driven from the host; H2D token ids + D2H sampled ids every step -----------------------
    if with_e2e:
        model.reset_cache()
        model.kv_caches = model.build_kv_caches(torch.zeros(B, 1, device=device), ctx)
        for k, v in model.kv_caches:
            k[:, :, :start].normal_(0, 1)
            v[:, :, :start].normal_(0, 1)
        h_idx = tok0.clone().pin_memory()
        h_pos = torch.tensor([start], dtype=torch.int64).pin_memory()
        h_out = torch.empty(B, dtype=torch.int32).pin_memory()
        # host-side bookkeeping through numpy views of the pinned buffers; the device-side input tensors are allocated once
        # (a caller's serving loop does the same): per step the timed region still holds the H2D copy of the ids and the position,
        # the forward through the public API, the sampler, the D2H copy of the sampled ids and the synchronisation
        # token ids and the position share ONE pinned int64 buffer and one device buffer: a single H2D copy per step
        h_in = torch.empty(B + 1, dtype=torch.int64).pin_memory()
        h_in[:B] = h_idx[:, 0]
        h_in[B] = h_pos[0]
        d_in = torch.empty(B + 1, dtype=torch.int64, device=device)
        x, p = d_in[:B].view(B, 1), d_in[B:]
        n_in, n_out = h_in.numpy(), h_out.numpy()
        n_idx, n_pos = n_in[:B].reshape(B, 1), n_in[B:]
        cur = torch.cuda.current_stream(device)

        def e2e_step():
            d_in.copy_(h_in, non_blocking=True)
            lg = model(x, ctx, p)
            t = lp.sample(lg[:, -1], 1.0, 1)
            h_out.copy_(t, non_blocking=True)
            cur.synchronize()
            n_idx[:, 0] = n_out
            n_pos[0] += 1

        for _ in range(warmup):
            e2e_step()
        sync_barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        sync_barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tdt = torch.tensor([dt], device=device)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt)
        res["e2e"] = {"value": replicas * B * steps / dt, "unit": "tok/s", "h2d_bytes_per_step": int(B * 8 + 8),
                      "d2h_bytes_per_step": int(B * 4)}

    # ---- dominant kernel alone: lp_linear on the MLP up-projection, rotating over the layers' weights (>> L2) ------
    if with_kernel:
        eng = model._get_engine(device)
        if B == 1 and any(v is not None for v in eng._steps.values()):
            res["kernel"] = time_step_kernel(eng, model, cfg, device, kv_b)
        else:
            res["kernel"] = time_dominant_kernel(eng, cfg, B, device)
    del model
    torch.cuda.empty_cache()
    return res


def run_prefill(preset, T, device, reps=3):
    """BASELINE configs[3]: prompt prefill through the public API (bf16-faithful mode: one bf16 term per activation, the
    reference's bf16-true arithmetic), tensor-core roofline.  FLOPs = 2 T sum(out*in) (lm_head on the last row only, as
    generate() consumes it) + causal attention L * 4 H hs T (T + 1) / 2."""
    import torch

    import lit_parrot_b200 as lp

    cfg = lp.Config.from_name(preset)
    torch.manual_seed(1234)
    with torch.device(device):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            model = lp.GPT(cfg)
        finally:
            torch.set_default_dtype(prev)
    model.apply(model._init_weights)
    model.eval().set_precision("bf16")
    max_seq = cfg.block_size
    idx = torch.randint(0, cfg.vocab_size, (1, T), generator=torch.Generator().manual_seed(1)).to(device)
    pos = torch.arange(T, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for i in range(reps + 1):
        model.reset_cache()
        model.kv_caches = model.build_kv_caches(idx, max_seq)
        torch.cuda.synchronize(device)
        e0.record()
        model._forward_impl(idx, max_seq, pos, last_only=True, raw_logits=True)
        e1.record()
        torch.cuda.synchronize(device)
        if i:
            times.append(e0.elapsed_time(e1))
    ms = min(times)
    eng = model._get_engine(device)
    lin = sum(m.N * m.K for L in eng.layers for m in (L.qkv, L.proj, L.fc, L.mlp_proj))
    flops = 2.0 * T * lin + 2.0 * cfg.padded_vocab_size * cfg.n_embd + cfg.n_layer * 4.0 * cfg.n_head * cfg.head_size * T * (T + 1) / 2
    pk = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    peak = float(pk.get("bf16_tflops_sustained", 1400.0))
    tf = flops / ms / 1e9
    del model
    torch.cuda.empty_cache()
    return {"workload": f"{preset} bf16 prefill T={T} (block_size {cfg.block_size}), batch 1", "ms": ms, "prefill_tok_s": T / ms * 1e3,
            "tflops": tf, "tensor_frac_of_sustained_peak": tf / peak, "flops": flops}


def run_gptq_quantizer(device, n_samples=8):
    """the device GPTQ quantiser (reference quantize/gptq.py, published: 850 s for the 32 layers of falcon-7b with 128
    samples on an A100, tutorials/quantize.md:114-117) on ONE Llama-2-7b-width block + lm_head, random-init bf16 weights,
    `n_samples` synthetic calibration sequences of 2048 tokens, group size 128."""
    import torch

    import lit_parrot_b200 as lp
    from lit_parrot_b200 import gptq

    cfg = lp.Config.from_name("Llama-2-7b-hf", n_layer=1, block_size=2048)
    torch.manual_seed(1234)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            m = lp.GPT(cfg)
    finally:
        torch.set_default_dtype(prev)
    m.apply(m._init_weights)
    m.eval()
    samples = torch.randint(0, cfg.vocab_size, (n_samples, cfg.block_size), generator=torch.Generator().manual_seed(1))
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    gptq.blockwise_quantization(m, samples, device, bits=4, groupsize=128, batch=8, verbose=False)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    del m
    torch.cuda.empty_cache()
    return {"workload": f"GPTQ quantiser: 1 Llama-2-7b block (5 linear layers) + lm_head, int4 g128, {n_samples} x 2048 calibration tokens",
            "ms": dt * 1e3}


def run_config0(device):
    """BASELINE configs[0]: pythia-70m random init, greedy generate() of 128 tokens from a 16-token prompt.  The reference runs it
    as-is on the host cores (fp32, baseline/_ref); beside it the same call through this repo's generate() on the GPU (bf16 weights,
    fp32 activations; token parity of this configuration is pinned by tests/test_model_gpu.py against the reference's golden ids)."""
    import torch

    import lit_parrot_b200 as lp

    ref = ref_run(["cpu-generate", "--preset", "pythia-70m", "--prompt", "16", "--tokens", "128", "--reps", "3"])
    cfg = lp.Config.from_name("pythia-70m")
    torch.manual_seed(1234)
    with torch.device(device):
        model = lp.GPT(cfg)
    model.apply(model._init_weights)
    model = model.to(torch.bfloat16).eval()
    prompt = torch.randint(0, cfg.vocab_size, (16,), generator=torch.Generator().manual_seed(1)).to(torch.int32).to(device)
    times = []
    for i in range(4):
        model.reset_cache()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        out = lp.generate(model, prompt, 128, 128, temperature=1.0, top_k=1)
        torch.cuda.synchronize(device)
        if i:
            times.append(time.perf_counter() - t0)
    del model
    torch.cuda.empty_cache()
    return {"workload": "pythia-70m random init, greedy generate() 16 -> 128 tokens (BASELINE configs[0])", "reference_cpu": ref,
            "ours_gpu": {"tok_s": 112 / min(times), "seconds": min(times), "new_tokens": 112, "tokens_head": out[:24].tolist(),
                         "what": "lit_parrot_b200.generate on cuda:0, bf16 weights, wall clock incl. the 16-token prefill, best of 3"}}


def time_step_kernel(eng, model, cfg, device, kv_elem_bytes, iters=64):
    """Batch-1 decode: the dominant kernel is the persistent step kernel (lp_decode_step = prologue + decode_step_kernel, one
    launch pair per token).  Timed alone with CUDA events on the launching stream, position fixed at the last one reached by the
    timed loop (the same KV slot is rewritten), algorithmic bytes = all weights + the KV rows read + written."""
    import torch

    st = eng._gen
    b = eng.buffers(1, 1, 1, model.kv_caches[0][0].size(2))
    stream = torch.cuda.current_stream(device).cuda_stream
    pos = int(st["pos"][0])
    run = lambda: eng._run(b, st["seq"].data_ptr(), 0, st["pos"].data_ptr(), st["pos"].data_ptr(), model.kv_caches, 1, 1, stream)  # noqa: E731
    for _ in range(4):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize(device)
    us = e0.elapsed_time(e1) * 1e3 / iters
    wbytes, kvbytes = algorithmic_bytes_per_step(eng, cfg, 1, pos + 1, kv_elem_bytes)
    nbytes = wbytes + kvbytes
    return {"name": f"lp_decode_step (decode_step_kernel<{cfg.head_size}>, {cfg.n_layer} layers + lm_head, kv length {pos + 1})",
            "us": us, "bytes": int(nbytes), "gbs": nbytes / us / 1e3}


def time_dominant_kernel(eng, cfg, B, device, iters=64):
    """Batches > 1: the MLP up-projection alone, rotating over all layers' weights (>> L2), through the kernel the step itself
    uses for this batch size: the streaming GEMV (lp_linear) up to 8 rows, the swap-AB tcgen05 GEMM (lp_gemm_bf16_tc on the bf16
    terms of the activations) from 9 rows on."""
    import torch

    from lit_parrot_b200 import _lib

    lib = eng.lib
    x = torch.randn(B, cfg.n_embd, device=device)
    out = torch.empty(B, cfg.intermediate_size, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    layers = eng.layers
    if B <= 8:
        name = "lp_linear(mlp.fc)"

        def run(L):
            _lib.check(lib.lp_linear(x.data_ptr(), B, L.fc.ref, eng.act, None, out.data_ptr(), eng.round, stream))
    else:
        nt = 1 if eng.round else (3 if B <= 16 else 2)  # as Engine._run_tc
        name = f"lp_gemm_bf16_tc(mlp.fc, swap-AB, {nt} bf16 terms)"
        terms = torch.empty(nt, B, cfg.n_embd, dtype=torch.bfloat16, device=device)
        _lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), B, cfg.n_embd, nt, -1, None, None, 0.0, eng.round, stream))

        def run(L):
            W = L.fc
            _lib.check(lib.lp_gemm_bf16_tc(terms.data_ptr(), nt, B, eng._bf16_weight(W, stream), W.N, W.K, W.rec.bias, eng.act, None,
                                           out.data_ptr(), None, 0, eng.round, stream))
    for i in range(8):
        run(layers[i % len(layers)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for i in range(iters):
        run(layers[i % len(layers)])
    e1.record()
    torch.cuda.synchronize(device)
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = layers[0].fc.stored_bytes
    return {"name": name, "N": layers[0].fc.N, "K": layers[0].fc.K, "us": us, "bytes": int(nbytes), "gbs": nbytes / us / 1e3}


# ------------------------------------------------------------------------------------------------------------------
def ref_available():
    return os.path.isdir(os.path.join(REPO, "baseline", "_ref", "lit_gpt"))


def ref_run(argv, timeout=600):
    """The UNMODIFIED reference (baseline/_ref) in a subprocess (baseline/ref_runner.py): one JSON dict, or {"error": ...}."""
    if not ref_available():
        return {"error": "baseline/_ref is not installed (baseline/install_ref.sh needs /root/reference)"}
    try:
        r = subprocess.run([sys.executable, os.path.join(REPO, "baseline", "ref_runner.py"), *argv], capture_output=True, text=True,
                           timeout=timeout, cwd=os.path.join(REPO, "baseline"))
        lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def cpu_reference_decode(workload, max_steps, warmup, budget_s=75.0):
    """CPU arm: the reference's own CPU path for this workload on the box's host cores.  The unmodified reference
    (baseline/_ref, kind "reference") where it can run the workload (float weights); the oracle port (kind "port") for the
    quantised workloads (the reference builds per-row GPTQ layers only, lit_gpt/utils.py:72-74, and bitsandbytes is absent) and
    wherever baseline/_ref is missing.  Returns (tok/s, steps, threads, description, kind)."""
    preset, quant, tile, B, ctx = WORKLOADS[workload]
    big = preset.endswith("70b-hf")  # fp32 70B does not fit host RAM: the port times a 4-layer slice
    if quant is None and B == 1 and not big and ref_available():
        r = ref_run(["cpu-decode", "--preset", preset, "--ctx", str(ctx), "--steps", str(max_steps), "--warmup", str(max(1, min(warmup, 2))),
                     "--budget", str(budget_s)], timeout=budget_s * 4 + 600)
        if "error" not in r and r.get("steps", 0) > 0:
            return r["tok_s"], r["steps"], r["threads"], r["what"], "reference"
    v, n, threads, desc = cpu_oracle_decode(workload, max_steps, warmup, budget_s)
    return v, n, threads, desc, "port"


def cpu_oracle_decode(workload, max_steps, warmup, budget_s=75.0):
    """The reference's CPU path (oracle port, torch CPU fp32, all host threads) on a bounded sample: short prefill,
    then decode steps at short context.  Returns (tok/s, steps timed, threads, description)."""
    import torch

    import lit_parrot_b200 as lp
    from oracle import lit_oracle as O

    preset, quant, tile, B, ctx = WORKLOADS[workload]
    cfg = lp.Config.from_name(preset)
    full_layers = cfg.n_layer
    if cfg.n_layer * cfg.n_embd * cfg.n_embd * 12 * 4 > 100e9:
        # Llama-2-70b in fp32 does not fit the host (276 GB): time a 4-layer slice of it and scale the per-token time by
        # n_layer / 4 (lm_head + embedding are counted once); stated in `sample`
        cfg = lp.Config.from_name(preset, n_layer=4)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t_build = time.perf_counter()
    g = torch.Generator().manual_seed(1234)
    pool = torch.empty(1 << 24).normal_(0, 0.02, generator=g)  # timing does not depend on the values: cycle a random pool
    sd = {}
    for k, shp in O.state_dict_shapes(cfg).items():
        if len(shp) == 2:
            n = shp[0] * shp[1]
            sd[k] = pool.repeat(-(-n // pool.numel()))[:n].view(shp).clone() if n > pool.numel() else pool[:n].view(shp).clone()
        else:
            sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
    if quant == "gptq.int4":
        for k in [k for k in sd if sd[k].dim() == 2 and "wte" not in k]:
            base = k[: -len(".weight")]
            sd[base + ".quant_weight"], sd[base + ".scales"], sd[base + ".zeros"] = O.gptq_rtn_quantize(sd.pop(k), tile)
    m = O.OracleGPT(cfg, sd)
    t_build = time.perf_counter() - t_build
    T0 = 16
    idx = torch.randint(0, cfg.vocab_size, (B, T0), generator=g)
    pos = torch.arange(T0)
    with torch.no_grad():
        lg = m(idx, 64, pos)
        tok = lg[:, -1].argmax(-1, keepdim=True)
        for _ in range(max(1, min(warmup, 2))):
            pos = pos[-1:] + 1
            tok = m(tok, 64, pos)[:, -1].argmax(-1, keepdim=True)
        n, t0 = 0, time.perf_counter()
        while n < max_steps and pos[-1] < 62 and (time.perf_counter() - t0) < budget_s:
            pos = pos[-1:] + 1
            tok = m(tok, 64, pos)[:, -1].argmax(-1, keepdim=True)
            n += 1
        dt = time.perf_counter() - t0
        if cfg.n_layer != full_layers and n:
            t_head, h0 = 0.0, time.perf_counter()  # lm_head share: time it alone on the same shapes
            for _ in range(3):
                torch.nn.functional.linear(torch.randn(B, cfg.n_embd), m.sd["lm_head.weight"])
            t_head = (time.perf_counter() - h0) / 3
            per_layer = max(dt / n - t_head, 0.0) / cfg.n_layer
            dt = n * (t_head + per_layer * full_layers)
    desc = (f"{preset} fp32 weights on CPU ({'int4 g128 dequant+F.linear per call, ' if quant else ''}oracle port of the reference, "
            f"torch {torch.__version__}), batch {B}, 16-token prefill then {n} greedy decode steps at context 17-{17 + n} "
            f"(not 2k: bounded sample), {threads} threads; weight build {t_build:.0f}s not timed"
            + (f"; {cfg.n_layer}-layer slice timed and scaled to {full_layers} layers (fp32 70B does not fit host RAM)"
               if cfg.n_layer != full_layers else ""))
    return B * n / dt, n, threads, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help=f"default: {DEFAULT} at 1 GPU; llama2-70b-bf16-b1-tp (tensor parallel over all ranks) at N > 1")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        # BASELINE configs[1] at every N: a 3B model does not shard, so N > 1 runs N independent replicas (weak scaling, no
        # collective) and the line stays comparable across N.  The configuration of the path that DOES shard — configs[4],
        # Llama-2-70b tensor parallel over the N ranks, strong scaling — is measured in the same run and reported under "tp".
        args.workload = DEFAULT
    is_tp = args.workload.endswith("-tp")
    preset, quant, tile, B, ctx = WORKLOADS[args.workload]
    base = {"metric": "decode_tokens_per_second", "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if is_tp else "weak", "vs_baseline": None,
            "dtype": "bf16" if quant is None else f"bf16 activations-in-fp32 / {quant} weights", "data": "synthetic",
            "config": {"workload": f"{preset} {'bf16' if quant is None else quant} decode, batch {B}, context {ctx} "
                                   f"(positions {ctx - args.steps - args.warmup - 1}..{ctx - 1})",
                       "weights": "random init N(0,0.02) (_init_weights), seed 1234", "kv_cache": "synthetic N(0,1) prefix, bf16, compact (B,G,ctx,hs)",
                       "l2": "inputs larger than L2 (weights streamed per step >> 126 MB)",
                       "parallelism": ("1 GPU" if world == 1 else
                                       (f"tensor parallel tp={world}: weights sharded by query group / MLP column, one-shot NVLink all-reduce "
                                        "fused with the residual add (2 exchanges per layer)" if is_tp else
                                        f"{world} independent replicas (no collective)"))}}

    if args.impl == "reference":
        if rank != 0:
            return 0
        v, n, threads, desc, kind = cpu_reference_decode(args.workload, args.steps, args.warmup)
        base["config"]["workload"] += f" — CPU arm sample: {desc}"
        line = dict(base, impl="reference", value=v, steps=n, ms_per_step=1e3 * B / v, n_gpus=args.gpus,
                    cpu_baseline={"value": v, "unit": "tok/s", "cores": threads, "kind": kind, "sample": desc},
                    e2e={"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
        line["dtype"] = "f32"
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    res = run_workload(args.workload, args.steps, args.warmup, device, rank, world)
    extras = []
    tp_res = None
    if not args.no_extras and not is_tp:
        try:  # Llama-2-70b bf16, batch 1, 2k context over all `world` ranks (world == 1: the unsharded model, 137 GB of HBM)
            tp_res = run_workload("llama2-70b-bf16-b1-tp", min(args.steps, 32), min(args.warmup, 4), device, rank, world, with_e2e=False,
                                  with_kernel=False)
        except Exception as e:
            tp_res = {"workload": "llama2-70b-bf16-b1-tp", "error": repr(e)[:200]}
        tp_n1 = None
        if world > 1:
            # the same model unsharded on ONE GPU (137 GB of 180) is the N = 1 point of the strong-scaling curve: rank 0 measures it
            # in the same run (the other ranks wait), so that speed-up and efficiency of the sharded path are in the line itself
            if rank == 0:
                try:
                    tp_n1 = run_workload("llama2-70b-bf16-b1-tp", min(args.steps, 32), min(args.warmup, 4), device, 0, 1, with_e2e=False,
                                         with_kernel=False)
                except Exception as e:
                    tp_n1 = {"error": repr(e)[:200]}
            dist.barrier()
    if not args.no_extras and world == 1:
        for w in EXTRAS:
            if w != args.workload:
                try:
                    extras.append(run_workload(w, args.steps, args.warmup, device, with_e2e=(w == "llama2-7b-int4g128-b1")))
                except Exception as e:  # an extra must never cost the headline line
                    extras.append({"workload": w, "error": repr(e)[:200]})
        try:
            extras.append(run_prefill("falcon-7b", 1792, device))  # 1792 + 256 decode tokens = block_size 2048
        except Exception as e:
            extras.append({"workload": "falcon-7b prefill", "error": repr(e)[:200]})
        try:
            extras.append(run_gptq_quantizer(device))
        except Exception as e:
            extras.append({"workload": "GPTQ quantiser", "error": repr(e)[:200]})
        try:
            extras.append(run_config0(device))  # BASELINE configs[0]: the reference as-is on CPU beside the same generate() here
        except Exception as e:
            extras.append({"workload": "pythia-70m configs[0]", "error": repr(e)[:200]})
        # the library path on the same GPU: the unmodified reference, eager PyTorch CUDA bf16, same model / batch / context
        r = ref_run(["cuda-decode", "--preset", preset, "--ctx", str(ctx), "--steps", str(min(args.steps, 64)), "--warmup", "8"])
        extras.append({"workload": f"reference eager CUDA: {preset} bf16 decode, batch 1, context {ctx}", "reference_eager_cuda": r})
    if rank == 0:
        peak, peak_src = peaks()
        k = res["kernel"]
        line = dict(base, value=res["tok_s"], ms_per_step=res["ms_per_step"], e2e=res["e2e"],
                    gpu_launches=res["launches_per_step"] * args.steps, clocks=res["clocks"],
                    roofline={"bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
                              "traffic": ncu_traffic(args.workload) if k["name"].startswith("lp_decode_step") else None,
                              "kernel": f"{k['name']}{' N=%d K=%d' % (k['N'], k['K']) if 'N' in k else ''}: {k['bytes']} B in {k['us']:.2f} us",
                              "peak_source": peak_src,
                              "whole_step": {"achieved": res["step_gbs"], "frac": res["step_gbs"] / peak,
                                             "bytes": res["bytes_per_step"], "launches": res["launches_per_step"]}})
        if tp_res is not None:
            line["tp"] = {"workload": f"Llama-2-70b-hf bf16 decode, batch 1, context 2048, tensor parallel over {world} GPU(s)",
                          "scaling": "strong", "n_gpus": world,
                          **{kk: tp_res[kk] for kk in ("tok_s", "ms_per_step", "launches_per_step", "bytes_per_step", "step_gbs", "error")
                             if kk in tp_res}}
            if "step_gbs" in tp_res:
                line["tp"]["per_gpu_hbm_frac"] = tp_res["step_gbs"] / peak  # bytes of ONE rank's shard / step time
            if "tokens_agree" in tp_res:
                line["tp"]["all_ranks_same_tokens"] = tp_res["tokens_agree"]
            if world > 1 and tp_n1 is not None:
                if "tok_s" in tp_n1 and "tok_s" in tp_res:
                    line["tp"]["n1_tok_s"] = tp_n1["tok_s"]
                    line["tp"]["speedup_vs_n1"] = tp_res["tok_s"] / tp_n1["tok_s"]
                    line["tp"]["efficiency"] = tp_res["tok_s"] / tp_n1["tok_s"] / world
                else:
                    line["tp"]["n1_error"] = tp_n1.get("error")
        if extras:
            line["also"] = [{kk: e[kk] for kk in e if kk in ("workload", "tok_s", "ms_per_step", "step_gbs", "launches_per_step",
                                                              "bytes_per_step", "kernel", "error", "ms", "prefill_tok_s", "tflops",
                                                              "tensor_frac_of_sustained_peak", "e2e", "reference_eager_cuda",
                                                              "reference_cpu", "ours_gpu")} for e in extras]
            for e in line["also"]:
                if "kernel" in e and e["kernel"].get("name", "").startswith("lp_decode_step"):
                    e["kernel"]["traffic"] = ncu_traffic(e["workload"])  # DRAM bytes of the committed ncu capture of the default build
            for e in line["also"]:
                if "step_gbs" in e:
                    e["step_frac"] = e["step_gbs"] / peak
        if not args.no_cpu_baseline and world == 1:
            v, n, threads, desc, kind = cpu_reference_decode(args.workload, 24, 2, budget_s=25.0)
            line["cpu_baseline"] = {"value": v, "unit": "tok/s", "cores": threads, "kind": kind, "sample": desc}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
