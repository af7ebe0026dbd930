"""Checkpoint interop: tests/golden/ckpt_tiny_llama was WRITTEN BY THE REFERENCE (oracle/make_golden.py:
checkpoint_cases) — lit_config.json, lit_model.pth and the column-major int4 file — and re-read there with the reference's own
lazy_load; the stored logits / state-dict digests are those of the reloaded reference model."""
import hashlib
import os

import numpy as np
import pytest
import torch

import lit_parrot_b200 as lp
from lit_parrot_b200 import checkpoint as ck
from oracle import lit_oracle as O
from helpers import GOLDEN, cosine, t

CKPT = os.path.join(GOLDEN, "ckpt_tiny_llama")


def expected():
    return np.load(os.path.join(GOLDEN, "ckpt_tiny_llama_expected.npz"), allow_pickle=False)


def digest(sd) -> str:  # same rule as oracle/make_golden.py:sd_digest
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_lazy_load_yields_the_saved_tensors():
    for fname in ("lit_model.pth", "lit_model_gptq.4bit.pth"):
        eager = torch.load(os.path.join(CKPT, fname), map_location="cpu", weights_only=True)
        with ck.lazy_load(os.path.join(CKPT, fname)) as sd:
            sd = sd.get("model", sd)
            assert list(sd) == list(eager)
            for k, v in eager.items():
                assert sd[k].dtype == v.dtype and sd[k].shape == v.shape and sd[k].stride() == v.stride(), k
                assert torch.equal(sd[k], v), k
    with ck.lazy_load(os.path.join(CKPT, "lit_model_gptq.4bit.pth")) as sd:
        qw = sd["lm_head.quant_weight"]
        assert qw.dtype == torch.uint8 and qw.stride() == (1, qw.shape[0])  # quantize/gptq.py:216-222: column-major bytes
    with pytest.raises(FileNotFoundError):
        ck.lazy_load(os.path.join(CKPT, "nope.pth"))


def test_config_from_lit_config_json():
    cfg = ck.load_config(CKPT)
    assert (cfg.n_layer, cfg.n_head, cfg.n_embd, cfg.intermediate_size) == (2, 4, 64, 192)
    assert cfg._norm_class == "RMSNorm" and cfg._mlp_class == "LLaMAMLP" and cfg.padded_vocab_size == 96


@pytest.mark.parametrize("quantize,tag", [(None, "fp32"), ("gptq.int4", "int4")])
def test_load_checkpoint_state_dict_is_the_reference_models(quantize, tag):
    model = ck.load_checkpoint(CKPT, quantize=quantize, device=None)
    assert not model.training
    assert digest(model.state_dict()) == str(expected()[f"digest_{tag}"])


@pytest.mark.parametrize("fname,tag", [("lit_model.pth", "fp32"), ("lit_model_gptq.4bit.pth", "int4")])
def test_oracle_on_the_checkpoint_files(fname, tag):
    z = expected()
    with ck.lazy_load(os.path.join(CKPT, fname)) as sd:
        out = O.OracleGPT(ck.load_config(CKPT), dict(sd))(t(z["idx"]))
    assert torch.allclose(out, t(z[f"logits_{tag}"]), atol=1e-5, rtol=0)


def test_model_file_selection_and_errors(tmp_path):
    assert ck.checkpoint_file(CKPT).name == "lit_model.pth"
    assert ck.checkpoint_file(CKPT, "bnb.nf4").name == "lit_model.pth"  # bnb modes quantise the fp checkpoint on load
    assert ck.checkpoint_file(CKPT, "gptq.int4").name == "lit_model_gptq.4bit.pth"
    with pytest.raises(ValueError, match="quantize/gptq.py"):
        ck.checkpoint_file(tmp_path, "gptq.int4")
    ck.check_valid_checkpoint_dir(CKPT)  # complete since the tokenizer files were added (oracle/make_golden.py::cli_cases)
    import shutil

    part = tmp_path / "partial"
    part.mkdir()
    for f in ("lit_model.pth", "lit_config.json"):
        shutil.copy(os.path.join(CKPT, f), part / f)
    with pytest.raises(SystemExit) as e:
        ck.check_valid_checkpoint_dir(part)  # no tokenizer files
    assert "tokenizer_config.json" in str(e.value) and "lit_model.pth" not in str(e.value).split("missing the files")[1]
    with pytest.raises(SystemExit, match="is not a checkpoint directory"):
        ck.check_valid_checkpoint_dir(tmp_path / "absent")
    for f in ("lit_model.pth", "lit_config.json", "tokenizer.json", "tokenizer_config.json"):
        (tmp_path / f).write_text("{}")
    ck.check_valid_checkpoint_dir(tmp_path)


def test_save_load_round_trip(tmp_path):
    torch.manual_seed(3)
    cfg = lp.Config(block_size=32, vocab_size=64, padding_multiple=32, n_layer=1, n_head=2, n_embd=32, n_query_groups=1,
                    parallel_residual=True, shared_attention_norm=True, bias=False, rotary_percentage=1.0)
    m = lp.GPT(cfg)
    m.apply(m._init_weights)
    ck.save_checkpoint(m, tmp_path)
    m2 = ck.load_checkpoint(tmp_path, device=None)
    assert vars(m2.config) == vars(m.config)
    assert digest(m2.state_dict()) == digest(m.state_dict())
    import lit_gpt.utils as shim  # the drop-in module exports the loader under the reference's names

    assert shim.lazy_load is ck.lazy_load and shim.check_valid_checkpoint_dir is ck.check_valid_checkpoint_dir


@pytest.mark.gpu
@pytest.mark.parametrize("quantize,tag", [(None, "fp32"), ("gptq.int4", "int4")])
def test_checkpoint_through_the_cuda_path(quantize, tag):
    z = expected()
    model = ck.load_checkpoint(CKPT, quantize=quantize, device="cuda")
    logits = model(t(z["idx"]).cuda()).float().cpu()
    ref = t(z[f"logits_{tag}"])
    assert (logits - ref).abs().max().item() < 1e-4  # fp32-activation mode: far inside the 2e-2 north-star tolerance
    assert cosine(logits, ref) > 0.99999
