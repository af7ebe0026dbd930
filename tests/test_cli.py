"""the callers on the other side of generate() — Tokenizer, the streaming chat generate with stop sequences,
prompt_config and the two CLI mains — against golden data produced by the UNMODIFIED reference (oracle/make_golden.py::cli_cases:
lit_gpt/tokenizer.py, chat/base.py:20-95, 202-290) and, on the GPU, against the oracle's greedy tokens."""
import io
import json
import os
from contextlib import redirect_stderr, redirect_stdout
from pathlib import Path

import pytest
import torch

import lit_parrot_b200 as lp
from lit_parrot_b200 import chat, cli
from lit_parrot_b200.tokenizer import Tokenizer

GOLDEN = Path(os.path.dirname(os.path.abspath(__file__))) / "golden"
CKPT = GOLDEN / "ckpt_tiny_llama"
G = json.load(open(GOLDEN / "cli_golden.json"))


def test_tokenizer_matches_reference():
    t = Tokenizer(CKPT)
    assert (t.vocab_size, t.bos_id, t.eos_id, t.backend) == (G["vocab_size"], G["bos_id"], G["eos_id"], G["backend"])
    for case in G["encode"]:
        ids = t.encode(case["text"], **case["kw"])
        assert ids.tolist() == case["ids"] and str(ids.dtype) == case["dtype"]
    for case in G["decode"]:
        assert t.decode(torch.tensor(case["ids"])) == case["text"]
    with pytest.raises(ValueError):
        t.token_to_id("not-a-token")
    lp.check_valid_checkpoint_dir(CKPT)  # tokenizer files complete the directory


class Scripted(torch.nn.Module):
    def __init__(self, script, vocab=96):
        super().__init__()
        self.script, self.vocab, self.calls = script, vocab, 0

    def forward(self, idx, max_seq_length, input_pos):
        lg = torch.zeros(1, idx.size(1), self.vocab)
        lg[0, -1, self.script[self.calls]] = 1e4
        self.calls += 1
        return lg


@pytest.mark.parametrize("case", G["chat"], ids=[f"stream{i}" for i in range(len(G["chat"]))])
def test_chat_generate_stop_sequences_match_reference(case):
    """Same scripted next tokens -> same yields (values, 0-dim vs 1-D, and what is held back / dropped) as chat/base.py:20-95."""
    stops = tuple(case["stops"])
    prompt = torch.tensor([15, 16], dtype=torch.int32)
    m = Scripted(case["script"])
    got = list(chat.generate(m, prompt, case["max_returned_tokens"], case["max_returned_tokens"], temperature=1.0, top_k=None,
                             stop_tokens=stops))
    assert [y.tolist() for y in got] == case["yields"] and [y.ndim for y in got] == case["ndims"]
    # the filter alone, on plain ints
    n_new = case["max_returned_tokens"] - 2
    got2 = list(chat._stop_filter(case["script"][:n_new], stops, "cpu"))
    assert [y.tolist() for y in got2] == case["yields"]


def test_prompt_config_matches_reference():
    t = Tokenizer(CKPT)
    for case in G["prompt_config"]:
        sp, stops = chat.prompt_config(Path(case["name"]), t)
        assert sp == case["system_prompt"] and [list(s) for s in stops] == case["stop_tokens"], case["name"]


def test_cli_argument_surface_and_errors(tmp_path):
    import inspect

    names = list(inspect.signature(cli.main).parameters)
    assert names == ["prompt", "num_samples", "max_new_tokens", "top_k", "temperature", "checkpoint_dir", "quantize", "strategy",
                     "devices", "precision"]  # generate/base.py:162-174
    assert list(inspect.signature(chat.main).parameters) == ["top_k", "temperature", "checkpoint_dir", "quantize", "precision"]
    with pytest.raises(SystemExit):  # not a checkpoint directory: the reference exits the CLI (lit_gpt/utils.py:253-262)
        cli.load_model(tmp_path / "nope", None, "bf16-true", torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        cli.main("x", checkpoint_dir=CKPT, devices=2)
    import generate.base as gb
    import chat.base as cb

    assert gb.main is cli.main and cb.main is chat.main and cb.generate is chat.generate


def test_cli_helper_parses_like_jsonargparse():
    """`CLI(main)`: every parameter is a `--name value` option (jsonargparse turns parameters with defaults into options, so the
    reference is run as `python generate/base.py --prompt "Hello" --checkpoint_dir ...`); the first one may also be positional."""
    seen = {}

    def fn(prompt: str = "Hello", *, max_new_tokens: int = 50, temperature: float = 0.8, checkpoint_dir: Path = Path("x")):
        """doc"""
        seen.update(prompt=prompt, max_new_tokens=max_new_tokens, temperature=temperature, checkpoint_dir=checkpoint_dir)

    cli.CLI(fn, ["--prompt", "a b", "--max_new_tokens", "7", "--checkpoint_dir", "ckpt/dir"])
    assert seen == dict(prompt="a b", max_new_tokens=7, temperature=0.8, checkpoint_dir=Path("ckpt/dir"))
    cli.CLI(fn, ["positional prompt", "--temperature", "0.5"])
    assert seen["prompt"] == "positional prompt" and seen["temperature"] == 0.5 and seen["max_new_tokens"] == 50
    with pytest.raises(SystemExit):
        cli.CLI(fn, ["--no_such_option", "1"])


@pytest.mark.gpu
@pytest.mark.parametrize("quantize", [None, "gptq.int4"])
def test_generate_cli_main_on_tiny_checkpoint(quantize):
    """`python generate/base.py --checkpoint_dir tests/golden/ckpt_tiny_llama --top_k 1 ...`: the printed sample is the decoded
    greedy continuation the oracle computes from the same files; the timing / memory lines go to stderr in the reference's format."""
    from oracle import lit_oracle as O

    t = Tokenizer(CKPT)
    prompt = "w20 w21 human : w95"
    out, err = io.StringIO(), io.StringIO()
    with redirect_stdout(out), redirect_stderr(err):
        cli.main(prompt, num_samples=2, max_new_tokens=20, top_k=1, temperature=1.0, checkpoint_dir=CKPT, quantize=quantize,
                 precision="32-true")
    cfg = lp.Config(**json.load(open(CKPT / "lit_config.json")))
    sd = torch.load(CKPT / ("lit_model_gptq.4bit.pth" if quantize else "lit_model.pth"))
    ids = t.encode(prompt)
    want = O.generate(O.OracleGPT(cfg, sd), ids, ids.numel() + 20, ids.numel() + 20, top_k=1, argmax_ties=True)
    lines = out.getvalue().strip().splitlines()
    assert lines == [t.decode(want)] * 2
    e = err.getvalue()
    assert "Loading model" in e and "Time to instantiate model:" in e and "Time to load the model weights:" in e
    assert "Time for inference 1:" in e and "Time for inference 2:" in e and "tokens/sec" in e and "Memory used:" in e


@pytest.mark.gpu
def test_chat_streaming_on_device_matches_generate():
    """chat.generate on our GPT (device-resident loop, one token read back per replay) yields the same greedy tokens as
    generate() up to the stop sequence, and `decode` prints them as they arrive."""
    from oracle import lit_oracle as O

    dev = "cuda:0"
    model = cli.load_model(CKPT, None, "32-true", torch.device(dev))
    t = Tokenizer(CKPT)
    ids = t.encode("w30 w31 w32").to(dev)
    full = lp.generate(model, ids, 40, 40, top_k=1).cpu().tolist()
    new = full[ids.numel():]
    model.reset_cache()
    stop = new[10:12]  # a 2-token stop sequence taken from the continuation itself
    first_hit = next(i for i in range(len(new) - 1) if new[i:i + 2] == stop)
    got = list(chat.generate(model, ids, 40, 40, temperature=1.0, top_k=1, stop_tokens=([t.eos_id], stop)))
    flat = [v for y in got for v in (y.view(-1).tolist())]
    assert flat == new[:first_hit]
    model.reset_cache()
    buf = io.StringIO()
    n = chat.decode(t, chat.generate(model, ids, 30, 30, temperature=1.0, top_k=1, stop_tokens=([t.eos_id],)), out=buf)
    want = O.generate(O.OracleGPT(model.config, torch.load(CKPT / "lit_model.pth")), ids.cpu(), 30, 30, top_k=1, argmax_ties=True)
    stream = want[ids.numel():].tolist()
    upto = stream.index(t.eos_id) if t.eos_id in stream else len(stream)  # 1-token stops: nothing is held back
    assert n == upto and buf.getvalue() == "".join(t.decode(torch.tensor(x)) for x in stream[:upto])
