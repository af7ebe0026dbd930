"""Config presets: every derived value equals the reference's (fixture made from lit_gpt/config.py by oracle/make_golden.py)."""
import numpy as np

from lit_parrot_b200 import Config, name_to_config


def test_config_reference_cases():
    # reference tests/test_config.py:1-14
    assert Config().block_size == 4096
    assert Config(block_size=2048).block_size == 2048
    assert Config.from_name("pythia-70m").block_size == 2048
    assert Config.from_name("pythia-70m", block_size=4096).block_size == 4096


def test_presets_match_reference(golden_dir):
    z = np.load(f"{golden_dir}/presets.npz")
    names = z["names"].tolist()
    assert sorted(name_to_config) == names
    for name, row, org in zip(names, z["table"], z["orgs"].tolist()):
        c = Config.from_name(name)
        got = [c.block_size, c.vocab_size, c.padded_vocab_size, c.n_layer, c.n_head, c.n_embd, c.n_query_groups,
               c.intermediate_size, int(c.rotary_percentage * 1000), int(c.parallel_residual), int(c.bias),
               int(c.shared_attention_norm), int(c._norm_class == "RMSNorm"), int(c._mlp_class == "LLaMAMLP"),
               c.condense_ratio, int(round(c.norm_eps * 1e9)), c.head_size]
        assert got == row.tolist(), name
        assert c.org == org and c.name == name


def test_derived_shapes():
    c = Config.from_name("falcon-7b")
    assert (c.q_per_kv, c.qkv_rows, c.rope_n_elem) == (71, 4672, 64)
    c = Config.from_name("Llama-2-70b-hf")
    assert (c.q_per_kv, c.qkv_rows, c.head_size) == (8, 10240, 128)
    c = Config.from_name("stablelm-base-alpha-3b")
    assert (c.rope_n_elem, c.padded_vocab_size, c.intermediate_size) == (32, 50688, 16384)
