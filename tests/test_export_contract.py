"""The exported ``.pth`` must strict-load into the reference's own
``torch_compatability.GPT2.model_getter`` and give the same logits (the reference
only checks tensor equality, never outputs -- SURVEY 4).  ``omegaconf`` is not
installed, so a tiny shim (OmegaConf.load -> yaml) is injected for the import."""
import importlib.util
import os
import sys
import types

import pytest
import torch
import yaml

from zero_transformer_b200.models.gpt import GPTConfig, Transformer
from zero_transformer_b200.utils.export import (from_reference_state_dict, load_flax_msgpack, save_reference_pth,
                                                to_reference_state_dict)

REF = "/root/reference/torch_compatability/GPT2.py"


def _load_reference_gpt2():
    if "omegaconf" not in sys.modules:
        shim = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def load(path):
                with open(path) as f:
                    return yaml.safe_load(f)

        shim.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = shim
    spec = importlib.util.spec_from_file_location("ref_gpt2", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _tiny(gelu="erf"):
    torch.manual_seed(0)
    return Transformer(GPTConfig(embedding_dim=64, vocab_size=256, num_head=4, block_size=8, N=2, dropout=0.0, gelu=gelu)).eval()


def test_state_dict_contract_shape():
    m = _tiny()
    sd = to_reference_state_dict(dict(m.named_parameters()), num_head=4, num_ctx=8)
    assert len(sd) == 18 * 2 + 4
    assert sd["blocks.0.attn.mask"].dtype == torch.uint8 and sd["blocks.0.attn.mask"].shape == (1, 1, 8, 8)
    assert sd["blocks.1.attn.slopes"].shape == (4,)
    assert all(float(v.abs().max()) == 0 for k, v in sd.items() if k.endswith(".bias"))
    assert torch.equal(sd["wte.weight"], sd["lm_head.weight"])
    back = from_reference_state_dict(sd)
    for n, p in m.named_parameters():
        torch.testing.assert_close(back[n], p.detach())


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not mounted")
def test_pth_strict_loads_in_reference_and_logits_match(tmp_path):
    ref = _load_reference_gpt2()
    m = _tiny(gelu="erf")  # the torch consumer uses erf GELU (GPT2.py:66); training default is tanh
    out = tmp_path / "test.pth"
    save_reference_pth(dict(m.named_parameters()), str(out), num_head=4, num_ctx=8)
    cfgp = tmp_path / "model_config.yaml"
    cfgp.write_text("test:\n  embedding_dim: 64\n  vocab_size: 256\n  num_head: 4\n  num_ctx: 8\n  N: 2\n")
    their = ref.model_getter("test", config_path=str(cfgp), model_checkpoint=str(out)).eval()  # strict load_state_dict
    x = torch.randint(0, 256, (2, 8))
    with torch.no_grad():
        torch.testing.assert_close(their(x), m(x), rtol=1e-4, atol=1e-4)
        lt, losst = their(x, labels=x)
        lo, losso = m(x, labels=x)
        torch.testing.assert_close(losst, losso, rtol=1e-4, atol=1e-5)


def test_flax_msgpack_import(tmp_path):
    """Decode the flax.serialization msgpack layout (ndarray = ExtType(1, packb((shape, dtype, bytes))))."""
    import msgpack
    import numpy as np

    rng = np.random.default_rng(0)
    D, V = 8, 16

    def arr(*shape):
        return rng.standard_normal(shape).astype(np.float32)

    blk = lambda: {"LayerNorm_0": {"scale": arr(D)}, "LayerNorm_1": {"scale": arr(D)},
                   "CausalAttention_0": {k: {"kernel": arr(D, D)} for k in ("query_proj", "key_proj", "value_proj", "residual_out")},
                   "MLPBlock_0": {"fc_in": {"kernel": arr(D, 4 * D)}, "fc_residual": {"kernel": arr(4 * D, D)}}}
    tree = {"params": {"wte": {"embedding": arr(V, D)}, "LayerNorm_0": {"scale": arr(D)},
                       "TransformerBlock_0": blk(), "TransformerBlock_1": blk()}}

    def default(o):
        if isinstance(o, np.ndarray):
            return msgpack.ExtType(1, msgpack.packb((o.shape, o.dtype.name, o.tobytes()), use_bin_type=True))
        raise TypeError

    p = tmp_path / "model_params_0.msgpack"
    p.write_bytes(msgpack.packb(tree, default=default, use_bin_type=True))
    ours = load_flax_msgpack(str(p))
    a = tree["params"]["TransformerBlock_1"]["CausalAttention_0"]
    np.testing.assert_array_equal(ours["blocks.1.attn.qkv.weight"][D:2 * D].numpy(), a["key_proj"]["kernel"].T)
    np.testing.assert_array_equal(ours["blocks.0.mlp.fc_in.weight"].numpy(), tree["params"]["TransformerBlock_0"]["MLPBlock_0"]["fc_in"]["kernel"].T)
    m = Transformer(GPTConfig(embedding_dim=D, vocab_size=V, num_head=2, block_size=8, N=2))
    m.load_state_dict(ours, strict=True)


def test_torch_compat_model_getter_loads_reference_pth(tmp_path):
    from zero_transformer_b200.models.torch_compat import model_getter
    from zero_transformer_b200.parallel.partition import create_opt_spec, layout_for, set_partitions_zero

    m = _tiny(gelu="erf")
    out = tmp_path / "t.pth"
    save_reference_pth(dict(m.named_parameters()), str(out), num_head=4, num_ctx=8)
    cfgp = tmp_path / "model_config.yaml"
    cfgp.write_text("test:\n  embedding_dim: 64\n  vocab_size: 256\n  num_head: 4\n  num_ctx: 8\n  N: 2\n")
    m2 = model_getter("test", str(cfgp), str(out))
    x = torch.randint(0, 256, (1, 8))
    with torch.no_grad():
        torch.testing.assert_close(m2(x), m(x))
    spec = set_partitions_zero({n: tuple(p.shape) for n, p in m.named_parameters()}, 4)
    assert spec["wte.weight"] == ("dp", None) and spec["ln_f.weight"] == ("dp",)
    assert set(create_opt_spec(spec)) == {"master", "mu", "nu", "count"}
    L = layout_for(m, 4)
    assert L.total == L.shard_total * 4
