"""Export to / import from the reference's PyTorch ``.pth`` contract, and
import of the reference's Flax msgpack parameter files.

Contract (produced upstream by ``torch_compatability/flax_to_pytorch.py:70-117``
and consumed by ``torch_compatability/GPT2.py:448-474`` with a *strict*
``load_state_dict``): a flat fp32 ``state_dict`` with ``18*N + 4`` entries (measured on the reference class; SURVEY 3.4 miscounts it as 19N+3) --

    wte.weight, lm_head.weight (same values), norm.{weight,bias}
    blocks.i.ln1.{weight,bias}  blocks.i.ln2.{weight,bias}
    blocks.i.attn.{key,query,value,fc_resid}.{weight,bias}
    blocks.i.attn.slopes (H,)   blocks.i.attn.mask (1,1,ctx,ctx) uint8 tril
    blocks.i.mlp.{fc1,fc_resid}.{weight,bias}

All biases are zero (the trained model has none).  Our fused ``qkv`` weight
``[Wq; Wk; Wv]`` is split on export and re-fused on import.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from ..ops.reference import get_slopes


def to_reference_state_dict(params: Mapping[str, torch.Tensor], *, num_head: int, num_ctx: int,
                            vocab_size: Optional[int] = None) -> "OrderedDict[str, torch.Tensor]":
    """``params``: our parameter names -> tensors (any dtype).  Returns the
    reference-format state dict (fp32)."""
    f = lambda t: t.detach().to("cpu", torch.float32).contiguous().clone()
    wte = f(params["wte.weight"])
    if vocab_size is not None:
        wte = wte[:vocab_size].contiguous()
    D = wte.shape[1]
    n_blocks = len({k.split(".")[1] for k in params if k.startswith("blocks.")})
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["wte.weight"] = wte
    slopes = torch.tensor(get_slopes(num_head), dtype=torch.float32)
    mask = torch.tril(torch.ones(num_ctx, num_ctx, dtype=torch.uint8)).view(1, 1, num_ctx, num_ctx)
    for i in range(n_blocks):
        p = f"blocks.{i}."
        sd[p + "ln1.weight"] = f(params[p + "ln1.weight"])
        sd[p + "ln1.bias"] = torch.zeros(D)
        sd[p + "ln2.weight"] = f(params[p + "ln2.weight"])
        sd[p + "ln2.bias"] = torch.zeros(D)
        sd[p + "attn.slopes"] = slopes.clone()
        sd[p + "attn.mask"] = mask.clone()
        wq, wk, wv = f(params[p + "attn.qkv.weight"]).chunk(3, dim=0)
        for name, w in (("key", wk), ("query", wq), ("value", wv), ("fc_resid", f(params[p + "attn.out.weight"]))):
            sd[p + f"attn.{name}.weight"] = w.contiguous()
            sd[p + f"attn.{name}.bias"] = torch.zeros(w.shape[0])
        w1 = f(params[p + "mlp.fc_in.weight"])
        w2 = f(params[p + "mlp.fc_out.weight"])
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = w1, torch.zeros(w1.shape[0])
        sd[p + "mlp.fc_resid.weight"], sd[p + "mlp.fc_resid.bias"] = w2, torch.zeros(w2.shape[0])
    sd["norm.weight"] = f(params["ln_f.weight"])
    sd["norm.bias"] = torch.zeros(D)
    sd["lm_head.weight"] = sd["wte.weight"]  # tied: same storage, like the reference exporter
    return sd


def save_reference_pth(params: Mapping[str, torch.Tensor], path: str, *, num_head: int, num_ctx: int,
                       vocab_size: Optional[int] = None) -> None:
    torch.save(to_reference_state_dict(params, num_head=num_head, num_ctx=num_ctx, vocab_size=vocab_size), path)


def from_reference_state_dict(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference ``.pth`` -> our parameter names (biases must be zero: our model has none)."""
    out: Dict[str, torch.Tensor] = {"wte.weight": sd["wte.weight"].float().clone(), "ln_f.weight": sd["norm.weight"].float().clone()}
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("blocks.")})
    for k, v in sd.items():
        if k.endswith(".bias") and float(v.abs().max()) != 0.0:
            raise ValueError(f"{k} is non-zero; this architecture is bias-free")
    for i in range(n_blocks):
        p = f"blocks.{i}."
        out[p + "ln1.weight"] = sd[p + "ln1.weight"].float().clone()
        out[p + "ln2.weight"] = sd[p + "ln2.weight"].float().clone()
        out[p + "attn.qkv.weight"] = torch.cat([sd[p + "attn.query.weight"], sd[p + "attn.key.weight"], sd[p + "attn.value.weight"]], 0).float()
        out[p + "attn.out.weight"] = sd[p + "attn.fc_resid.weight"].float().clone()
        out[p + "mlp.fc_in.weight"] = sd[p + "mlp.fc1.weight"].float().clone()
        out[p + "mlp.fc_out.weight"] = sd[p + "mlp.fc_resid.weight"].float().clone()
    return out


def load_reference_pth(model: torch.nn.Module, path: str) -> None:
    sd = torch.load(path, map_location="cpu", weights_only=False)
    ours = from_reference_state_dict(sd)
    missing = model.load_state_dict(ours, strict=True)
    return missing


# ---------------------------------------------------------------------------
# Flax msgpack import (capability of torch_compatability/extract_msgpack.py +
# flax_to_pytorch.py, without needing flax): decodes flax.serialization's
# msgpack encoding (ndarray = ExtType(1, packb((shape, dtype, bytes))); arrays
# above 2^30 bytes are stored chunked under '__msgpack_chunked_array__').
# ---------------------------------------------------------------------------
def _decode_flax_msgpack(raw: bytes):
    import msgpack

    def ext_hook(code, data):
        if code == 1:
            shape, dtype, buf = msgpack.unpackb(data, raw=False)
            if dtype == "bfloat16":
                arr = np.frombuffer(buf, dtype=np.uint16).reshape(shape)
                return torch.from_numpy(arr.copy()).view(torch.bfloat16).float().numpy()
            return np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape).copy()
        if code == 3:  # native numpy scalar
            shape, dtype, buf = msgpack.unpackb(data, raw=False)
            return np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape)[()]
        return msgpack.ExtType(code, data)

    tree = msgpack.unpackb(raw, ext_hook=ext_hook, raw=False, strict_map_key=False)

    def unchunk(node):
        if isinstance(node, dict):
            if node.get("__msgpack_chunked_array__"):
                chunks = [node["chunks"][str(i)] for i in range(len(node["chunks"]))]
                return np.concatenate([c.reshape(-1) for c in chunks]).reshape(node["shape"])
            return {k: unchunk(v) for k, v in node.items()}
        return node

    return unchunk(tree)


def from_flax_params(tree: Mapping) -> Dict[str, torch.Tensor]:
    """Flax parameter pytree of the reference model (``wte/embedding``,
    ``TransformerBlock_i/{LayerNorm_0,LayerNorm_1,CausalAttention_0/{query_proj,key_proj,value_proj,residual_out},MLPBlock_0/{fc_in,fc_residual}}``,
    ``LayerNorm_0/scale``; name map of ``flax_to_pytorch.py:6-35``) -> our names.
    2-D kernels are ``(in, out)`` in flax and transposed to ``(out, in)`` here."""
    if "params" in tree:
        tree = tree["params"]
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    out: Dict[str, torch.Tensor] = {"wte.weight": t(tree["wte"]["embedding"]), "ln_f.weight": t(tree["LayerNorm_0"]["scale"])}
    i = 0
    while f"TransformerBlock_{i}" in tree:
        b = tree[f"TransformerBlock_{i}"]
        a, m = b["CausalAttention_0"], b["MLPBlock_0"]
        p = f"blocks.{i}."
        out[p + "ln1.weight"] = t(b["LayerNorm_0"]["scale"])
        out[p + "ln2.weight"] = t(b["LayerNorm_1"]["scale"])
        out[p + "attn.qkv.weight"] = torch.cat([t(a[k]["kernel"]).T for k in ("query_proj", "key_proj", "value_proj")], 0).contiguous()
        out[p + "attn.out.weight"] = t(a["residual_out"]["kernel"]).T.contiguous()
        out[p + "mlp.fc_in.weight"] = t(m["fc_in"]["kernel"]).T.contiguous()
        out[p + "mlp.fc_out.weight"] = t(m["fc_residual"]["kernel"]).T.contiguous()
        i += 1
    return out


def load_flax_msgpack(path: str) -> Dict[str, torch.Tensor]:
    with open(path, "rb") as f:
        return from_flax_params(_decode_flax_msgpack(f.read()))
