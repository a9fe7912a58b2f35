"""Checkpoint conversion CLI.

Covers the reference's two converter scripts: ``torch_compatability/extract_msgpack.py:28-62``
(training checkpoint -> bare parameter file) and ``torch_compatability/convert_to_torch.py:26-31``
(``--model-name --flax-path --torch-path [--vocab-size --seq-len]`` -> ``.pth``).

    python -m zero_transformer_b200.utils.convert extract --ckpt-dir checkpoints --step 1000 --out params.pt
    python -m zero_transformer_b200.utils.convert to-torch --model-name 1_3b --params params.pt --torch-path out.pth
    python -m zero_transformer_b200.utils.convert to-torch --model-name 1_3b --flax-path model_params_1000.msgpack --torch-path out.pth
    python -m zero_transformer_b200.utils.convert from-torch --torch-path out.pth --out params.pt
"""
from __future__ import annotations

import argparse

import torch

from ..config import DEFAULT_MODEL_CFG, load_yaml
from . import checkpoint as ckpt
from .export import from_reference_state_dict, load_flax_msgpack, save_reference_pth


def main(argv=None):
    p = argparse.ArgumentParser()
    sub = p.add_subparsers(dest="cmd", required=True)
    e = sub.add_parser("extract")
    e.add_argument("--ckpt-dir", required=True)
    e.add_argument("--step", type=int, default=None)
    e.add_argument("--out", required=True)
    t = sub.add_parser("to-torch")
    t.add_argument("--model-name", required=True)
    t.add_argument("--model-cfg", default=DEFAULT_MODEL_CFG)
    t.add_argument("--params", default=None, help="our params_<step>.pt / extracted file")
    t.add_argument("--flax-path", default=None, help="reference flax msgpack parameter file")
    t.add_argument("--torch-path", required=True)
    t.add_argument("--vocab-size", type=int, default=None)
    t.add_argument("--seq-len", type=int, default=None, help="num_ctx of the exported causal mask buffer")
    f = sub.add_parser("from-torch")
    f.add_argument("--torch-path", required=True)
    f.add_argument("--out", required=True)
    a = p.parse_args(argv)

    if a.cmd == "extract":
        params, step = ckpt.restore_param_checkpoint(a.ckpt_dir, a.step)
        torch.save({"step": step, "params": params}, a.out)
        print(f"wrote {a.out} (step {step}, {len(params)} tensors)")
    elif a.cmd == "to-torch":
        entry = load_yaml(a.model_cfg)[a.model_name]
        if a.flax_path:
            params = load_flax_msgpack(a.flax_path)
        else:
            params = torch.load(a.params, map_location="cpu", weights_only=False)["params"]
        ctx = a.seq_len or int(entry.get("block_size", entry.get("num_ctx", 1024)))
        save_reference_pth(params, a.torch_path, num_head=int(entry["num_head"]), num_ctx=ctx, vocab_size=a.vocab_size)
        print(f"wrote {a.torch_path}")
    else:
        sd = torch.load(a.torch_path, map_location="cpu", weights_only=False)
        torch.save({"step": 0, "params": from_reference_state_dict(sd)}, a.out)
        print(f"wrote {a.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
