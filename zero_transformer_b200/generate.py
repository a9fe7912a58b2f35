"""Text generation -- the ``app.py`` of this framework.

Capabilities of the reference app (``app.py:30-189``): load a model by size +
checkpoint path, move to GPU in half precision, stream tokens with a KV cache,
temperature, repetition penalty, top-k, nucleus (top-p), greedy, EOS stop,
context cropping to the model's window.  The reference serves this through
Gradio; this module is the CLI and the generator API, ``serve.py`` is the browser UI (FastAPI + uvicorn, same controls)
and ``utils/tokenizer.py`` the GPT-NeoX-compatible byte-level BPE tokenizer.
"""
from __future__ import annotations

import argparse
import sys
from typing import Iterator, List, Optional

import torch
import torch.nn.functional as F

from .config import DEFAULT_MODEL_CFG
from .models.gpt import model_getter
from .utils import checkpoint as ckpt
from .utils.export import load_reference_pth


from .utils.tokenizer import BPETokenizer, ByteTokenizer, load_tokenizer  # noqa: E402,F401  (re-exported: public API of this module)


def model_creator(size: str, path: Optional[str], model_cfg: str = DEFAULT_MODEL_CFG, device: Optional[str] = None):
    """Build + load + move to device in inference precision (reference ``model_creator``, ``app.py:30-39``).
    ``path`` may be a reference-format ``.pth`` or one of our ``params_<step>.pt`` files/checkpoint dirs."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    model = model_getter(size, model_cfg, dtype=torch.float32)
    if path:
        if path.endswith(".pth"):
            load_reference_pth(model, path)
        else:
            import os

            if os.path.isdir(path):
                params, _ = ckpt.restore_param_checkpoint(path)
            else:
                params = torch.load(path, map_location="cpu", weights_only=False)["params"]
            model.load_state_dict(params, strict=True)
    model = model.to(dev)
    if dev.type == "cuda":
        model = model.to(torch.bfloat16)
    return model.eval()


def process_logits(logits: torch.Tensor, generated: torch.Tensor, temperature: float, repetition_penalty: float) -> torch.Tensor:
    """Temperature + CTRL-style repetition penalty on already-generated ids
    (reference ``process_logits``, ``app.py:97-108``)."""
    logits = logits.float()
    if repetition_penalty != 1.0 and generated.numel() > 0:
        idx = torch.unique(generated)
        picked = logits[..., idx]
        logits[..., idx] = torch.where(picked < 0, picked * repetition_penalty, picked / repetition_penalty)
    return logits / max(temperature, 1e-6)


def top_k_logits(logits: torch.Tensor, k: int) -> torch.Tensor:
    if k <= 0 or k >= logits.shape[-1]:
        return logits
    kth = torch.topk(logits, k, dim=-1).values[..., -1:]
    return logits.masked_fill(logits < kth, float("-inf"))


def top_p_logits(logits: torch.Tensor, p: float) -> torch.Tensor:
    """Nucleus filtering: keep the smallest prefix of the sorted distribution with mass >= p."""
    if p <= 0.0 or p >= 1.0:
        return logits
    srt, order = torch.sort(logits, descending=True, dim=-1)
    cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
    drop = cum - F.softmax(srt, dim=-1) >= p  # tokens strictly after the nucleus
    srt = srt.masked_fill(drop, float("-inf"))
    return torch.full_like(logits, float("-inf")).scatter(-1, order, srt)


@torch.no_grad()
def generate_from_prompt(model, prompt_ids: List[int], steps: int, *, temperature: float = 1.0, top_k: int = 0,
                         top_p: float = 0.0, repetition_penalty: float = 1.0, sample: bool = True,
                         eos_token_id: Optional[int] = None, seed: Optional[int] = None) -> Iterator[int]:
    """Yield one token id at a time (reference generator ``app.py:42-94``)."""
    dev = next(model.parameters()).device
    ctx = model.cfg.block_size
    gen = torch.Generator(device=dev)
    if seed is not None:
        gen.manual_seed(seed)
    x = torch.tensor(prompt_ids[-ctx:], dtype=torch.long, device=dev).view(1, -1)
    generated = x.clone()
    past = None
    inp = x
    for _ in range(steps):
        if past is not None and past[0][0].shape[2] >= ctx:
            # ALiBi extrapolates, but we keep the reference's window crop: restart the cache on the last ctx-1 tokens
            inp, past = generated[:, -(ctx - 1):], None
        logits, past = model(inp, use_cache=True, past_states=past)
        logits = process_logits(logits[:, -1, :], generated, temperature, repetition_penalty)
        logits = top_p_logits(top_k_logits(logits, top_k), top_p)
        probs = F.softmax(logits, dim=-1)
        nxt = torch.multinomial(probs, 1, generator=gen) if sample else torch.argmax(probs, dim=-1, keepdim=True)
        tok = int(nxt.item())
        generated = torch.cat([generated, nxt], dim=1)
        inp = nxt
        yield tok
        if eos_token_id is not None and tok == eos_token_id:
            return


def generate_text(model, tokenizer, prompt: str, steps: int = 64, **kw) -> str:
    ids = tokenizer.encode(prompt)
    out = list(generate_from_prompt(model, ids, steps, eos_token_id=getattr(tokenizer, "eos_token_id", None), **kw))
    return tokenizer.decode(out)


def launch_ui(model, tokenizer, model_name: str = "model", share: bool = False, port: int = 7860):
    """Serve the browser UI (the reference's Gradio Blocks page, ``app.py:199-261``) with FastAPI + uvicorn."""
    from .serve import launch

    launch(model, tokenizer, model_name, share=share, port=port)


def parse(argv=None):
    p = argparse.ArgumentParser(description="Sample from a trained model")
    p.add_argument("--model-size", default="417m")
    p.add_argument("--model-path", default=None)
    p.add_argument("--model-cfg", default=DEFAULT_MODEL_CFG)
    p.add_argument("--tokenizer", default=None, help="tokenizer.json / tokenizer directory (default: $ZTB_TOKENIZER, cached gpt-neox-20b, else bytes)")
    p.add_argument("--port", type=int, default=7860)
    p.add_argument("--prompt", default="Hello")
    p.add_argument("--steps", type=int, default=64)
    p.add_argument("--temperature", type=float, default=0.8)
    p.add_argument("--top-k", type=int, default=40)
    p.add_argument("--top-p", type=float, default=0.95)
    p.add_argument("--repetition-penalty", type=float, default=1.2)
    p.add_argument("--greedy", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--share", action="store_true", help="serve the browser UI on 0.0.0.0 (reference flag name)")
    p.add_argument("--ui", action="store_true")
    return p.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    model = model_creator(a.model_size, a.model_path, a.model_cfg)
    tok = load_tokenizer(a.tokenizer)
    if a.ui or a.share:
        return launch_ui(model, tok, a.model_size, a.share, a.port)
    sys.stdout.write(a.prompt)
    for t in generate_from_prompt(model, tok.encode(a.prompt), a.steps, temperature=a.temperature, top_k=a.top_k,
                                  top_p=a.top_p, repetition_penalty=a.repetition_penalty, sample=not a.greedy,
                                  eos_token_id=getattr(tok, "eos_token_id", None), seed=a.seed):
        sys.stdout.write(tok.decode([t]))
        sys.stdout.flush()
    sys.stdout.write("\n")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
