"""Text generation -- the ``app.py`` of this framework.

Capabilities of the reference app (``app.py:30-189``): load a model by size +
checkpoint path, move to GPU in half precision, stream tokens with a KV cache,
temperature, repetition penalty, top-k, nucleus (top-p), greedy, EOS stop,
context cropping to the model's window.  The reference serves this through
Gradio (not installed here); this module is a CLI and a generator API, and
``launch_gradio`` wraps the same generator when gradio is importable.
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


class ByteTokenizer:
    """Fallback tokenizer (UTF-8 bytes) for models with a >=256 vocabulary when no
    HF tokenizer directory is available offline (the reference downloads
    ``EleutherAI/gpt-neox-20b``, ``app.py:27``)."""

    eos_token_id = 0

    def encode(self, s: str) -> List[int]:
        return list(s.encode("utf-8"))

    def decode(self, ids) -> str:
        return bytes(int(i) % 256 for i in ids).decode("utf-8", errors="replace")


def load_tokenizer(path: Optional[str]):
    if path:
        from transformers import AutoTokenizer

        return AutoTokenizer.from_pretrained(path, local_files_only=True)
    return ByteTokenizer()


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


def launch_gradio(model, tokenizer, share: bool = False):  # pragma: no cover - gradio not installed in this image
    import gradio as gr

    def fn(prompt, steps, temperature, top_k, top_p, rep, sampling):
        text = ""
        for tok in generate_from_prompt(model, tokenizer.encode(prompt), int(steps), temperature=temperature, top_k=int(top_k),
                                        top_p=top_p, repetition_penalty=rep, sample=(sampling == "sample"),
                                        eos_token_id=getattr(tokenizer, "eos_token_id", None)):
            text += tokenizer.decode([tok])
            yield text

    ui = gr.Interface(fn, [gr.Textbox(lines=4, label="prompt"), gr.Slider(1, 512, 64, step=1, label="steps"),
                           gr.Slider(0.1, 2.0, 0.8, label="temperature"), gr.Slider(0, 200, 40, step=1, label="top-k"),
                           gr.Slider(0.0, 1.0, 0.95, label="top-p"), gr.Slider(1.0, 2.0, 1.2, label="repetition penalty"),
                           gr.Radio(["sample", "greedy"], value="sample")], gr.Textbox(label="completion"))
    ui.queue().launch(share=share)


def parse(argv=None):
    p = argparse.ArgumentParser(description="Sample from a trained model")
    p.add_argument("--model-size", default="417m")
    p.add_argument("--model-path", default=None)
    p.add_argument("--model-cfg", default=DEFAULT_MODEL_CFG)
    p.add_argument("--tokenizer", default=None, help="local HF tokenizer directory (default: byte tokenizer)")
    p.add_argument("--prompt", default="Hello")
    p.add_argument("--steps", type=int, default=64)
    p.add_argument("--temperature", type=float, default=0.8)
    p.add_argument("--top-k", type=int, default=40)
    p.add_argument("--top-p", type=float, default=0.95)
    p.add_argument("--repetition-penalty", type=float, default=1.2)
    p.add_argument("--greedy", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--share", action="store_true", help="serve the Gradio UI (needs gradio)")
    p.add_argument("--ui", action="store_true")
    return p.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    model = model_creator(a.model_size, a.model_path, a.model_cfg)
    tok = load_tokenizer(a.tokenizer)
    if a.ui or a.share:
        return launch_gradio(model, tok, a.share)
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
