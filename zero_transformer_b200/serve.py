"""Browser UI + HTTP API for text generation -- the reference's Gradio app (``/root/reference/app.py:199-261``) rebuilt on
FastAPI / uvicorn (both ship in this image; gradio does not).

Same controls as the reference ``gr.Blocks`` layout: prompt box, number of tokens, temperature, top-k, top-p (nucleus),
repetition penalty, sampling method (Top-k / Nucleus / Greedy), terminate-on-EOS; the completion streams into the page
token by token.

    python app.py --model-size 417m --model-path out.pth --ui [--share] [--port 7860]

``--share`` binds 0.0.0.0 instead of 127.0.0.1 (there is no tunnelling service offline).

Endpoints
    GET  /            the page
    GET  /healthz     {"ok": true, "model": ..., "vocab": ...}
    POST /generate    JSON {prompt, steps, temperature, top_k, top_p, repetition_penalty, method, stop_on_eos, seed}
                      -> text/plain stream of decoded pieces   (``"stream": false`` -> JSON {"completion": ...})
"""
from __future__ import annotations

import threading
from typing import Iterator, Optional

from fastapi import FastAPI, Request
from fastapi.responses import HTMLResponse, JSONResponse, StreamingResponse

from .generate import generate_from_prompt

PAGE = """<!doctype html>
<html><head><meta charset="utf-8"><title>zero_transformer_b200</title>
<style>
 body{font-family:system-ui,sans-serif;max-width:1100px;margin:2em auto;padding:0 1em;color:#222}
 .row{display:flex;gap:2em}.col{flex:1}textarea{width:100%;height:14em;font:inherit}
 label{display:block;margin:.6em 0 .2em;font-size:.9em}input[type=range]{width:100%}
 #out{white-space:pre-wrap;border:1px solid #ccc;border-radius:6px;padding:1em;min-height:14em}
 #out .gen{color:#1d4ed8}button{margin-top:1em;padding:.6em 1.4em;font:inherit}
 details{margin-top:1em}
</style></head><body>
<h2>zero_transformer_b200 &mdash; __MODEL__</h2>
<div class="row"><div class="col">
 <label>Enter your text here</label><textarea id="prompt"></textarea>
 <label>Number of tokens to generate: <span id="v_steps">100</span></label><input id="steps" type="range" min="1" max="1000" value="100">
 <details><summary>Generation Parameters</summary>
  <label>Temperature: <span id="v_temperature">0.8</span></label><input id="temperature" type="range" min="0.05" max="2" step="0.05" value="0.8">
  <label>k (Top-k Sampling): <span id="v_top_k">40</span></label><input id="top_k" type="range" min="0" max="50" value="40">
  <label>p (Nucleus Sampling): <span id="v_top_p">0.96</span></label><input id="top_p" type="range" min="0" max="1" step="0.01" value="0.96">
  <label>Repetition Penalty: <span id="v_repetition_penalty">1.2</span></label><input id="repetition_penalty" type="range" min="1" max="1.3" step="0.01" value="1.2">
  <label>Sampling Method</label><select id="method"><option>Nucleus</option><option>Top-k</option><option>Greedy</option></select>
  <label><input id="stop_on_eos" type="checkbox" checked> Terminate generation on EOS token.</label>
 </details>
 <button id="go">Generate Text</button>
</div><div class="col"><label>Generated Text</label><div id="out"></div></div></div>
<script>
for (const id of ["steps","temperature","top_k","top_p","repetition_penalty"]) {
  const el = document.getElementById(id); el.oninput = () => document.getElementById("v_"+id).textContent = el.value;
}
document.getElementById("go").onclick = async () => {
  const g = id => document.getElementById(id);
  const body = {prompt: g("prompt").value, steps: +g("steps").value, temperature: +g("temperature").value, top_k: +g("top_k").value,
                top_p: +g("top_p").value, repetition_penalty: +g("repetition_penalty").value, method: g("method").value,
                stop_on_eos: g("stop_on_eos").checked};
  const out = g("out"); out.textContent = body.prompt; const span = document.createElement("span"); span.className = "gen"; out.appendChild(span);
  const resp = await fetch("generate", {method: "POST", headers: {"Content-Type": "application/json"}, body: JSON.stringify(body)});
  const reader = resp.body.getReader(); const dec = new TextDecoder();
  for (;;) { const {done, value} = await reader.read(); if (done) break; span.textContent += dec.decode(value, {stream: true}); }
};
</script></body></html>"""


def sampling_kwargs(method: str, top_k: int, top_p: float) -> dict:
    """The reference's sampling-method dropdown (``app.py:233-238``): Top-k ignores p, Nucleus ignores k, Greedy ignores both."""
    m = (method or "Nucleus").lower()
    if m.startswith("greedy"):
        return dict(sample=False, top_k=0, top_p=0.0)
    if m.startswith("top"):
        return dict(sample=True, top_k=int(top_k), top_p=0.0)
    return dict(sample=True, top_k=0, top_p=float(top_p))


def stream_completion(model, tokenizer, req: dict) -> Iterator[str]:
    """Decoded text pieces, one per generated token.  Byte-level BPE tokens can end in the middle of a UTF-8 character:
    decode the running id list and emit only the new, complete suffix."""
    ids = tokenizer.encode(str(req.get("prompt", "")).strip()) or [getattr(tokenizer, "eos_token_id", 0) or 0]
    kw = sampling_kwargs(req.get("method", "Nucleus"), int(req.get("top_k", 40)), float(req.get("top_p", 0.96)))
    eos = getattr(tokenizer, "eos_token_id", None) if req.get("stop_on_eos", True) else None
    gen, sent = [], 0
    for tok in generate_from_prompt(model, ids, int(req.get("steps", 100)), temperature=float(req.get("temperature", 0.8)),
                                    repetition_penalty=float(req.get("repetition_penalty", 1.2)), eos_token_id=eos,
                                    seed=req.get("seed"), **kw):
        gen.append(tok)
        text = tokenizer.decode(gen)
        stable = len(text) - (1 if text.endswith("�") else 0)   # hold back an incomplete multi-byte character
        if stable > sent:
            yield text[sent:stable]
            sent = stable
    text = tokenizer.decode(gen)
    if len(text) > sent:
        yield text[sent:]


def create_app(model, tokenizer, model_name: str = "model"):
    app = FastAPI(title="zero_transformer_b200")
    lock = threading.Lock()   # one generation at a time: the KV cache and the sampling RNG live on one device

    @app.get("/", response_class=HTMLResponse)
    def index():
        return PAGE.replace("__MODEL__", model_name)

    @app.get("/healthz")
    def healthz():
        return {"ok": True, "model": model_name, "vocab": int(model.cfg.vocab_size), "context": int(model.cfg.block_size)}

    @app.post("/generate")
    async def generate(request: Request):
        req = await request.json()
        if not isinstance(req, dict):
            return JSONResponse({"error": "expected a JSON object"}, status_code=400)
        try:
            steps = int(req.get("steps", 100))
        except (TypeError, ValueError):
            return JSONResponse({"error": "steps must be an integer"}, status_code=400)
        if not 1 <= steps <= 4096:
            return JSONResponse({"error": "steps must be in [1, 4096]"}, status_code=400)

        def run():
            with lock:
                yield from stream_completion(model, tokenizer, req)

        if req.get("stream", True):
            return StreamingResponse(run(), media_type="text/plain; charset=utf-8")
        return {"completion": "".join(run())}

    return app


def launch(model, tokenizer, model_name: str = "model", share: bool = False, port: int = 7860, host: Optional[str] = None) -> None:  # pragma: no cover
    import uvicorn

    uvicorn.run(create_app(model, tokenizer, model_name), host=host or ("0.0.0.0" if share else "127.0.0.1"), port=port, log_level="info")
