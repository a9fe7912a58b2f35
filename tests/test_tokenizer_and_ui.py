"""Byte-level BPE tokenizer (GPT-NeoX file format) and the FastAPI generation UI, on CPU."""
import json

import pytest
import torch

from zero_transformer_b200.utils.tokenizer import BPETokenizer, ByteTokenizer, load_tokenizer, train_bpe

CORPUS = [
    "The quick brown fox jumps over the lazy dog. The dog barks; the fox runs away!",
    "ZeRO stage one shards the optimizer state: fp32 master weights and Adam moments live on 1/N of the ranks.",
    "def reduce_scatter(x):\n    return [sum(c) for c in zip(*x)]  # héllo wörld — naïve café 你好",
    "Tensor cores multiply 128x256 tiles;   whitespace   runs\tand\ttabs\n\nnewlines are tokens too.",
] * 20


def test_trained_bpe_roundtrips_and_compresses():
    tok = train_bpe(CORPUS, vocab_size=600)
    assert 300 < tok.vocab_size <= 600
    for s in CORPUS[:4] + ["unseen words: tokenizer!?", "", "   ", "emoji 🙂 ok"]:
        ids = tok.encode(s)
        assert tok.decode(ids) == s
    assert len(tok.encode(CORPUS[0])) < len(CORPUS[0].encode()) * 0.6
    eos = tok.eos_token_id
    ids = tok.encode("a<|endoftext|>b")
    assert eos in ids and tok.decode(ids) == "a<|endoftext|>b" and tok.decode(ids, skip_special=True) == "ab"


def test_matches_hf_tokenizers_on_the_same_file(tmp_path):
    """Train with the `tokenizers` library (the GPT-NeoX toolchain), save tokenizer.json, load it with our reader:
    identical ids, and our writer's file loads back into `tokenizers` with identical ids."""
    tk = pytest.importorskip("tokenizers")
    t = tk.Tokenizer(tk.models.BPE())
    t.pre_tokenizer = tk.pre_tokenizers.ByteLevel(add_prefix_space=False)
    t.decoder = tk.decoders.ByteLevel()
    t.train_from_iterator(CORPUS, tk.trainers.BpeTrainer(vocab_size=500, special_tokens=["<|endoftext|>"],
                                                          initial_alphabet=tk.pre_tokenizers.ByteLevel.alphabet()))
    path = str(tmp_path / "tokenizer.json")
    t.save(path)
    ours = BPETokenizer.from_file(path)
    probes = CORPUS[:4] + ["Completely new sentence, with numbers 12345 and   spaces.", "x<|endoftext|>y"]
    for s in probes:
        assert ours.encode(s) == t.encode(s).ids, s
        assert ours.decode(ours.encode(s)) == s
    out = str(tmp_path / "resaved" / "tokenizer.json")
    ours.save(out)
    back = tk.Tokenizer.from_file(out)
    for s in probes:
        assert back.encode(s).ids == ours.encode(s), s
    assert load_tokenizer(str(tmp_path)).encode(probes[0]) == ours.encode(probes[0])   # directory form


def test_load_tokenizer_fallback_is_bytes(monkeypatch, tmp_path, capsys):
    monkeypatch.delenv("ZTB_TOKENIZER", raising=False)
    monkeypatch.setenv("HF_HOME", str(tmp_path))
    monkeypatch.setenv("HOME", str(tmp_path))
    tok = load_tokenizer(None)
    assert isinstance(tok, ByteTokenizer)
    assert "falling back" in capsys.readouterr().err


def _tiny_model():
    from zero_transformer_b200.models.gpt import GPTConfig, Transformer

    torch.manual_seed(0)
    return Transformer(GPTConfig(embedding_dim=32, vocab_size=600, num_head=2, block_size=32, N=1, dropout=0.0)).eval()


def test_ui_endpoints_stream_a_completion():
    from starlette.testclient import TestClient

    from zero_transformer_b200.serve import create_app, sampling_kwargs

    tok = train_bpe(CORPUS, vocab_size=600)
    app = create_app(_tiny_model(), tok, "tiny")
    c = TestClient(app)
    page = c.get("/")
    assert page.status_code == 200 and "Generate Text" in page.text and "Repetition Penalty" in page.text
    assert c.get("/healthz").json() == {"ok": True, "model": "tiny", "vocab": 600, "context": 32}
    req = {"prompt": "The quick", "steps": 12, "temperature": 0.9, "top_k": 20, "top_p": 0.9, "repetition_penalty": 1.1,
           "method": "Greedy", "stop_on_eos": False}
    with c.stream("POST", "/generate", json=req) as r:
        assert r.status_code == 200
        streamed = "".join(r.iter_text())
    whole = c.post("/generate", json=dict(req, stream=False)).json()["completion"]
    assert streamed == whole and len(tok.encode(whole)) >= 1        # greedy: deterministic
    assert c.post("/generate", json=dict(req, steps=0)).status_code == 400
    assert sampling_kwargs("Top-k", 7, 0.5) == dict(sample=True, top_k=7, top_p=0.0)
    assert sampling_kwargs("Nucleus", 7, 0.5) == dict(sample=True, top_k=0, top_p=0.5)
    assert sampling_kwargs("Greedy", 7, 0.5)["sample"] is False
    sampled = c.post("/generate", json=dict(req, method="Nucleus", seed=3, stream=False)).json()["completion"]
    again = c.post("/generate", json=dict(req, method="Nucleus", seed=3, stream=False)).json()["completion"]
    assert sampled == again
