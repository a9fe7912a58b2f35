"""Data pipeline (untested in the reference)."""
import io
import tarfile

import numpy as np
import pytest
import torch

from zero_transformer_b200.data.pipeline import (DevicePrefetcher, StreamingTokenDataset, SyntheticTokenDataset,
                                                 batch_iterator, read_index, rechunk, split_by_rank, to_micro_batches)


def _write_shards(tmp_path, n_shards=4, rows=10, ctx=16):
    names = []
    for s in range(n_shards):
        arr = (np.arange(rows * ctx).reshape(rows, ctx) + 1000 * s).astype(np.int32)
        p = tmp_path / f"shard{s}.npy"
        np.save(p, arr)
        names.append(p.name)
    idx = tmp_path / "train.index"
    idx.write_text("# comment\n" + "\n".join(names) + "\n\n")
    return str(idx)


def test_index_and_rank_striding(tmp_path):
    idx = _write_shards(tmp_path)
    shards = read_index(idx)
    assert len(shards) == 4 and all(s.endswith(".npy") for s in shards)
    assert split_by_rank(shards, 1, 2) == shards[1::2]
    a = {int(r[0]) for r in StreamingTokenDataset(idx, max_context=16, rank=0, world=2, shuffle_buffer=4)}
    b = {int(r[0]) for r in StreamingTokenDataset(idx, max_context=16, rank=1, world=2, shuffle_buffer=4)}
    assert len(a) == len(b) == 20 and not (a & b)


def test_shuffle_is_seeded(tmp_path):
    idx = _write_shards(tmp_path)
    f = lambda seed: [int(r[0]) for r in StreamingTokenDataset(idx, max_context=16, shuffle_buffer=8, seed=seed)]
    assert f(23) == f(23) and f(23) != f(24) and sorted(f(23)) == sorted(f(24))


def test_tar_shard_with_pth_members_and_corrupt_sample(tmp_path):
    p = tmp_path / "s.tar"
    with tarfile.open(p, "w") as tf:
        for i in range(3):
            buf = io.BytesIO()
            torch.save(torch.arange(20) + i, buf)
            data = buf.getvalue()
            ti = tarfile.TarInfo(f"{i:06d}.input_id.pth")
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
        bad = b"not a tensor"
        ti = tarfile.TarInfo("bad.input_id.pth")
        ti.size = len(bad)
        tf.addfile(ti, io.BytesIO(bad))
    (tmp_path / "v.index").write_text("s.tar\n")
    rows = list(StreamingTokenDataset(str(tmp_path / "v.index"), max_context=16, shuffle_buffer=1))
    assert len(rows) == 3 and rows[0].shape == (16,) and rows[0].dtype == np.int32


def test_rechunk_and_micro_batches():
    b = torch.arange(4 * 8).view(4, 8)
    r = rechunk(b, 4)
    assert r.shape == (8, 4) and torch.equal(r[1], torch.tensor([4, 5, 6, 7]))
    mb = to_micro_batches(r, 2)
    assert mb.shape == (2, 4, 4)
    with pytest.raises(AssertionError):
        to_micro_batches(r, 3)


def test_synthetic_is_deterministic_and_prefetcher_runs():
    a = [x.copy() for _, x in zip(range(3), SyntheticTokenDataset(100, 8, rank=0, seed=1))]
    b = [x.copy() for _, x in zip(range(3), SyntheticTokenDataset(100, 8, rank=0, seed=1))]
    c = [x.copy() for _, x in zip(range(3), SyntheticTokenDataset(100, 8, rank=1, seed=1))]
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not np.array_equal(a[0], c[0])
    pf = DevicePrefetcher(batch_iterator(SyntheticTokenDataset(100, 8, length=10), 4), "cpu")
    got = list(pf)
    assert len(got) == 2 and got[0].shape == (4, 8) and got[0].max() < 100


def test_make_shards_roundtrip(tmp_path):
    from zero_transformer_b200.data import make_shards

    toks = np.arange(16 * 50, dtype=np.int32)
    np.save(tmp_path / "toks.npy", toks)
    idx = tmp_path / "index" / "train.index"
    assert make_shards.main(["--tokens", str(tmp_path / "toks.npy"), "--out", str(tmp_path / "shards"), "--index", str(idx),
                             "--max-context", "16", "--samples-per-shard", "20"]) == 0
    rows = list(StreamingTokenDataset(str(idx), max_context=16, shuffle_buffer=1))
    assert len(rows) == 50 and sorted(int(r[0]) for r in rows) == [16 * i for i in range(50)]


def test_shard_urls_go_through_openers_and_fail_per_shard(tmp_path, capsys):
    """Index files may list URLs (the reference's were GCS paths, main_zero.py:368-421): they are kept verbatim, opened
    through the registered transport, and a transport that is unavailable only costs that shard."""
    import io

    from zero_transformer_b200.data import pipeline as P

    arr = np.arange(4 * 16, dtype=np.int32).reshape(4, 16)
    np.save(tmp_path / "a.npy", arr)
    buf = io.BytesIO()
    np.save(buf, arr + 1000)
    P.register_opener("mem", lambda url: io.BytesIO(buf.getvalue()))
    idx = tmp_path / "t.index"
    idx.write_text(f"a.npy\nfile://{tmp_path}/a.npy\nmem://bucket/b.npy\npipe:cat {tmp_path}/a.npy\ngs://nope/c.tar\nhttps://127.0.0.1:9/d.tar\nftp://x/e.tar\n")
    shards = P.read_index(str(idx))
    assert shards[2] == "mem://bucket/b.npy" and shards[4].startswith("gs://") and shards[0] == str(tmp_path / "a.npy")
    ds = P.StreamingTokenDataset(str(idx), max_context=16, shuffle_buffer=1, epochs=1)
    rows = list(ds)
    assert len(rows) == 16                                   # 4 readable shards x 4 rows; the 3 unreachable ones are skipped
    assert sum(int(r[0] >= 1000) for r in rows) == 4         # the mem:// shard was read through the custom opener
    err = capsys.readouterr().out
    assert "gs://nope/c.tar" in err and "ftp" in err


def test_skip_batches_fast_forwards_the_stream():
    from zero_transformer_b200.data.pipeline import SyntheticTokenDataset, batch_iterator, skip_batches

    def stream():
        return batch_iterator(SyntheticTokenDataset(100, 8, rank=0, seed=1, length=40), 4)

    full = list(stream())
    assert len(full) == 10
    rest = list(skip_batches(stream(), 3))
    assert len(rest) == 7 and all(np.array_equal(a, b) for a, b in zip(rest, full[3:]))
    assert list(skip_batches(stream(), 50)) == []


def test_prefetcher_close_releases_the_producer_thread():
    import time

    from zero_transformer_b200.data.pipeline import DevicePrefetcher, SyntheticTokenDataset, batch_iterator

    closed = []

    def gen():
        try:
            yield from batch_iterator(SyntheticTokenDataset(100, 8, rank=0, seed=1), 2)   # infinite
        finally:
            closed.append(True)

    pf = DevicePrefetcher(gen(), "cpu", depth=2)
    assert next(pf).shape == (2, 8)
    time.sleep(0.3)            # the producer is now parked on a full queue
    pf.close()
    assert not pf.t.is_alive() and closed == [True]
