"""Token data pipeline.

Capabilities of the reference pipeline (``main_zero.py:368-421,477-493``):
index-file driven shard list, deterministic per-process shard striding, a
seeded shuffle buffer (seed ``23 + resume_step``), fixed ``max_context``-token
samples truncated/cast to int32, fixed-size batches with ``drop_last``, and
re-chunking of ``max_context`` rows into ``train_context`` rows -- plus what the
benchmark protocol needs: a synthetic random-token source and a pinned-memory
prefetcher that overlaps the host->device copy with the previous step.

Shard locations: plain paths and ``file://`` are opened directly; any other ``scheme://`` URL goes through a
pluggable *opener* (``register_opener``).  Built in: ``http(s)://`` (urllib), ``gs://`` (``gcloud storage cat`` /
``gsutil cat`` subprocess pipe, the reference's GCS buckets, ``main_zero.py:368-421``) and ``pipe:<command>``
(WebDataset's convention).  Index files may therefore be copied over unchanged from the reference; a scheme whose
transport is unavailable (no network, no gcloud) fails when the shard is *opened*, with the URL in the message, and the
loader skips that shard like any other unreadable one.

Shard formats:
  ``*.npy``  2-D integer array ``(n_samples, >=max_context)``
  ``*.bin``  raw little-endian uint16 token stream, cut into ``max_context`` rows
  ``*.tar``  members ending in ``.npy`` or ``.pth`` (e.g. ``<key>.input_id.pth``,
             the reference's member name), one sample each
"""
from __future__ import annotations

import io
import os
import queue
import random
import tarfile
import threading
from typing import BinaryIO, Callable, Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch


# scheme -> callable(url) -> binary file object.  Looked up when a shard is opened, never when the index is read.
_OPENERS: Dict[str, Callable[[str], BinaryIO]] = {}


def register_opener(scheme: str, fn: Callable[[str], BinaryIO]) -> None:
    """Plug in a transport for ``<scheme>://...`` shard URLs (``fn(url)`` returns a readable binary stream)."""
    _OPENERS[scheme.lower()] = fn


def _open_http(url: str) -> BinaryIO:
    import urllib.request

    return urllib.request.urlopen(url, timeout=60)  # noqa: S310 - user-supplied shard list


def _open_pipe(cmd: str) -> BinaryIO:
    import subprocess

    proc = subprocess.Popen(cmd, shell=True, stdout=subprocess.PIPE, bufsize=1 << 20)  # noqa: S602 - WebDataset "pipe:" convention
    return proc.stdout


def _open_gs(url: str) -> BinaryIO:
    import shutil

    for tool in (["gcloud", "storage", "cat"], ["gsutil", "cat"]):
        if shutil.which(tool[0]):
            return _open_pipe(" ".join(tool) + " '" + url.replace("'", "") + "'")
    raise OSError(f"cannot open {url}: neither gcloud nor gsutil is installed (register_opener('gs', ...) to plug in another client)")


register_opener("http", _open_http)
register_opener("https", _open_http)
register_opener("gs", _open_gs)


def _scheme(loc: str) -> str:
    if loc.startswith("pipe:"):
        return "pipe"
    return loc.split("://", 1)[0].lower() if "://" in loc else ""


def open_shard(loc: str) -> BinaryIO:
    """Binary stream of a shard location (local path, ``file://``, ``pipe:cmd`` or a registered ``scheme://``)."""
    sch = _scheme(loc)
    if sch == "":
        return open(loc, "rb")
    if sch == "file":
        return open(loc[len("file://"):], "rb")
    if sch == "pipe":
        return _open_pipe(loc[len("pipe:"):])
    if sch not in _OPENERS:
        raise OSError(f"no opener registered for '{sch}://' shard URLs ({loc})")
    return _OPENERS[sch](loc)


def read_index(path: str) -> List[str]:
    """One shard path/URL per line; blank lines and ``#`` comments ignored.
    Relative paths resolve against the index file's directory; URLs are kept verbatim."""
    base = os.path.dirname(os.path.abspath(path))
    out = []
    with open(path, "r") as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            if _scheme(s) == "file":
                s = s[len("file://"):]
            if _scheme(s):
                out.append(s)
            else:
                out.append(s if os.path.isabs(s) else os.path.join(base, s))
    return out


def split_by_rank(shards: Sequence[str], rank: int, world: int) -> List[str]:
    """Deterministic striding (reference ``split_by_jax_process``, ``main_zero.py:377-387``)."""
    if len(shards) < world:
        # fewer shards than ranks: every rank reads every shard but strides samples instead
        return list(shards)
    return list(shards[rank::world])


def _suffix(loc: str) -> str:
    return os.path.splitext(loc.split("?", 1)[0].rstrip("'\" "))[1].lower()


def _iter_shard(path: str, max_context: int) -> Iterator[np.ndarray]:
    remote = _scheme(path) != ""
    try:
        if remote and _suffix(path) in (".npy", ".bin"):
            # array shards need random access: spool the stream to memory first
            raw = open_shard(path).read()
            if _suffix(path) == ".npy":
                arr = np.load(io.BytesIO(raw))
            else:
                arr = np.frombuffer(raw, dtype=np.uint16)
            if arr.ndim == 1:
                arr = arr[: (arr.shape[0] // max_context) * max_context].reshape(-1, max_context)
            for row in arr:
                if row.shape[0] >= max_context:
                    yield np.asarray(row[:max_context], dtype=np.int32)
        elif path.endswith(".npy"):
            arr = np.load(path, mmap_mode="r")
            if arr.ndim == 1:
                arr = arr[: (arr.shape[0] // max_context) * max_context].reshape(-1, max_context)
            for row in arr:
                if row.shape[0] >= max_context:
                    yield np.asarray(row[:max_context], dtype=np.int32)
        elif path.endswith(".bin"):
            arr = np.memmap(path, dtype=np.uint16, mode="r")
            n = arr.shape[0] // max_context
            for i in range(n):
                yield np.asarray(arr[i * max_context:(i + 1) * max_context], dtype=np.int32)
        elif _suffix(path) == ".tar":
            # "r|" = sequential stream: works for local files and for sockets / subprocess pipes alike
            with tarfile.open(fileobj=open_shard(path), mode="r|") as tf:
                for m in tf:
                    if not m.isfile():
                        continue
                    try:
                        raw = tf.extractfile(m).read()
                        if m.name.endswith(".npy"):
                            row = np.load(io.BytesIO(raw))
                        elif m.name.endswith(".pth"):
                            row = torch.load(io.BytesIO(raw), map_location="cpu", weights_only=False)
                            row = row.numpy() if isinstance(row, torch.Tensor) else np.asarray(row)
                        else:
                            continue
                        row = np.asarray(row).reshape(-1)
                        if row.shape[0] >= max_context:
                            yield row[:max_context].astype(np.int32)
                    except Exception as e:  # corrupt sample: warn and continue (reference: wds.warn_and_continue)
                        print(f"[data] skipping corrupt sample {m.name} in {path}: {e}")
        else:
            raise ValueError(f"unknown shard format: {path}")
    except (OSError, ValueError, tarfile.TarError) as e:
        print(f"[data] skipping unreadable shard {path}: {e}")


class StreamingTokenDataset:
    """Infinite-epoch capable iterable of ``(max_context,) int32`` samples."""

    def __init__(self, index_path: str, *, max_context: int, rank: int = 0, world: int = 1,
                 shuffle_buffer: int = 100000, seed: int = 23, epochs: Optional[int] = 1):
        self.all_shards = read_index(index_path)
        if not self.all_shards:
            raise ValueError(f"{index_path}: no shards listed")
        self.shards = split_by_rank(self.all_shards, rank, world)
        self.stride_samples = len(self.all_shards) < world
        self.rank, self.world = rank, world
        self.max_context = max_context
        self.shuffle_buffer = max(int(shuffle_buffer), 1)
        self.seed = seed
        self.epochs = epochs

    def _raw(self, rng: random.Random) -> Iterator[np.ndarray]:
        ep = 0
        while self.epochs is None or ep < self.epochs:
            shards = list(self.shards)
            rng.shuffle(shards)
            k = 0
            for sh in shards:
                for row in _iter_shard(sh, self.max_context):
                    if self.stride_samples and (k % self.world) != self.rank:
                        k += 1
                        continue
                    k += 1
                    yield row
            ep += 1

    def __iter__(self) -> Iterator[np.ndarray]:
        rng = random.Random(self.seed)
        buf: List[np.ndarray] = []
        for row in self._raw(rng):
            if len(buf) < self.shuffle_buffer:
                buf.append(row)
                continue
            i = rng.randrange(len(buf))
            out, buf[i] = buf[i], row
            yield out
        rng.shuffle(buf)
        yield from buf


class SyntheticTokenDataset:
    """Uniform random tokens of the named shape -- the benchmark's data source
    (``"data": "synthetic"``).  Deterministic in ``(seed, rank, sample index)``."""

    def __init__(self, vocab_size: int, max_context: int, *, rank: int = 0, seed: int = 0, length: Optional[int] = None):
        self.vocab_size, self.max_context = vocab_size, max_context
        self.rank, self.seed, self.length = rank, seed, length

    def __iter__(self) -> Iterator[np.ndarray]:
        g = np.random.default_rng([self.seed, self.rank])
        i = 0
        while self.length is None or i < self.length:
            yield g.integers(0, self.vocab_size, size=(self.max_context,), dtype=np.int32)
            i += 1


def numpy_collate(rows: Sequence[np.ndarray]) -> np.ndarray:
    """Stack samples into ``(B, max_context)`` (reference ``src/utils/dataloader.py:9-16``)."""
    return np.stack(rows, axis=0)


def batch_iterator(samples: Iterable[np.ndarray], batch_rows: int, drop_last: bool = True) -> Iterator[np.ndarray]:
    buf: List[np.ndarray] = []
    for s in samples:
        buf.append(s)
        if len(buf) == batch_rows:
            yield numpy_collate(buf)
            buf = []
    if buf and not drop_last:
        yield numpy_collate(buf)


def skip_batches(batches: Iterable[np.ndarray], n: int) -> Iterator[np.ndarray]:
    """Drop the first ``n`` batches: the resume fast-forward of the reference loop (``if i < iterator_resume_step: continue``,
    ``main_zero.py:437,470-471``), done on the host before anything is copied to the device."""
    it = iter(batches)
    for _ in range(max(int(n), 0)):
        if next(it, None) is None:
            return
    yield from it


def rechunk(batch: torch.Tensor, train_context: int) -> torch.Tensor:
    """``(B, max_context)`` -> ``(B * max_context/train_context, train_context)``
    (reference ``main_zero.py:425-428,477-478``)."""
    B, L = batch.shape
    assert L % train_context == 0, "max_context must be a multiple of train_context"
    return batch.reshape(B * (L // train_context), train_context)


def to_micro_batches(batch: torch.Tensor, accum_steps: int) -> torch.Tensor:
    """``(R, T)`` -> ``(accum, R/accum, T)`` (reference layout prep ``main_zero.py:480-493``)."""
    R, T = batch.shape
    assert R % accum_steps == 0, f"{R} rows per rank do not divide into {accum_steps} accumulation steps"
    return batch.view(accum_steps, R // accum_steps, T)


class DevicePrefetcher:
    """Background thread: numpy batch -> pinned host tensor -> async H2D copy on
    a side stream, ``depth`` batches ahead.  ``next()`` makes the consumer stream
    wait on the copy event only (no host sync)."""

    def __init__(self, batches: Iterable[np.ndarray], device: torch.device, depth: int = 2, dtype=torch.int32):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.dtype = dtype
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._stop = False
        self.h2d_bytes = 0
        self.t = threading.Thread(target=self._run, args=(iter(batches),), daemon=True)
        self.t.start()

    def _put(self, item) -> bool:
        """Blocking put that gives up when the consumer closed the prefetcher (so the thread, the dataset generator and
        its shuffle buffer are released instead of staying parked in ``q.put`` forever)."""
        while not self._stop:
            try:
                self.q.put(item, timeout=0.2)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, it):
        try:
            for b in it:
                if self._stop:
                    return
                host = torch.from_numpy(np.ascontiguousarray(b)).to(self.dtype)
                if self.cuda:
                    host = host.pin_memory()
                    with torch.cuda.stream(self.stream):
                        dev = host.to(self.device, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.stream)
                    if not self._put((dev, ev, host)):
                        return
                elif not self._put((host, None, host)):
                    return
        except Exception as e:  # surface loader errors in the consumer
            self._put(e)
            return
        finally:
            close = getattr(it, "close", None)   # generator: run its cleanup now, not at garbage collection
            if close is not None:
                try:
                    close()
                except Exception:
                    pass
        self._put(None)

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        item = self.q.get()
        if item is None:
            raise StopIteration
        if isinstance(item, Exception):
            raise item
        dev, ev, host = item
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            dev.record_stream(torch.cuda.current_stream(self.device))
        self.h2d_bytes = host.numel() * host.element_size()
        return dev

    def close(self):
        self._stop = True
        try:
            while True:
                self.q.get_nowait()
        except queue.Empty:
            pass
        if self.t.is_alive() and threading.current_thread() is not self.t:
            self.t.join(timeout=5.0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
