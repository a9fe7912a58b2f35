"""Node health probe -- the analogue of the reference's ``src/utils/pod_test.py:17-34``
(``pmap(psum)`` over all devices, printed on host 0).

Every rank contributes ``rank+1``; the all-reduced value must be ``n(n+1)/2`` on
every rank.  On GPUs it additionally checks the NVLink peer path used by the fused
ZeRO-1 kernels (symmetric-memory rendezvous, a peer write/read round trip and a
bandwidth estimate) so a bad link shows up before a long run does.

    torchrun --nproc-per-node 8 -m zero_transformer_b200.utils.selftest
"""
from __future__ import annotations

import json
import time

import torch
import torch.distributed as dist

from ..parallel import dist as pdist


def run(verbose: bool = True) -> dict:
    env = pdist.setup_dp()
    out = {"world": env.world, "rank": env.rank, "device": str(env.device)}
    x = torch.full((1024,), float(env.rank + 1), device=env.device)
    if env.world > 1:
        dist.all_reduce(x)
    expect = env.world * (env.world + 1) / 2
    out["allreduce_ok"] = bool(torch.all(x == expect).item())
    if env.device.type == "cuda":
        out["gpu"] = torch.cuda.get_device_name(env.device)
        if env.world > 1:
            try:
                from ..parallel.fused import p2p_probe

                out.update(p2p_probe(env))
            except Exception as e:  # report, do not crash the probe
                out["p2p_error"] = repr(e)
        n = 1 << 28
        a = torch.empty(n, dtype=torch.uint8, device=env.device)
        b = torch.empty_like(a)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            b.copy_(a)
        torch.cuda.synchronize()
        out["hbm_copy_gbs"] = round(10 * 2 * n / (time.perf_counter() - t0) / 1e9, 1)
    oks = [None] * env.world
    if env.world > 1:
        dist.all_gather_object(oks, out)
    else:
        oks = [out]
    if env.is_main and verbose:
        for o in oks:
            print(json.dumps(o))
        print("SELFTEST", "PASS" if all(o["allreduce_ok"] and "p2p_error" not in o for o in oks) else "FAIL")
    return out


if __name__ == "__main__":
    run()
    pdist.shutdown()
