"""Health probe (pod_test analogue, reference src/utils/pod_test.py:17-34): 2-rank gloo run on CPU."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_selftest_two_ranks_gloo():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="", PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-m", "zero_transformer_b200.utils.selftest"], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = [l for l in outs[0].splitlines() if l.startswith("{")]
    recs = [json.loads(l) for l in lines]
    assert len(recs) == 2 and all(r["allreduce_ok"] and r["world"] == 2 for r in recs)
    assert "SELFTEST PASS" in outs[0]


def test_selftest_single_process():
    from zero_transformer_b200.utils import selftest

    out = selftest.run(verbose=False)
    assert out["allreduce_ok"] and out["world"] == 1
