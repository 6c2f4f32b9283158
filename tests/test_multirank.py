"""CPU, world size 2, gloo: the host logic of the multi-GPU path -- batch sharding and
the all-reduce of [sum loss, n] give exactly the single-process answer.  The per-shard
losses come from the oracle here (no GPU in this container); on the GPU box the same
functions are driven by bench.py with NCCL."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from asr_dfcnn_transformer_b200 import pipeline
    from oracle import ctc_ref, synth
    rng = np.random.default_rng(7)                       # every rank builds the same global batch
    x, labels, ll, il = synth.ctc_batch(rng, [17, 9, 30, 12, 25, 21, 8], 40, 1, 6)
    lo, hi = pipeline.shard_bounds(len(il), rank, world)
    loss, _, _ = ctc_ref.ctc_loss_grad_batch(np.ascontiguousarray(x[:, lo:hi]), labels[lo:hi], ll[lo:hi],
                                             il[lo:hi], 39)
    t = torch.tensor([float(loss.sum()), float(hi - lo)], dtype=torch.float64)
    pipeline.all_reduce_loss(t)
    q.put((rank, lo, hi, t.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_equals_single_process():
    from oracle import ctc_ref, synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    x, labels, ll, il = synth.ctc_batch(rng, [17, 9, 30, 12, 25, 21, 8], 40, 1, 6)
    loss, _, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 39)
    shards = sorted((lo, hi) for _, lo, hi, _ in got)
    assert shards[0][0] == 0 and shards[-1][1] == len(il)
    assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))          # contiguous, disjoint, complete
    for _, _, _, t in got:
        assert t[1] == len(il)
        assert abs(t[0] - float(loss.sum())) <= 1e-9 * abs(float(loss.sum()))


def test_shard_bounds():
    from asr_dfcnn_transformer_b200 import pipeline
    for n in (0, 1, 7, 256, 1000003):
        for world in (1, 2, 3, 8):
            b = [pipeline.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pipeline.shard_bounds(4, 2, 2)


def test_reference_arm_under_the_multi_rank_launch():
    """`bench.py --impl reference --gpus 2` (CPU only): started by hand it relaunches itself under
    torch.distributed.run; rank 0 alone times the reference algorithm and prints the line, rank 1 exits 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--gpus", "2", "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # one line, from rank 0
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
