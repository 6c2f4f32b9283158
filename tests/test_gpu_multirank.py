"""GPU, world size 2, NCCL (needs two devices; skipped otherwise): the batch-sharded hot path gives exactly
the single-GPU answer -- the all-reduced [sum loss, n] equals the single-GPU sums, the per-rank greedy-decode
tokens concatenated are bit-identical to the single-GPU tokens, and so are the features and gradients
(SURVEY.md section 4 "Multi-GPU" tier; utterances are independent, no tensor crosses GPUs)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _batch():
    from oracle import synth
    rng = np.random.default_rng(5005)
    lens = synth.ragged_lengths(rng, 64, 3.0, 7.0)
    pcm = [synth.g2_voiced(rng, int(n)) for n in lens]
    il = np.array([synth.t_ctc(synth.n_frames(int(n))) for n in lens], dtype=np.int32)
    x, labels, ll, il = synth.ctc_batch(rng, il, synth.VOCAB_DICT_TXT, 8, 24, lmax=64)
    return pcm, x, labels, ll, il


def _run_shard(pcm, x, labels, ll, il, dev):
    from asr_dfcnn_transformer_b200 import ctc, features
    V = x.shape[2]
    fb = features.compute_features(pcm, mode="fbank", device=dev)
    T = int(il.max())
    r = ctc.ctc_loss_grad(torch.as_tensor(np.ascontiguousarray(x[:T])).to(dev), labels, ll, il, V - 1, decode=True)
    red = ctc.loss_sum(r.loss, r.row_status)
    return fb, r, red


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from asr_dfcnn_transformer_b200 import ctc, pipeline
    pcm, x, labels, ll, il = _batch()
    lo, hi = pipeline.shard_bounds(len(il), rank, world)
    fb, r, red = _run_shard(pcm[lo:hi], np.ascontiguousarray(x[:, lo:hi]), labels[lo:hi], ll[lo:hi], il[lo:hi], dev)
    pipeline.all_reduce_loss(red)
    torch.cuda.synchronize()
    q.put((rank, lo, hi, red.cpu().tolist(), ctc.tokens_to_lists(r.tokens, r.token_len), r.loss.cpu().numpy(),
           fb.features.cpu().numpy(), r.grad[:, :, ::97].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from asr_dfcnn_transformer_b200 import ctc
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(world)], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pcm, x, labels, ll, il = _batch()
    dev = torch.device("cuda", 0)
    fb, r, red = _run_shard(pcm, x, labels, ll, il, dev)
    red = red.cpu().tolist()
    tokens = ctc.tokens_to_lists(r.tokens, r.token_len)
    loss = r.loss.cpu().numpy()
    feats = fb.features.cpu().numpy()
    for rank, lo, hi, t, tok, ls, ft, gr in got:
        assert t[1] == len(il) == red[1]
        assert abs(t[0] - red[0]) <= 1e-12 * abs(red[0])              # float64 sums of the same float32 losses
        assert tok == tokens[lo:hi]                                    # bit-identical decode
        assert np.array_equal(ls, loss[lo:hi])
        assert np.array_equal(ft, feats[fb.frame_offsets[lo]:fb.frame_offsets[hi]])
        T = gr.shape[0]
        assert np.array_equal(gr, r.grad[:T, lo:hi, ::97].cpu().numpy())
    assert [g[1] for g in got] == [0, 32] and got[-1][2] == 64
