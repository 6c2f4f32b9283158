"""Several C2 steps in flight: does the next batch's transform fill the SMs that the previous batch's HBM-bound tail
(CTC stragglers) leaves idle?  One HotPathStep (own streams, own workspaces) and one CUDA graph per resident batch;
the graphs are replayed round-robin on 1, 2 or 3 lane streams.  Every variant is checked bit for bit against the
serial run (features, gradient, loss) before it is timed.

  python tools/time_overlap.py [replays]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multiprocessing as mp  # noqa: E402
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 240
POOL = 6


def _mk(i):
    return bench.make_batch(2000 + i, 256, "c2")


with mp.get_context("fork").Pool(POOL) as _p:        # (forked before CUDA is initialised)
    host_batches = _p.map(_mk, range(POOL))
import torch  # noqa: E402
from asr_dfcnn_transformer_b200 import pipeline  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
V = bench.V
pool = [bench.DeviceBatch(hb, dev, torch, "c2", "logits") for hb in host_batches]
audio = sum(d.audio_s for d in pool) / POOL


def args_of(db):
    return (db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels, db.label_len, db.input_len, V - 1)


def kw_of(db):
    return dict(feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale, ctc_bounds=db.ctc_bounds)


# serial reference results (eager, one stream)
hp0 = pipeline.HotPathStep(dev)
refs = []
for db in pool:
    f, r = hp0(*args_of(db), **kw_of(db))
    torch.cuda.synchronize()
    refs.append((f.clone(), r.grad.clone(), r.loss.clone()))


def run(lanes, feature_ctas=0, merged=False, prio=0, label=""):
    hps, graphs, accs, res = [], [], [], []
    for db in pool:
        hp = pipeline.HotPathStep(dev, feature_ctas=feature_ctas, merged_tail=merged, feature_priority=prio)
        hp.reserve(db.B, db.total_frames, db.logits.shape[0], db.labels.shape[1])
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        g, f, r = hp.capture(*args_of(db), loss_acc=acc, **kw_of(db))
        hps.append(hp), graphs.append(g), accs.append(acc), res.append(r)
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    main = torch.cuda.current_stream(dev)
    ev = torch.cuda.Event()

    def go(n):
        ev.record(main)
        for s in streams:
            s.wait_event(ev)
        for i in range(n):
            with torch.cuda.stream(streams[i % lanes]):
                graphs[i % POOL].replay()
        for s in streams:
            main.wait_stream(s)

    # correctness under overlap: poison the outputs, run two rounds, compare with the serial results
    for db in pool:
        db.feat.zero_()
        db.grad.zero_()
    go(2 * POOL)
    torch.cuda.synchronize()
    ok = all(torch.equal(pool[i].feat[:pool[i].total_frames], refs[i][0]) and torch.equal(res[i].grad, refs[i][1]) and
             torch.equal(res[i].loss, refs[i][2]) for i in range(POOL))
    go(POOL)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        go(N)
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / N)
    print("lanes %d  feature_ctas %3d  merged %-5s  prio %2d  %s: %.1f us per step = %.3f M audio-s/s  bit-identical %s"
          % (lanes, feature_ctas or 148, merged, prio, label, best, audio / best, ok), flush=True)
    del graphs
    torch.cuda.synchronize()


if os.environ.get("ASRK_OVERLAP_SWEEP", "1") == "1":
    # first sweep (gpurun_out/r2_overlap39.log)
    run(1)
    for lanes in (2, 3):
        for prio in (0, -1):
            for n in (0, 140, 132, 124, 116, 108, 100):
                run(lanes, feature_ctas=n, prio=prio)
    for n in (0, 132, 116):
        for prio in (0, -1):
            run(2, feature_ctas=n, merged=True, prio=prio)
    run(1, feature_ctas=132)
    run(1)
else:
    # second sweep: narrower transforms, more lanes
    run(1)
    for lanes in (2, 3, 6):
        for n in (112, 108, 104, 100, 96, 92, 88, 80, 72):
            run(lanes, feature_ctas=n)
    for n in (108, 100, 92, 84):
        run(2, feature_ctas=n, merged=True)
    run(1)
