#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the hot path (spectrogram features + CTC
loss + gradient) on N B200s, with the HBM roofline of the dominant kernel and the
CPU baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): per GPU an AISHELL-shaped synthetic
batch of 256 utterances of U(3,7) s 16 kHz int16 audio (generator G2), CTC logits
[Tmax, 256, 1424] fp32 with T_ctc = min(200, n_frames//8+1) and labels ~ U{8..24}.
One step = features (frames -> Hamming -> 400-pt FFT -> log(|X|+1) -> z-score) of
the whole batch + CTC loss and gradient w.r.t. the logits of the whole batch.
Utterances are sharded by batch across ranks (weak scaling: every rank owns its own
256 utterances); the only collective is the all-reduce of [sum loss, n].
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402  (pure data generation, no reference arithmetic)

FS = 16000
V = synth.VOCAB_DICT_TXT
BATCH = 256
POOL = 3          # distinct device-resident batches rotated through the timed loop
METRIC = "audio-sec/sec (fbank+CTC loss+grad)"
WORKLOAD = ("C2: AISHELL-shaped 256 utt/GPU x U(3,7) s int16 16 kHz (G2), fbank z-scored + CTC loss/grad, V=1424, "
            "T_ctc=min(200,n_frames//8+1), L~U{8..24}")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from
    the committed `ncu --set full` capture of this workload (profiles/traffic.json)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]["dram_bytes"])
    except Exception:
        return None


def ncu_fp64_pct(kernel):
    """fp64 pipe utilisation of that kernel in the same capture (the spectrogram kernel is bound by the
    fp64 pipe and its feeding, not by HBM: the HBM fraction alone would misread it)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]["fp64_pipe_pct"])
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------
def make_batch(seed, batch=BATCH):
    """One C2 batch as host arrays (numpy)."""
    cache = os.environ.get("ASRK_BENCH_CACHE")
    cpath = os.path.join(cache, "c2_%d_%d.npz" % (seed, batch)) if cache else None
    if cpath and os.path.isfile(cpath):
        d = np.load(cpath)
        lens = d["lens"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        pcm = [d["pcm"][offs[i]:offs[i + 1]] for i in range(len(lens))]
        return dict(pcm=pcm, lens=lens, nfr=d["nfr"], logits=d["logits"], labels=d["labels"],
                    label_len=d["label_len"], input_len=d["input_len"])
    rng = np.random.default_rng(seed)
    lens = synth.ragged_lengths(rng, batch, 3.0, 7.0)
    # G2 synthesis of 256 x 5 s costs seconds per batch on the host; the bench
    # needs the *shape* of the signal (harmonics + AM + floor), so build it
    # vectorised here with the same recipe.
    pcm = []
    for n in lens:
        pcm.append(synth.g2_voiced(rng, int(n)))
    nfr = np.array([synth.n_frames(int(n)) for n in lens], dtype=np.int64)
    il = np.array([synth.t_ctc(int(f)) for f in nfr], dtype=np.int32)
    x, labels, ll, il = synth.ctc_batch(rng, il, V, 8, 24, lmax=64, scale=3.0)
    if cpath:
        os.makedirs(cache, exist_ok=True)
        np.savez(cpath, pcm=np.concatenate(pcm), lens=lens, nfr=nfr, logits=x, labels=labels, label_len=ll,
                 input_len=il)
    return dict(pcm=pcm, lens=lens, nfr=nfr, logits=x, labels=labels, label_len=ll, input_len=il)


def algorithmic_bytes(b):
    """features 2N + 800 n_frames per utterance; CTC 8 T_ctc V."""
    feat = int(2 * b["lens"].sum() + 800 * b["nfr"].sum())
    ctc = int(8 * V * b["input_len"].astype(np.int64).sum())
    return feat, ctc


class DeviceBatch:
    """A batch resident in HBM plus its pinned host copy (for the e2e leg)."""

    def __init__(self, hb, dev, torch):
        from asr_dfcnn_transformer_b200 import features
        self.hb = hb
        pk = features.pack_host(hb["pcm"], FS, "fbank")
        self.pk = pk
        self.B = len(hb["pcm"])
        self.h_samples = pk.samples                       # pinned int16
        self.h_logits = torch.from_numpy(hb["logits"]).pin_memory()
        self.h_labels = torch.from_numpy(hb["labels"]).pin_memory()
        self.samples = pk.samples.to(dev)
        self.so = torch.from_numpy(pk.sample_offsets).to(dev)
        self.sc = torch.from_numpy(pk.sample_counts).to(dev)
        self.fo = torch.from_numpy(pk.frame_offsets).to(dev)
        self.total_frames = pk.total_frames
        self.logits = self.h_logits.to(dev)
        self.labels = self.h_labels.to(dev)
        self.label_len = torch.from_numpy(hb["label_len"]).to(dev)
        self.input_len = torch.from_numpy(hb["input_len"]).to(dev)
        self.feat = torch.empty((pk.total_frames, 200), dtype=torch.float32, device=dev)
        self.grad = torch.empty_like(self.logits)
        self.grad_scale = torch.full((self.B,), 1.0 / self.B, dtype=torch.float32, device=dev)
        self.audio_s = float(hb["lens"].sum()) / FS
        # the loader's host-side length vectors bound the batch: every lattice fits the fused CTC kernel
        self.ctc_bounds = (int(hb["input_len"].max()), int(hb["label_len"].max()))
        self.bytes_feat, self.bytes_ctc = algorithmic_bytes(hb)


_STEP = {}


def hot_path(dev):
    from asr_dfcnn_transformer_b200 import pipeline
    if dev not in _STEP:
        _STEP[dev] = pipeline.HotPathStep(dev, feature_ctas=int(os.environ.get("ASRK_FEATURE_CTAS", "0")))
    return _STEP[dev]


def run_step(db, phases_timer=None):
    """The hot path on device-resident inputs.  Returns the CtcResult."""
    from asr_dfcnn_transformer_b200 import _lib, ctc, features
    if phases_timer is None:
        _, r = hot_path(db.samples.device)(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames,
                                           db.logits, db.labels, db.label_len, db.input_len, V - 1,
                                           feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                                           ctc_bounds=db.ctc_bounds)
        return r
    # same launches, issued phase by phase with CUDA events between them
    ev = phases_timer
    ev.mark()
    for ph in (_lib.PHASE_SPEC_SETUP, _lib.PHASE_SPEC_MAIN, _lib.PHASE_SPEC_NORMALIZE):
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank",
                                    out=db.feat, phases=ph)
        ev.mark()
    r = None
    for ph in (_lib.PHASE_CTC_PREP, _lib.PHASE_CTC_FUSED, _lib.PHASE_CTC_ROWS, _lib.PHASE_CTC_LATTICE,
               _lib.PHASE_CTC_GRAD):
        r = ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1,
                              grad_scale=db.grad_scale, grad_out=db.grad, phases=ph,
                              outputs=None if r is None else (r.loss, db.grad, r.row_status, None, None, None))
        ev.mark()
    return r


KERNELS = ["spec_setup", "spec_main", "spec_normalize", "ctc_prep", "ctc_fused", "ctc_rows", "ctc_lattice",
           "ctc_grad"]


class PhaseTimer:
    def __init__(self, torch):
        self.torch = torch
        self.steps = []
        self.cur = None

    def begin(self):
        self.cur = []

    def mark(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.cur.append(e)

    def end(self):
        self.steps.append(self.cur)

    def summary(self):
        acc = np.zeros(len(KERNELS))
        for evs in self.steps:
            for k in range(len(KERNELS)):
                acc[k] += evs[k].elapsed_time(evs[k + 1])
        return {n: float(a / max(len(self.steps), 1)) for n, a in zip(KERNELS, acc)}


# ----------------------------------------------------------------------------
# CPU baseline (oracle port / reference arm)
# ----------------------------------------------------------------------------
def _cpu_features_worker(sig):
    from oracle import fbank_ref
    return fbank_ref.compute_fbank(sig).shape[0]


_POOL = {}


def _pool(cores):
    """One worker pool for the whole run (forked before CUDA is initialised)."""
    import multiprocessing as mp
    if cores not in _POOL:
        _POOL[cores] = mp.get_context("fork").Pool(cores)
    return _POOL[cores]


def cpu_sample(hb, n_utt, cores):
    """Time the oracle (port of the reference's CPU path) on the first n_utt
    utterances of the workload: per-frame scipy FFT features fanned out over the
    host cores + the float32 C restatement of TF's CTC loss/grad (OpenMP)."""
    import multiprocessing as mp
    from oracle import build_c
    build_c.lib()
    sigs = hb["pcm"][:n_utt]
    audio_s = sum(len(s) for s in sigs) / FS
    t0 = time.perf_counter()
    if cores > 1:
        _pool(cores).map(_cpu_features_worker, sigs, chunksize=1)
    else:
        for s in sigs:
            _cpu_features_worker(s)
    t_feat = time.perf_counter() - t0
    il = hb["input_len"][:n_utt]
    T = int(il.max())
    x = np.ascontiguousarray(hb["logits"][:T, :n_utt])
    t0 = time.perf_counter()
    build_c.ctc_loss_grad(x, hb["labels"][:n_utt], hb["label_len"][:n_utt], il, V - 1, real="f32",
                          threads=cores)
    t_ctc = time.perf_counter() - t0
    return audio_s, t_feat, t_ctc


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the oracle
    port: /root/reference is Python+TensorFlow and does not exist on the GPU box)
    on the host cores, same workload/metric, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    hb = make_batch(2000)                      # one full C2 batch per step (~1 s of wall time on 16 cores)
    n_utt = len(hb["pcm"])
    for _ in range(args.warmup):
        cpu_sample(hb, min(n_utt, 2 * cores), cores)
    tot_audio, tot_t = 0.0, 0.0
    for _ in range(args.steps):
        a, tf_, tc_ = cpu_sample(hb, n_utt, cores)
        tot_audio += a
        tot_t += tf_ + tc_
    val = tot_audio / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "audio-sec/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_gpu": BATCH,
                   "sample": "%d utterances of the workload per step" % n_utt, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": "audio-sec/sec", "cores": cores, "kind": "port",
                         "sample": "%d C2 utterances per step: oracle/fbank_ref.py (per-frame scipy FFT, "
                                   "multiprocessing) + oracle/ctc_ref.c float32 (OpenMP)" % n_utt},
        "e2e": {"value": val, "unit": "audio-sec/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# main
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started by hand with --gpus N: become the launch the driver uses (one rank per GPU, NCCL, loopback)
        import socket
        with socket.socket() as so:
            so.bind(("127.0.0.1", 0))
            port = so.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                  str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", str(port),
                                  os.path.abspath(__file__)] + sys.argv[1:])

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    # ---- (0) CPU baseline on rank 0 (N = 1 only), BEFORE CUDA is initialised so that the
    # worker processes can be forked safely ---------------------------------------------
    cpu = None
    hb0 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        hb0 = make_batch(2000)
        n_utt = len(hb0["pcm"])                             # the whole C2 batch, twice: ~10 s of CPU work
        cpu_sample(hb0, min(n_utt, cores), cores)          # warm the pool / page cache
        a = tf_ = tc_ = 0.0
        for _ in range(2):
            a1, t1, t2 = cpu_sample(hb0, n_utt, cores)
            a, tf_, tc_ = a + a1, tf_ + t1, tc_ + t2
        cpu = {"value": a / (tf_ + tc_), "unit": "audio-sec/sec", "cores": cores, "kind": "port",
               "sample": "2 passes over the %d utterances of one C2 batch (%.0f audio-s): oracle/fbank_ref.py "
                         "features (per-frame scipy FFT, %d processes, %.2f s) + oracle/ctc_ref.c float32 CTC "
                         "loss/grad (OpenMP, %.2f s)" % (n_utt, a, cores, tf_, tc_)}

    import torch
    import torch.distributed as dist
    from asr_dfcnn_transformer_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: POOL distinct batches per rank, resident in HBM --------------
    pool = [DeviceBatch(hb0 if (i == 0 and rank == 0 and hb0 is not None) else make_batch(2000 + 100 * rank + i),
                        dev, torch) for i in range(POOL)]
    audio_per_step = float(np.mean([d.audio_s for d in pool]))
    red = torch.zeros(2, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)

    def reduce_loss(r, db):
        # the path's only collective: all-reduce of [sum loss, n] (one tiny kernel fills
        # the operand), on a side stream so that it overlaps the next step
        from asr_dfcnn_transformer_b200 import ctc, pipeline
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)      # the previous step's all-reduce is done with `red`
        ctc.loss_sum(r.loss, r.row_status, out=red)
        if world > 1:
            pipeline.all_reduce_loss(red, stream=side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident throughput ------------------------------------
    for i in range(args.warmup):
        r = run_step(pool[i % POOL])
        reduce_loss(r, pool[i % POOL])
    barrier()
    loss_check = red.clone()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    audio = 0.0
    for i in range(args.steps):
        db = pool[i % POOL]
        r = run_step(db)
        reduce_loss(r, db)
        audio += db.audio_s
    if world > 1:
        torch.cuda.current_stream().wait_stream(side)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, audio], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm[0:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tm[1:2], op=dist.ReduceOp.SUM)
        ms, audio_all = float(tm[0]), float(tm[1])
    else:
        audio_all = audio
    value = audio_all / (ms * 1e-3)

    # ---- (2) per-kernel durations, live, same launches ----------------------
    pt = PhaseTimer(torch)
    for i in range(args.steps):
        pt.begin()
        run_step(pool[i % POOL], pt)
        pt.end()
    torch.cuda.synchronize()
    kms = pt.summary()
    # a noise-free batch launches no kernel in the setup phase (one 16-byte memset): what the events
    # bracket there is the host's enqueue latency, not device work
    kms.pop("spec_setup", None)
    clocks = sampler.stop() if rank == 0 else None

    # ---- (3) end to end through the public API with host buffers ------------
    from asr_dfcnn_transformer_b200 import ctc, features
    e2e_steps = max(3, min(args.steps, 10))
    h2d = d2h = 0

    def e2e_step(db):
        # the public host-buffer entry point: PCM + labels host -> device, logits staged without
        # their padding (only rows t < input_len cross PCIe), kernels, loss back to the host
        nonlocal h2d, d2h
        _, r, nbytes = hot_path(dev).from_host(db.h_samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.h_logits,
                                               db.h_labels, db.label_len, db.input_len, V - 1, logits_dev=db.logits,
                                               feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                                               ctc_bounds=db.ctc_bounds)
        loss_host = r.loss.cpu()          # device -> host read of the step's result (synchronises)
        h2d = nbytes + 4 * V * int(db.hb["input_len"].astype(np.int64).sum())
        d2h = loss_host.numel() * 4
        return loss_host

    for i in range(2):
        e2e_step(pool[i % POOL])
    barrier()
    t0 = time.perf_counter()
    e0.record()
    a2 = 0.0
    for i in range(e2e_steps):
        e2e_step(pool[i % POOL])
        a2 += pool[i % POOL].audio_s
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    t = torch.tensor([ms2, a2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
    e2e_value = float(t[1]) / (float(t[0]) * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(("spec_main", "ctc_fused", "ctc_rows", "ctc_grad"), key=lambda k: kms[k])
        bf = float(np.mean([d.bytes_feat for d in pool]))
        bc = float(np.mean([d.bytes_ctc for d in pool]))
        alg = {"spec_main": bf, "ctc_fused": bc, "ctc_rows": bc / 2, "ctc_grad": bc / 2}[dom]
        ach = alg / (kms[dom] * 1e-3) / 1e9
        step_alg = bf + bc
        line = {
            "metric": METRIC, "value": value, "unit": "audio-sec/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 FFT / f32 log, CTC",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "utterances_per_gpu": BATCH, "audio_s_per_step_per_gpu": audio_per_step,
                       "l2": "inputs larger than L2: %d distinct batches rotated, ~%.0f MB touched per step"
                             % (POOL, (step_alg + bc / 2) / 1e6)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic(dom), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "fp64_pipe_pct_ncu": ncu_fp64_pct(dom),
                         "step_frac": (step_alg / (ms / args.steps * 1e-3) / 1e9) / peak / world},
            "kernel_ms": kms,
            "loss_mean": float(loss_check[0] / max(float(loss_check[1]), 1.0)),
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "audio-sec/sec", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            # per step: spectrogram, stats, z-score, fused CTC (prepares its own utterance; the batch is
            # bounded by the host's length vectors, so no generic CTC kernels are launched), loss sum.
            # kernel_ms times the CTC phases one by one, i.e. with the separate prep kernel and the
            # three generic kernels that find nothing to do.
            "gpu_launches": 5 * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
