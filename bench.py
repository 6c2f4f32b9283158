#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the hot path (spectrogram features + CTC
loss + gradient) on N B200s, with the HBM roofline of the dominant kernel and the
CPU baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4|c5] [--surface logits|keras]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json `configs`; C2 is the default and the headline -- the configuration the metric is
quoted on; C1 is the reference's own CPU-runnable case and a parity-test case only):
  c2  per GPU 256 utterances of U(3,7) s 16 kHz int16 audio (generator G2), z-scored features + CTC loss and
      gradient on logits [Tmax,256,1424] with T_ctc = min(200, n_frames//8+1), labels ~ U{8..24}
      (--surface keras: the CTC part through K.ctc_batch_cost's input, probabilities [B,T,V])
  c3  64 utterances of 20 s, frame-rate CTC (T = 1998, labels ~ U{280..320}): the generic long-lattice kernels
  c4  noise-augmented features only: 512 utterances of U(3,7) s float32 signal + noise at 5..10 dB, gains on
      the device, mix fused into the frame load
  c5  the sweep shape: C2 batches from a larger rotating pool, greedy decode fused into the CTC pass and the
      label error (device edit distance) computed per step
One step = the whole batch through the path.  For c2 / c5 two steps (of different resident batches) are in flight
at a time (pipeline.StepsInFlight, --steps-in-flight 1 for strictly serial steps): the next batch's transform runs
next to the previous batch's HBM-bound z-score and CTC kernels; every step's work lies inside the timed region.  Utterances are sharded by batch across ranks (weak scaling:
every rank owns its own batch); the only collective is the all-reduce of [sum loss, n], issued every second
step on the accumulated sums (the reference prints its mean loss every second step, train.py:71-73).
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
METRIC = "audio-sec/sec (fbank+CTC loss+grad)"
REDUCE_EVERY = 2
WORKLOADS = {
    "c2": dict(batch=256, pool=3, ctc=True,
               name="C2: AISHELL-shaped 256 utt/GPU x U(3,7) s int16 16 kHz (G2), fbank z-scored + CTC loss/grad, "
                    "V=1424, T_ctc=min(200,n_frames//8+1), L~U{8..24}"),
    "c3": dict(batch=64, pool=2, ctc=True,
               name="C3: long utterances 64 utt/GPU x 20 s int16 (G2), fbank z-scored + frame-rate CTC loss/grad, "
                    "V=1424, T=1998, L~U{280..320} (generic lattice kernels)"),
    "c4": dict(batch=512, pool=2, ctc=False,
               name="C4: noise-augmented features 512 utt/GPU x U(3,7) s float32 signal + noise at 5..10 dB, gains "
                    "(SNR2K) on the device, mix fused into the frame load, fbank z-scored; no CTC"),
    "c5": dict(batch=256, pool=8, ctc=True,
               name="C5: sweep of C2-shaped batches (256 utt/GPU per step from a rotating pool of 8), fbank z-scored "
                    "+ CTC loss/grad with the greedy decode fused + device label error"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture(kernel, key):
    """per-launch figures of a kernel from the committed `ncu --set full` capture (profiles/traffic.json)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel][key])
    except Exception:
        return None


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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
# workloads (host arrays)
# ----------------------------------------------------------------------------
def _cached(name, build):
    cache = os.environ.get("ASRK_BENCH_CACHE")
    path = os.path.join(cache, name + ".npz") if cache else None
    if path and os.path.isfile(path):
        d = dict(np.load(path))
        offs = np.concatenate([[0], np.cumsum(d["lens"])])
        for k in ("pcm", "noise"):
            if k in d:
                d[k] = [d[k][offs[i]:offs[i + 1]] for i in range(len(d["lens"]))]
        return d
    d = build()
    if path:
        os.makedirs(cache, exist_ok=True)
        flat = dict(d)
        for k in ("pcm", "noise"):
            if k in flat:
                flat[k] = np.concatenate(flat[k])
        np.savez(path, **flat)
    return d


def make_batch(seed, batch=256, workload="c2"):
    """One batch of the workload as host arrays (numpy).  c2 / c5: the AISHELL shape; c3: 20 s utterances;
    c4: float32 signal + noise + SNR."""
    def build():
        rng = np.random.default_rng(seed)
        if workload == "c3":
            lens = np.full(batch, 320000, dtype=np.int64)
        else:
            lens = synth.ragged_lengths(rng, batch, 3.0, 7.0)
        pcm = [synth.g2_voiced(rng, int(n)) for n in lens]
        nfr = np.array([synth.n_frames(int(n)) for n in lens], dtype=np.int64)
        if workload == "c4":
            sig = [(p.astype(np.float32) / np.float32(32768.0)) for p in pcm]
            noise = []
            for i, n in enumerate(lens):
                x = rng.standard_normal(int(n))          # coloured-noise stand-in: the generator is an input of the mix
                nz = np.cumsum(x) if i % 2 else x
                nz = nz - nz.mean()
                noise.append((nz / nz.max()).astype(np.float32))
            return dict(pcm=sig, noise=noise, snr_db=rng.integers(5, 11, batch).astype(np.int32), lens=lens, nfr=nfr)
        if workload == "c3":
            il = nfr.astype(np.int32)
            x, labels, ll, il = synth.ctc_batch(rng, il, V, 280, 320)
        else:
            il = np.array([synth.t_ctc(int(f)) for f in nfr], dtype=np.int32)
            x, labels, ll, il = synth.ctc_batch(rng, il, V, 8, 24, lmax=64, scale=3.0)
        return dict(pcm=pcm, lens=lens, nfr=nfr, logits=x, labels=labels, label_len=ll, input_len=il)
    tag = "c2" if workload == "c5" else workload
    return _cached("%s_%d_%d" % (tag, seed, batch), build)


def algorithmic_bytes(b, workload):
    """features 2N + 800 n_frames per utterance (noise mode: 4N + 4N + 800 n_frames);
    CTC loss+grad 8 T_ctc V (the fused greedy decode of c5 shares the logits read: 0 extra)."""
    n, f = int(b["lens"].sum()), int(b["nfr"].sum())
    feat = (8 * n if workload == "c4" else 2 * n) + 800 * f
    ctc = int(8 * V * b["input_len"].astype(np.int64).sum()) if "input_len" in b else 0
    return feat, ctc


class DeviceBatch:
    """A batch resident in HBM plus its pinned host copy (for the e2e leg)."""

    def __init__(self, hb, dev, torch, workload, surface):
        from asr_dfcnn_transformer_b200 import features
        self.hb, self.workload = hb, workload
        noises = hb.get("noise")
        pk = features.pack_host(hb["pcm"], FS, "fbank", noises=noises)
        self.pk = pk
        self.B = len(hb["pcm"])
        self.h_samples = pk.samples                       # pinned
        self.samples = pk.samples.to(dev)
        self.noise = pk.noise.to(dev) if pk.noise is not None else None
        self.h_noise = pk.noise
        self.snr_db = torch.from_numpy(hb["snr_db"]).to(dev) if "snr_db" in hb else None
        self.so = torch.from_numpy(pk.sample_offsets).to(dev)
        self.sc = torch.from_numpy(pk.sample_counts).to(dev)
        self.fo = torch.from_numpy(pk.frame_offsets).to(dev)
        self.total_frames = pk.total_frames
        self.feat = torch.empty((pk.total_frames, 200), dtype=torch.float32, device=dev)
        self.audio_s = float(hb["lens"].sum()) / FS
        self.bytes_feat, self.bytes_ctc = algorithmic_bytes(hb, workload)
        if "logits" in hb:
            self.h_logits = torch.from_numpy(hb["logits"]).pin_memory()
            self.h_labels = torch.from_numpy(hb["labels"]).pin_memory()
            self.logits = self.h_logits.to(dev)
            if surface == "keras":                        # what the Keras model hands to K.ctc_batch_cost: softmax [B,T,V]
                self.probs = torch.softmax(self.logits.permute(1, 0, 2), -1).contiguous()
                self.grad = torch.empty_like(self.probs)
            else:
                self.grad = torch.empty_like(self.logits)
            self.labels = self.h_labels.to(dev)
            self.label_len = torch.from_numpy(hb["label_len"]).to(dev)
            self.input_len = torch.from_numpy(hb["input_len"]).to(dev)
            self.grad_scale = torch.full((self.B,), 1.0 / self.B, dtype=torch.float32, device=dev)
            # the loader's host-side length vectors bound the batch (data_loader.py:132-148)
            self.ctc_bounds = (int(hb["input_len"].max()), int(hb["label_len"].max()))


_STEP = {}


MERGED_TAIL = False    # --merged-tail: the z-score as co-work of the fused CTC kernel (one kernel for the step's tail)


def hot_path(dev):
    from asr_dfcnn_transformer_b200 import pipeline
    if dev not in _STEP:
        _STEP[dev] = pipeline.HotPathStep(dev, feature_ctas=int(os.environ.get("ASRK_FEATURE_CTAS", "0")),
                                          merged_tail=MERGED_TAIL)
    return _STEP[dev]


def run_step(db, surface="logits"):
    """The hot path on device-resident inputs.  Returns the CtcResult (None for c4)."""
    from asr_dfcnn_transformer_b200 import ctc, features, utils
    w = db.workload
    if w == "c4":
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", noise=db.noise,
                                    snr_db=db.snr_db, out=db.feat)
        return None
    if surface == "keras":
        hp = hot_path(db.samples.device)
        torch = hp.torch
        cur = torch.cuda.current_stream(hp.device)
        hp._ev_fork.record(cur)
        hp.side.wait_event(hp._ev_fork)
        from asr_dfcnn_transformer_b200 import _lib
        merged = MERGED_TAIL and db.ctc_bounds is not None
        with torch.cuda.stream(hp.side):
            features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                        stream=hp.side,
                                        phases=(_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN | _lib.PHASE_SPEC_STATS)
                                        if merged else _lib.PHASE_ALL)
            hp._ev_join.record(hp.side)
        zw = None
        if merged:       # the CTC kernel normalises the feature rows as co-work: it runs behind the transform
            cur.wait_event(hp._ev_join)
            zw = features.zscore_work(db.feat, db.fo, db.B, db.total_frames, stream=hp.side)
        r = ctc.ctc_loss_grad(db.probs, db.labels, db.label_len, db.input_len, V - 1, layout="btv",
                              grad_scale=db.grad_scale, grad_out=db.grad, bounds=db.ctc_bounds, input_kind="prob",
                              zscore=zw)
        cur.wait_event(hp._ev_join)
        return r
    _, r = hot_path(db.samples.device)(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames,
                                       db.logits, db.labels, db.label_len, db.input_len, V - 1,
                                       feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                                       ctc_bounds=db.ctc_bounds, decode=(w == "c5"))
    if w == "c5":
        db.label_err = utils.edit_distance(r.tokens, r.token_len, db.labels, db.label_len)
    return r


KERNELS = ["spec_setup", "spec_main", "spec_normalize", "ctc_prep", "ctc_fused", "ctc_rows", "ctc_lattice",
           "ctc_grad"]


def run_step_phases(db, ev):
    """the same launches issued phase by phase with CUDA events between them (c4: feature phases only)"""
    from asr_dfcnn_transformer_b200 import _lib, ctc, features
    ev.mark()
    if MERGED_TAIL and db.workload in ("c2", "c5") and db.ctc_bounds is not None:
        # the step's three kernels: transform, statistics, fused CTC with the z-score as co-work
        for ph in (_lib.PHASE_SPEC_SETUP, _lib.PHASE_SPEC_MAIN, _lib.PHASE_SPEC_STATS):
            features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                        phases=ph)
            ev.mark()
        ev.mark()                                   # (no separate prep kernel)
        zw = features.zscore_work(db.feat, db.fo, db.B, db.total_frames)
        ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale,
                          grad_out=db.grad, bounds=db.ctc_bounds, decode=(db.workload == "c5"), zscore=zw)
        for _ in range(4):
            ev.mark()                               # ctc_fused = the merged kernel; no generic kernels
        return
    for ph in (_lib.PHASE_SPEC_SETUP, _lib.PHASE_SPEC_MAIN, _lib.PHASE_SPEC_NORMALIZE):
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                    noise=db.noise, snr_db=db.snr_db, phases=ph)
        ev.mark()
    r = None
    for ph in (_lib.PHASE_CTC_PREP, _lib.PHASE_CTC_FUSED, _lib.PHASE_CTC_ROWS, _lib.PHASE_CTC_LATTICE,
               _lib.PHASE_CTC_GRAD):
        if db.workload != "c4":
            r = ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1,
                                  grad_scale=db.grad_scale, grad_out=db.grad if db.grad.shape == db.logits.shape else None,
                                  phases=ph,
                                  outputs=None if r is None else (r.loss, r.grad, r.row_status, None, None, None))
        ev.mark()


class PhaseTimer:
    def __init__(self, torch):
        self.torch = torch
        self.steps = []
        self.cur = None

    def begin(self):
        # plug the stream for ~2 ms first: the step's launches and event records are then all enqueued while the GPU
        # is busy, so an interval between two events is device time only (without it every interval also holds the
        # host's enqueue latency of the next launch whenever the GPU has run dry: +20 us on a 125 us kernel; with a
        # 0.4 ms plug the last call of a step -- the Python-heavy CTC entry -- still arrived late)
        self.torch.cuda._sleep(4000000)
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
def _cpu_features_worker(arg):
    from oracle import fbank_ref
    sig, nz, db = arg
    if nz is not None:
        sig = fbank_ref.mix_noise(sig, nz, db)          # noise.py:48-52,108 on the host, then the features
    return fbank_ref.compute_fbank(sig).shape[0]


_POOL = {}


def _pool(cores):
    """One worker pool for the whole run (forked before CUDA is initialised)."""
    import multiprocessing as mp
    if cores not in _POOL:
        import atexit
        _POOL[cores] = mp.get_context("fork").Pool(cores)
        atexit.register(_close_pool, cores)
    return _POOL[cores]


def _close_pool(cores):
    p = _POOL.pop(cores, None)
    if p is not None:
        p.terminate()
        p.join()


def cpu_sample(hb, n_utt, cores):
    """Time the oracle (port of the reference's CPU path) on the first n_utt utterances of the workload:
    per-frame scipy FFT features (+ the numpy noise mix for c4) fanned out over the host cores + the float32 C
    restatement of TF's CTC loss/grad (OpenMP)."""
    from oracle import build_c
    build_c.lib()
    sigs = hb["pcm"][:n_utt]
    noises = hb["noise"][:n_utt] if "noise" in hb else [None] * n_utt
    dbs = hb["snr_db"][:n_utt] if "snr_db" in hb else [0] * n_utt
    audio_s = sum(len(s) for s in sigs) / FS
    args = list(zip(sigs, noises, [int(d) for d in dbs]))
    t0 = time.perf_counter()
    if cores > 1:
        _pool(cores).map(_cpu_features_worker, args, chunksize=1)
    else:
        for a in args:
            _cpu_features_worker(a)
    t_feat = time.perf_counter() - t0
    t_ctc = 0.0
    if "logits" in hb:
        il = hb["input_len"][:n_utt]
        T = int(il.max())
        x = np.ascontiguousarray(hb["logits"][:T, :n_utt])
        t0 = time.perf_counter()
        build_c.ctc_loss_grad(x, hb["labels"][:n_utt], hb["label_len"][:n_utt], il, V - 1, real="f32", threads=cores)
        t_ctc = time.perf_counter() - t0
    return audio_s, t_feat, t_ctc


def cpu_sample_size(workload, n_utt, cores):
    # bounded: ~1 s per pass on 16 cores
    return {"c2": n_utt, "c5": n_utt, "c3": min(n_utt, 16), "c4": min(n_utt, 256)}[workload]


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: /root/reference is
    Python+TensorFlow and does not exist on the GPU box) on the host cores, same workload/metric, bounded
    sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    hb = make_batch(2000, w["batch"], args.workload)
    n_utt = cpu_sample_size(args.workload, len(hb["pcm"]), cores)
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
        "config": {"workload": w["name"], "utterances_per_gpu": w["batch"],
                   "sample": "%d utterances of the workload per step" % n_utt, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": "audio-sec/sec", "cores": cores, "kind": "port",
                         "sample": "%d %s utterances per step: oracle/fbank_ref.py (per-frame scipy FFT, "
                                   "multiprocessing) + oracle/ctc_ref.c float32 (OpenMP)" % (n_utt, args.workload.upper())},
        "e2e": {"value": val, "unit": "audio-sec/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def pin_to_gpu_numa_node(local_rank):
    """Keep this rank's threads (and therefore its first-touched pinned buffers) on the NUMA node of its GPU:
    eight ranks staging through one node's memory is what limited the end-to-end scaling in round 1."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        devid = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, devid)
        node = int(open(path + "numa_node").read())
        if node < 0:
            return None
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------
# main
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--surface", default="logits", choices=["logits", "keras"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merged-tail", action="store_true",
                    help="one kernel for the step's tail: the z-score pass as co-work of the fused CTC kernel "
                         "(bit-identical, measured no faster: profiles/r2_tail.md)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--steps-in-flight", type=int, default=None,
                    help="device-resident steps of different batches in flight at once (pipeline.StepsInFlight); "
                         "default 2 for c2 / c5 with graphs, 1 (strictly serial steps) otherwise")
    ap.add_argument("--feature-ctas", type=int, default=None,
                    help="CTAs of the persistent transform kernel (default: 104 with several steps in flight, else one per SM)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global MERGED_TAIL
    MERGED_TAIL = bool(args.merged_tail)
    if args.surface == "keras" and args.workload != "c2":
        ap.error("--surface keras is a variant of the c2 workload")

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started by hand with --gpus N: become the launch the driver uses (one rank per GPU, NCCL, loopback)
        import socket
        with socket.socket() as so:
            so.bind(("127.0.0.1", 0))
            port = so.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                  str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", str(port),
                                  os.path.abspath(__file__)] + sys.argv[1:])

    if os.environ.get("ASRK_BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["ASRK_BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    wl = WORKLOADS[args.workload]
    BATCH, POOL = wl["batch"], wl["pool"]
    graphable = (not args.no_graph) and args.surface == "logits" and args.workload in ("c2", "c3", "c5")
    in_flight = args.steps_in_flight
    if in_flight is None:
        in_flight = 2 if (graphable and args.workload in ("c2", "c5")) else 1
    if in_flight > 1 and not graphable:
        ap.error("--steps-in-flight > 1 needs the graph path (c2 / c3 / c5, logits surface, no --no-graph)")
    if in_flight > 1 and POOL % in_flight:
        POOL += in_flight - POOL % in_flight          # every resident batch keeps its lane
    feature_ctas = args.feature_ctas if args.feature_ctas is not None else (104 if in_flight > 1 else 0)

    # ---- (0) CPU baseline on rank 0 (N = 1 only), BEFORE CUDA is initialised so that the
    # worker processes can be forked safely ---------------------------------------------
    cpu = None
    hb0 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        hb0 = make_batch(2000, BATCH, args.workload)
        n_utt = cpu_sample_size(args.workload, len(hb0["pcm"]), cores)
        cpu_sample(hb0, min(n_utt, cores), cores)          # warm the pool / page cache
        a = tf_ = tc_ = 0.0
        passes = 0
        # bounded sample: whole passes over the batch until ~5 s of wall clock (= cores x 5 s of CPU work) are spent
        while passes < 2 or (tf_ + tc_ < 5.0 and passes < 64):
            a1, t1, t2 = cpu_sample(hb0, n_utt, cores)
            a, tf_, tc_ = a + a1, tf_ + t1, tc_ + t2
            passes += 1
        cpu = {"value": a / (tf_ + tc_), "unit": "audio-sec/sec", "cores": cores, "kind": "port",
               "sample": "%d passes over %d utterances of one %s batch (%.0f audio-s): oracle/fbank_ref.py features "
                         "(per-frame scipy FFT, %d processes, %.2f s) + oracle/ctc_ref.c float32 CTC loss/grad "
                         "(OpenMP, %.2f s)" % (passes, n_utt, args.workload.upper(), a, cores, tf_, tc_)}

    import torch
    import torch.distributed as dist
    from asr_dfcnn_transformer_b200 import _lib, ctc, pipeline
    L = _lib.lib()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: POOL distinct batches per rank, resident in HBM --------------
    pool = [DeviceBatch(hb0 if (i == 0 and rank == 0 and hb0 is not None)
                        else make_batch(2000 + 100 * rank + i, BATCH, args.workload), dev, torch, args.workload,
                        args.surface) for i in range(POOL)]
    audio_per_step = float(np.mean([d.audio_s for d in pool]))
    acc = torch.zeros(2, dtype=torch.float64, device=dev)     # running [sum loss, n] of this rank
    red = torch.zeros(2, dtype=torch.float64, device=dev)     # operand / result of the all-reduce
    side = torch.cuda.Stream(device=dev)
    step_no = [0]

    def reduce_loss(r, in_graph=False):
        # the path's only collective: all-reduce of [sum loss, n].  The sums accumulate on the device (one tiny
        # kernel per step, inside the step's graph) and are all-reduced every REDUCE_EVERY steps on a side stream,
        # off the SMs' critical path
        if r is None:
            return
        if not in_graph:
            ctc.loss_sum(r.loss, r.row_status, out=acc, accumulate=True)
        step_no[0] += 1
        if step_no[0] % REDUCE_EVERY == 0:
            cur = torch.cuda.current_stream()
            if world > 1:
                cur.wait_stream(side)                      # the previous all-reduce is done with `red`
            red.copy_(acc)
            acc.zero_()
            if world > 1:
                pipeline.all_reduce_loss(red, stream=side)

    # steady state: one CUDA graph per resident batch (the step's launches, stream forks and joins replayed without
    # any host-side enqueue); c4 (one call) and the keras surface stay on the plain path
    use_graph = graphable
    launches_per_step = None
    flight = None
    if in_flight > 1:
        # several steps in flight: one HotPathStep + graph per resident batch, replayed round-robin on `in_flight`
        # lane streams; every replay writes its own [sum loss, n], summed over REDUCE_EVERY steps on the side stream
        flight = pipeline.StepsInFlight(dev, lanes=in_flight, feature_ctas=feature_ctas, merged_tail=MERGED_TAIL)
        for db in pool:
            n0 = L.asrk_launch_count()
            db.slot = flight.add(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels,
                                 db.label_len, db.input_len, V - 1, feat_out=db.feat, grad_out=db.grad,
                                 grad_scale=db.grad_scale, ctc_bounds=db.ctc_bounds, decode=(args.workload == "c5"))
            db.res = db.slot.result
            launches_per_step = int(L.asrk_launch_count() - n0) // 2
    elif use_graph:
        hot_path(dev).reserve(BATCH, max(d.total_frames for d in pool), max(d.logits.shape[0] for d in pool),
                              pool[0].labels.shape[1])
        for db in pool:
            n0 = L.asrk_launch_count()
            db.graph, _, db.res = hot_path(dev).capture(
                db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels, db.label_len,
                db.input_len, V - 1, feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                ctc_bounds=db.ctc_bounds, decode=(args.workload == "c5"), loss_acc=acc)
            launches_per_step = int(L.asrk_launch_count() - n0) // 2      # (capture() runs the step once before capturing)
        acc.zero_()

    pending = []

    def step_in_flight(db):
        lane = flight.launch(db.slot)
        if args.workload == "c5":
            from asr_dfcnn_transformer_b200 import utils
            with torch.cuda.stream(lane):
                db.label_err = utils.edit_distance(db.res.tokens, db.res.token_len, db.labels, db.label_len)
                db.slot.done.record(lane)
        pending.append(db.slot)
        step_no[0] += 1
        if step_no[0] % REDUCE_EVERY == 0:
            for sl in pending:
                side.wait_event(sl.done)
            with torch.cuda.stream(side):
                red.copy_(pending[0].loss_sum)
                for sl in pending[1:]:
                    red.add_(sl.loss_sum)
                ev = torch.cuda.Event()
                ev.record(side)
                for sl in pending:
                    sl.reusable = ev              # the slot's next replay overwrites loss_sum: behind this read
                if world > 1:
                    dist.all_reduce(red, op=dist.ReduceOp.SUM)
            pending.clear()

    def step(db):
        if flight is not None:
            step_in_flight(db)
        elif use_graph:
            db.graph.replay()
            if args.workload == "c5":
                from asr_dfcnn_transformer_b200 import utils
                db.label_err = utils.edit_distance(db.res.tokens, db.res.token_len, db.labels, db.label_len)
            reduce_loss(db.res, in_graph=True)
        else:
            reduce_loss(run_step(db, args.surface))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident throughput ------------------------------------
    for i in range(args.warmup):
        step(pool[i % POOL])
    barrier()
    acc.zero_()
    step_no[0] = 0
    pending.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.asrk_launch_count()
    e0.record()
    audio = 0.0
    for i in range(args.steps):
        db = pool[i % POOL]
        step(db)
        audio += db.audio_s
    if flight is not None:
        flight.join()
    if world > 1 or flight is not None:
        torch.cuda.current_stream().wait_stream(side)
    e1.record()
    barrier()
    launches = int(L.asrk_launch_count() - launches0)
    if use_graph:
        # kernels of the library inside one step's graph (counted while it was captured) x replays, plus whatever was
        # launched outside the graphs in the timed region
        launches += launches_per_step * args.steps
    loss_check = red.clone()
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
    kms = None
    narrow_ms = None
    if args.surface == "logits":
        pt = PhaseTimer(torch)
        for i in range(args.steps):
            pt.begin()
            run_step_phases(pool[i % POOL], pt)
            pt.end()
        torch.cuda.synchronize()
        kms = pt.summary()
        if flight is not None and feature_ctas:
            # the transform as the in-flight step launches it (fewer CTAs than SMs), alone, behind the same plug
            from asr_dfcnn_transformer_b200 import features as _features
            tot = 0.0
            for i in range(min(args.steps, 12)):
                db = pool[i % POOL]
                torch.cuda._sleep(4000000)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                _features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                             phases=_lib.PHASE_SPEC_SETUP)
                a0.record()
                _features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                             phases=_lib.PHASE_SPEC_MAIN, cta_limit=feature_ctas)
                a1.record()
                torch.cuda.synchronize()
                tot += a0.elapsed_time(a1)
            narrow_ms = tot / min(args.steps, 12)
        if args.workload != "c4":
            # a noise-free batch launches no kernel in the setup phase (one small memset): what the events
            # bracket there is the host's enqueue latency, not device work
            kms.pop("spec_setup", None)
    clocks = sampler.stop() if rank == 0 else None

    # ---- (3) end to end through the public API with host buffers ------------
    e2e = None
    if args.workload in ("c2", "c5") and args.surface == "logits":
        e2e_steps = max(8, min(2 * args.steps, 48))      # ~0.2 s: the pipeline's fill and drain are inside the timed region
        d0 = pool[0]
        T = max(d.logits.shape[0] for d in pool)
        # every batch of the pool has its own T_max: the round-trip buffers take the largest
        # how the logits come in next to the returning results (profiles/r2_e2e.md): at N <= 2 the x16 link is the
        # limit and a DMA of the padded tensor avoids the SM-read / device->host conflict; from N = 4 the HOST's memory
        # system is (94 GB/s of device->host writes for the whole box), and the 29 % fewer bytes of the zero-copy staging win
        logits_in = os.environ.get("ASRK_BENCH_LOGITS_IN") or ("dma" if world <= 2 else "zero_copy")

        def mk(ret):
            return pipeline.HostRoundTrip(hot_path(dev), max(d.total_frames for d in pool),
                                          max(d.h_samples.numel() for d in pool), d0.h_samples.dtype, T, BATCH, V,
                                          d0.h_labels.shape[1], slots=3, return_outputs=ret,
                                          logits_in=logits_in if ret else None)

        def leg(rt, steps):
            h2d = d2h = 0
            barrier()
            e0.record()
            a2 = 0.0
            for i in range(steps):
                db = pool[i % POOL]
                s = rt.submit(db.h_samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.h_logits, db.h_labels,
                              db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, ctc_bounds=db.ctc_bounds,
                              valid_rows=int(db.hb["input_len"].astype(np.int64).sum()))
                h2d += s.h2d_bytes
                d2h += s.d2h_bytes
                a2 += db.audio_s
            rt.drain()
            e1.record()
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1), a2], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt[0:1], op=dist.ReduceOp.MAX)
                dist.all_reduce(tt[1:2], op=dist.ReduceOp.SUM)
            return float(tt[1]) / (float(tt[0]) * 1e-3), float(tt[0]), h2d / steps, d2h / steps

        # (the round-trip buffers take the pool's largest T_max; a batch with a shorter T_max uses their first rows)
        full = mk(True)
        leg(full, 2)
        v_full, ms_full, h2d, d2h = leg(full, e2e_steps)
        del full
        torch.cuda.empty_cache()
        lo = mk(False)
        leg(lo, 2)
        v_lo, ms_lo, h2d_lo, d2h_lo = leg(lo, e2e_steps)
        del lo
        n_used = e2e_steps
        e2e = {"value": v_full, "unit": "audio-sec/sec", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_used,
               "what": "full host->host round trip through pipeline.HostRoundTrip: PCM, labels (DMA) and logits (%s) in "
                       "from pinned host buffers; loss, features (DMA) and gradient (valid rows, written by the SMs) out "
                       "into pinned host buffers, three steps in flight"
                       % ("padded tensor by DMA" if logits_in == "dma" else "valid rows, read by the SMs"),
               "pcie_gbs_per_gpu": (h2d + d2h) * n_used / (ms_full * 1e-3) / 1e9,
               "results_stay_on_device": {"value": v_lo, "unit": "audio-sec/sec", "h2d_bytes_per_step": int(h2d_lo),
                                          "d2h_bytes_per_step": int(d2h_lo),
                                          "what": "same inputs from host buffers, only the per-utterance loss comes "
                                                  "back (features and gradient stay in HBM for the consumer)"}}
    elif args.workload in ("c3", "c4"):
        e2e = None

    if rank == 0:
        peak, peak_src = peaks()
        bf = float(np.mean([d.bytes_feat for d in pool]))
        bc = float(np.mean([d.bytes_ctc for d in pool]))
        step_alg = bf + bc
        roof = None
        if kms is not None:
            cands = [k for k in ("spec_main", "ctc_fused", "ctc_rows", "ctc_lattice", "ctc_grad") if k in kms]
            dom = max(cands, key=lambda k: kms[k])
            # algorithmic bytes of one launch of that kernel: the feature kernel moves the feature bytes; the fused
            # CTC kernel the logits in and the gradient out; of the generic kernels rows reads the logits, grad reads
            # them again (non-algorithmic) and writes the gradient, the lattice kernel moves no algorithmic bytes
            algs = {"spec_main": bf, "ctc_fused": bc, "ctc_rows": bc / 2, "ctc_grad": bc / 2, "ctc_lattice": 0.0}
            note = None
            if algs[dom] == 0.0:
                # C3: the longest kernel is the lattice sweep, a dependency chain of T steps that moves no algorithmic
                # bytes; the HBM roofline is quoted for the longest kernel that does
                note = "longest kernel: %s %.3f ms (latency bound, no algorithmic bytes)" % (dom, kms[dom])
                dom = max((k for k in cands if algs[k] > 0), key=lambda k: kms[k])
            alg = algs[dom]
            ach = alg / (kms[dom] * 1e-3) / 1e9
            names = {"spec_main": "spectrogram_kernel", "ctc_fused": "fused_small_kernel"}
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_capture({"c4": {"spec_main": "spec_main_f32"}}.get(args.workload, {}).get(dom, names.get(dom, dom)),
                                           "dram_bytes"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                    "fp64_pipe_pct_ncu": ncu_capture(names.get(dom, dom), "fp64_pipe_pct") if dom == "spec_main" else None,
                    "step_algorithmic_bytes": step_alg,
                    "step_frac": (step_alg / (ms / args.steps * 1e-3) / 1e9) / peak}
            if note:
                roof["note"] = note
            if flight is not None:
                roof["in_step"] = {
                    "transform_ctas": feature_ctas or None, "transform_alone_ms": narrow_ms,
                    "note": "kernel_ms and frac are each kernel ALONE on the whole chip; in the timed step the transform "
                            "runs on transform_ctas SMs next to the previous batch's z-score and CTC kernels "
                            "(steps_in_flight), which is what step_frac measures"}
        line = {
            "metric": METRIC, "value": value, "unit": "audio-sec/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 FFT / f32 log, CTC",
            "data": "synthetic",
            "config": {"workload": wl["name"], "surface": args.surface,
                       "utterances_per_gpu": BATCH, "audio_s_per_step_per_gpu": audio_per_step,
                       "all_reduce": "[sum loss, n] accumulated on the device, all-reduced every %d steps" % REDUCE_EVERY,
                       "numa_node_rank0": numa, "cuda_graph": bool(use_graph),
                       "steps_in_flight": in_flight,
                       "feature_ctas": feature_ctas or "one per SM",
                       "tail": ("one kernel: fused CTC with the z-score as co-work" if MERGED_TAIL and args.workload in ("c2", "c5")
                                else "z-score and CTC kernels separate"),
                       "l2": "inputs larger than L2: %d distinct batches rotated, ~%.0f MB touched per step"
                             % (POOL, (step_alg + bc / 2) / 1e6)},
            "roofline": roof,
            "kernel_ms": kms,
            "loss_mean": float(loss_check[0] / max(float(loss_check[1]), 1.0)),
            "cpu_baseline": cpu,
            "e2e": e2e,
            # measured: kernels the library launched inside the timed region (asrk_launch_count)
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if args.workload == "c5":
            line["label_error_mean"] = float(pool[0].label_err.mean())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
