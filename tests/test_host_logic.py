"""CPU: host-side packing / frame rules / error behaviour (no device needed)."""
import os

import numpy as np
import pytest

from oracle import synth


def test_pack_host_alignment_and_offsets():
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(0)
    sigs = [synth.g1_white(rng, n) for n in (16000, 401, 9999, 300)]
    pk = features.pack_host(sigs, pin=False)
    assert pk.sample_counts.tolist() == [16000, 401, 9999, 300]
    assert all(o % 8 == 0 for o in pk.sample_offsets)         # 16-byte aligned int16 starts
    assert pk.n_frames.tolist() == [98, 1, 60, 0]
    assert pk.frame_offsets.tolist() == [0, 98, 99, 159, 159]
    buf = pk.samples.numpy()
    for s, o in zip(sigs, pk.sample_offsets):
        assert np.array_equal(buf[o:o + len(s)], s)
    f32 = [s.astype(np.float32) / 32768 for s in sigs[:2]]
    pk = features.pack_host(f32, pin=False, noises=[np.zeros(16000, np.float32), np.ones(401, np.float32)])
    assert all(o % 4 == 0 for o in pk.sample_offsets) and pk.noise is not None


def test_pack_host_errors():
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):
        features.pack_host([], pin=False)
    with pytest.raises(ValueError):                       # short last frame, like the reference
        features.pack_host([synth.g1_white(rng, 399)], pin=False)
    with pytest.raises(ValueError):                       # 8 kHz audio: frame rule asks for too many frames
        features.pack_host([synth.g1_white(rng, 8000)], fs=8000, pin=False)
    with pytest.raises(ValueError):
        features.pack_host([np.zeros((2, 100), np.int16)], pin=False)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from asr_dfcnn_transformer_b200 import ctc, features
    with pytest.raises(RuntimeError):
        features.compute_features([np.zeros(16000, np.int16)])
    with pytest.raises(RuntimeError):
        ctc.greedy_decode(torch.zeros(3, 1, 4), [3])
    from asr_dfcnn_transformer_b200 import pipeline
    with pytest.raises(RuntimeError):
        pipeline.HotPathStep()
    with pytest.raises(RuntimeError):
        pipeline.StepsInFlight(lanes=2)


def test_snr2k_length_limit_is_loud():
    """the one limit the reference's SNR2K does not have (device summation tree: 2**22 samples) raises at the
    Python surface instead of returning the C ABI's NaN gain"""
    from asr_dfcnn_transformer_b200 import noise
    big = np.zeros(noise.SNR2K_MAX_SAMPLES + 1, np.float32)
    with pytest.raises(ValueError, match="2\\*\\*22"):
        noise.SNR2K(big, big, 5)


def test_bench_workload_shapes():
    import bench
    hb = bench.make_batch(2000, batch=8)
    assert len(hb["pcm"]) == 8 and hb["logits"].shape[1:] == (8, synth.VOCAB_DICT_TXT)
    assert all(48000 <= len(p) <= 112000 for p in hb["pcm"])
    assert (hb["label_len"] < hb["input_len"]).all()
    feat, ctc_b = bench.algorithmic_bytes(hb, "c2")
    n = sum(len(p) for p in hb["pcm"])
    assert feat == 2 * n + 800 * int(hb["nfr"].sum())
    assert ctc_b == 8 * synth.VOCAB_DICT_TXT * int(hb["input_len"].sum())
    # the other workloads (BASELINE.json configs[2..4])
    h3 = bench.make_batch(3000, batch=2, workload="c3")
    assert all(len(p) == 320000 for p in h3["pcm"]) and h3["logits"].shape == (1998, 2, synth.VOCAB_DICT_TXT)
    assert h3["label_len"].min() >= 280
    h4 = bench.make_batch(4000, batch=4, workload="c4")
    assert h4["pcm"][0].dtype == np.float32 and len(h4["noise"]) == 4 and "logits" not in h4
    f4, c4 = bench.algorithmic_bytes(h4, "c4")
    assert c4 == 0 and f4 == 8 * int(h4["lens"].sum()) + 800 * int(h4["nfr"].sum())


def test_loader_rules(tmp_path):
    """data_loader.py:63-71 vocabulary, :44-60 pny2id, :132/:231 lengths, :139-141 rejects."""
    from asr_dfcnn_transformer_b200 import data_loader as dl
    d = tmp_path / "dict.txt"
    d.write_text("a1\tx\nb2\ty\na1\tz\nc3\tw\n", encoding="utf-8")
    size, s2i, i2s = dl.load_acoustic_vocab(str(d))
    assert size == 5 and s2i["_"] == 4 and i2s[4] == "_"
    assert s2i["a1"] == 2                      # duplicate symbol -> the LATER index
    assert dl.pny2id(" b2 c3 ", s2i) == [1, 3]
    with pytest.raises(ValueError):
        dl.pny2id("b2 zz", s2i)
    assert dl.ctc_input_length(998) == 125 and dl.ctc_input_length(1600) == 200
    assert dl.ctc_input_length(1600, capped=False) == 201
    # every row rejected -> no device work at all: too long, label >= input length, unknown symbol
    rng = np.random.default_rng(0)
    sigs = [synth.g1_white(rng, 16000 * 17), synth.g1_white(rng, 4000), synth.g1_white(rng, 16000)]
    wav, il, lab, ll, keep = dl.data_generation(sigs, ["b2", "b2 c3 b2 c3", "b2 nope"], s2i)
    assert wav is None and keep == [] and il.shape == (0,) and lab.shape == (0, 64)
    ref = os.path.join("/root/reference", "dict.txt")
    if os.path.isfile(ref):
        size, s2i, _ = dl.load_acoustic_vocab(ref)
        assert size == synth.VOCAB_DICT_TXT and s2i["_"] == size - 1
        assert s2i["heng"] == 1364                                   # duplicate key: the later index
        pd = pytest.importorskip("pandas")                            # the reference's own recipe (:64-69)
        symbol_list = pd.read_table(ref, header=None).iloc[:, 0].tolist()
        symbol_list.append("_")
        assert dict([p, i] for i, p in enumerate(symbol_list)) == s2i


def test_bench_rejects_steps_in_flight_without_the_graph_path():
    """several steps in flight are replayed CUDA graphs: workloads / surfaces that stay on the plain path refuse the
    flag before any work is done (argparse error, exit status 2)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in (["--workload", "c4"], ["--surface", "keras"], ["--no-graph"]):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps-in-flight", "2"] + extra,
                             capture_output=True, text=True, timeout=120)
        assert out.returncode == 2 and "steps-in-flight" in out.stderr, (extra, out.stderr[-500:])
