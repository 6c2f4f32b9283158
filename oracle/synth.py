"""Seeded synthetic inputs for the five BASELINE.json configurations (SURVEY.md
section 8d).  Used by tests, bench.py and tools/make_golden.py -- pure data
generation, no reference arithmetic, so both the product-side bench and the
oracle-side tests may import it.

Audio generators (int16, 16 kHz):
  G1  white Gaussian, sigma 3000
  G2  "voiced": 20 harmonics of f0~U(120,250) Hz with 1/k amplitudes, 3 Hz AM
      envelope, Gaussian floor sigma 50 (high dynamic range: exposes fp32 FFTs)
  G3  edge cases: silence, single click, full-scale square wave, float-hazard
      lengths of the frame-count expression (wav_util.py:61)
"""
import numpy as np

FS = 16000
HAZARD_LENGTHS = (16080, 16240, 64240, 64880, 65520, 129040)
VOCAB_DICT_TXT = 1424        # dict.txt rows (1423) + '_' blank  (data_loader.py:86-88)
VOCAB_MIXDICT = 1536


def g1_white(rng, n, sigma=3000.0):
    return np.clip(np.rint(rng.normal(0.0, sigma, n)), -32768, 32767).astype(np.int16)


def g2_voiced(rng, n, amp=8000.0, floor=50.0):
    t = np.arange(n, dtype=np.float64) / FS
    f0 = rng.uniform(120.0, 250.0)
    x = np.zeros(n)
    for k in range(1, 21):
        x += (amp / k) * np.sin(2 * np.pi * k * f0 * t + rng.uniform(0, 2 * np.pi))
    x *= 0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t + rng.uniform(0, 2 * np.pi))
    x += rng.normal(0.0, floor, n)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def g3_edge_cases(rng):
    """name -> int16 signal."""
    out = {}
    out["silence"] = np.zeros(16000, dtype=np.int16)
    click = np.zeros(16000, dtype=np.int16)
    click[5000] = 32767
    out["click"] = click
    n = 24000
    sq = np.where((np.arange(n) // 37) % 2 == 0, 32767, -32767).astype(np.int16)
    out["square_p74"] = sq           # period 74 samples: does not divide the hop
    out["min_length"] = g1_white(rng, 400)
    for h in HAZARD_LENGTHS:
        out[f"hazard_{h}"] = g2_voiced(rng, h)
    return out


def ragged_lengths(rng, batch, lo_s, hi_s):
    """utterance lengths in samples, duration ~ U(lo_s, hi_s) seconds."""
    return np.rint(rng.uniform(lo_s, hi_s, batch) * FS).astype(np.int64)


def labels_with_repeats(rng, length, vocab, p_repeat=0.1):
    """iid uniform on [0, vocab-2] with ~10 % forced immediate repeats."""
    lab = rng.integers(0, vocab - 1, size=length)
    for i in range(1, length):
        if rng.random() < p_repeat:
            lab[i] = lab[i - 1]
    return lab.astype(np.int32)


def n_repeats(lab):
    lab = np.asarray(lab)
    return int(np.sum(lab[1:] == lab[:-1])) if len(lab) > 1 else 0


def logits_tbv(rng, T, B, V, scale=3.0):
    return (scale * rng.standard_normal((T, B, V))).astype(np.float32)


def ctc_batch(rng, input_len, V, lab_lo, lab_hi, lmax=None, scale=3.0):
    """Logits [T,B,V] fp32 + padded labels for given per-utterance CTC lengths.
    Label lengths ~ U{lab_lo..lab_hi}, clipped so that every row is feasible
    (len + repeats <= input_len, and len < input_len as data_loader.py:141)."""
    input_len = np.asarray(input_len, dtype=np.int32)
    B = len(input_len)
    T = int(input_len.max())
    labs = []
    for b in range(B):
        L = int(rng.integers(lab_lo, lab_hi + 1))
        L = max(0, min(L, int(input_len[b]) - 1))
        lab = labels_with_repeats(rng, L, V)
        while L > 0 and L + n_repeats(lab) > int(input_len[b]):
            L -= 1
            lab = lab[:L]
        labs.append(lab)
    lmax = max([len(l) for l in labs] + [1]) if lmax is None else lmax
    labels = np.zeros((B, lmax), dtype=np.int32)
    label_len = np.zeros(B, dtype=np.int32)
    for b, lab in enumerate(labs):
        labels[b, : len(lab)] = lab
        label_len[b] = len(lab)
    x = logits_tbv(rng, T, B, V, scale)
    return x, labels, label_len, input_len


def n_frames(n_samples, fs=FS):
    """wav_util.py:61, the identical Python float expression."""
    return int(n_samples / fs * 1000 - 25) // 10 + 1


def t_ctc(nf):
    """data_loader.py:132."""
    return min(200, nf // 8 + 1)


# ---- the five configurations ------------------------------------------------
def config_c1(seed=1000):
    """8 x 10 s, dict.txt vocab, L ~ U{25..45}."""
    rng = np.random.default_rng(seed)
    pcm = [g2_voiced(rng, 160000) for _ in range(8)]
    il = np.array([t_ctc(n_frames(len(p))) for p in pcm], dtype=np.int32)
    x, labels, ll, il = ctc_batch(rng, il, VOCAB_DICT_TXT, 25, 45, lmax=64)
    return dict(pcm=pcm, logits=x, labels=labels, label_len=ll, input_len=il, V=VOCAB_DICT_TXT)


def config_c2(seed=2000, batch=256, gen=g2_voiced):
    """256 x ~5 s (U(3,7) s), T_ctc = min(200, n_frames//8+1), L ~ U{8..24}."""
    rng = np.random.default_rng(seed)
    lens = ragged_lengths(rng, batch, 3.0, 7.0)
    pcm = [gen(rng, int(n)) for n in lens]
    il = np.array([t_ctc(n_frames(int(n))) for n in lens], dtype=np.int32)
    x, labels, ll, il = ctc_batch(rng, il, VOCAB_DICT_TXT, 8, 24, lmax=64)
    return dict(pcm=pcm, logits=x, labels=labels, label_len=ll, input_len=il, V=VOCAB_DICT_TXT)


def config_c3(seed=3000, batch=64, seconds=20.0):
    """64 x 20 s, frame-rate CTC (T = n_frames = 1998), L ~ U{280..320}."""
    rng = np.random.default_rng(seed)
    n = int(seconds * FS)
    pcm = [g2_voiced(rng, n) for _ in range(batch)]
    il = np.array([n_frames(n)] * batch, dtype=np.int32)
    x, labels, ll, il = ctc_batch(rng, il, VOCAB_DICT_TXT, 280, 320)
    return dict(pcm=pcm, logits=x, labels=labels, label_len=ll, input_len=il, V=VOCAB_DICT_TXT)


def config_c4(seed=4000, batch=512, noise_fn=None):
    """noise-augmented features: fp32 signal (G2/32768) + coloured noise, SNR in
    {5..10} dB, colour in {-1.0..1.0 step 0.1}.  ``noise_fn(x_random, colour)``
    shapes a N(0,1) draw (the oracle's or the product's colour filter)."""
    rng = np.random.default_rng(seed)
    lens = ragged_lengths(rng, batch, 3.0, 7.0)
    sig, noi, db, col = [], [], [], []
    for n in lens:
        s = (g2_voiced(rng, int(n)).astype(np.float32) / np.float32(32768.0)).astype(np.float32)
        c = int(rng.integers(-10, 11)) / 10
        xr = rng.normal(0.0, 1.0, int(n))
        nz = noise_fn(xr, c) if noise_fn is not None else (xr / np.abs(xr).max()).astype(np.float32)
        sig.append(s)
        noi.append(nz.astype(np.float32))
        db.append(int(rng.integers(5, 11)))
        col.append(c)
    return dict(signal=sig, noise=noi, snr_db=np.array(db, dtype=np.int32), colour=col)
