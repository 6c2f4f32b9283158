"""GPU parity on the BASELINE.json configurations at FULL size, EVERY row against the oracle at the
contract tolerances (features 1e-4, CTC loss / gradient 1e-3, tokens bit-exact).  The worst errors are
written to gpurun_out/parity.json, including the worst UN-widened z-score error and how
many (utterance, column) pairs needed the conditioning allowance of tests/util.py::zscore_feature_err."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from oracle import build_c, ctc_ref, fbank_ref, synth
from tests.util import (CTC_ATOL, CTC_RTOL, FEATURE_TOL, assert_ctc_grad_close, record_parity,
                        zscore_err_report)

pytestmark = pytest.mark.gpu


def _ref_features(sig):
    raw = fbank_ref.compute_fbank_unnormalised(sig)
    return raw, fbank_ref.zscore_columns(raw)


def _ref_features_many(sigs):
    n = min(os.cpu_count() or 1, 16)
    if n <= 1 or len(sigs) < 8:
        return [_ref_features(s) for s in sigs]
    with mp.get_context("fork").Pool(n) as pool:       # forked before this process touches CUDA state in the children
        return pool.map(_ref_features, sigs, chunksize=4)


def _features_all_rows(sigs, out, fo, refs):
    worst_plain = worst_wide = 0.0
    needed = 0
    for i in range(len(sigs)):
        raw, z = refs[i]
        got = out[fo[i]:fo[i + 1]]
        assert got.shape == z.shape, (i, got.shape, z.shape)
        a, b, c = zscore_err_report(got, z, raw)
        worst_plain, worst_wide, needed = max(worst_plain, a), max(worst_wide, b), needed + c
    return worst_plain, worst_wide, needed


def test_c2_features_all_256_utterances():
    import bench
    from asr_dfcnn_transformer_b200 import features
    hb = bench.make_batch(2000)
    refs = _ref_features_many(hb["pcm"])
    fb = features.compute_features(hb["pcm"], mode="fbank")
    out = fb.features.cpu().numpy()
    plain, wide, needed = _features_all_rows(hb["pcm"], out, fb.frame_offsets, refs)
    record_parity("C2_features", {"utterances": 256, "worst_err_unwidened": plain, "worst_err_widened": wide,
                                  "columns_needing_allowance": needed, "tolerance": FEATURE_TOL})
    assert plain <= FEATURE_TOL, (plain, wide, needed)      # no allowance used on C2
    # the raw log-spectrogram too (isolates the transform)
    fr = features.compute_features(hb["pcm"], mode="fbank_raw").features.cpu().numpy()
    worst = 0.0
    for i in range(256):
        raw = refs[i][0]
        worst = max(worst, float(np.max(np.abs(fr[fb.frame_offsets[i]:fb.frame_offsets[i + 1]] - raw) / np.maximum(np.abs(raw), 1.0))))
    record_parity("C2_log_spectrogram", {"worst_err": worst, "tolerance": FEATURE_TOL})
    assert worst <= FEATURE_TOL


def _ctc_all_rows(x, labels, ll, il, V, r, chunk):
    """loss / gradient of every row against the float64 C oracle, `chunk` utterances at a time."""
    loss = r.loss.cpu().numpy()
    worst_loss = 0.0
    B = x.shape[1]
    for b0 in range(0, B, chunk):
        sl = slice(b0, min(b0 + chunk, B))
        xs = np.ascontiguousarray(x[:, sl])
        rl, rg, st = build_c.ctc_loss_grad(xs, labels[sl], ll[sl], il[sl], V - 1, real="f64")
        assert not st.any()
        np.testing.assert_allclose(loss[sl], rl, rtol=CTC_RTOL, atol=CTC_ATOL)
        worst_loss = max(worst_loss, float(np.max(np.abs(loss[sl] - rl) / np.maximum(np.abs(rl), 1e-30))))
        assert_ctc_grad_close(r.grad[:, sl].cpu().numpy(), rg, xs, il[sl])
    return worst_loss


def test_c2_ctc_all_256_rows():
    import torch
    import bench
    from asr_dfcnn_transformer_b200 import ctc
    hb = bench.make_batch(2001)
    x, labels, ll, il = hb["logits"], hb["labels"], hb["label_len"], hb["input_len"]
    V = x.shape[2]
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1, decode=True)
    assert int(r.row_status.max()) == 0
    worst = _ctc_all_rows(x, labels, ll, il, V, r, 64)
    ref_tok, _ = build_c.greedy_decode(x, il, V - 1)
    assert ctc.tokens_to_lists(r.tokens, r.token_len) == ref_tok          # bit-exact, all rows
    record_parity("C2_ctc", {"rows": 256, "worst_loss_rel_err": worst, "tokens_bit_exact_rows": 256,
                             "rtol": CTC_RTOL, "atol": CTC_ATOL})


def test_c3_ctc_all_64_rows():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(3001)
    V = synth.VOCAB_DICT_TXT
    il = np.full(64, 1998, dtype=np.int32)
    il[5], il[17] = 1500, 1001
    x, labels, ll, il = synth.ctc_batch(rng, il, V, 280, 320)
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1, decode=True)
    assert int(r.row_status.max()) == 0
    worst = _ctc_all_rows(x, labels, ll, il, V, r, 8)
    ref_tok, _ = build_c.greedy_decode(x, il, V - 1)
    assert ctc.tokens_to_lists(r.tokens, r.token_len) == ref_tok
    record_parity("C3_ctc", {"rows": 64, "T": 1998, "worst_loss_rel_err": worst, "tokens_bit_exact_rows": 64,
                             "rtol": CTC_RTOL, "atol": CTC_ATOL})


def _c4_batch(n_utt, seed=4000):
    rng = np.random.default_rng(seed)
    lens = synth.ragged_lengths(rng, n_utt, 3.0, 7.0)
    sigs, noises, dbs = [], [], []
    for i, n in enumerate(lens):
        n = int(n)
        sigs.append((synth.g2_voiced(rng, n).astype(np.float32) / 32768.0))
        x = rng.standard_normal(n)                       # coloured noise stand-in (the generator is an input of the mix)
        nz = np.cumsum(x) if i % 2 else x
        nz = nz - nz.mean()
        noises.append((nz / nz.max()).astype(np.float32))
        dbs.append(int(rng.integers(5, 11)))
    return sigs, noises, dbs


def _c4_ref(args):
    s, nz, db = args
    K = fbank_ref.snr2k(s, nz, db)                       # the oracle's own gain (noise.py:48-52, float32)
    mixed = fbank_ref.mix_noise(s, nz, db)
    raw = fbank_ref.compute_fbank_unnormalised(mixed)
    return np.float32(K), raw, fbank_ref.zscore_columns(raw)


def test_c4_noise_augmented_batch_full_size():
    """BASELINE.json configs[3] at full size: 512 utterances of U(3,7) s, float32 signal + noise at 5..10 dB,
    gains on the device, mix fused into the frame load; EVERY utterance against the oracle of the mixed signal
    (oracle gain = fbank_ref.snr2k) at 1 x FEATURE_TOL."""
    from asr_dfcnn_transformer_b200 import features, noise
    sigs, noises, dbs = _c4_batch(512)
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        refs = pool.map(_c4_ref, list(zip(sigs, noises, dbs)), chunksize=4)
    fb = features.compute_features(sigs, noises=noises, snr_db=dbs)
    out = fb.features.cpu().numpy()
    plain, wide, needed = _features_all_rows(sigs, out, fb.frame_offsets, [(r[1], r[2]) for r in refs])
    record_parity("C4_noise_features", {"utterances": 512, "worst_err_unwidened": plain, "worst_err_widened": wide,
                                        "columns_needing_allowance": needed, "tolerance": FEATURE_TOL})
    assert plain <= FEATURE_TOL, (plain, wide, needed)
    # the device gains are the oracle's, bit for bit
    for i in range(0, 512, 37):
        assert noise.SNR2K(sigs[i], noises[i], dbs[i]) == refs[i][0], i


def test_long_utterances_and_hazard_lengths_zscored():
    """10 s and 20 s utterances (998 / 1998 frames: the audio of C1 and C3) and the float-hazard lengths of
    the frame-count expression beyond 16240, through the z-scored ``fbank`` mode."""
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(31)
    sigs = [synth.g2_voiced(rng, 160000), synth.g1_white(rng, 160000), synth.g2_voiced(rng, 320000),
            synth.g1_white(rng, 320000)]
    sigs += [synth.g2_voiced(rng, n) for n in (64240, 64880, 65520, 129040)]
    refs = _ref_features_many(sigs)
    fb = features.compute_features(sigs, mode="fbank")
    assert fb.n_frames[:4].tolist() == [998, 998, 1998, 1998]
    out = fb.features.cpu().numpy()
    plain, wide, needed = _features_all_rows(sigs, out, fb.frame_offsets, refs)
    record_parity("long_and_hazard_features", {"utterances": len(sigs), "worst_err_unwidened": plain,
                                               "worst_err_widened": wide, "columns_needing_allowance": needed})
    assert plain <= FEATURE_TOL, (plain, wide, needed)
    # C1 shape through the loader's padded layout: 8 x 10 s
    c1 = [synth.g2_voiced(rng, 160000) for _ in range(8)]
    fb = features.compute_features(c1, mode="fbank", padded_rows=1600)
    out = fb.features.cpu().numpy()
    for i, (raw, z) in enumerate(_ref_features_many(c1)):
        a, b, c = zscore_err_report(out[i, :998], z, raw)
        assert a <= FEATURE_TOL and not out[i, 998:].any()
