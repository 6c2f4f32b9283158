"""Shared by the CPU and GPU loader tests: rebuild the corpus of tests/golden/loader_case.npz (produced by
the reference's own DataLoader, tools/make_golden_loader.py) in a temp dir and compare a loader against
everything the reference returned."""
import os
import types

import numpy as np
import scipy.io.wavfile as wavfile

from tests.util import FEATURE_TOL


def build_corpus(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "loader_case.npz"))
    d = str(tmp_path)
    offs = np.concatenate([[0], np.cumsum(g["pcm_len"])])
    names = []
    for i in range(len(g["pcm_len"])):
        name = "utt%d.wav" % i
        wavfile.write(os.path.join(d, name), 16000, g["pcm"][offs[i]:offs[i + 1]])
        names.append(name)
    dict_path = os.path.join(d, "dict.txt")
    open(dict_path, "w", encoding="utf-8").write("\n".join(g["dict_lines"].tolist()) + "\n")
    hanzi_path = os.path.join(d, "hanzi.txt")
    open(hanzi_path, "w", encoding="utf-8").write("\n".join(g["hanzi_lines"].tolist()) + "\n")
    data_util = types.SimpleNamespace(path_lst=np.array(names), pny_lst=g["pny"], han_lst=g["han"], shuffle=False)
    data_args = types.SimpleNamespace(pinyin_dict=dict_path, hanzi_dict=hanzi_path, lfr_m=4, lfr_n=3)
    train_args = types.SimpleNamespace(am_batch_size=int(g["am_batch_size"]), lm_batch_size=2, feature_dim=200,
                                       feature_max_length=int(g["feature_max_length"]))
    return g, d, data_util, data_args, train_args


def _live_columns(wav):
    """Columns of the mel features that are not one of the 43 empty filters (constant columns: the
    reference holds zeros or float64 rounding noise there, we return exact zeros: INTEGRATION.md)."""
    from oracle import psf_ref
    return psf_ref.get_filterbanks(200, 512, 16000).sum(axis=1) > 0


def check_loader(loader, g, feature_tol=FEATURE_TOL):
    assert loader.acoustic_vocab_size == int(g["acoustic_vocab_size"])
    assert loader.language_vocab_size == int(g["language_vocab_size"])
    assert [str(k) for k in loader.pinyin2index.keys()] == [str(k) for k in g["pinyin_keys"].tolist()]
    assert list(loader.pinyin2index.values()) == g["pinyin_vals"].tolist()
    assert len(loader) == int(g["n_batches"])
    worst = 0.0
    for b, item in enumerate(loader.am_generator()):
        wav, il, lab, ll, han, wl = item
        ref = g["b%d_wav" % b]
        assert isinstance(wav, np.ndarray) and wav.dtype == np.float64 and wav.shape == ref.shape, (wav.shape, ref.shape)
        assert np.array_equal(il, g["b%d_input_length" % b])
        assert lab.dtype == np.int32 and np.array_equal(lab, g["b%d_label" % b])
        assert np.array_equal(ll, g["b%d_label_length" % b])
        assert han.dtype == np.int32 and np.array_equal(han, g["b%d_han" % b])
        assert np.array_equal(wl, g["b%d_word_length" % b])
        live = _live_columns(ref)
        err = np.abs(wav[..., 0][:, :, live] - ref[..., 0][:, :, live]) / np.maximum(np.abs(ref[..., 0][:, :, live]), 1.0)
        worst = max(worst, float(err.max()))
        pad = np.abs(ref[..., 0]).max(axis=2) == 0            # rows past the utterance: zero padding on both sides
        assert not wav[..., 0][pad].any()
    assert worst <= feature_tol, worst
    for i, ok in enumerate(g["u_ok"].tolist()):
        if not ok:
            try:
                loader.get_fbank_and_pinyin_data(i)
            except ValueError:
                continue
            raise AssertionError("utterance %d: the reference raises ValueError" % i)
        wd, dl, label, len_label = loader.get_fbank_and_pinyin_data(i)
        assert wd.shape == (1, loader.feature_max_length, 200, 1) and wd.dtype == np.float64
        assert np.array_equal(dl, g["u%d_data_length" % i]) and np.array_equal(label, g["u%d_label" % i])
        assert len_label == int(g["u%d_len_label" % i])
        if ("u%d_wav" % i) in g.files:
            ref = g["u%d_wav" % i]
            live = _live_columns(ref)
            err = np.abs(wd[0, :, live, 0] - ref[0, :, live, 0]) / np.maximum(np.abs(ref[0, :, live, 0]), 1.0)
            worst = max(worst, float(err.max()))
            assert err.max() <= feature_tol, (i, float(err.max()))
    return worst
