"""Generate tests/golden/loader_case.npz by RUNNING THE REFERENCE'S OWN ``DataLoader``
(/root/reference/lm_and_am/data_loader.py, unmodified, imported through oracle/ref_import.py).

Run in the build container only (needs /root/reference):
    python tools/make_golden_loader.py

A small temporary corpus (wav files + manifest lists + dictionaries) is fed to the reference class;
the fixture holds the corpus and everything the reference returned from ``__getitem__`` (the 6-tuple of
``data_generation``, data_loader.py:105-162) and from ``get_fbank_and_pinyin_data`` (:213-244).  What is
pinned to the reference's own code is the loader's control logic; the mel features inside come from
``python_speech_features.logfbank``, which is not installable here and is stood in for by
oracle/psf_ref.py (see ref_import.load_data_loader).
"""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.io.wavfile as wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, synth  # noqa: E402

DICT_LINES = ["a1\t阿啊", "b2\t拔", "c3\t擦", "d4\t大", "a1\t吖", "e5\t鹅", "null\t空"]   # duplicate key, pandas-NaN word
HANZI_LINES = ["阿", "拔", "擦", "大", "鹅", "空"]
FEATURE_MAX_LENGTH = 160      # small, so that the "too long" rule is exercised by a 2 s utterance
BATCH = 4


def corpus(rng):
    """(pcm list, pinyin lines, hanzi lines): two batches of four."""
    n = [9000, 16080, 4000, 33000,      # batch 0: ok, ok (float-hazard length), label >= T_ctc, too long (205 frames)
         12345, 8000, 8000, 20000]      # batch 1: unknown pinyin, unknown hanzi, hanzi line > 64 chars, ok
    pcm = [synth.g2_voiced(rng, k) if i % 2 else synth.g1_white(rng, k) for i, k in enumerate(n)]
    pny = ["a1 b2 c3", "d4 a1", "a1 b2 c3 d4", "a1",
           "a1 zz9", "b2", "c3 d4", "e5 a1 b2 c3 d4 e5"]
    han = ["阿拔擦", "大阿", "阿拔擦大", "阿",
           "阿拔", "拔龘", "擦" * 65, "鹅阿拔擦大鹅"]
    return pcm, pny, han


def main():
    mod = ref_import.load_data_loader()
    rng = np.random.default_rng(20261019)
    pcm, pny, han = corpus(rng)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        names = []
        for i, s in enumerate(pcm):
            name = "utt%d.wav" % i
            wavfile.write(os.path.join(d, name), 16000, s)
            names.append(name)
        dict_path = os.path.join(d, "dict.txt")
        open(dict_path, "w", encoding="utf-8").write("\n".join(DICT_LINES) + "\n")
        hanzi_path = os.path.join(d, "hanzi.txt")
        open(hanzi_path, "w", encoding="utf-8").write("\n".join(HANZI_LINES) + "\n")
        mod.Const.SpeechDataPath = d
        mod.Const.NoiseOutPath = os.path.join(d, "noise")
        data_util = types.SimpleNamespace(path_lst=np.array(names), pny_lst=np.array(pny), han_lst=np.array(han),
                                          shuffle=False)
        data_args = types.SimpleNamespace(pinyin_dict=dict_path, hanzi_dict=hanzi_path, lfr_m=4, lfr_n=3)
        train_args = types.SimpleNamespace(am_batch_size=BATCH, lm_batch_size=2, feature_dim=200,
                                           feature_max_length=FEATURE_MAX_LENGTH)
        loader = mod.DataLoader(data_util, data_args, train_args)
        out["acoustic_vocab_size"] = np.int64(loader.acoustic_vocab_size)
        out["language_vocab_size"] = np.int64(loader.language_vocab_size)
        out["pinyin_keys"] = np.array(list(loader.pinyin2index.keys()), dtype=object).astype(str)
        out["pinyin_vals"] = np.array(list(loader.pinyin2index.values()), dtype=np.int64)
        out["n_batches"] = np.int64(len(loader))
        for b in range(len(loader)):
            wavd, il, lab, ll, hd, wl = loader[b]
            out["b%d_wav" % b] = wavd.astype(np.float32)
            out["b%d_input_length" % b] = np.asarray(il)
            out["b%d_label" % b] = lab
            out["b%d_label_length" % b] = np.asarray(ll)
            out["b%d_han" % b] = hd
            out["b%d_word_length" % b] = np.asarray(wl)
            print("batch", b, wavd.shape, wavd.dtype, il, ll, wl, lab.shape, hd.shape)
        ok = []
        for i in range(len(names)):
            try:
                wd, dl, label, len_label = loader.get_fbank_and_pinyin_data(i)
                ok.append(1)
                out["u%d_data_length" % i] = np.asarray(dl)
                out["u%d_label" % i] = np.asarray(label)
                out["u%d_len_label" % i] = np.int64(len_label)
                if i in (0, 1):
                    out["u%d_wav" % i] = wd.astype(np.float32)
                print("utt", i, wd.shape, dl, label, len_label)
            except ValueError:
                ok.append(0)
                print("utt", i, "ValueError")
        out["u_ok"] = np.array(ok, dtype=np.int64)
    out["pcm"] = np.concatenate(pcm)
    out["pcm_len"] = np.array([len(s) for s in pcm], dtype=np.int64)
    out["pny"] = np.array(pny)
    out["han"] = np.array(han)
    out["dict_lines"] = np.array(DICT_LINES)
    out["hanzi_lines"] = np.array(HANZI_LINES)
    out["feature_max_length"] = np.int64(FEATURE_MAX_LENGTH)
    out["am_batch_size"] = np.int64(BATCH)
    path = os.path.join(ROOT, "tests", "golden", "loader_case.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
