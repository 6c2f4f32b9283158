"""Host-side batch rules of the reference's loader, on top of the GPU feature kernel.

Mirrors /root/reference/lm_and_am/data_loader.py:
  * vocabulary      get_acoustic_vocab_list (:63-71): first column of the dict file,
                    blank '_' appended LAST; duplicate symbols map to the LATER index
  * pny2id          (:44-60): any failure becomes ValueError (the row is dropped)
  * data_generation (:105-162): zero-padded ``(B, 1600, 200, 1)`` features,
                    ``(B, 64)`` int32 labels, ``input_length = min(200, n_frames//8+1)``
                    (:132), reject ``n_frames > 1600`` (:139), ``len_label > 64 or
                    len_label >= input_length`` (:141), rejected rows deleted (:153-156)
  * get_fbank_and_pinyin_data (:213-244): ``input_length = n_frames//8+1`` (no cap),
                    reject ``len_label > input_length`` (strict '>')
The reference reads files one by one on the host; here the features of the whole
batch are computed in one GPU pass and stay on the device: by default with the mel
front end the live loader calls (``compute_fbank_from_api``, data_loader.py:129), or
with the in-repo FFT spectrogram (``front_end="spectrogram"``).
Pure control logic stays on the host, as in the reference.
"""
import numpy as np

from . import features, wav_util

FEATURE_MAX_LENGTH = 1600     # util/hparams.py feature_max_length
LABEL_MAX = 64                # data_loader.py:109,141
T_CTC_CAP = 200               # data_loader.py:132


def load_acoustic_vocab(dict_path):
    """data_loader.py:63-71.  Returns (size, symbol->index, index->symbol)."""
    symbols = []
    with open(dict_path, encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if line == "":
                continue
            symbols.append(line.split("\t")[0])
    symbols.append("_")
    sym2idx = {}
    for i, s in enumerate(symbols):          # dict([...]) keeps the LAST index of a duplicate
        sym2idx[s] = i
    idx2sym = dict(enumerate(symbols))
    return len(symbols), sym2idx, idx2sym


def pny2id(line, sym2idx):
    """data_loader.py:44-60."""
    try:
        return [sym2idx[p] for p in line.strip().split(" ")]
    except Exception:
        raise ValueError("unknown symbol in %r" % (line,))


def ctc_input_length(n_frames, capped=True):
    """data_loader.py:132 (capped) / :231 (uncapped)."""
    t = n_frames // 8 + 1
    return min(T_CTC_CAP, t) if capped else t


def data_generation(signals, py_labels, sym2idx, fs=16000, batch_size=None, device=None, front_end="logfbank"):
    """Batch assembly of data_loader.py:105-162 for in-memory utterances.

    signals: list of 1-D int16 / float32 arrays; py_labels: list of pinyin strings.
    Returns (wav [B',1600,200,1] float32 device tensor, input_length [B'] int64,
    labels [B',64] int32, label_length [B'] int64, kept) where kept lists the indices
    of the rows that survived the reference's reject rules, in order."""
    B = len(signals) if batch_size is None else batch_size
    if len(signals) != len(py_labels):
        raise ValueError("signals / labels length mismatch")
    keep, labels, in_len, lab_len = [], [], [], []
    for i, (sig, py) in enumerate(zip(signals, py_labels)):
        try:
            if front_end == "logfbank":
                n = wav_util.logfbank_frames(len(sig), wav_util._round_half_up(0.025 * fs),
                                             wav_util._round_half_up(0.01 * fs))
            else:
                n = features.n_frames_for(len(sig), fs, "fbank")      # the reference's float expression
            if n < 1:
                raise ValueError
            ids = pny2id(py, sym2idx)
            data_length = ctc_input_length(n, capped=True)
            if n > FEATURE_MAX_LENGTH:
                raise ValueError
            if len(ids) > LABEL_MAX or len(ids) >= data_length:
                raise ValueError
        except ValueError:
            continue                                                  # data_loader.py:150-152
        keep.append(i)
        labels.append(ids)
        in_len.append(data_length)
        lab_len.append(len(ids))
    batch_label = np.zeros((len(keep), LABEL_MAX), dtype=np.int32)
    for r, ids in enumerate(labels):
        batch_label[r, :len(ids)] = ids
    if keep:
        if front_end == "logfbank":
            fb = wav_util.compute_fbank_from_api_batch([signals[i] for i in keep], fs, 200, device=device,
                                                       padded_rows=FEATURE_MAX_LENGTH)
        else:
            fb = features.compute_features([signals[i] for i in keep], fs=fs, mode="fbank", device=device,
                                           padded_rows=FEATURE_MAX_LENGTH)
        wav = fb.features.reshape(len(keep), FEATURE_MAX_LENGTH, 200, 1)
    else:
        wav = None
    del B
    return wav, np.array(in_len, dtype=np.int64), batch_label, np.array(lab_len, dtype=np.int64), keep
