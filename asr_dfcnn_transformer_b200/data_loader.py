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
import math
import os

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


# ----------------------------------------------------------------------------------------------
# The reference's loader class (lm_and_am/data_loader.py:19-280), same constructor and methods
# ----------------------------------------------------------------------------------------------
class Const:
    """The constants of util/const.py the loader reads (const.py:32-78).  The two corpus roots are
    per-host settings in the reference (edited in source through ``ServerId``); set them here or
    pass ``speech_data_path`` / ``noise_out_path`` to ``DataLoader``."""
    PAD, SOS, EOS = 0, 1, 2
    PAD_FLAG, SOS_FLAG, EOS_FLAG = "<pad>", "<sos>", "</sos>"
    SpeechDataPath = "../../../speech_data/"
    NoiseOutPath = "/usr/corpus/noise_data/"


def _sf_read(path):
    """``soundfile.read(path)`` (data_loader.py:123,125): float64 in [-1, 1] and the sample rate.
    soundfile is optional: 16-bit PCM wavs are read with the wave module and scaled by 1/32768,
    which is exactly what soundfile returns for them."""
    try:
        import soundfile as sf
    except ImportError:
        sig, fs = wav_util.read_wav_data(path)
        if sig.shape[0] != 1:
            return sig.T.astype(np.float64) / 32768.0, fs
        return sig[0].astype(np.float64) / 32768.0, fs
    return sf.read(path)


class DataLoader:
    """Drop-in for ``lm_and_am.data_loader.DataLoader`` on the acoustic-model path:
    ``DataLoader(data_util, data_args, train_args)`` with ``data_generation``, ``__getitem__``,
    ``__len__``, ``am_generator``, ``end2end_generator`` and ``get_fbank_and_pinyin_data``
    returning what the reference returns (numpy, same shapes / dtypes / error behaviour).  The
    per-utterance Python feature loop of data_loader.py:117-151 becomes ONE batched GPU call per
    batch (the mel front end ``compute_fbank_from_api`` of :129).  ``data_util`` is any object with
    ``path_lst``, ``pny_lst``, ``han_lst`` and ``shuffle``; ``data_args`` / ``train_args`` any
    namespace-like objects with the attributes read below (util/hparams.py).

    ``output="device"`` (an addition) keeps the padded features on the GPU as a float32 torch
    tensor instead of copying a float64 array back to the host."""

    def __init__(self, data_util, data_args, train_args, speech_data_path=None, noise_out_path=None,
                 output="numpy", device=None):
        self.am_batch_size = train_args.am_batch_size
        self.lm_batch_size = getattr(train_args, "lm_batch_size", None)
        self.feature_dim = train_args.feature_dim
        self.feature_max_length = train_args.feature_max_length
        self.pinyin_dict = data_args.pinyin_dict
        self.hanzi_dict = data_args.hanzi_dict
        self.lfr_m = getattr(data_args, "lfr_m", 4)
        self.lfr_n = getattr(data_args, "lfr_n", 3)
        self.speech_data_path = Const.SpeechDataPath if speech_data_path is None else speech_data_path
        self.noise_out_path = Const.NoiseOutPath if noise_out_path is None else noise_out_path
        if output not in ("numpy", "device"):
            raise ValueError("output must be 'numpy' or 'device'")
        self.output = output
        self.device = device
        self.acoustic_vocab_size, self.pinyin2index, self.index2pinyin = self.get_acoustic_vocab_list()
        self.language_vocab_size, self.word2index, self.index2word = self.get_language_vocab_list()
        self.data = data_util
        self.path_lst = self.data.path_lst
        self.pny_lst = self.data.pny_lst
        self.han_lst = self.data.han_lst
        self.shuffle = data_util.shuffle
        self.indexes = [i for i in range(len(self.path_lst))]

    # ---- vocabularies (data_loader.py:44-103) ------------------------------------------------
    def pny2id(self, line):
        return pny2id(line, self.pinyin2index)

    def han2id(self, line):
        """data_loader.py:62-82: one id per character, the three flags map to PAD/SOS/EOS ids; any
        failure becomes ValueError."""
        try:
            line = line.strip()
            res = []
            for han in line:
                if han == Const.PAD_FLAG:
                    res.append(Const.PAD)
                elif han == Const.SOS_FLAG:
                    res.append(Const.SOS)
                elif han == Const.EOS_FLAG:
                    res.append(Const.EOS)
                else:
                    res.append(self.word2index[han])
            return res
        except Exception:
            raise ValueError("unknown character in %r" % (line,))

    def get_acoustic_vocab_list(self):
        """data_loader.py:85-92 (pandas ``read_table``: first column, '_' appended last)."""
        import pandas as pd
        text = pd.read_table(os.path.join(self.pinyin_dict), header=None)
        symbol_list = text.iloc[:, 0].tolist()
        symbol_list.append("_")
        pinyin2index = dict([pinyin, index] for index, pinyin in enumerate(symbol_list))
        index2pinyin = dict([index, pinyin] for index, pinyin in enumerate(symbol_list))
        return len(symbol_list), pinyin2index, index2pinyin

    def get_language_vocab_list(self):
        """data_loader.py:95-103 (one character per line; '<pad>' is index 0)."""
        import pandas as pd
        pd_data = pd.read_csv(os.path.join(os.getcwd(), self.hanzi_dict), header=None)
        hanzi_list = pd_data.T.values.tolist()[0]
        word_list = [Const.PAD_FLAG]
        word_list.extend(hanzi_list)
        word2index = dict([word, index] for index, word in enumerate(word_list))
        index2word = dict([index, word] for index, word in enumerate(word_list))
        return len(word_list), word2index, index2word

    # ---- batch assembly (data_loader.py:105-162) ---------------------------------------------
    def data_generation(self, batch_datas, py_label_datas, han_label_datas):
        """Returns the 6-tuple ``(batch_wav_data [B',feature_max_length,200,1], input_length [B'],
        batch_label_data [B',64] int32, label_length [B'], batch_han_data [B',64] int32,
        word_length [B'])``; a missing file prints "file path Error" and returns 0 (:126-128);
        rows that raise ValueError in the reference are deleted (:149-156)."""
        B = self.am_batch_size
        batch_label_data = np.zeros((B, 64), dtype=np.int32)
        batch_han_data = np.zeros((B, 64), dtype=np.int32)
        input_length, label_length, word_length, error_count = [], [], [], []
        kept, kept_sigs, kept_fs = [], [], []
        for i, path in enumerate(batch_datas):
            try:
                file1 = os.path.join(self.speech_data_path, path)
                file2 = os.path.join(self.noise_out_path, path)
                if os.path.isfile(file1):
                    signal, sample_rate = _sf_read(file1)
                elif os.path.isfile(file2):
                    signal, sample_rate = _sf_read(file2)
                else:
                    print("file path Error")
                    return 0
                signal = np.asarray(signal)
                if signal.ndim != 1:
                    raise ValueError("mono audio expected")
                # frames of compute_fbank_from_api (:129) without computing them yet
                wav_length = wav_util.logfbank_frames(len(signal), wav_util._round_half_up(0.025 * sample_rate),
                                                      wav_util._round_half_up(0.01 * sample_rate))
                data_length = min(T_CTC_CAP, math.ceil(wav_length // 8 + 1))          # :132
                seq_ids = np.array(self.han2id(han_label_datas[i]))                   # :133-134
                py_label_ids = np.array(self.pny2id(py_label_datas[i]))               # :135-136
                len_label = len(py_label_ids)
                if wav_length > self.feature_max_length:                              # :139
                    raise ValueError
                if len_label > 64 or len_label >= data_length:                        # :141
                    raise ValueError
                # :143-148 -- the lengths are appended BEFORE the row assignments, so a hanzi line longer
                # than 64 (numpy broadcast ValueError at :148) drops the row but keeps its lengths: kept as is
                input_length.append(data_length)
                label_length.append(len_label)
                word_length.append(len_label)
                batch_label_data[i, 0:len(py_label_ids)] = py_label_ids
                batch_han_data[i, 0:len(seq_ids)] = seq_ids
                kept.append(i)
                kept_sigs.append(signal)
                kept_fs.append(sample_rate)
            except ValueError:
                error_count.append(i)
                continue
        wav = self._features(kept, kept_sigs, kept_fs, B, error_count)
        if error_count != []:
            batch_label_data = np.delete(batch_label_data, error_count, axis=0)
            batch_han_data = np.delete(batch_han_data, error_count, axis=0)
        return (wav, np.array(input_length), batch_label_data, np.array(label_length), batch_han_data,
                np.array(word_length))

    def _features(self, kept, sigs, rates, B, error_count):
        """Zero-padded ``[B - len(error_count), feature_max_length, 200, 1]`` features: one batched GPU
        call per sample rate present in the batch (normally one)."""
        from . import _lib
        torch = _lib.require_cuda()
        rows = B - len(error_count)
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        out = torch.zeros((rows, self.feature_max_length, 200), dtype=torch.float32, device=dev)
        # position of batch row i after the deletion of the error rows (np.delete keeps the order)
        errs = set(error_count)
        pos, nxt = {}, 0
        for i in range(B):
            if i in errs:
                continue
            pos[i] = nxt
            nxt += 1
        for fs in sorted(set(rates)):
            idx = [k for k, r in enumerate(rates) if r == fs]
            fb = wav_util.compute_fbank_from_api_batch([sigs[k] for k in idx], fs, 200, device=dev,
                                                       padded_rows=self.feature_max_length)
            dst = torch.as_tensor([pos[kept[k]] for k in idx], dtype=torch.long, device=dev)
            out.index_copy_(0, dst, fb.features)
        out = out.reshape(rows, self.feature_max_length, 200, 1)
        if self.output == "device":
            return out
        return out.cpu().numpy().astype(np.float64)

    # ---- one utterance (data_loader.py:213-244) ----------------------------------------------
    def get_fbank_and_pinyin_data(self, index):
        """``(wav_data [1,feature_max_length,200,1], data_length [1], label, len_label)``;
        ``data_length = n_frames // 8 + 1`` without the cap of 200 (:231) and the label test is a
        strict '>' (:238).  An over-long utterance raises ValueError (the reference fails at the slice
        assignment of :230; its ``shape[0]`` test at :236 looks at the batch axis and never fires)."""
        try:
            file = os.path.join(self.speech_data_path, self.path_lst[index])
            noise_file = os.path.join(self.noise_out_path, self.path_lst[index])
            fbank = wav_util.compute_fbank_from_file(file) if os.path.exists(file) else \
                wav_util.compute_fbank_from_file(noise_file)
            wav_data = np.zeros((1, self.feature_max_length, 200, 1), dtype=np.float64)
            input_data = fbank.reshape([fbank.shape[0], fbank.shape[1], 1])
            wav_data[0, 0:len(input_data)] = input_data          # ValueError when longer than the buffer
            data_length = input_data.shape[0] // 8 + 1
            label = np.array(self.pny2id(self.pny_lst[index]))
            len_label = len(label)
            if wav_data.shape[0] > self.feature_max_length:
                raise ValueError
            if len_label > 64 or len_label > data_length:
                raise ValueError
            return wav_data, np.array([data_length]), label, len_label
        except ValueError:
            raise ValueError

    # ---- generators (data_loader.py:246-280) -------------------------------------------------
    def am_generator(self):
        for i in range(len(self)):
            yield self.__getitem__(i)

    def end2end_generator(self):
        for i in range(len(self)):
            yield self.__getitem__(i)

    def __getitem__(self, index):
        batch_indexs = self.indexes[index * self.am_batch_size:(index + 1) * self.am_batch_size]
        batch_datas = [self.path_lst[k] for k in batch_indexs]
        py_label_datas = [self.pny_lst[k] for k in batch_indexs]
        han_label_datas = [self.han_lst[k] for k in batch_indexs]
        return self.data_generation(batch_datas, py_label_datas, han_label_datas)

    def __len__(self):
        return len(self.path_lst) // self.am_batch_size
