"""The reference's loader class surface (lm_and_am/data_loader.py:19-280) against the fixture produced by
the reference's OWN DataLoader (tests/golden/loader_case.npz, tools/make_golden_loader.py).

CPU: the control logic (vocabularies, lengths, label / hanzi ids, reject rules, row deletion, the
lengths-kept-for-a-dropped-row quirk of :143-148, the uncapped / strict rules of :231,:238) with the
feature call replaced by the oracle.  GPU: the real thing, features from the CUDA kernels."""
import numpy as np
import pytest

from tests.loader_case import build_corpus, check_loader


class _OracleFeatureLoader:
    """mixin: features from oracle/psf_ref.py instead of the GPU (CPU test only)."""

    def _features(self, kept, sigs, rates, B, error_count):
        from oracle import psf_ref
        rows = B - len(error_count)
        out = np.zeros((rows, self.feature_max_length, 200, 1), dtype=np.float64)
        errs = set(error_count)
        pos = [i for i in range(B) if i not in errs]
        for k, i in enumerate(kept):
            fb = psf_ref.compute_fbank_from_api(sigs[k], rates[k])
            out[pos.index(i), :fb.shape[0], :, 0] = fb
        return out


def test_loader_class_control_logic_cpu(golden_dir, tmp_path, monkeypatch):
    from asr_dfcnn_transformer_b200 import data_loader as dl, wav_util
    from oracle import psf_ref
    g, d, data_util, data_args, train_args = build_corpus(golden_dir, tmp_path)

    class Loader(_OracleFeatureLoader, dl.DataLoader):
        pass

    def api(signal, sample_rate, nfilt=200):       # compute_fbank_from_api on the oracle (2-D input included)
        return psf_ref.compute_fbank_from_api(np.asarray(signal), sample_rate, nfilt)
    monkeypatch.setattr(wav_util, "compute_fbank_from_api", api)
    loader = Loader(data_util, data_args, train_args, speech_data_path=d, noise_out_path=d + "/noise")
    check_loader(loader, g, feature_tol=1e-6)
    # missing file: prints and returns 0 (data_loader.py:126-128)
    assert loader.data_generation(["nope.wav"], ["a1"], ["阿"]) == 0
    # the noise directory is the fall-back root (:121-125)
    import os
    import shutil
    os.makedirs(d + "/noise")
    shutil.move(d + "/utt0.wav", d + "/noise/utt0.wav")
    wav, il, lab, ll, han, wl = loader[0]
    assert np.array_equal(il, g["b0_input_length"]) and wav.shape == g["b0_wav"].shape


@pytest.mark.gpu
def test_loader_class_gpu(golden_dir, tmp_path):
    from asr_dfcnn_transformer_b200 import data_loader as dl
    g, d, data_util, data_args, train_args = build_corpus(golden_dir, tmp_path)
    loader = dl.DataLoader(data_util, data_args, train_args, speech_data_path=d, noise_out_path=d + "/noise")
    worst = check_loader(loader, g)
    print("loader class: worst feature error vs the reference's DataLoader output", worst)
    assert loader.data_generation(["nope.wav"], ["a1"], ["阿"]) == 0
    # device output (addition): same numbers, float32 torch tensor on the GPU
    dev_loader = dl.DataLoader(data_util, data_args, train_args, speech_data_path=d, noise_out_path=d + "/noise",
                               output="device")
    wav, il, lab, ll, han, wl = dev_loader[0]
    assert wav.is_cuda and tuple(wav.shape) == g["b0_wav"].shape
    ref, _, _, _, _, _ = loader[0]
    assert np.allclose(wav.cpu().numpy(), ref, atol=1e-6)
