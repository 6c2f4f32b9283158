"""asr_dfcnn_transformer_b200 -- B200-native hot path of ASR_DFCNN_Transformer.

Spectrogram features (+ fused noise mix), CTC loss/gradient and greedy CTC decode
as hand-written sm_100a CUDA kernels behind a C ABI (include/asrk.h), exposed
through the reference's own Python call surface.  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["features", "wav_util", "noise", "ctc", "data_loader", "pipeline", "utils"]
__version__ = "0.1.0"
