"""ctypes binding of libasrk.so (include/asrk.h).

There is deliberately no fallback: if the CUDA library is missing or fails to
load, every product entry point raises.  ``torch`` is used only as the carrier of
device memory and streams.
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libasrk%s.so" % os.environ.get("ASRK_LIB_SUFFIX", ""))

OK = 0
E_BADARG, E_SHAPE, E_ALIGN, E_WORKSPACE, E_CUDA = -1, -2, -3, -4, -5
ROW_OK, ROW_INFEASIBLE, ROW_NOT_ENOUGH_TIME, ROW_BAD_LENGTH, ROW_NOT_SMALL = 0, 1, 2, 3, 4
SPEC_FBANK, SPEC_ASRT, SPEC_FBANK_RAW = 0, 1, 2
DTYPE_I16, DTYPE_F32 = 0, 1
LABELS_BY_LENGTH, LABELS_DROP_ZEROS = 0, 1
PHASE_ALL = 0xffff
CTC_SMALL_ONLY = 0x10000
CTC_INPUT_PROB = 0x20000
PHASE_SPEC_SETUP, PHASE_SPEC_MAIN, PHASE_SPEC_NORMALIZE, PHASE_SPEC_STATS = 1, 2, 4, 8
PHASE_CTC_PREP, PHASE_CTC_ROWS, PHASE_CTC_LATTICE, PHASE_CTC_GRAD, PHASE_CTC_COLLAPSE = 1, 2, 4, 8, 16
PHASE_CTC_FUSED = 32

# every symbol include/asrk.h declares: (restype, argtypes)
_vp, _i, _ll, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t
SIGNATURES = {
    "asrk_version": (_i, []),
    "asrk_error_string": (ctypes.c_char_p, [_i]),
    "asrk_launch_count": (ctypes.c_ulonglong, []),
    "asrk_spectrogram_workspace_bytes": (_sz, [_i, _ll]),
    "asrk_spectrogram_run": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp, _vp, _sz,
                                  _vp]),
    "asrk_snr2k_run": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "asrk_ctc_workspace_bytes": (_sz, [_i, _i, _i]),
    "asrk_ctc_fits_fused": (_i, [_i, _i]),
    "asrk_ctc_loss_grad_run": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp,
                                    _ll, _ll, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "asrk_spectrogram_run_phases": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp, _vp,
                                         _sz, _vp, _i]),
    "asrk_ctc_loss_grad_run_phases": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp,
                                           _vp, _ll, _ll, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _i]),
    "asrk_ctc_loss_grad_zscore_run": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp,
                                           _vp, _ll, _ll, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _i,
                                           _vp, _vp, _vp, _vp, _i, _ll, _vp]),
    "asrk_spectrogram_zscore_handles": (_i, [_vp, _sz, _i, _ll, ctypes.POINTER(ctypes.c_void_p),
                                             ctypes.POINTER(ctypes.c_void_p)]),
    "asrk_ctc_batch_cost_run": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _ll, _ll, _vp, _vp,
                                     _sz, _vp, _i]),
    "asrk_ctc_stage_logits_run": (_i, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _i, _i, _i, _vp]),
    "asrk_ctc_loss_sum_run": (_i, [_vp, _vp, _i, _vp, _vp]),
    "asrk_ctc_loss_sum_acc_run": (_i, [_vp, _vp, _i, _vp, _vp]),
    "asrk_ctc_unstage_rows_run": (_i, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _vp, _i, _i, _i, _vp]),
    "asrk_color_noise_workspace_bytes": (_sz, [_i, _ll]),
    "asrk_color_noise_run": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _sz, _vp]),
    "asrk_logfbank_run": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _i, ctypes.c_double, _i, _vp, _vp]),
    "asrk_lfr_run": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _vp]),
    "asrk_edit_distance_run": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "asrk_ctc_decode_workspace_bytes": (_sz, [_i, _i]),
    "asrk_ctc_greedy_decode_run": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _vp,
                                        _sz, _vp]),
}


class AsrkError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = "asrk status %d" % code
        try:
            msg = lib().asrk_error_string(code).decode()
        except Exception:
            pass
        super().__init__("%s: %s (%d)" % (where, msg, code))


_lib = None


def lib():
    """Load libasrk.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libasrk.so is not built (%s). Run `python -m asr_dfcnn_transformer_b200._build`; "
            "there is no CPU fallback." % LIB_PATH)
    try:
        import torch  # noqa: F401  (loads libcudart.so.12 so the library resolves it)
    except Exception:
        pass
    handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(code, where):
    if code != OK:
        raise AsrkError(code, where)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("asr_dfcnn_transformer_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return torch


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)
