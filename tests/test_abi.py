"""CPU: the C-ABI library loads and exports every symbol include/asrk.h declares
(no compute calls without a GPU), argument validation that needs no device, and
the FFT codelets on the host."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "asrk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asrk_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from asr_dfcnn_transformer_b200 import _build, _lib
    path = _build.build()
    assert os.path.isfile(path)
    names = _header_functions()
    assert len(names) >= 11
    handle = ctypes.CDLL(path)
    for n in names:
        assert hasattr(handle, n), n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table covers the header
    L = _lib.lib()
    assert L.asrk_version() >= 100
    assert b"workspace" in L.asrk_error_string(_lib.E_WORKSPACE)


def test_cuobjdump_shows_sm100a_only():
    from asr_dfcnn_transformer_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation_without_device():
    from asr_dfcnn_transformer_b200 import _lib
    L = _lib.lib()
    assert L.asrk_spectrogram_workspace_bytes(256, 130000) > 0
    assert L.asrk_ctc_workspace_bytes(88, 256, 64) > 0
    assert L.asrk_ctc_decode_workspace_bytes(88, 256) > 0
    # batch == 0 is a no-op; bad enums / null pointers are rejected before any CUDA call
    assert L.asrk_spectrogram_run(None, 0, None, None, None, None, None, None, None, 0, 0, 0, None, None, 0,
                                  None) == _lib.OK
    assert L.asrk_spectrogram_run(None, 0, None, None, None, None, None, None, None, 4, 10, 0, None, None, 0,
                                  None) == _lib.E_BADARG
    assert L.asrk_ctc_loss_grad_run(None, 0, 0, 5, 2, 10, None, 4, None, None, 9, 0, None, None, None, 0, 0,
                                    None, None, 0, None, None, None, 0, None) == _lib.E_BADARG
    assert L.asrk_ctc_greedy_decode_run(None, 0, 0, 5, 0, 10, None, 9, 1, None, 0, None, None, None, 0,
                                        None) == _lib.OK


@pytest.mark.parametrize("harness", ["fft_codelets_host.cpp", "fft_lanes_host.cpp"])
def test_fft_codelets_on_host(tmp_path, harness):
    """Both pass-2 formulations (role-uniform warps / lane-uniform selects) with the generated tables."""
    exe = str(tmp_path / "fft_host")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", harness)])
    rng = np.random.default_rng(1)
    w = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(400) / 399)
    for trial in range(3):
        x = rng.integers(-32768, 32767, 400).astype(float)
        if trial == 2:
            x = np.round(np.sin(2 * np.pi * 1000 * np.arange(400) / 16000) * 30000)
        out = subprocess.run([exe], input="\n".join(repr(float(v)) for v in x), capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        p4 = np.array([float(s) for s in out.stdout.split()])
        ref = np.abs(np.fft.fft(x * w)[:200])
        assert np.abs(np.sqrt(p4 / 4) - ref).max() <= 1e-12 * ref.max()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "asr_dfcnn_transformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_python_constants_match_the_header():
    """Every numeric ASRK_* #define of include/asrk.h that the Python side mirrors has the same value there."""
    import re
    from asr_dfcnn_transformer_b200 import _lib
    text = open(os.path.join(ROOT, "include", "asrk.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+ASRK_(\w+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))\)?", text)}
    mirrored = 0
    for name, value in defs.items():
        for cand in (name, name.replace("PHASE_CTC_", "PHASE_CTC_"), name.replace("SPEC_", "MODE_")):
            if hasattr(_lib, cand):
                assert getattr(_lib, cand) == value, (name, getattr(_lib, cand), value)
                mirrored += 1
                break
    assert mirrored >= 15, mirrored
    assert _lib.CTC_SMALL_ONLY == defs["CTC_SMALL_ONLY"] and _lib.ROW_NOT_SMALL == defs["ROW_NOT_SMALL"]


def test_dft16_and_the_16x16_split_on_host(tmp_path):
    """The mel front end's 256-point complex transform: two passes of the dft16 codelet with the W256
    twiddle between them, index maps exactly as in csrc/logfbank.cu, against numpy."""
    exe = str(tmp_path / "fft16")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "fft16x16_host.cpp")])
    rng = np.random.default_rng(16)
    for trial in range(3):
        z = rng.standard_normal(256) + 1j * rng.standard_normal(256)
        if trial == 2:
            z = np.exp(2j * np.pi * 37 * np.arange(256) / 256) * 1e3 + 1e-6 * z
        text = "\n".join("%r %r" % (float(v.real), float(v.imag)) for v in z)
        out = subprocess.run([exe], input=text, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = np.array([float(s) for s in out.stdout.split()]).reshape(256, 2)
        ref = np.fft.fft(z)
        assert np.abs(got[:, 0] + 1j * got[:, 1] - ref).max() <= 1e-12 * np.abs(ref).max()
