"""Build libasrk.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m asr_dfcnn_transformer_b200._build [--force] [-v]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored and travels
to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# ASRK_LIB_SUFFIX builds / loads a variant library next to the default one (experiments: other -D switches)
SUFFIX = os.environ.get("ASRK_LIB_SUFFIX", "")
LIB_PATH = os.path.join(PKG_DIR, "libasrk%s.so" % SUFFIX)
# ASRK_SPEC_SOURCE: another implementation file of the spectrogram entry points (A/B experiments)
SOURCES = ["asrk_api.cu", os.environ.get("ASRK_SPEC_SOURCE", "spectrogram.cu"), "noise.cu", "ctc.cu", "post.cu", "logfbank.cu",
           "color_noise.cu"]
HEADERS = ["asrk_common.cuh", "asrk_fft.cuh", "asrk_tables.inc", os.path.join("..", "..", "include", "asrk.h")]
NVCC_FLAGS = (os.environ.get("ASRK_EXTRA_NVCC", "").split()) + [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


FLAGS_STAMP = os.path.join(PKG_DIR, "build" + SUFFIX, "flags.txt")


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    # a library built with other flags (e.g. a -D debug switch through ASRK_EXTRA_NVCC) is stale;
    # without a stamp (a library that travelled alone) the time stamps decide
    if os.path.isfile(FLAGS_STAMP) and open(FLAGS_STAMP).read() != " ".join(NVCC_FLAGS):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    obj_dir = os.path.join(PKG_DIR, "build" + SUFFIX)
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(obj_dir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "shared"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    os.replace(tmp, LIB_PATH)
    open(FLAGS_STAMP, "w").write(" ".join(NVCC_FLAGS))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
