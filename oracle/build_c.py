"""Compile oracle/ctc_ref.c (float = TF-faithful, double = checking) into
oracle/_ref/libctc_ref.so with gcc + OpenMP.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libctc_ref.so")
SRC = os.path.join(HERE, "ctc_ref.c")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    objs = []
    for suffix, real in (("f32", "float"), ("f64", "double")):
        o = os.path.join(OUT_DIR, "ctc_ref_%s.o" % suffix)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-fopenmp", "-std=c99", "-DCTC_REAL=%s" % real,
                               "-DCTC_NAME(x)=x##_%s" % suffix, "-c", SRC, "-o", o])
        objs.append(o)
    subprocess.check_call(["gcc", "-shared", "-fopenmp", "-o", LIB] + objs + ["-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def ctc_loss_grad(logits_tbv, labels, label_len, input_len, blank, want_grad=True, real="f32", threads=0):
    import numpy as np
    x = np.ascontiguousarray(logits_tbv, dtype=np.float32)
    T, B, V = x.shape
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    ll = np.ascontiguousarray(label_len, dtype=np.int32)
    il = np.ascontiguousarray(input_len, dtype=np.int32)
    loss = np.zeros(B, dtype=np.float32)
    status = np.zeros(B, dtype=np.int32)
    grad = np.zeros_like(x) if want_grad else None
    fn = getattr(lib(), "ctc_loss_grad_" + real)
    vp = ctypes.c_void_p
    fn.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, vp, ctypes.c_int,
                   vp, vp, vp, ctypes.c_int]
    fn(x.ctypes.data, T, B, V, labels.ctypes.data, labels.shape[1], ll.ctypes.data, il.ctypes.data,
       int(blank), loss.ctypes.data, grad.ctypes.data if want_grad else None, status.ctypes.data, threads)
    return loss, grad, status


def greedy_decode(logits_tbv, input_len, blank, merge_repeated=True, threads=0):
    import numpy as np
    x = np.ascontiguousarray(logits_tbv, dtype=np.float32)
    T, B, V = x.shape
    il = np.ascontiguousarray(input_len, dtype=np.int32)
    tokens = np.zeros((B, T), dtype=np.int32)
    tlen = np.zeros(B, dtype=np.int32)
    nsl = np.zeros(B, dtype=np.float32)
    fn = lib().ctc_greedy_decode_f32
    vp = ctypes.c_void_p
    fn.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp,
                   ctypes.c_int]
    fn(x.ctypes.data, T, B, V, il.ctypes.data, int(blank), 1 if merge_repeated else 0, tokens.ctypes.data,
       tlen.ctypes.data, nsl.ctypes.data, threads)
    return [tokens[b, :tlen[b]].tolist() for b in range(B)], nsl


if __name__ == "__main__":
    print(build(force=True))
