"""CTC loss / gradient and greedy decode: host side of asrk_ctc_*.

Mirrors the call surface the reference's models use (names, argument meaning,
error behaviour):
  * ``ctc_batch_cost(y_true, y_pred, input_length, label_length)`` -- Keras
    ``K.ctc_batch_cost`` as called through ``ctc_lambda``
    (lm_and_am/model/cnn_ctc.py:149-152): softmax input ``[B,T,V]``, blank =
    V-1, ``log(y_pred + 1e-7)``, labels masked by ``label_length``, ``[B,1]``.
  * ``ctc_loss_v2(labels, logits, label_length, logit_length, blank_index)`` --
    ``tf.nn.ctc_loss_v2`` as called at lm_and_am/model/acoustic_model2.py:79-80
    (time-major logits ``[T,B,V]``).  ``labels`` may be dense ``[B,L]`` (masked by
    label_length) or a ``SparseLabels`` made by ``dense_to_sparse`` (which drops
    every 0, acoustic_model2.py:71).
  * ``ctc_greedy_decoder(inputs, sequence_length)`` -- ``tf.nn.ctc_greedy_decoder``
    (acoustic_model2.py:69); ``decode_ctc(num_result, input_length)`` --
    util/utils.py:57-66.
Both losses are differentiable (``torch.autograd.Function``): the gradient w.r.t.
the logits is produced by the same kernel pass as the loss.
"""
from collections import namedtuple

import ctypes

import numpy as np

from . import _lib
from .features import workspace

SparseTensorValue = namedtuple("SparseTensorValue", "indices values dense_shape")
SparseLabels = namedtuple("SparseLabels", "dense")   # result of dense_to_sparse: zeros are dropped
CtcResult = namedtuple("CtcResult", "loss grad row_status tokens token_len neg_sum_logits")


class InvalidArgumentError(ValueError):
    """What TensorFlow raises from CTCLossOp ("Not enough time for target
    transition sequence"), aborting ``sess.run`` (lm_and_am/train.py:58-69)."""


def _strides(t, layout):
    # element strides of (time, batch); V must be contiguous
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError("logits must be a 3-D tensor with a contiguous last dimension")
    if layout == "tbv":
        return t.shape[0], t.shape[1], t.shape[2], t.stride(0), t.stride(1)
    if layout == "btv":
        return t.shape[1], t.shape[0], t.shape[2], t.stride(1), t.stride(0)
    raise ValueError("layout must be 'tbv' or 'btv'")


def _i32(x, dev, torch):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.int32).contiguous().view(-1)
    return torch.as_tensor(np.asarray(x).reshape(-1).astype(np.int32)).to(dev)


class CoWorkNotEligible(RuntimeError):
    """ctc_loss_grad(zscore=...) on a batch that cannot take the fused kernel; nothing was launched."""


def ctc_loss_grad(logits, labels, label_len, input_len, blank=None, label_mode="by_length",
                  layout="tbv", grad_scale=None, want_grad=True, decode=False, grad_out=None,
                  stream=None, phases=_lib.PHASE_ALL, outputs=None, bounds=None, input_kind="logits", zscore=None):
    """Raw op: one fused pass.  logits float32 CUDA tensor ``[T,B,V]`` ('tbv') or
    ``[B,T,V]`` ('btv'); labels int32 ``[B,Lmax]``.  Returns CtcResult of device
    tensors (no synchronisation, statuses are NOT checked here).

    ``bounds = (max input_len, max label_len)`` of the batch, when the host knows them (the
    reference's loader builds both length vectors on the host, data_loader.py:132-148): a batch
    whose largest lattice fits the fused kernel is then ONE launch -- the three generic kernels
    that would otherwise be launched to find nothing to do are skipped.  Lengths given as host
    arrays are bounded here; a row that breaks the promise gets row_status ROW_NOT_SMALL.

    ``input_kind="prob"``: ``logits`` holds the softmax output p that Keras hands to ``K.ctc_batch_cost``
    (cnn_ctc.py:149-152); the op's input ``log(p + 1e-7)`` is formed inside the kernel and ``grad`` is the
    gradient w.r.t. p (no element-wise pass before or after the kernel).

    ``zscore`` (``features.ZScoreWork``): the z-score pass of the feature path rides on the fused kernel as co-work
    (one kernel for the step's HBM-bound tail).  Raises ``CoWorkNotEligible`` -- before anything is launched -- when
    the batch cannot take the fused kernel alone (not bounded to small lattices, vector path not applicable)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    if logits.dtype != torch.float32 or not logits.is_cuda:
        raise TypeError("logits must be a float32 CUDA tensor")
    dev = logits.device
    T, B, V, st, sb = _strides(logits, layout)
    if blank is None:
        blank = V - 1
    labels = labels.to(device=dev, dtype=torch.int32) if isinstance(labels, torch.Tensor) else \
        torch.as_tensor(np.asarray(labels).astype(np.int32)).to(dev)
    labels = labels.reshape(B, -1).contiguous()
    Ls = labels.shape[1]
    # length vectors that live on the host (numpy, lists, CPU tensors -- what Keras / the loader hand over)
    if isinstance(input_len, torch.Tensor) and not input_len.is_cuda:
        input_len = input_len.numpy()
    if isinstance(label_len, torch.Tensor) and not label_len.is_cuda:
        label_len = label_len.numpy()
    if bounds is None and not isinstance(input_len, torch.Tensor) and np.size(input_len):
        lmax = Ls if (label_len is None or isinstance(label_len, torch.Tensor) or not np.size(label_len)) \
            else int(np.max(label_len))
        bounds = (int(np.max(input_len)), lmax)
    if bounds is not None and phases == _lib.PHASE_ALL and \
            L.asrk_ctc_fits_fused(max(int(bounds[0]), 0), max(int(bounds[1]), 0)):
        phases = phases | _lib.CTC_SMALL_ONLY
    if input_kind == "prob":
        if decode:
            raise ValueError("the greedy decode is defined on the op's input, not on probabilities")
        phases = phases | _lib.CTC_INPUT_PROB
    elif input_kind != "logits":
        raise ValueError("input_kind must be 'logits' or 'prob'")
    input_len = _i32(input_len, dev, torch)
    label_len = _i32(label_len, dev, torch) if label_len is not None else None
    mode = _lib.LABELS_BY_LENGTH if label_mode == "by_length" else _lib.LABELS_DROP_ZEROS
    if outputs is not None:      # re-issue of a later phase on the same buffers
        loss, grad_out, status, tokens_, tlen_, nsl_ = outputs
    else:
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        tokens_ = tlen_ = nsl_ = None
    grad = None
    gt = gb = 0
    if want_grad:
        grad = grad_out if grad_out is not None else torch.empty_like(logits)
        _, _, _, gt, gb = _strides(grad, layout)
    tokens = tlen = nsl = None
    if decode:
        tokens = tokens_ if tokens_ is not None else torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
        tlen = tlen_ if tlen_ is not None else torch.empty(B, dtype=torch.int32, device=dev)
        nsl = nsl_ if nsl_ is not None else torch.empty(B, dtype=torch.float32, device=dev)
    if grad_scale is not None:
        grad_scale = grad_scale.to(device=dev, dtype=torch.float32).contiguous()
    nbytes = L.asrk_ctc_workspace_bytes(T, B, Ls)
    ws = workspace(nbytes, dev, "ctc", stream)
    if zscore is not None:
        if (int(phases) & 0xffff) != _lib.PHASE_ALL or not (int(phases) & _lib.CTC_SMALL_ONLY):
            raise CoWorkNotEligible("the z-score co-work needs a whole-op call on a batch bounded to small lattices")
        z = zscore
        rc = L.asrk_ctc_loss_grad_zscore_run(_lib.ptr(logits), st, sb, T, B, V, _lib.ptr(labels), Ls,
                                             _lib.ptr(label_len), _lib.ptr(input_len), int(blank), mode,
                                             _lib.ptr(grad_scale), _lib.ptr(loss), _lib.ptr(grad), gt, gb,
                                             _lib.ptr(status), _lib.ptr(tokens), max(T, 1), _lib.ptr(tlen),
                                             _lib.ptr(nsl), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(stream),
                                             int(phases), _lib.ptr(z.features), z.stats, _lib.ptr(z.frame_offsets),
                                             _lib.ptr(z.row_offsets), z.batch, z.total_frames, z.ticket)
        if rc == _lib.E_SHAPE:
            raise CoWorkNotEligible("this batch does not take the fused CTC kernel (nothing was launched)")
        _lib.check(rc, "asrk_ctc_loss_grad_zscore_run")
        return CtcResult(loss, grad, status, tokens, tlen, nsl)
    rc = L.asrk_ctc_loss_grad_run_phases(_lib.ptr(logits), st, sb, T, B, V, _lib.ptr(labels), Ls,
                                         _lib.ptr(label_len), _lib.ptr(input_len), int(blank), mode,
                                         _lib.ptr(grad_scale), _lib.ptr(loss), _lib.ptr(grad), gt, gb,
                                         _lib.ptr(status), _lib.ptr(tokens), max(T, 1), _lib.ptr(tlen),
                                         _lib.ptr(nsl), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(stream),
                                         int(phases))
    _lib.check(rc, "asrk_ctc_loss_grad_run")
    return CtcResult(loss, grad, status, tokens, tlen, nsl)


def stage_logits(host_logits, input_len, out=None, layout="tbv", stream=None):
    """Host -> device copy of the logits WITHOUT their padding: ``host_logits`` is a pinned host
    tensor ([T,B,V] or [B,T,V]); only the rows t < input_len[b] cross PCIe (the SMs read the
    mapped host memory directly).  ``input_len``: int32 device tensor.  Returns the device tensor
    (rows past input_len are left as they were: the CTC kernels never read them)."""
    torch = _lib.require_cuda()
    if not host_logits.is_pinned():
        raise ValueError("stage_logits needs pinned (page-locked) host memory")
    dev = input_len.device
    if out is None:
        out = torch.empty(host_logits.shape, dtype=torch.float32, device=dev)
    T, B, V, st, sb = _strides(host_logits, layout)
    _, _, _, dt, db = _strides(out, layout)
    rc = _lib.lib().asrk_ctc_stage_logits_run(ctypes.c_void_p(host_logits.data_ptr()), st, sb, _lib.ptr(out), dt, db,
                                              _lib.ptr(input_len), T, B, V, _lib.stream_ptr(stream))
    _lib.check(rc, "asrk_ctc_stage_logits_run")
    return out


def loss_sum(loss, row_status=None, out=None, stream=None, accumulate=False):
    """[sum of the accepted rows' losses, their number] as a float64 device tensor of 2
    (the operand of the batch mean / of the cross-rank all-reduce).  ``accumulate=True`` adds to ``out``
    (mean over several steps with one all-reduce, like the reference's print every second step)."""
    torch = _lib.require_cuda()
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs the running tensor in `out`")
        out = torch.empty(2, dtype=torch.float64, device=loss.device)
    fn = _lib.lib().asrk_ctc_loss_sum_acc_run if accumulate else _lib.lib().asrk_ctc_loss_sum_run
    st = fn(_lib.ptr(loss), _lib.ptr(row_status), int(loss.numel()), _lib.ptr(out), _lib.stream_ptr(stream))
    _lib.check(st, "asrk_ctc_loss_sum_run")
    return out


def unstage_rows(dev_tensor, input_len, host_out, layout="tbv", stream=None, stale_len=None):
    """Device -> host copy of a ``[T,B,V]`` / ``[B,T,V]`` tensor (the gradient) WITHOUT its padding: only rows
    t < input_len[b] cross PCIe, written by the SMs into the pinned host tensor ``host_out``.  ``stale_len`` (int32
    device tensor [B], updated in place; start it at zero over a cleared buffer): the lengths of the batch that used
    ``host_out`` before -- the rows it wrote and this batch does not are set back to zero, so the buffer always equals
    the device tensor.  Without it the padding rows keep what the buffer held."""
    if not host_out.is_pinned():
        raise ValueError("unstage_rows needs pinned (page-locked) host memory")
    T, B, V, st, sb = _strides(dev_tensor, layout)
    _, _, _, dt, db = _strides(host_out, layout)
    rc = _lib.lib().asrk_ctc_unstage_rows_run(_lib.ptr(dev_tensor), st, sb, ctypes.c_void_p(host_out.data_ptr()), dt, db,
                                              _lib.ptr(input_len), _lib.ptr(stale_len), T, B, V, _lib.stream_ptr(stream))
    _lib.check(rc, "asrk_ctc_unstage_rows_run")
    return host_out


def _raise_on_status(status):
    st = status.cpu().numpy()
    bad = np.nonzero(st == _lib.ROW_NOT_ENOUGH_TIME)[0]
    if len(bad):
        raise InvalidArgumentError("Not enough time for target transition sequence "
                                   "(required: label length + repeats) in batch rows %s" % bad.tolist())
    bad = np.nonzero(st == _lib.ROW_BAD_LENGTH)[0]
    if len(bad):
        raise InvalidArgumentError("sequence_length / label_length / label value out of range in "
                                   "batch rows %s" % bad.tolist())
    bad = np.nonzero(st == _lib.ROW_NOT_SMALL)[0]
    if len(bad):
        raise ValueError("ctc_loss_grad(bounds=...) promised small lattices, batch rows %s are larger"
                         % bad.tolist())


def _make_fn():
    torch = _lib.require_cuda()

    class _CTCLoss(torch.autograd.Function):
        """loss = CTC(logits); the gradient w.r.t. the logits comes out of the same kernel pass, already
        multiplied by ``upstream`` (what the caller states the gradient of the final objective w.r.t. every
        loss[b] will be: 1 for ``.sum()``, 1/B for ``.mean()``; None = 1).  backward returns that tensor as it
        is when autograd's upstream gradient is the stated one -- a broadcast scalar is read back from the
        device (4 bytes, no kernel) to check -- and only otherwise pays an element-wise pass."""

        @staticmethod
        def forward(ctx, logits, labels, label_len, input_len, blank, label_mode, layout, check, kind, upstream):
            need = logits.requires_grad
            B = logits.shape[1] if layout == "tbv" else logits.shape[0]
            gs = None
            if upstream is not None and need:
                gs = torch.full((B,), float(upstream), dtype=torch.float32, device=logits.device)
            r = ctc_loss_grad(logits.detach(), labels, label_len, input_len, blank, label_mode,
                              layout, want_grad=need, grad_scale=gs, input_kind=kind)
            if check:
                _raise_on_status(r.row_status)
            ctx.layout = layout
            ctx.upstream = 1.0 if upstream is None else float(upstream)
            if need:
                ctx.save_for_backward(r.grad)
            return r.loss

        @staticmethod
        def backward(ctx, go):
            (g,) = ctx.saved_tensors
            go = go.reshape(-1)
            if go.numel() == 1 or go.stride(0) == 0:
                c = float(go[0].item()) / ctx.upstream      # broadcast upstream gradient (.sum(), .mean())
                gi = g if c == 1.0 else g * c
            else:
                w = go / ctx.upstream
                gi = g * (w.view(1, -1, 1) if ctx.layout == "tbv" else w.view(-1, 1, 1))
            return gi, None, None, None, None, None, None, None, None, None

    return _CTCLoss


_fn_cache = []


def _fn():
    if not _fn_cache:
        _fn_cache.append(_make_fn())
    return _fn_cache[0]


def dense_to_sparse(target):
    """tf.contrib.layers.dense_to_sparse (acoustic_model2.py:71): every entry
    equal to 0 is dropped, including a genuine label id 0."""
    return SparseLabels(target)


def ctc_loss_v2(labels, logits, label_length, logit_length, logits_time_major=True, blank_index=None,
                check=True, upstream=None):
    """tf.nn.ctc_loss_v2 (acoustic_model2.py:79-80).  Returns loss ``[B]``.  ``blank_index=None`` follows
    TensorFlow: 0 for dense labels, an error for sparse labels (the reference passes V-1 explicitly).
    ``upstream``: see ``ctc_batch_cost``."""
    layout = "tbv" if logits_time_major else "btv"
    if isinstance(labels, SparseLabels):
        if blank_index is None:
            raise ValueError("blank_index must be given when using SparseTensor labels.")
        return _fn().apply(logits, labels.dense, None, logit_length, int(blank_index), "drop_zeros", layout, check,
                           "logits", upstream)
    if blank_index is None:
        blank_index = 0
    V = logits.shape[2]
    if blank_index < 0:
        blank_index += V
    return _fn().apply(logits, labels, label_length, logit_length, int(blank_index), "by_length", layout, check,
                       "logits", upstream)


def ctc_batch_cost(y_true, y_pred, input_length, label_length, check=True, upstream=None):
    """Keras K.ctc_batch_cost (cnn_ctc.py:149-152): ``y_pred`` softmax output ``[B,T,V]``; returns ``[B,1]``.
    ONE kernel: ``log(y_pred + 1e-7)`` is formed when a row is loaded (the transpose is a stride swap) and the
    gradient is written w.r.t. ``y_pred``.  ``upstream`` states the gradient the caller's objective will send
    back into every loss entry (1 for ``.sum()`` -- the default --, ``1/B`` for Keras' batch mean): the kernel
    scales by it and ``backward`` hands the tensor on without another pass."""
    loss = _fn().apply(y_pred, y_true, label_length, input_length, None, "by_length", "btv", check, "prob", upstream)
    return loss.unsqueeze(1)


def greedy_decode(logits, input_len, blank=None, merge_repeated=True, layout="tbv", stream=None):
    """Raw op.  Returns (tokens int32 [B,T], token_len int32 [B], neg_sum_logits
    float32 [B]) as device tensors; only ``tokens[b, :token_len[b]]`` is defined."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    if logits.dtype != torch.float32 or not logits.is_cuda:
        raise TypeError("logits must be a float32 CUDA tensor")
    dev = logits.device
    T, B, V, st, sb = _strides(logits, layout)
    if blank is None:
        blank = V - 1
    input_len = _i32(input_len, dev, torch)
    tokens = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
    tlen = torch.empty(B, dtype=torch.int32, device=dev)
    nsl = torch.empty(B, dtype=torch.float32, device=dev)
    nbytes = L.asrk_ctc_decode_workspace_bytes(max(T, 1), B)
    ws = workspace(nbytes, dev, "ctc", stream)
    rc = L.asrk_ctc_greedy_decode_run(_lib.ptr(logits), st, sb, T, B, V, _lib.ptr(input_len), int(blank),
                                      1 if merge_repeated else 0, _lib.ptr(tokens), max(T, 1),
                                      _lib.ptr(tlen), _lib.ptr(nsl), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(stream))
    _lib.check(rc, "asrk_ctc_greedy_decode_run")
    return tokens, tlen, nsl


def tokens_to_lists(tokens, token_len):
    tk = tokens.cpu().numpy()
    tl = token_len.cpu().numpy()
    return [tk[b, : tl[b]].astype(np.int64).tolist() for b in range(tk.shape[0])]


def ctc_greedy_decoder(inputs, sequence_length, merge_repeated=True):
    """tf.nn.ctc_greedy_decoder (acoustic_model2.py:69): inputs ``[T,B,V]``.
    Returns ``([SparseTensorValue], neg_sum_logits[B,1])`` like ``sess.run`` of
    the TF op gives."""
    tokens, tlen, nsl = greedy_decode(inputs, sequence_length, None, merge_repeated, "tbv")
    seqs = tokens_to_lists(tokens, tlen)
    idx = [(b, j) for b, s in enumerate(seqs) for j in range(len(s))]
    vals = [v for s in seqs for v in s]
    width = max([len(s) for s in seqs] + [0])
    sp = SparseTensorValue(np.asarray(idx, dtype=np.int64).reshape(-1, 2),
                           np.asarray(vals, dtype=np.int64),
                           np.asarray([len(seqs), width], dtype=np.int64))
    return [sp], nsl.cpu().numpy().reshape(-1, 1)


def sparse_tensor_to_dense(sp, default_value=0):
    """tf.sparse_tensor_to_dense (lm_and_am/test.py:51 pads with 0)."""
    out = np.full(tuple(int(v) for v in sp.dense_shape), default_value, dtype=np.int64)
    if len(sp.values):
        out[sp.indices[:, 0], sp.indices[:, 1]] = sp.values
    return out


def decode_ctc(num_result, input_length):
    """util/utils.py:57-66: greedy decode of one utterance ``[1,T,V]`` (Keras
    ``K.ctc_decode(greedy=True)``); returns the 1-D id array of row 0."""
    torch = _lib.require_cuda()
    x = num_result if isinstance(num_result, torch.Tensor) else torch.as_tensor(np.asarray(num_result, dtype=np.float32))
    x = x[:, :, :].to("cuda", dtype=torch.float32)
    in_len = np.zeros((1), dtype=np.int32)
    in_len[0] = input_length
    # K.ctc_decode takes log(transpose(y_pred) + epsilon); the arg-max is unchanged
    # by the monotone map, but ties created by the float32 log are not, so apply it
    xl = torch.log(x + 1e-7)
    tokens, tlen, _ = greedy_decode(xl, in_len, None, True, "btv")
    return np.asarray(tokens_to_lists(tokens, tlen)[0], dtype=np.int64)
