"""oracle/ — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``asr_dfcnn_transformer_b200/`` may import this package.  The only
callers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and there only as the checker or
as the timed CPU baseline -- never as the thing shipped.

Pinning status (see DESIGN.md "Oracle"):
  * features / noise: PINNED -- ``oracle/fbank_ref.py`` is checked against the
    reference's own ``util/wav_util.py`` and ``util/noise.py`` imported from
    ``/root/reference`` (``oracle/ref_import.py``) and against the vectors that
    import produced, committed under ``tests/golden/``.
  * CTC loss/grad and greedy decode: the arithmetic lives in TensorFlow 1.14 /
    Keras 2.3.1 C++ kernels that are not vendored in the reference and cannot be
    installed here.  The reference holds no test or golden vector for them, so
    these two are "parity unpinned" against the reference itself; they are
    cross-checked against an independent float64 implementation
    (``torch.nn.functional.ctc_loss``) and hand-computed known answers.
"""
