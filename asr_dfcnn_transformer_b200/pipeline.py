"""The fused training-side step of the hot path: spectrogram features of a batch
and CTC loss + gradient of a batch, issued on two CUDA streams.

The feature kernel is bound by the fp64 pipe, the CTC kernels by HBM bandwidth, so
they overlap well: the persistent feature kernel is capped to ``feature_ctas`` SMs
and the CTC kernels fill the rest.  Nothing here synchronises with the host.
"""
from . import _lib, ctc, features


class HotPathStep:
    def __init__(self, device=None, feature_ctas=0):
        torch = _lib.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.feature_ctas = int(feature_ctas)
        self.side = torch.cuda.Stream(device=self.device)
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()

    def __call__(self, samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames,
                 logits, labels, label_len, input_len, blank=None, feat_out=None, grad_out=None,
                 grad_scale=None, mode="fbank", decode=False, layout="tbv"):
        """Returns (features, CtcResult).  Both are complete, in stream order, on the
        current stream when this returns (no host synchronisation)."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        self._ev_fork.record(cur)
        self.side.wait_event(self._ev_fork)
        with torch.cuda.stream(self.side):
            feats = features.spectrogram_device(samples, sample_offsets, sample_counts, frame_offsets, batch,
                                                total_frames, mode, out=feat_out, cta_limit=self.feature_ctas)
            self._ev_join.record(self.side)
        res = ctc.ctc_loss_grad(logits, labels, label_len, input_len, blank, layout=layout,
                                grad_scale=grad_scale, grad_out=grad_out, decode=decode)
        cur.wait_event(self._ev_join)
        return feats, res
