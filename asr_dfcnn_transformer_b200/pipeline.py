"""The fused training-side step of the hot path: spectrogram features of a batch
and CTC loss + gradient of a batch, issued on two CUDA streams.

The persistent feature kernel (enqueued first, on a side stream) owns every SM while it
runs; the CTC kernels (main stream) fill the SMs as its CTAs retire, and the small z-score
CTAs fit next to the CTC kernel's two CTAs per SM.  Measured both ways round: features
first is 1-2 % faster than CTC first.  Nothing here synchronises with the host.
``shard_bounds`` / ``all_reduce_loss`` are the multi-GPU plumbing: contiguous batch shards,
one SUM all-reduce of [sum loss, n] per step.
"""
from . import _lib, ctc, features


def shard_bounds(n_utterances, rank, world):
    """Contiguous batch shard of rank ``rank``: utterances are
    independent in all three parts, so no tensor ever crosses GPUs."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    lo = (n_utterances * rank) // world
    hi = (n_utterances * (rank + 1)) // world
    return lo, hi


def all_reduce_loss(loss_sum_and_count, group=None, stream=None):
    """The path's only collective: SUM all-reduce of the 2-element tensor
    ``[sum of per-utterance losses, number of utterances]`` (float64), mirroring
    ``tf.reduce_mean(self.loss)`` (acoustic_model2.py:83) across ranks.  Works with any
    ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests); when a CUDA
    ``stream`` is given the collective is enqueued there so that it overlaps the next
    step.  Returns the (in-place reduced) tensor; mean = t[0] / t[1]."""
    import torch
    import torch.distributed as dist
    t = loss_sum_and_count
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    if stream is not None and t.is_cuda:
        stream.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(stream):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class HotPathStep:
    def __init__(self, device=None, feature_ctas=0):
        torch = _lib.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.feature_ctas = int(feature_ctas)
        self.side = torch.cuda.Stream(device=self.device)
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()

    def from_host(self, h_samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames,
                  h_logits, h_labels, label_len, input_len, blank=None, logits_dev=None, **kw):
        """The same step with the per-step inputs in PINNED HOST memory: PCM and labels are copied
        host -> device, the logits are staged without their padding (``ctc.stage_logits``: only rows
        t < input_len[b] cross PCIe), then the kernels run.  Lengths / offsets are device tensors.
        Returns (features, CtcResult, h2d_bytes)."""
        torch = self.torch
        samples = h_samples.to(self.device, non_blocking=True)
        labels = h_labels.to(self.device, non_blocking=True)
        logits = ctc.stage_logits(h_logits, input_len, out=logits_dev, layout=kw.get("layout", "tbv"))
        feats, res = self(samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames, logits, labels,
                          label_len, input_len, blank, **kw)
        return feats, res, samples.numel() * samples.element_size() + labels.numel() * labels.element_size()

    def __call__(self, samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames,
                 logits, labels, label_len, input_len, blank=None, feat_out=None, grad_out=None,
                 grad_scale=None, mode="fbank", decode=False, layout="tbv", ctc_bounds=None):
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
                                grad_scale=grad_scale, grad_out=grad_out, decode=decode, bounds=ctc_bounds)
        cur.wait_event(self._ev_join)
        return feats, res
