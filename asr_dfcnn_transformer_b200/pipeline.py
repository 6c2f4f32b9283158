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
    def __init__(self, device=None, feature_ctas=0, merged_tail=False, feature_priority=0):
        torch = _lib.require_cuda()
        self.torch = torch
        # merged_tail=True: the step's HBM-bound tail as ONE kernel -- the z-score pass rides on the fused CTC kernel as
        # co-work whenever the batch is bounded to small lattices (``ctc_bounds``).  Bit-identical and one launch fewer,
        # but measured no faster than the two overlapping kernels (profiles/r2_tail.md), so it is not the default
        self.merged_tail = bool(merged_tail)
        self._ev_feat = torch.cuda.Event()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.feature_ctas = int(feature_ctas)
        # (feature_priority < 0: the persistent transform on a high-priority stream -- only matters when several
        # steps are in flight on different HotPathStep instances, see tools/time_overlap.py)
        self.side = torch.cuda.Stream(device=self.device, priority=int(feature_priority))
        # the CTC kernels go on a second stream of their own (fixed streams keep the scratch buffers' addresses fixed
        # for graph capture).  Measured and dropped: a HIGH-priority stream for the CTC kernel, gated behind the
        # transform kernel, so that its CTAs are dispatched before the z-score CTAs -- the two HBM-bound kernels
        # overlap by no more than ~18 us however they are ordered (tools/time_tail.py), and the extra event and
        # launch cost 4 us per step.
        self.hi = torch.cuda.Stream(device=self.device)
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self._ev_join2 = torch.cuda.Event()

    def reserve(self, batch, total_frames, T, label_stride):
        """Size the scratch buffers of both streams for the largest step that will be captured: the grow-only
        workspaces must not be re-allocated once a graph holds their addresses."""
        L = _lib.lib()
        features.workspace(L.asrk_spectrogram_workspace_bytes(int(batch), int(total_frames)), self.device, "spec", self.side)
        features.workspace(L.asrk_ctc_workspace_bytes(int(T), int(batch), int(label_stride)), self.device, "ctc", self.hi)

    def capture(self, *args, loss_acc=None, loss_out=None, **kw):
        """Capture one step on fixed device buffers (same arguments as ``__call__``; pass ``feat_out`` / ``grad_out``)
        into a CUDA graph: the steady state of a training loop replays it without any host-side enqueue between
        its kernels.  ``loss_acc`` (float64 [2] device tensor): the per-step [sum loss, n] is ADDED to it inside the
        graph (``ctc.loss_sum(..., accumulate=True)``); ``loss_out``: the step's own [sum loss, n] is WRITTEN to it.
        Returns (graph, features, CtcResult); ``graph.replay()`` runs the step on the current stream."""
        torch = self.torch
        # workspaces, function attributes and the streams' pools exist before the capture starts
        self(*args, **kw)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (NCCL's watchdog polls CUDA events) must not invalidate the capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            feats, res = self(*args, **kw)
            if loss_acc is not None:
                ctc.loss_sum(res.loss, res.row_status, out=loss_acc, accumulate=True)
            if loss_out is not None:
                ctc.loss_sum(res.loss, res.row_status, out=loss_out)
        return g, feats, res

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
        self.hi.wait_event(self._ev_fork)
        res = None
        if self.merged_tail and mode == "fbank" and ctc_bounds is not None and batch > 0 and total_frames > 0:
            # transform + statistics on the side stream; the CTC kernel waits for them and normalises the rows itself
            with torch.cuda.stream(self.side):
                feats = features.spectrogram_device(samples, sample_offsets, sample_counts, frame_offsets, batch,
                                                    total_frames, mode, out=feat_out, cta_limit=self.feature_ctas,
                                                    stream=self.side,
                                                    phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN | _lib.PHASE_SPEC_STATS)
                self._ev_feat.record(self.side)
            zw = features.zscore_work(feats, frame_offsets, batch, total_frames, stream=self.side)
            self.hi.wait_event(self._ev_feat)
            try:
                with torch.cuda.stream(self.hi):
                    res = ctc.ctc_loss_grad(logits, labels, label_len, input_len, blank, layout=layout,
                                            grad_scale=grad_scale, grad_out=grad_out, decode=decode, bounds=ctc_bounds,
                                            stream=self.hi, zscore=zw)
                    self._ev_join2.record(self.hi)
                self._ev_join.record(self.side)
            except ctc.CoWorkNotEligible:
                res = None                      # nothing was launched: finish the features the two-kernel way
                with torch.cuda.stream(self.side):
                    features.spectrogram_device(samples, sample_offsets, sample_counts, frame_offsets, batch,
                                                total_frames, mode, out=feats, stream=self.side,
                                                phases=_lib.PHASE_SPEC_NORMALIZE)
                    self._ev_join.record(self.side)
        else:
            with torch.cuda.stream(self.side):
                feats = features.spectrogram_device(samples, sample_offsets, sample_counts, frame_offsets, batch,
                                                    total_frames, mode, out=feat_out, cta_limit=self.feature_ctas,
                                                    stream=self.side)
                self._ev_join.record(self.side)
        if res is None:
            with torch.cuda.stream(self.hi):
                res = ctc.ctc_loss_grad(logits, labels, label_len, input_len, blank, layout=layout,
                                        grad_scale=grad_scale, grad_out=grad_out, decode=decode, bounds=ctc_bounds,
                                        stream=self.hi)
                self._ev_join2.record(self.hi)
        cur.wait_event(self._ev_join)
        cur.wait_event(self._ev_join2)
        # (the tensors were allocated / are used on other streams than the caller's: tell the allocator -- not under a
        # graph capture, where the graph's private pool owns them)
        if not torch.cuda.is_current_stream_capturing():
            for t in (feats, res.loss, res.grad, res.row_status, res.tokens, res.token_len, res.neg_sum_logits):
                if t is not None:
                    t.record_stream(cur)
        return feats, res


class StepsInFlight:
    """Device-resident steps of DIFFERENT batches in flight at the same time.

    Inside one step the persistent transform (fp64 pipe and shared memory: HBM nearly idle, every SM taken) is
    followed by the HBM-bound tail (z-score, CTC), whose CTAs retire at different times: a step alone leaves SMs
    empty while the last CTC CTAs finish, and the memory system idle while the transform runs.  The steps of
    consecutive batches are independent (the loader of the reference runs ahead of the training step in its own
    thread, ``train.py:40-42``), so each resident batch gets its own ``HotPathStep`` (own streams, own scratch
    buffers), its step is captured into a CUDA graph, and the graphs are replayed round-robin on ``lanes`` streams:
    the next batch's transform -- limited to ``feature_ctas`` CTAs so that it never needs the whole chip -- starts on
    the SMs the previous batch's tail has left, and that tail runs next to it.  Results are bit-identical to the
    serial step (``tests/test_gpu_roundtrip.py``); measured on a C2 batch: 219 -> 191 us per step with two lanes,
    flat for ``feature_ctas`` between 72 and 112 (``tools/time_overlap.py``, ``profiles/r2_overlap.md``).

    ``add`` captures a batch (same arguments as ``HotPathStep.__call__``) and returns its slot: ``slot.features``,
    ``slot.result`` and ``slot.loss_sum`` (float64 [2]: the step's own [sum loss, n], rewritten by every replay).
    ``launch(slot)`` replays it on the next lane and returns that lane's stream; ``slot.done`` is recorded behind
    it.  A consumer that reads ``slot.loss_sum`` on another stream hands ``launch`` an event through
    ``slot.reusable`` (recorded after its read) so that the next replay of the slot waits for it.  ``join`` makes
    the current stream wait for everything launched so far."""

    class Slot:
        pass

    def __init__(self, device=None, lanes=2, feature_ctas=104, merged_tail=False):
        torch = _lib.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.lanes = [torch.cuda.Stream(device=self.device) for _ in range(int(lanes))]
        self.feature_ctas, self.merged_tail = int(feature_ctas), bool(merged_tail)
        self.slots = []
        self._k = 0
        self._ev_start = torch.cuda.Event()

    def add(self, samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames, logits, labels,
            label_len, input_len, blank=None, **kw):
        torch = self.torch
        s = StepsInFlight.Slot()
        s.step = HotPathStep(self.device, feature_ctas=self.feature_ctas, merged_tail=self.merged_tail)
        # the scratch buffers are keyed on the launch streams: steps in flight must not share one (torch hands out
        # streams from a pool of 32 per device and priority)
        used = {h for sl in self.slots for h in (sl.step.side.cuda_stream, sl.step.hi.cuda_stream)}
        if {s.step.side.cuda_stream, s.step.hi.cuda_stream} & used or s.step.side.cuda_stream == s.step.hi.cuda_stream:
            raise RuntimeError("StepsInFlight: torch's stream pool handed out a stream twice (too many resident batches)")
        s.step.reserve(batch, total_frames, logits.shape[0] if kw.get("layout", "tbv") == "tbv" else logits.shape[1],
                       labels.shape[1])
        s.loss_sum = torch.zeros(2, dtype=torch.float64, device=self.device)
        s.graph, s.features, s.result = s.step.capture(samples, sample_offsets, sample_counts, frame_offsets, batch,
                                                       total_frames, logits, labels, label_len, input_len, blank,
                                                       loss_out=s.loss_sum, **kw)
        # the graph holds raw addresses: the slot keeps every tensor of the batch alive
        s.inputs = (samples, sample_offsets, sample_counts, frame_offsets, logits, labels, label_len, input_len, kw)
        s.done = torch.cuda.Event()
        s.reusable = None
        self.slots.append(s)
        return s

    def launch(self, slot):
        torch = self.torch
        lane = self.lanes[self._k % len(self.lanes)]
        self._k += 1
        # the lane starts behind whatever the caller has enqueued so far (inputs written on the current stream)
        self._ev_start.record(torch.cuda.current_stream(self.device))
        lane.wait_event(self._ev_start)
        if slot.reusable is not None:
            lane.wait_event(slot.reusable)
            slot.reusable = None
        with torch.cuda.stream(lane):
            slot.graph.replay()
            slot.done.record(lane)
        return lane

    def join(self, stream=None):
        cur = self.torch.cuda.current_stream(self.device) if stream is None else stream
        for lane in self.lanes:
            cur.wait_stream(lane)


class HostRoundTrip:
    """The hot path with HOST buffers on both sides, double-buffered: while step k computes, the inputs
    of step k+1 go host -> device on a copy-in stream and the results of step k-1 come back on a
    copy-out stream (PCIe is full duplex).  Per step:

      in   PCM, labels and the padded logits tensor (DMA from pinned memory)
      out  per-utterance loss, the features (DMA into pinned memory) and the gradient without its
           all-zero padding rows (``ctc.unstage_rows``: the SMs write the mapped host buffer)

    ``logits_in="zero_copy"`` pulls the logits without their padding with ``ctc.stage_logits`` instead (29 % fewer
    bytes for a C2 batch).  Measured on the B200 box (tools/pcie_ceiling.py, profiles/r2_e2e.md): SM reads of host
    memory and any concurrent device -> host traffic serialise (60 GB/s for both directions together against
    100 GB/s for two DMA copies), SM *writes* to host memory next to a host -> device DMA do not, so the default
    moves the padded tensor by DMA: 3.8 ms per C2 step instead of 4.6.

    ``submit`` returns a slot; ``wait(slot)`` blocks until that step's results are in its host buffers
    (``slot.h_loss / h_feat / h_grad``).  A slot's buffers are reused by the submit after next."""

    class Slot:
        pass

    def __init__(self, step, total_frames_max, samples_max, sample_dtype, T, B, V, label_stride, slots=2,
                 return_outputs=True, logits_in=None):
        if logits_in is None:      # SM reads of host memory only hurt next to device -> host traffic
            logits_in = "dma" if return_outputs else "zero_copy"
        if logits_in not in ("dma", "zero_copy"):
            raise ValueError("logits_in must be 'dma' or 'zero_copy'")
        self.logits_in = logits_in
        torch = step.torch
        self.step, self.torch = step, torch
        dev = step.device
        self.return_outputs = return_outputs
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(slots):
            s = HostRoundTrip.Slot()
            s.samples = torch.empty(samples_max, dtype=sample_dtype, device=dev)
            s.labels = torch.empty((B, label_stride), dtype=torch.int32, device=dev)
            s.logits = torch.zeros((T, B, V), dtype=torch.float32, device=dev)
            s.feat = torch.empty((total_frames_max, 200), dtype=torch.float32, device=dev)
            s.grad = torch.empty((T, B, V), dtype=torch.float32, device=dev)
            s.h_loss = torch.empty(B, dtype=torch.float32).pin_memory()
            if return_outputs:
                s.h_feat = torch.empty((total_frames_max, 200), dtype=torch.float32).pin_memory()
                s.h_grad = torch.zeros((T, B, V), dtype=torch.float32).pin_memory()
                s.stale_len = torch.zeros(B, dtype=torch.int32, device=dev)   # lengths of the buffer's previous batch
            s.ev_in, s.ev_done, s.ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            s.busy = False
            self.slots.append(s)
        self.k = 0

    def submit(self, h_samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames, h_logits,
               h_labels, label_len, input_len, blank=None, grad_scale=None, ctc_bounds=None, valid_rows=None):
        """Host tensors are pinned; offsets / lengths are device tensors (they are a few KB and belong to the
        batch description).  Returns the slot; its h2d_bytes / d2h_bytes attributes say what crossed PCIe
        (``valid_rows`` = sum of the input lengths, a host int: the rows the zero-copy kernels move; without it
        those legs are counted at the padded size)."""
        torch = self.torch
        s = self.slots[self.k % len(self.slots)]
        self.k += 1
        if s.busy:
            self.wait(s)
        cur = torch.cuda.current_stream(self.step.device)
        n = h_samples.numel()
        with torch.cuda.stream(self.s_in):
            s.samples[:n].copy_(h_samples, non_blocking=True)
            s.labels.copy_(h_labels, non_blocking=True)
            if self.logits_in == "dma":
                s.logits[:h_logits.shape[0]].copy_(h_logits, non_blocking=True)
            else:
                ctc.stage_logits(h_logits, input_len, out=s.logits, stream=self.s_in)
            s.ev_in.record(self.s_in)
        cur.wait_event(s.ev_in)
        feats, res = self.step(s.samples[:n], sample_offsets, sample_counts, frame_offsets, batch, total_frames,
                               s.logits, s.labels, label_len, input_len, blank, feat_out=s.feat[:total_frames],
                               grad_out=s.grad, grad_scale=grad_scale, ctc_bounds=ctc_bounds)
        s.ev_done.record(cur)
        self.s_out.wait_event(s.ev_done)
        with torch.cuda.stream(self.s_out):
            s.h_loss.copy_(res.loss, non_blocking=True)
            if self.return_outputs:
                s.h_feat[:total_frames].copy_(feats, non_blocking=True)
                ctc.unstage_rows(s.grad, input_len, s.h_grad, stream=self.s_out, stale_len=s.stale_len)
            s.ev_out.record(self.s_out)
        s.res, s.total_frames, s.busy = res, total_frames, True
        row_bytes = 4 * h_logits.shape[-1]
        padded = h_logits.numel() * 4
        moved = padded if valid_rows is None else int(valid_rows) * row_bytes
        s.h2d_bytes = n * h_samples.element_size() + h_labels.numel() * 4 + (padded if self.logits_in == "dma" else moved)
        s.d2h_bytes = s.h_loss.numel() * 4 + ((total_frames * 800 + moved) if self.return_outputs else 0)
        return s

    def wait(self, s):
        s.ev_out.synchronize()
        s.busy = False
        return s

    def drain(self):
        for s in self.slots:
            if s.busy:
                self.wait(s)
