"""Host <-> device transfer ceilings of the box, next to the transfers the end-to-end path makes (C2 batch):
DMA copies of pinned buffers in each direction and in both at once, the zero-copy staging kernels
(ctc.stage_logits / ctc.unstage_rows: the SMs read / write mapped host memory, padding rows never cross PCIe),
and the DMA alternative for the same tensors.  Prints one line per leg (GB/s of the bytes that crossed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import ctc  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
hb = bench.make_batch(2000, 256, "c2")
T, B, V = hb["logits"].shape
il = torch.as_tensor(hb["input_len"]).to(dev)
valid = int(hb["input_len"].sum()) * V * 4
h_logits = torch.as_tensor(hb["logits"]).pin_memory()
d_logits = torch.zeros((T, B, V), dtype=torch.float32, device=dev)
d_grad = torch.randn((T, B, V), dtype=torch.float32, device=dev)
h_grad = torch.zeros((T, B, V), dtype=torch.float32).pin_memory()
nf = int(hb["nfr"].sum())
d_feat = torch.randn((nf, 200), dtype=torch.float32, device=dev)
h_feat = torch.empty((nf, 200), dtype=torch.float32).pin_memory()
big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
big_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big_h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
big_d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(name, nbytes, fn, n=8):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-58s %7.3f ms  %6.1f GB/s  (%.1f MB)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e6), flush=True)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        big_d.copy_(big_h, non_blocking=True)
    with torch.cuda.stream(s2):
        big_h2.copy_(big_d2, non_blocking=True)


def both_zero_copy():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ctc.stage_logits(h_logits, il, out=d_logits, stream=s1)
    ctc.unstage_rows(d_grad, il, h_grad, stream=s2)


def both_e2e_like():
    """what one step of pipeline.HostRoundTrip moves: logits in (zero copy) || features out (DMA) + gradient out (zero copy)"""
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ctc.stage_logits(h_logits, il, out=d_logits, stream=s1)
    with torch.cuda.stream(s2):
        h_feat.copy_(d_feat, non_blocking=True)
    ctc.unstage_rows(d_grad, il, h_grad, stream=s2)


def both_dma_like():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_logits.copy_(h_logits, non_blocking=True)
    with torch.cuda.stream(s2):
        h_feat.copy_(d_feat, non_blocking=True)
        h_grad.copy_(d_grad, non_blocking=True)


def mixed_x():
    """logits in by DMA (padded) || features DMA + gradient out zero copy"""
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_logits.copy_(h_logits, non_blocking=True)
    with torch.cuda.stream(s2):
        h_feat.copy_(d_feat, non_blocking=True)
    ctc.unstage_rows(d_grad, il, h_grad, stream=s2)


def mixed_y():
    """logits in zero copy || features DMA + gradient DMA (padded)"""
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ctc.stage_logits(h_logits, il, out=d_logits, stream=s1)
    with torch.cuda.stream(s2):
        h_feat.copy_(d_feat, non_blocking=True)
        h_grad.copy_(d_grad, non_blocking=True)


def dma_in_zc_out():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        big_d.copy_(big_h, non_blocking=True)
    ctc.unstage_rows(d_grad, il, h_grad, stream=s2)


def zc_in_dma_out():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ctc.stage_logits(h_logits, il, out=d_logits, stream=s1)
    with torch.cuda.stream(s2):
        big_h2.copy_(big_d2, non_blocking=True)


timed("DMA host -> device, 256 MiB pinned", big_h.numel(), lambda: big_d.copy_(big_h, non_blocking=True))
timed("DMA device -> host, 256 MiB pinned", big_h.numel(), lambda: big_h.copy_(big_d, non_blocking=True))
timed("DMA both directions at once, 2 x 256 MiB", 2 * big_h.numel(), both)
timed("logits in: DMA of the padded [T,B,V] tensor", T * B * V * 4, lambda: d_logits.copy_(h_logits, non_blocking=True))
timed("logits in: stage_logits (zero copy, valid rows only)", valid, lambda: ctc.stage_logits(h_logits, il, out=d_logits))
timed("gradient out: DMA of the padded [T,B,V] tensor", T * B * V * 4, lambda: h_grad.copy_(d_grad, non_blocking=True))
timed("gradient out: unstage_rows (zero copy, valid rows only)", valid, lambda: ctc.unstage_rows(d_grad, il, h_grad))
timed("features out: DMA", nf * 800, lambda: h_feat.copy_(d_feat, non_blocking=True))
timed("stage_logits || unstage_rows", 2 * valid, both_zero_copy)
timed("step-like: stage_logits || (features DMA + unstage_rows)", 2 * valid + nf * 800, both_e2e_like)
timed("step-like, all DMA (padded tensors)", 2 * T * B * V * 4 + nf * 800, both_dma_like)
timed("step-like: logits DMA (padded) || (features DMA + unstage_rows)", T * B * V * 4 + valid + nf * 800, mixed_x)
timed("step-like: stage_logits || (features DMA + gradient DMA padded)", T * B * V * 4 + valid + nf * 800, mixed_y)
timed("DMA in 256 MiB || unstage_rows", big_h.numel() + valid, dma_in_zc_out)
timed("stage_logits || DMA out 256 MiB", big_h.numel() + valid, zc_in_dma_out)
