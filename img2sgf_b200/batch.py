"""Batched / multi-GPU driver of the hot path.

Images are independent units (img2sgf.py:117-204 handles one region at a time), so a batch is
sharded by contiguous image-index ranges, one process per GPU, with NO data-path collective.
The only exchange is one all-gather of the fixed 384-byte per-image record (19x19 board +
grid verdict + counts) after the last kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

RECORD_BYTES = N.RECORD_DTYPE.itemsize


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous image-index range [start, end) of `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _copy_limits(lim):
    return N.Limits(lim.cand_cap, lim.circle_cap, lim.line_cap, lim.hyst_passes)


class Engine:
    """Owns the device workspace for up to `n` images of h x w and runs i2s_pipeline on them."""

    def __init__(self, n: int, h: int, w: int, limits: N.Limits | None = None, taps: bool = False):
        if not torch.cuda.is_available():
            raise N.NativeError("img2sgf_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = N.lib()
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.lim = _copy_limits(limits) if limits is not None else N.default_limits()
        self.ws_bytes = int(self.lib.i2s_pipeline_workspace_bytes(self.n, self.h, self.w, C.byref(self.lim)))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device="cuda")
        self.records = torch.zeros((self.n, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        self.taps = None
        if taps:
            lc, cc = self.lim.line_cap, self.lim.circle_cap
            e = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
            self.taps = {
                "grey": e((n, h, w), torch.uint8), "edges": e((n, h, w), torch.uint8),
                "masked": e((n, h, w), torch.uint8), "circles": e((n, cc, 3), torch.float32),
                "counts": e((n,), torch.int32), "rho": e((n, 2, lc), torch.float32),
                "line_counts": e((n, 2), torch.int32), "grids": e((n, N.GRID_DTYPE.itemsize), torch.uint8),
            }

    def run(self, rgb: torch.Tensor, line_threshold: int, black_threshold: int = 128, n: int | None = None,
            records_out: torch.Tensor | None = None) -> torch.Tensor:
        """Enqueue the whole path for rgb [n,h,w,3] u8 (device) on the current stream.  Returns the
        device record buffer [n,384] u8 (no synchronisation)."""
        n = self.n if n is None else int(n)
        assert 0 <= n <= self.n
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        assert rgb.numel() == n * self.h * self.w * 3
        rec = self.records if records_out is None else records_out
        t = self.taps
        p = lambda k: C.c_void_p(t[k].data_ptr()) if t is not None else None
        rc = self.lib.i2s_pipeline(
            C.c_void_p(rgb.data_ptr()), n, self.h, self.w, int(line_threshold), int(black_threshold),
            C.c_void_p(rec.data_ptr()), p("grey"), p("edges"), p("masked"), p("circles"), p("counts"), p("rho"),
            p("line_counts"), p("grids"), C.byref(self.lim), C.c_void_p(self.ws.data_ptr()), self.ws_bytes,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        N.check(rc, "i2s_pipeline")
        return rec[:n]

    def run_host(self, rgb_host, line_threshold: int, black_threshold: int = 128) -> np.ndarray:
        """Host buffers in, host records out (H2D + kernels + D2H, synchronous)."""
        if isinstance(rgb_host, np.ndarray):
            rgb_host = torch.from_numpy(np.ascontiguousarray(rgb_host, np.uint8))
        n = rgb_host.shape[0]
        dev = rgb_host.cuda(non_blocking=True)
        rec = self.run(dev, line_threshold, black_threshold, n=n)
        return rec.cpu().numpy().view(N.RECORD_DTYPE).reshape(n)

    def taps_host(self) -> dict:
        out = {k: v.cpu().numpy() for k, v in self.taps.items()}
        out["grids"] = out["grids"].view(N.GRID_DTYPE).reshape(self.n)
        return out


class BatchRunner:
    """Runs `total` same-sized images through Engines in chunks of `chunk` images.

    With `streams` > 1 consecutive chunks alternate between that many CUDA streams (one Engine and
    workspace each), so the short low-occupancy kernels at the end of one chunk (circle sort and
    suppression, line peaks, clustering, classification) overlap the wide kernels of the next."""

    def __init__(self, h: int, w: int, chunk: int, limits: N.Limits | None = None, streams: int = 1):
        self.h, self.w, self.chunk = h, w, chunk
        self.engines = [Engine(chunk, h, w, limits) for _ in range(max(1, streams))]
        self.engine = self.engines[0]
        self.streams = [torch.cuda.Stream() for _ in self.engines] if streams > 1 else None

    def run(self, rgb: torch.Tensor, line_threshold: int, black_threshold: int = 128,
            records: torch.Tensor | None = None) -> torch.Tensor:
        """rgb: [total,h,w,3] u8 on the device.  Returns device records [total,384] (async)."""
        total = rgb.shape[0]
        if records is None:
            records = torch.zeros((total, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        cur = torch.cuda.current_stream()
        if self.streams:
            for st in self.streams:
                st.wait_stream(cur)
        for k, s in enumerate(range(0, total, self.chunk)):
            e = min(total, s + self.chunk)
            eng = self.engines[k % len(self.engines)]
            if self.streams:
                with torch.cuda.stream(self.streams[k % len(self.streams)]):
                    eng.run(rgb[s:e], line_threshold, black_threshold, n=e - s, records_out=records[s:e])
            else:
                eng.run(rgb[s:e], line_threshold, black_threshold, n=e - s, records_out=records[s:e])
        if self.streams:
            for st in self.streams:
                cur.wait_stream(st)
        return records

    def launches_per_chunk(self) -> int:
        """Kernel + memset launches i2s_pipeline enqueues for one chunk (for bench.py's gpu_launches)."""
        passes = self.engine.lim.hyst_passes
        # grey, sobel_nms(rgb), hysteresis passes + check, state->edges, gauss, 3 medians,
        # sobel_nms(8 maps), hysteresis passes + check, vote, peaks, radius, finish, stack, mask,
        # line vote, line peaks, cluster, validate, classify
        return 1 + 1 + (passes + 1) + 1 + 1 + 3 + 1 + (passes + 1) + 4 + 1 + 1 + 2 + 1 + 1 + 1


def records_to_numpy(records: torch.Tensor) -> np.ndarray:
    return records.cpu().numpy().view(N.RECORD_DTYPE).reshape(-1)


def gather_records(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-image records of every rank's shard into [total,384] (same on every rank).

    `local` holds this rank's shard_range(total, rank, world) records.  Works with NCCL (device
    tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = -(-total // world)
    pad = torch.zeros((per, RECORD_BYTES), dtype=torch.uint8, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per, RECORD_BYTES), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        s, e = shard_range(total, r, world)
        parts.append(out[r * per:r * per + (e - s)])
    assert shard_range(total, rank, world)[1] - shard_range(total, rank, world)[0] == local.shape[0]
    return torch.cat(parts, 0)


def failed_images(records_np: np.ndarray) -> np.ndarray:
    bits = N.ST_CAND_OVERFLOW | N.ST_CIRCLE_OVERFLOW | N.ST_LINE_OVERFLOW | N.ST_HYST_NOT_CONVERGED
    return np.nonzero(records_np["status"] & bits)[0]


def run_with_retry(runner: BatchRunner, rgb: torch.Tensor, line_threshold: int, black_threshold: int = 128):
    """Batch run, then re-run any image whose status word reports an exceeded limit, one at a time
    with enlarged limits, so the returned records are always valid."""
    from .api import _retrying
    rec = records_to_numpy(runner.run(rgb, line_threshold, black_threshold))
    for i in failed_images(rec):
        def one(lim, i=i):
            eng = Engine(1, runner.h, runner.w, limits=lim)
            r = records_to_numpy(eng.run(rgb[i:i + 1].contiguous(), line_threshold, black_threshold))
            return r[0], int(r[0]["status"])
        rec[i] = _retrying(one, runner.engine.lim)
    return rec
