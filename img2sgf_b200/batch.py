"""Batched / multi-GPU driver of the hot path.

Images are independent units (img2sgf.py:117-204 handles one region at a time), so a batch is
sharded across GPUs, one process per GPU, with NO data-path collective.  The only exchange is
one all-gather of the fixed 384-byte per-image record (19x19 board + grid verdict + counts)
after the last kernel.

    Engine        one device workspace; runs i2s_pipeline on a batch (one size or ragged)
    BatchRunner   same-sized images in chunks over several CUDA streams:
                    run()       device-resident input  -> device records
                    run_host()  pinned host input      -> host records, H2D copies pipelined with compute
    RaggedRunner  images of different sizes (the reference's test images are 110x102 .. 1265x1245):
                    process_images(list of arrays) -> records, grouped by size, balanced by pixel count
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


def shard_by_pixels(sizes, world: int) -> list[list[int]]:
    """Ragged batches are balanced by pixel count, not image count: largest
    image first, each to the rank with the fewest pixels so far.  Returns the image indices of
    every rank (ascending); deterministic, so every rank computes the same assignment."""
    px = [int(h) * int(w) for h, w in sizes]
    order = sorted(range(len(px)), key=lambda i: (-px[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += px[i]
    return [sorted(o) for o in out]


RETRY_BITS = N.ST_CAND_OVERFLOW | N.ST_CIRCLE_OVERFLOW | N.ST_LINE_OVERFLOW | N.ST_HYST_NOT_CONVERGED


def group_by_size(sizes, max_group: int = 32, max_waste: float = 0.35) -> list[list[int]]:
    """Groups of images that share one canvas (max h x max w of the group): images sorted by area,
    a group is closed when adding the next image would leave more than `max_waste` of the canvas
    area unused, or at `max_group` images."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i][0]) * int(sizes[i][1]), i))
    groups, cur, ch, cw, used = [], [], 0, 0, 0
    for i in order:
        h, w = int(sizes[i][0]), int(sizes[i][1])
        nh, nw = max(ch, h), max(cw, w)
        if cur and (len(cur) >= max_group or (used + h * w) < (1.0 - max_waste) * nh * nw * (len(cur) + 1)):
            groups.append(cur)
            cur, nh, nw, used = [], h, w, 0
        cur.append(i)
        ch, cw, used = nh, nw, used + h * w
    if cur:
        groups.append(cur)
    return groups


def _copy_limits(lim):
    return N.Limits(lim.cand_cap, lim.circle_cap, lim.line_cap, lim.hyst_passes)


def make_params(line_threshold=None, black_threshold: int = 128, canny=(50, 200), contrast_factor: float = 1.0,
                brightness_factor: float = 1.0) -> N.Params:
    """line_threshold None / 0 = choose_threshold() per image (img2sgf.py:606-613)."""
    p = N.default_params()
    p.line_threshold = int(line_threshold or 0)
    p.black_threshold = int(black_threshold)
    p.canny_low, p.canny_high = int(canny[0]), int(canny[1])
    p.contrast_factor, p.brightness_factor = float(contrast_factor), float(brightness_factor)
    return p


class Engine:
    """Owns the device workspace for up to `n` images on an h x w canvas and runs i2s_pipeline."""

    def __init__(self, n: int, h: int, w: int, limits: N.Limits | None = None, taps: bool = False):
        if not torch.cuda.is_available():
            raise N.NativeError("img2sgf_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = N.lib()
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.pitch = N.canvas_pitch(self.w)
        self.lim = _copy_limits(limits) if limits is not None else N.default_limits()
        self.ws_bytes = int(self.lib.i2s_pipeline_workspace_bytes(self.n, self.h, self.w, C.byref(self.lim)))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device="cuda")
        self.records = torch.zeros((self.n, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        self.taps = None
        self._taps_struct = None
        if taps:
            lc, cc, P = self.lim.line_cap, self.lim.circle_cap, self.pitch
            e = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
            t = self.taps = {
                "grey": e((n, h, P), torch.uint8), "edges": e((n, h, P), torch.uint8),
                "masked": e((n, h, P), torch.uint8), "circles": e((n, cc, 3), torch.float32),
                "counts": e((n,), torch.int32), "rho": e((n, 2, lc), torch.float32),
                "line_counts": e((n, 2), torch.int32), "grids": e((n, N.GRID_DTYPE.itemsize), torch.uint8),
                "brightness": e((n, N.BOARD_SIZE * N.BOARD_SIZE), torch.float64),
            }
            self._taps_struct = N.Taps(P, 0, *(t[k].data_ptr() for k in ("grey", "edges", "masked", "circles", "counts",
                                                                         "rho", "line_counts", "grids", "brightness")))

    def run(self, src: torch.Tensor, params: N.Params | int, black_threshold: int = 128, n: int | None = None,
            records_out: torch.Tensor | None = None, channels: int = 3, images: torch.Tensor | None = None,
            pitch: int = 0) -> torch.Tensor:
        """Enqueue the whole path on the current stream (no synchronisation).

        Uniform batch: src = [n,h,w,3] (or [n,h,w] with channels=1) u8 on the device.  Ragged batch:
        src = the packed bytes, `images` = device tensor holding n i2s_image_t descriptors.  `params` is
        an i2s_params_t or, for short, the line threshold.  Returns the device record buffer [n,384]."""
        n = self.n if n is None else int(n)
        assert 0 <= n <= self.n
        assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
        if images is None:
            assert src.numel() >= n * self.h * max(pitch, self.w * channels)
        if not isinstance(params, N.Params):
            params = make_params(params, black_threshold)
        rec = self.records if records_out is None else records_out
        batch = N.Batch(n, channels, self.h, self.w, int(pitch), 0, images.data_ptr() if images is not None else None)
        rc = self.lib.i2s_pipeline(
            C.c_void_p(src.data_ptr()), C.byref(batch), C.byref(params), C.c_void_p(rec.data_ptr()),
            C.byref(self._taps_struct) if self._taps_struct is not None else None, C.byref(self.lim),
            C.c_void_p(self.ws.data_ptr()), self.ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        N.check(rc, "i2s_pipeline")
        return rec[:n]

    def run_host(self, rgb_host, params, black_threshold: int = 128, channels: int = 3) -> np.ndarray:
        """Host array in, host records out (one H2D, the kernels, one D2H; synchronous)."""
        if isinstance(rgb_host, np.ndarray):
            rgb_host = torch.from_numpy(np.ascontiguousarray(rgb_host, np.uint8))
        n = rgb_host.shape[0]
        dev = rgb_host.cuda(non_blocking=True)
        rec = self.run(dev, params, black_threshold, n=n, channels=channels)
        return rec.cpu().numpy().view(N.RECORD_DTYPE).reshape(n)

    def taps_host(self) -> dict:
        out = {k: v.cpu().numpy() for k, v in self.taps.items()}
        out["grids"] = out["grids"].view(N.GRID_DTYPE).reshape(self.n)
        for k in ("grey", "edges", "masked"):
            out[k] = out[k][:, :, :self.w]
        return out


class BatchRunner:
    """Runs `total` same-sized images through Engines in chunks of `chunk` images.

    With `streams` > 1 consecutive chunks alternate between that many CUDA streams (one Engine and
    workspace each), so the short low-occupancy kernels at the end of one chunk (circle sort and
    suppression, line peaks, clustering, classification) overlap the wide kernels of the next."""

    def __init__(self, h: int, w: int, chunk: int, limits: N.Limits | None = None, streams: int = 1, channels: int = 3,
                 copy_streams: int = 2):
        self.h, self.w, self.chunk, self.channels = h, w, chunk, channels
        self.engines = [Engine(chunk, h, w, limits) for _ in range(max(1, streams))]
        self.engine = self.engines[0]
        self.streams = [torch.cuda.Stream() for _ in self.engines] if streams > 1 else None
        self.n_copy_streams = max(1, copy_streams)
        self._staging = None            # device staging buffers + events of run_host, created on first use

    def run(self, rgb: torch.Tensor, line_threshold, black_threshold: int = 128,
            records: torch.Tensor | None = None) -> torch.Tensor:
        """rgb: [total,h,w,3] (or [total,h,w] for channels=1) u8 on the device.  Returns device records
        [total,384] (asynchronous)."""
        total = rgb.shape[0]
        if records is None:
            records = torch.zeros((total, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        params = line_threshold if isinstance(line_threshold, N.Params) else make_params(line_threshold, black_threshold)
        cur = torch.cuda.current_stream()
        if self.streams:
            for st in self.streams:
                st.wait_stream(cur)
        for k, s in enumerate(range(0, total, self.chunk)):
            e = min(total, s + self.chunk)
            eng = self.engines[k % len(self.engines)]
            if self.streams:
                with torch.cuda.stream(self.streams[k % len(self.streams)]):
                    eng.run(rgb[s:e], params, n=e - s, records_out=records[s:e], channels=self.channels)
            else:
                eng.run(rgb[s:e], params, n=e - s, records_out=records[s:e], channels=self.channels)
        if self.streams:
            for st in self.streams:
                cur.wait_stream(st)
        return records

    # ---- host buffers in, host records out ---------------------------------------------------------
    def _ensure_staging(self):
        if self._staging is None:
            S = len(self.engines)
            nbuf = 2 * S                                    # two staging buffers per compute stream
            shape = (self.chunk, self.h, self.w, 3) if self.channels == 3 else (self.chunk, self.h, self.w)
            self._staging = {
                "bufs": [torch.empty(shape, dtype=torch.uint8, device="cuda") for _ in range(nbuf)],
                "ready": [torch.cuda.Event() for _ in range(nbuf)],
                "done": [torch.cuda.Event() for _ in range(nbuf)],
                "copy": [torch.cuda.Stream() for _ in range(self.n_copy_streams)],
                "records": None, "host_records": None,
            }
        return self._staging

    def run_host(self, host: torch.Tensor, line_threshold, black_threshold: int = 128,
                 gather=None, host_records: torch.Tensor | None = None) -> np.ndarray:
        """The reference-facing call for a batch: `host` = [total,h,w,3] (or [total,h,w]) u8 in HOST memory
        (pinned for full speed; a pageable tensor or numpy array is pinned first).  Chunks are copied
        host->device on dedicated copy streams into a ring of staging buffers while earlier chunks
        compute; the records come back in one device->host copy.  `gather`: optional callable
        records_device -> records_device (the multi-GPU all-gather).  Returns the host records
        (numpy structured array) after synchronising."""
        if isinstance(host, np.ndarray):
            host = torch.from_numpy(np.ascontiguousarray(host, np.uint8))
        if not host.is_pinned():
            host = host.pin_memory()
        total = host.shape[0]
        st = self._ensure_staging()
        bufs, ready, done, copies = st["bufs"], st["ready"], st["done"], st["copy"]
        nbuf, S = len(bufs), len(self.engines)
        if st["records"] is None or st["records"].shape[0] != total:
            st["records"] = torch.zeros((total, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        records = st["records"]
        params = line_threshold if isinstance(line_threshold, N.Params) else make_params(line_threshold, black_threshold)
        main = torch.cuda.current_stream()
        comp = self.streams
        for c in copies:
            c.wait_stream(main)
        if comp:
            for cs in comp:
                cs.wait_stream(main)
        for k, s in enumerate(range(0, total, self.chunk)):
            e = min(total, s + self.chunk)
            b = k % nbuf
            cstream = copies[k % len(copies)]
            with torch.cuda.stream(cstream):
                if k >= nbuf:
                    cstream.wait_event(done[b])            # the kernels that read this buffer have finished
                bufs[b][:e - s].copy_(host[s:e], non_blocking=True)
                ready[b].record(cstream)
            cs = comp[k % S] if comp else main
            cs.wait_event(ready[b])
            with torch.cuda.stream(cs):
                self.engines[k % S].run(bufs[b][:e - s], params, n=e - s, records_out=records[s:e], channels=self.channels)
                done[b].record(cs)
        if comp:
            for cs in comp:
                main.wait_stream(cs)
        full = gather(records) if gather is not None else records
        if host_records is None:
            if st["host_records"] is None or st["host_records"].shape[0] != full.shape[0]:
                st["host_records"] = torch.empty((full.shape[0], RECORD_BYTES), dtype=torch.uint8).pin_memory()
            host_records = st["host_records"]
        host_records.copy_(full, non_blocking=True)
        torch.cuda.synchronize()
        return host_records.numpy().view(N.RECORD_DTYPE).reshape(-1)

    def copy_only(self, host: torch.Tensor) -> None:
        """The host->device copies of run_host() alone (same staging ring, same copy streams, no kernels):
        the ceiling of the end-to-end rate on this box.  Asynchronous; synchronise to time it."""
        st = self._ensure_staging()
        bufs, copies = st["bufs"], st["copy"]
        main = torch.cuda.current_stream()
        for c in copies:
            c.wait_stream(main)
        for k, s in enumerate(range(0, host.shape[0], self.chunk)):
            e = min(host.shape[0], s + self.chunk)
            with torch.cuda.stream(copies[k % len(copies)]):
                bufs[k % len(bufs)][:e - s].copy_(host[s:e], non_blocking=True)
        for c in copies:
            main.wait_stream(c)


class RaggedRunner:
    """Images of different sizes.  Images are grouped by size (group_by_size); a group shares one
    canvas (max h x max w) in the library's workspace, is packed into one pinned staging buffer
    together with its i2s_image_t descriptors, copied with ONE host->device transfer and processed by
    ONE i2s_pipeline call; groups alternate between CUDA streams.  Engines are cached per canvas size."""

    def __init__(self, limits: N.Limits | None = None, streams: int = 2, max_group: int = 32, pack_threads: int = 8):
        if not torch.cuda.is_available():
            raise N.NativeError("img2sgf_b200 needs a CUDA device (there is no CPU fallback)")
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=pack_threads) if pack_threads > 1 else None
        self.limits = limits
        self.max_group = max_group
        self.streams = [torch.cuda.Stream() for _ in range(max(1, streams))]
        self._engines = {}              # stream slot -> Engine (grows to the largest group seen)
        self._pinned = {}               # stream slot -> pinned staging tensor
        self._device = {}               # stream slot -> device staging tensor
        self._hostrec = {}              # stream slot -> pinned record buffer

    @staticmethod
    def _round(v, a=64):
        return (int(v) + a - 1) // a * a

    def _engine(self, slot, n, h, w):
        """One engine per stream slot, grown (re-allocated) when a group needs more images or a larger
        canvas; a ragged batch may run on a canvas larger than its own."""
        eng = self._engines.get(slot)
        if eng is None or eng.n < n or eng.h < h or eng.w < w:
            need = (max(self._round(n, 8), eng.n if eng else 0), max(self._round(h), eng.h if eng else 0),
                    max(self._round(w), eng.w if eng else 0))
            self._engines.pop(slot, None)
            del eng
            eng = self._engines[slot] = Engine(need[0], need[1], need[2], self.limits)
            self._hostrec[slot] = torch.empty((need[0], RECORD_BYTES), dtype=torch.uint8).pin_memory()
        return eng

    def _buffers(self, slot, nbytes):
        if slot not in self._pinned or self._pinned[slot].numel() < nbytes:
            cap = max(nbytes, 1 << 20)
            self._pinned[slot] = torch.empty(cap, dtype=torch.uint8).pin_memory()
            self._device[slot] = torch.empty(cap, dtype=torch.uint8, device="cuda")
        return self._pinned[slot], self._device[slot]

    @staticmethod
    def pack(images, idx, thresholds=None):
        """Layout of one group: [descriptors | image 0 | image 1 ...]; rows padded to a multiple of 4 bytes
        and image starts to 16, so the kernels take their aligned paths.  Returns (nbytes, descriptors,
        channels, fill) where fill(buffer) writes the group into a uint8 numpy view."""
        ch = 1 if images[idx[0]].ndim == 2 else 3
        desc = np.zeros(len(idx), N.IMAGE_DTYPE)
        off = (desc.nbytes + 255) // 256 * 256
        for k, i in enumerate(idx):
            a = images[i]
            assert a.dtype == np.uint8 and (a.ndim == 2) == (ch == 1), "one channel layout per group"
            h, w = a.shape[:2]
            pitch = (w * ch + 3) // 4 * 4
            desc[k] = (off, h, w, pitch, int(thresholds[i]) if thresholds is not None else 0)
            off = (off + h * pitch + 15) // 16 * 16

        def fill_one(buf, k):
            a = images[idx[k]]
            h, w = a.shape[:2]
            o, pitch = int(desc[k]["offset"]), int(desc[k]["pitch"])
            buf[o:o + h * pitch].reshape(h, pitch)[:, :w * ch] = a.reshape(h, w * ch)

        def fill(buf, pool=None):
            buf[:desc.nbytes] = desc.view(np.uint8)
            if pool is None:
                for k in range(len(idx)):
                    fill_one(buf, k)
            else:                                          # numpy copies release the GIL: pack images in parallel
                list(pool.map(lambda k: fill_one(buf, k), range(len(idx))))
        return off, desc, ch, fill

    def process_images(self, images, line_threshold=None, black_threshold: int = 128, thresholds=None,
                       contrast_factor: float = 1.0, brightness_factor: float = 1.0) -> np.ndarray:
        """images: list of u8 arrays, [h,w,3] (RGB order) or [h,w] (greyscale sources), any sizes.
        Returns one record per image, in input order.  line_threshold None = choose_threshold() per image;
        `thresholds` = optional per-image slider values.

        Images whose status word reports an exceeded limit (candidates, circles, lines, hysteresis passes) are
        re-run with enlarged limits, and the runner keeps the enlarged limits for later calls, so the
        returned records are valid (status 0, or I2S_ST_GRID_OVERFLOW for a grid the record cannot hold)."""
        from .api import _grow
        kw = dict(line_threshold=line_threshold, black_threshold=black_threshold, contrast_factor=contrast_factor,
                  brightness_factor=brightness_factor)
        out = self._process(images, thresholds=thresholds, **kw)
        for _ in range(8):
            bad = np.nonzero(out["status"] & RETRY_BITS)[0]
            if len(bad) == 0:
                return out
            lim = self.limits if self.limits is not None else N.default_limits()
            self.limits = _grow(lim, int(np.bitwise_or.reduce(out["status"][bad]) & RETRY_BITS))
            self._engines.clear()
            torch.cuda.empty_cache()
            sub_thr = [thresholds[i] for i in bad] if thresholds is not None else None
            out[bad] = self._process([images[i] for i in bad], thresholds=sub_thr, **kw)
        raise N.NativeError("retry budget exhausted: " + N.describe_status(int(out["status"].max())))

    def _process(self, images, line_threshold=None, black_threshold: int = 128, thresholds=None,
                 contrast_factor: float = 1.0, brightness_factor: float = 1.0) -> np.ndarray:
        total = len(images)
        out = np.zeros(total, N.RECORD_DTYPE)
        if total == 0:
            return out
        rgb_idx = [i for i in range(total) if images[i].ndim == 3]
        grey_idx = [i for i in range(total) if images[i].ndim == 2]
        groups = []
        for sub in (rgb_idx, grey_idx):
            sizes = [images[i].shape[:2] for i in sub]
            groups += [[sub[j] for j in g] for g in group_by_size(sizes, self.max_group)]
        params = make_params(line_threshold, black_threshold, contrast_factor=contrast_factor,
                             brightness_factor=brightness_factor)
        main = torch.cuda.current_stream()
        inflight = {}                                    # slot -> (image indices, event)

        def drain(slot):
            idx, ev = inflight.pop(slot)
            ev.synchronize()
            out[idx] = self._hostrec[slot][:len(idx)].numpy().reshape(-1).view(N.RECORD_DTYPE)

        for gi, idx in enumerate(groups):
            slot = gi % len(self.streams)
            if slot in inflight:
                drain(slot)                              # the slot's staging and record buffers are free again
            nbytes, desc, ch, fill = self.pack(images, idx, thresholds)
            pinned, dev = self._buffers(slot, nbytes)
            fill(pinned.numpy(), self._pool)
            eng = self._engine(slot, len(idx), int(desc["h"].max()), int(desc["w"].max()))
            stream = self.streams[slot]
            stream.wait_stream(main)
            with torch.cuda.stream(stream):
                dev[:nbytes].copy_(pinned[:nbytes], non_blocking=True)
                rec = eng.run(dev, params, n=len(idx), channels=ch, images=dev[:desc.nbytes])
                self._hostrec[slot][:len(idx)].copy_(rec, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            inflight[slot] = (idx, ev)
        for slot in list(inflight):
            drain(slot)
        return out


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


def gather_ragged_records(local: np.ndarray, assignment: list[list[int]], group=None) -> np.ndarray:
    """Ragged multi-GPU batches: rank r processed the images `assignment[r]` (shard_by_pixels); returns
    the records of ALL images in input order on every rank (one all-gather of padded shards)."""
    import torch.distributed as dist
    total = sum(len(a) for a in assignment)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = np.zeros(total, N.RECORD_DTYPE)
    if world == 1:
        out[assignment[0]] = local
        return out
    rank = dist.get_rank(group)
    per = max(len(a) for a in assignment)
    use_cuda = dist.get_backend(group) == "nccl"
    pad = torch.zeros((per, RECORD_BYTES), dtype=torch.uint8)
    pad[:len(local)] = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(len(local), RECORD_BYTES))
    if use_cuda:
        pad = pad.cuda()
    buf = torch.empty((world * per, RECORD_BYTES), dtype=torch.uint8, device=pad.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    allrec = buf.cpu().numpy()
    for r, idx in enumerate(assignment):
        out[idx] = allrec[r * per:r * per + len(idx)].reshape(-1).view(N.RECORD_DTYPE)
    assert len(assignment[rank]) == len(local)
    return out


def failed_images(records_np: np.ndarray) -> np.ndarray:
    """Images whose record is not valid: a limit was exceeded (retry with larger limits), or the grid
    has more lines than the fixed-size record holds (I2S_ST_GRID_OVERFLOW)."""
    return np.nonzero(records_np["status"] & (RETRY_BITS | N.ST_GRID_OVERFLOW))[0]


def run_with_retry(runner: BatchRunner, rgb: torch.Tensor, line_threshold, black_threshold: int = 128):
    """Batch run, then re-run any image whose status word reports an exceeded limit, one at a time
    with enlarged limits, so the returned records are always valid.  A grid with more than 32 lines
    on an axis cannot be represented in the record: that raises."""
    from .api import _retrying
    rec = records_to_numpy(runner.run(rgb, line_threshold, black_threshold))
    for i in np.nonzero(rec["status"] & RETRY_BITS)[0]:
        def one(lim, i=i):
            eng = Engine(1, runner.h, runner.w, limits=lim)
            r = records_to_numpy(eng.run(rgb[i:i + 1].contiguous(), line_threshold, black_threshold,
                                         channels=runner.channels))
            return r[0], int(r[0]["status"])
        rec[i] = _retrying(one, runner.engine.lim)
    over = np.nonzero(rec["status"] & N.ST_GRID_OVERFLOW)[0]
    if len(over):
        raise N.NativeError(f"images {over.tolist()}: more than {N.MAX_GRID} grid lines on an axis "
                            "(the reference reports 'too many lines' for anything above 19, img2sgf.py:568-571)")
    return rec
