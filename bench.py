#!/usr/bin/env python
"""bench.py -- diagram images/s of the img2sgf hot path on N B200s, with the Hough-accumulator
roofline and the reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (CUDA kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference CPU path

Workload (BASELINE.json configs[3]): synthetic 1024x1024 diagrams (s=50, r=24, line threshold
pinned to 150), 1024 images per GPU (8192 over 8 GPUs), the whole path
RGB array -> 19x19 board record, weak scaling over image shards with one NCCL all-gather of
the 384-byte records.  A "step" is one pass of the path over the rank's whole batch, in chunks
of 64 images alternating between 8 CUDA streams.  `gpu_launches` counts this library's kernel launches in
the whole timed region (all K steps).

`value`  : images/s with the inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the package's host-buffer call, img2sgf_b200.batch.BatchRunner.run_host():
           pinned HOST RGB in, host records out; H2D of every image and D2H of the records inside the
           timed region.  `copy_ceiling` is the H2D copy of the same buffers alone (no kernels).
`e2e_grey`: beside (never instead of) the RGB line: the exact single-plane entry for greyscale sources.
`roofline`: Hough accumulator kernels (k_edge_list + k_vote_peaks), algorithmic bytes
           10*W*H per HoughCircles call over their CUDA-event time, taken in a
           second pass of the same K steps on ONE stream (`sections_pass`); `roofline_kernels`
           gives the same fraction for the blur / Canny kernels.
`roofline_config2`: the same measurement on BASELINE.json configs[2] (2048x2048 diagrams), N=1 only.
`robustness`: N=1 only -- noisy numbered diagrams (sigma 2) and the reference's 17 real test images
           (ragged batch), each with its CPU rate on the same data and an oracle check of a sample.
`cpu_baseline`: the reference's own cv2/sklearn calls (oracle/ref_replay.py) on all host cores,
           one process per core, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth config, per-GPU batch, chunk)
    "synth1024": ("synth1024", 1024, 64),
    "synth2048": ("synth2048", 512, 16),
}
METRIC = "diagram images/sec"
UNIT = "images/s"
FIXTURE_NAMES = [f"ex{i}" for i in range(1, 18)]          # BASELINE.json configs[1]: the 17 ex*.jpg


# ------------------------------------------------------------------ synthetic inputs (CPU, untimed)
def _gen_one(args):
    from img2sgf_b200 import synth
    config, seed, noise, numbered = args
    size, s, r, _ = synth.CONFIGS[config]
    return synth.diagram(size, s, r, seed, noise=noise, numbered=numbered)


def generate(config: str, start: int, count: int, procs: int, noise: float = 0.0, numbered: bool = False):
    from img2sgf_b200 import synth
    size = synth.CONFIGS[config][0]
    imgs = np.empty((count, size, size), np.uint8)
    truths = np.empty((count, 19, 19), np.int8)
    jobs = [(config, start + k, noise, numbered) for k in range(count)]
    if procs > 1 and count > 8:
        with mp.get_context("fork").Pool(procs) as pool:
            for k, (g, t) in enumerate(pool.imap(_gen_one, jobs, chunksize=4)):
                imgs[k], truths[k] = g, t
    else:
        for k, j in enumerate(jobs):
            imgs[k], truths[k] = _gen_one(j)
    return imgs, truths


def load_fixtures():
    """The 17 contrast-enhanced reference test images as the reference's np.array(region_PIL) gives them
    (RGB order, img2sgf.py:150), from the committed lossless copies under tests/golden/inputs."""
    from PIL import Image
    out = []
    for name in FIXTURE_NAMES:
        a = np.array(Image.open(os.path.join(ROOT, "tests", "golden", "inputs", name + ".png")))
        if a.ndim == 2:
            a = np.repeat(a[..., None], 3, axis=-1)
        out.append(np.ascontiguousarray(a[..., :3], np.uint8))
    return out


# ------------------------------------------------------------------ reference CPU path
def _cpu_runner():
    kind = "reference"
    try:
        import cv2
        cv2.setNumThreads(1)
        from oracle import ref_replay as R
        run = lambda a, thr: R.run(a, threshold=thr)
    except Exception:
        from oracle import oracle as O
        kind = "port"
        run = lambda a, thr: O.pipeline(a, thr)
    return run, kind


def _cpu_worker(spec, thr, barrier, out):
    """Process a list of images with the reference's library calls; puts (t_start, t_end, n, kind) on `out`."""
    from img2sgf_b200 import synth
    if spec[0] == "synth":
        _, config, seeds, noise, numbered = spec
        size, s, r, _ = synth.CONFIGS[config]
        imgs = [synth.to_rgb(synth.diagram(size, s, r, sd, noise=noise, numbered=numbered)[0]) for sd in seeds]
    else:
        imgs = spec[1]
    run, kind = _cpu_runner()
    if imgs:
        run(imgs[0], thr)                   # untimed: imports, page-in, one-off library initialisation
    barrier.wait()                          # common start line
    t0 = time.perf_counter()
    for a in imgs:
        run(a, thr)
    out.put((t0, time.perf_counter(), len(imgs), kind))


def _cpu_rate(specs, thr):
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(len(specs)), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(sp, thr, barrier, out)) for sp in specs]
    for p in procs:
        p.start()
    res = [out.get(timeout=1800) for _ in procs]
    for p in procs:
        p.join()
    t0 = min(r[0] for r in res)
    t1 = max(r[1] for r in res)
    n = sum(r[2] for r in res)
    return n / (t1 - t0), n, len(specs), res[0][3], t1 - t0


def cpu_reference_rate(config: str, thr: int, images: int, cores: int, first_seed: int = 0, noise: float = 0.0,
                       numbered: bool = False):
    """images/s of the reference CPU path with one single-threaded process per core."""
    per = max(1, images // cores)
    workers = min(cores, max(1, images // per))
    specs = [("synth", config, list(range(first_seed + k * per, first_seed + (k + 1) * per)), noise, numbered)
             for k in range(workers)]
    return _cpu_rate(specs, thr)


def cpu_fixture_rate(images, cores: int, rounds: int = 2):
    """The ragged fixture set on the host cores: images dealt round-robin, largest first, `rounds` times."""
    order = sorted(range(len(images)), key=lambda i: -images[i].size)
    workers = min(cores, len(images))
    specs = [("list", [images[i] for i in order[k::workers]] * rounds) for k in range(workers)]
    return _cpu_rate(specs, None)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    from img2sgf_b200 import synth
    config, per_gpu, _ = WORKLOADS[args.workload]
    thr = synth.CONFIGS[config][3]
    cores = os.cpu_count() or 1
    sample = max(8 * cores, 64) if args.cpu_images is None else args.cpu_images
    rates, secs = [], []
    used = cores
    kind = "reference"
    for step in range(args.warmup + args.steps):
        rate, n, used, kind, dt = cpu_reference_rate(config, thr, sample, cores, first_seed=0)
        if step >= args.warmup:
            rates.append(rate); secs.append(dt)
    value = statistics.mean(rates)
    size = synth.CONFIGS[config][0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * statistics.mean(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {size}x{size} synthetic diagrams, full path RGB->board record, "
                               f"line threshold {thr}", "images_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": f"{sample} images of the workload per step, one single-threaded process per core "
                                   f"(cv2 {_cv2_version()}); img2sgf.py:153-198,230-292,420-445,497-543 replayed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _cv2_version():
    try:
        import cv2
        return cv2.__version__
    except Exception:
        return "absent (C port used)"


class Sections:
    """The library's CUDA-event section timers around one or more passes."""

    def __init__(self, lib, N):
        import ctypes as C
        self.lib, self.N, self.C = lib, N, C
        self.n = lib.i2s_profile_enable(0)
        self.names = [lib.i2s_profile_section_name(i).decode() for i in range(self.n)]

    def __enter__(self):
        self.lib.i2s_profile_enable(1)
        return self

    def __exit__(self, *a):
        ms = (self.C.c_double * self.n)(); cnt = (self.C.c_longlong * self.n)()
        self.N.check(self.lib.i2s_profile_read(ms, cnt, self.n), "i2s_profile_read")
        self.lib.i2s_profile_enable(0)
        self.ms = {self.names[i]: ms[i] for i in range(self.n) if cnt[i]}
        self.launches = {self.names[i]: int(cnt[i]) for i in range(self.n) if cnt[i]}

    def per_step(self, steps):
        return {k: {"ms_per_step": v / steps, "launches": self.launches[k]} for k, v in self.ms.items()}


def kernel_rooflines(sections, pixels_per_step, peak):
    """Algorithmic bytes (P = pixels of one image, per step = P x images) over the section time.
    blur: 2 P per Gaussian / median output; Canny: read u8 + write state per map (8 maps; RGB variant reads
    3 P and writes state + grey); Hough accumulator: 10 P per HoughCircles call, 8 calls."""
    table = {"gauss357": (4, ("gauss357",)), "median357": (6, ("median",)), "sobel_nms_8maps": (16, ("sobel_nms",)),
             "sobel_nms_rgb_grey": (5, ("sobel_nms_rgb",)), "hough_accum": (80, ("edge_list", "vote")), "mask": (2, ("mask",)),
             "line_vote": (1, ("line_vote",))}
    out = {}
    for name, (bpp, secs) in table.items():
        ms = sum(sections.get(s, {"ms_per_step": 0})["ms_per_step"] for s in secs)
        if ms > 0:
            ach = bpp * pixels_per_step / (ms / 1000.0) / 1e9
            out[name] = {"bytes_per_pixel": bpp, "ms_per_step": ms, "achieved_gbs": ach, "frac": ach / peak}
    return out


def run_ours(args, rank, world, local_rank):
    from img2sgf_b200 import synth
    config, per_gpu, chunk = WORKLOADS[args.workload]
    if args.per_gpu:
        per_gpu = args.per_gpu
    if args.chunk:
        chunk = args.chunk
    size, _, _, thr = synth.CONFIGS[config]
    total = per_gpu * world
    cores = os.cpu_count() or 1
    extras = world == 1 and args.workload == "synth1024" and not args.no_extras
    gen_procs = max(1, min(32, cores // max(world, 1)))

    # ---- CPU work first (fork-based pools must not run after CUDA is initialised)
    cpu = None
    cpu_noise = cpu_fix = cpu_2k = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = max(8 * cores, 64) if args.cpu_images is None else args.cpu_images
        rate, n, used, kind, dt = cpu_reference_rate(config, thr, sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
               "sample": f"{n} images of the workload, one single-threaded process per core, {dt:.1f} s wall "
                         f"(cv2 {_cv2_version()})"}
    grey, truth = generate(config, rank * per_gpu, per_gpu, gen_procs)
    noisy = fixtures = grey2k = None
    if extras:
        noisy, _ = generate("synth1024", 100000, args.noisy_images, gen_procs, noise=2.0, numbered=True)
        fixtures = load_fixtures()
        grey2k, truth2k = generate("synth2048", 0, args.images2048, gen_procs)
        if not args.no_cpu_baseline:
            r, n, used, kind, dt = cpu_reference_rate("synth1024", 150, max(4 * cores, 32), cores, first_seed=100000,
                                                      noise=2.0, numbered=True)
            cpu_noise = {"value": r, "unit": UNIT, "cores": used, "kind": kind, "sample": f"{n} images, {dt:.1f} s wall"}
            r, n, used, kind, dt = cpu_fixture_rate(fixtures, cores)
            cpu_fix = {"value": r, "unit": UNIT, "cores": used, "kind": kind,
                       "sample": f"the 17 images x2 dealt over {used} processes, {dt:.1f} s wall"}
            r, n, used, kind, dt = cpu_reference_rate("synth2048", 176, max(2 * cores, 16), cores)
            cpu_2k = {"value": r, "unit": UNIT, "cores": used, "kind": kind, "sample": f"{n} images, full path, {dt:.1f} s wall"}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    from img2sgf_b200 import _native as N, batch as B
    lib = N.lib()

    host = torch.empty((per_gpu, size, size, 3), dtype=torch.uint8).pin_memory()
    host.copy_(torch.from_numpy(grey)[..., None].expand(-1, -1, -1, 3))
    dev = host.cuda()
    runner = B.BatchRunner(size, size, chunk, streams=args.streams, copy_streams=args.copy_streams)
    records = torch.zeros((per_gpu, B.RECORD_BYTES), dtype=torch.uint8, device="cuda")
    gather = (lambda r: B.gather_records(r, total)) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()

    def step_resident():
        runner.run(dev, thr, 128, records=records)
        return B.gather_records(records, total)

    for _ in range(args.warmup):
        full = step_resident()
    torch.cuda.synchronize()
    # ---- correctness of what is being timed (untimed): every shard valid, the gathered buffer identical
    # on every rank, boards against the generator's truth, and the oracle's verdict where they differ
    rec_np = B.records_to_numpy(records)
    bad_status = int((rec_np["status"] != 0).sum())
    differ = [i for i in range(per_gpu) if (rec_np[i]["board"].reshape(19, 19) != truth[i]).any()]
    assert full.shape[0] == total
    import hashlib
    digest = hashlib.sha1(full.cpu().numpy().tobytes()).digest()
    ranks_agree = True
    if world > 1:
        mine = torch.tensor(list(digest) + [bad_status], dtype=torch.int64, device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        ranks_agree = all(bool((v[:20] == allv[0][:20]).all()) for v in allv)
        bad_status = int(sum(int(v[20]) for v in allv))
        assert ranks_agree, "gathered record buffers differ between ranks"
    oracle_agrees = None
    if rank == 0 and differ:
        from oracle import oracle as O
        oracle_agrees = True
        for i in differ[:8]:
            res, _, _ = O.pipeline(synth.to_rgb(grey[i]), thr)
            want = np.zeros((19, 19), np.uint8)
            if res.board_ready:
                b = O.board_of(res)
                want[:b.shape[0], :b.shape[1]] = b
            oracle_agrees &= bool((rec_np[i]["board"].reshape(19, 19) == want).all())

    sampler = ClockSampler(local_rank)
    lib.i2s_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize(); barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.i2s_launch_count(1))
    elapsed_ms = e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = total * args.steps / (elapsed_ms / 1000.0)

    # ---- per-kernel-group section timers (roofline): the same K steps again on ONE stream with the
    # library's CUDA-event section timers on, so that a kernel's duration is not stretched by kernels
    # of another stream sharing the SMs.
    prof_runner = runner if args.streams == 1 else B.BatchRunner(size, size, chunk, streams=1)
    prof_runner.run(dev, thr, 128, records=records)
    torch.cuda.synchronize()
    sec = Sections(lib, N)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sec:
        p0.record()
        for _ in range(args.steps):
            prof_runner.run(dev, thr, 128, records=records)
        p1.record()
        torch.cuda.synchronize()
    prof_ms_per_step = p0.elapsed_time(p1) / args.steps
    sections = sec.per_step(args.steps)
    if prof_runner is not runner:
        del prof_runner
        torch.cuda.empty_cache()

    # ---- end to end through the package's host-buffer call
    def timed_host(fn, steps):
        fn()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e = ceiling = e2e_grey = None
    if not args.no_e2e:
        dt = timed_host(lambda: runner.run_host(host, thr, 128, gather=gather), args.steps)
        e2e = {"value": total * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(per_gpu * size * size * 3) * world,
               "d2h_bytes_per_step": int(total * B.RECORD_BYTES) * world,
               "api": "img2sgf_b200.batch.BatchRunner.run_host (pinned host RGB in, host records out)",
               "copy_streams": args.copy_streams}
        host_np = runner.run_host(host, thr, 128, gather=gather)
        assert hashlib.sha1(host_np.tobytes()).digest() == digest, "run_host records differ from the device-resident run"
        # the copies alone: what the box allows for these buffers with this many ranks feeding at once
        def copy_pass():
            runner.copy_only(host)
            torch.cuda.synchronize()
        dtc = timed_host(copy_pass, args.steps)
        gbs = per_gpu * size * size * 3 * args.steps / dtc / 1e9
        ceiling = {"h2d_gb_per_s_per_gpu": gbs, "h2d_gb_per_s_all": gbs * world,
                   "images_per_s_all": total * args.steps / dtc,
                   "note": "the H2D copies of run_host alone (same pinned buffers, staging ring and copy streams, no kernels), "
                           "all ranks at once, slowest rank"}
        if not args.no_grey:
            # beside the RGB line: the diagrams ARE greyscale (R = G = B), so the single-plane entry gives the same records
            hostg = torch.from_numpy(grey).pin_memory()
            grunner = B.BatchRunner(size, size, chunk, streams=args.streams, channels=1, copy_streams=args.copy_streams)
            rg = grunner.run_host(hostg, thr, 128, gather=gather)
            same = bool(rg.tobytes() == host_np.tobytes())
            dtg = timed_host(lambda: grunner.run_host(hostg, thr, 128, gather=gather), args.steps)
            e2e_grey = {"value": total * args.steps / dtg, "unit": UNIT, "h2d_bytes_per_step": int(per_gpu * size * size) * world,
                        "records_identical_to_rgb": same,
                        "note": "i2s_batch_t.channels = 1: exact for greyscale sources (8 of the reference's 17 test images "
                                "are mode L); reported beside, not instead of, the RGB line"}
            del grunner, hostg
            torch.cuda.empty_cache()

    # ---- robustness workloads and configs[2] (N = 1 only; untimed region of the headline)
    robustness = roof2k = None
    if extras:
        robustness = {}
        # (1) noisy numbered diagrams: none of the saturated-content shortcuts applies
        nd = torch.from_numpy(noisy).cuda()[..., None].expand(-1, -1, -1, 3).contiguous()
        nrec = torch.zeros((noisy.shape[0], B.RECORD_BYTES), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            runner.run(nd, 150, 128, records=nrec)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            runner.run(nd, 150, 128, records=nrec)
        q1.record()
        torch.cuda.synchronize()
        rate_noisy = noisy.shape[0] * args.steps / (q0.elapsed_time(q1) / 1000.0)
        one = B.BatchRunner(size, size, chunk, streams=1)
        one.run(nd, 150, 128, records=nrec)
        torch.cuda.synchronize()
        sn = Sections(lib, N)
        with sn:
            one.run(nd, 150, 128, records=nrec)
            torch.cuda.synchronize()
        nrec_np = B.records_to_numpy(nrec)
        from oracle import oracle as O
        def oracle_ok(rec, rgb, th):
            res, _, _ = O.pipeline(rgb, th)
            want = np.zeros((19, 19), np.uint8)
            if res.board_ready:
                b = O.board_of(res)
                want[:b.shape[0], :b.shape[1]] = b
            return bool(rec["status"] == 0 and rec["n_circles"] == res.n_circles and
                        (rec["board"].reshape(19, 19) == want).all())
        robustness["synth1024_noise2"] = {
            "workload": f"{noisy.shape[0]} numbered-stone 1024x1024 diagrams with Gaussian pixel noise sigma 2 (SURVEY 8d config 5 recipe), threshold 150",
            "images_per_s": rate_noisy, "cpu_baseline": cpu_noise, "bad_status": int((nrec_np["status"] != 0).sum()),
            "oracle_check": {"images": 4, "identical": all(oracle_ok(nrec_np[i], synth.to_rgb(noisy[i]), 150) for i in range(4))},
            "sections_ms_per_image": {k: v / noisy.shape[0] for k, v in sn.ms.items()}}
        del nd, nrec, one
        torch.cuda.empty_cache()
        # (2) the reference's 17 real test images, tiled, as ragged batches through process_images()
        reps = max(1, args.fixture_images // len(fixtures))
        tiled = fixtures * reps
        rr = B.RaggedRunner(streams=4, max_group=24)
        frec = rr.process_images(tiled)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            frec = rr.process_images(tiled)
        dtf = (time.perf_counter() - t0) / args.steps
        sf = Sections(lib, N)
        with sf:
            rr.process_images(fixtures)
            torch.cuda.synchronize()
        same_tiles = all(frec[k].tobytes() == frec[k % len(fixtures)].tobytes() for k in range(len(tiled)))
        robustness["fixtures"] = {
            "workload": f"the reference's 17 test images (239x175 .. 1265x1245, 9.6 Mpx) x{reps} = {len(tiled)} images, ragged batches, "
                        "auto line threshold, end to end (host arrays in, host records out, packing included)",
            "images_per_s": len(tiled) / dtf, "mpx_per_s": sum(a.shape[0] * a.shape[1] for a in tiled) / dtf / 1e6,
            "cpu_baseline": cpu_fix, "bad_status": int((frec["status"] != 0).sum()), "tiles_identical": same_tiles,
            "oracle_check": {"images": 17, "identical": all(oracle_ok(frec[i], fixtures[i], None) for i in range(len(fixtures)))},
            "sections_ms_per_17_images": dict(sf.ms)}
        del rr
        torch.cuda.empty_cache()
        # (3) BASELINE.json configs[2]: 2048x2048 diagrams, the Hough circle + line stages
        size2, _, _, thr2 = synth.CONFIGS["synth2048"]
        n2, chunk2 = grey2k.shape[0], WORKLOADS["synth2048"][2]
        dev2 = torch.from_numpy(grey2k).cuda()[..., None].expand(-1, -1, -1, 3).contiguous()
        run2 = B.BatchRunner(size2, size2, chunk2, streams=4)
        rec2 = torch.zeros((n2, B.RECORD_BYTES), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            run2.run(dev2, thr2, 128, records=rec2)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            run2.run(dev2, thr2, 128, records=rec2)
        q1.record()
        torch.cuda.synchronize()
        rate2 = n2 * args.steps / (q0.elapsed_time(q1) / 1000.0)
        del run2
        torch.cuda.empty_cache()
        run2 = B.BatchRunner(size2, size2, chunk2, streams=1)
        run2.run(dev2, thr2, 128, records=rec2)
        torch.cuda.synchronize()
        s2 = Sections(lib, N)
        with s2:
            for _ in range(args.steps):
                run2.run(dev2, thr2, 128, records=rec2)
            torch.cuda.synchronize()
        rec2_np = B.records_to_numpy(rec2)
        sec2 = s2.per_step(args.steps)
        wrong2 = int(sum((rec2_np[i]["board"].reshape(19, 19) != truth2k[i]).any() for i in range(n2)))
        roof2k = {"workload": f"synth2048: {size2}x{size2} synthetic diagrams (BASELINE.json configs[2]), {n2} images, chunk {chunk2}",
                  "images_per_s": rate2, "cpu_baseline": cpu_2k, "bad_status": int((rec2_np["status"] != 0).sum()),
                  "boards_not_equal_truth": wrong2, "sections": sec2,
                  "stages_a2_a8_ms_per_step": sum(v["ms_per_step"] for k, v in sec2.items()
                                                  if k not in ("cluster", "validate", "classify", "grey", "enhance"))}
        del dev2, run2, rec2
        torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        acc_ms = sum(sections.get(k, {"ms_per_step": 0})["ms_per_step"] for k in ("edge_list", "vote"))
        calls_per_step = 8 * per_gpu                     # unique HoughCircles inputs per image (SURVEY Fact 2)
        alg_bytes = 10.0 * size * size * calls_per_step  # 10*P per call: image+edges read, int32 acc store+load
        achieved = alg_bytes / (acc_ms / 1000.0) / 1e9 if acc_ms > 0 else None
        # DRAM traffic of the two kernels from the committed ncu capture (bytes per HoughCircles call,
        # scaled to the calls one launch of this run processes); null if the capture is absent
        traffic = traffic_src = None
        for name in ("r2_hough_accum_traffic.json", "r1_hough_accum_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic = float(tr["dram_bytes_per_call"]) * 8 * min(chunk, per_gpu)
                traffic_src = "profiles/" + name
                break
            except Exception:
                pass
        if roof2k:
            a2 = sum(roof2k["sections"].get(k, {"ms_per_step": 0})["ms_per_step"] for k in ("edge_list", "vote"))
            ach2 = 10.0 * 2048 * 2048 * 8 * args.images2048 / (a2 / 1000.0) / 1e9
            roof2k.update({"bound": "hbm", "achieved": ach2, "peak": peak, "frac": ach2 / peak, "unit": "GB/s",
                           "algorithmic_bytes_per_call": 10 * 2048 * 2048, "hough_accum_ms_per_step": a2,
                           "roofline_kernels": kernel_rooflines(roof2k["sections"], 2048.0 * 2048 * args.images2048, peak)})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {size}x{size} synthetic diagrams, full path RGB->board record, "
                                   f"line threshold {thr}", "images_per_gpu": per_gpu, "global_batch": total,
                       "chunk": chunk, "streams": args.streams, "parallelism": f"image shards x{world}, all-gather of 384-byte records",
                       "l2": "inputs (3 MiB/image x batch) far larger than the 126 MB L2; no flush needed"},
            "clocks": clocks, "e2e": e2e, "copy_ceiling": ceiling, "e2e_grey": e2e_grey, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "hough_accum (k_edge_list + k_vote_peaks: vote + peak find fused), 8 calls/image",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "traffic_note": f"dram__bytes_read+write of both kernels per launch (8 x chunk calls), from {traffic_src}",
                         "algorithmic_bytes_per_launch": 10.0 * size * size * 8 * min(chunk, per_gpu),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_call": 10 * size * size, "ms_per_step": acc_ms,
                         "share_of_step": acc_ms / prof_ms_per_step},
            "roofline_kernels": kernel_rooflines(sections, float(size) * size * per_gpu, peak),
            "roofline_config2": roof2k,
            "robustness": robustness,
            "cpu_baseline": cpu,
            "sections": sections,
            "sections_pass": {"streams": 1, "ms_per_step": prof_ms_per_step,
                              "note": "section timers and roofline come from a second pass of the same steps on one stream"},
            "check": {"bad_status": bad_status, "boards_not_equal_truth": len(differ), "differing_seeds": [rank * per_gpu + i for i in differ[:16]],
                      "oracle_agrees_on_differing": oracle_agrees, "images_checked": per_gpu, "ranks_agree": ranks_agree},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth1024", choices=sorted(WORKLOADS))
    ap.add_argument("--per-gpu", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--copy-streams", type=int, default=2)
    ap.add_argument("--cpu-images", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the robustness / configs[2] blocks (N=1)")
    ap.add_argument("--no-grey", action="store_true", help="skip the single-plane (greyscale source) end-to-end block")
    ap.add_argument("--images2048", type=int, default=512)
    ap.add_argument("--noisy-images", type=int, default=256)
    ap.add_argument("--fixture-images", type=int, default=2040)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
