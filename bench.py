#!/usr/bin/env python
"""bench.py -- diagram images/s of the img2sgf hot path on N B200s, with the Hough-accumulator
roofline and the reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (CUDA kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference CPU path

Workload (BASELINE.json configs[3]): synthetic 1024x1024 diagrams (s=50, r=24, line threshold
pinned to 150), 1024 images per GPU (8192 over 8 GPUs), the whole path
RGB array -> 19x19 board record, weak scaling over image shards with one NCCL all-gather of
the 384-byte records.  A "step" is one pass of the path over the rank's whole batch, in chunks
of 32 images alternating between 8 CUDA streams.

`value`  : images/s with the inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public API with HOST (pinned) buffers: H2D of every image and
           D2H of the records inside the timed region.
`roofline`: Hough accumulator kernels (k_edge_buckets16 + k_vote_peaks2), algorithmic bytes
           10*W*H per HoughCircles call over their CUDA-event time, taken in a
           second pass of the same K steps on ONE stream (`sections_pass`).
`roofline_config2`: the same measurement on BASELINE.json configs[2] (2048x2048 diagrams), N=1 only.
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
    "synth1024": ("synth1024", 1024, 32),
    "synth2048": ("synth2048", 512, 16),
}
METRIC = "diagram images/sec"
UNIT = "images/s"


# ------------------------------------------------------------------ synthetic inputs (CPU, untimed)
def _gen_one(args):
    from img2sgf_b200 import synth
    config, seed = args
    size, s, r, _ = synth.CONFIGS[config]
    return synth.diagram(size, s, r, seed)


def generate(config: str, start: int, count: int, procs: int):
    from img2sgf_b200 import synth
    size = synth.CONFIGS[config][0]
    imgs = np.empty((count, size, size), np.uint8)
    truths = np.empty((count, 19, 19), np.int8)
    jobs = [(config, start + k) for k in range(count)]
    if procs > 1 and count > 8:
        with mp.get_context("fork").Pool(procs) as pool:
            for k, (g, t) in enumerate(pool.imap(_gen_one, jobs, chunksize=8)):
                imgs[k], truths[k] = g, t
    else:
        for k, j in enumerate(jobs):
            imgs[k], truths[k] = _gen_one(j)
    return imgs, truths


# ------------------------------------------------------------------ reference CPU path
def _cpu_worker(config, seeds, thr, barrier, out):
    """Process `seeds` with the reference's library calls; puts (t_start, t_end, n, kind) on `out`."""
    from img2sgf_b200 import synth
    size, s, r, _ = synth.CONFIGS[config]
    imgs = [synth.to_rgb(synth.diagram(size, s, r, sd)[0]) for sd in seeds]
    kind = "reference"
    try:
        import cv2
        cv2.setNumThreads(1)
        from oracle import ref_replay as R
        run = lambda a: R.run(a, threshold=thr)
    except Exception:
        from oracle import oracle as O
        kind = "port"
        run = lambda a: O.pipeline(a, thr)
    run(imgs[0])                            # untimed: imports, page-in, one-off library initialisation
    barrier.wait()                          # common start line
    t0 = time.perf_counter()
    for a in imgs:
        run(a)
    out.put((t0, time.perf_counter(), len(imgs), kind))


def cpu_reference_rate(config: str, thr: int, images: int, cores: int, first_seed: int = 0):
    """images/s of the reference CPU path with one single-threaded process per core."""
    per = max(1, images // cores)
    workers = min(cores, max(1, images // per))
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(workers), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker,
                         args=(config, list(range(first_seed + k * per, first_seed + (k + 1) * per)), thr, barrier, out))
             for k in range(workers)]
    for p in procs:
        p.start()
    res = [out.get(timeout=1800) for _ in procs]
    for p in procs:
        p.join()
    t0 = min(r[0] for r in res)
    t1 = max(r[1] for r in res)
    n = sum(r[2] for r in res)
    return n / (t1 - t0), n, workers, res[0][3], t1 - t0


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

    # CPU work first (fork-based pools must not run after CUDA is initialised)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = max(8 * cores, 64) if args.cpu_images is None else args.cpu_images
        rate, n, used, kind, dt = cpu_reference_rate(config, thr, sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
               "sample": f"{n} images of the workload, one single-threaded process per core, {dt:.1f} s wall "
                         f"(cv2 {_cv2_version()})"}
    grey, truth = generate(config, rank * per_gpu, per_gpu, max(1, min(32, cores // max(world, 1))))
    # BASELINE.json configs[2] ("synthetic 2048x2048 ... 1 GPU HBM-roofline run"): a second, short
    # Hough-accum roofline measurement on 2048^2 diagrams next to the default workload's (N=1 only)
    grey2k = None
    if world == 1 and args.workload == "synth1024" and not args.no_roofline2048:
        grey2k, _ = generate("synth2048", 0, args.images2048, max(1, min(32, cores)))

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    from img2sgf_b200 import _native as N, batch as B
    import ctypes as C
    lib = N.lib()

    host = torch.empty((per_gpu, size, size, 3), dtype=torch.uint8).pin_memory()
    host.copy_(torch.from_numpy(grey)[..., None].expand(-1, -1, -1, 3))
    dev = host.cuda()
    runner = B.BatchRunner(size, size, chunk, streams=args.streams)
    records = torch.zeros((per_gpu, B.RECORD_BYTES), dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()

    def step_resident():
        runner.run(dev, thr, 128, records=records)
        return B.gather_records(records, total)

    for _ in range(args.warmup):
        full = step_resident()
    torch.cuda.synchronize()
    # correctness of what is being timed: every record valid and equal to the generator's truth
    rec_np = B.records_to_numpy(records)
    bad_status = int((rec_np["status"] != 0).sum())
    wrong = int(sum((rec_np[i]["board"].reshape(19, 19) != truth[i]).any() for i in range(per_gpu)))
    assert full.shape[0] == total

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
    # of another stream sharing the SMs.  (With --streams 1 this pass is identical to the timed one.)
    prof_runner = runner if args.streams == 1 else B.BatchRunner(size, size, chunk, streams=1)
    prof_runner.run(dev, thr, 128, records=records)
    torch.cuda.synchronize()
    nsec = lib.i2s_profile_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        prof_runner.run(dev, thr, 128, records=records)
    p1.record()
    torch.cuda.synchronize()
    ms = (C.c_double * nsec)(); cnt = (C.c_longlong * nsec)()
    N.check(lib.i2s_profile_read(ms, cnt, nsec), "i2s_profile_read")
    lib.i2s_profile_enable(0)
    prof_ms_per_step = p0.elapsed_time(p1) / args.steps

    roof2k = None
    if grey2k is not None:
        size2, _, _, thr2 = synth.CONFIGS["synth2048"]
        n2, chunk2 = grey2k.shape[0], WORKLOADS["synth2048"][2]
        dev2 = torch.from_numpy(grey2k)[..., None].expand(-1, -1, -1, 3).contiguous().cuda()
        run2 = B.BatchRunner(size2, size2, chunk2, streams=1)
        rec2 = torch.zeros((n2, B.RECORD_BYTES), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            run2.run(dev2, thr2, 128, records=rec2)
        torch.cuda.synchronize()
        lib.i2s_profile_enable(1)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            run2.run(dev2, thr2, 128, records=rec2)
        q1.record()
        torch.cuda.synchronize()
        ms2 = (C.c_double * nsec)(); cnt2 = (C.c_longlong * nsec)()
        N.check(lib.i2s_profile_read(ms2, cnt2, nsec), "i2s_profile_read")
        lib.i2s_profile_enable(0)
        names = [lib.i2s_profile_section_name(i).decode() for i in range(nsec)]
        acc2 = sum(ms2[i] for i in range(nsec) if names[i] in ("edge_list", "vote")) / args.steps
        bad2 = int((B.records_to_numpy(rec2)["status"] != 0).sum())
        alg2 = 10.0 * size2 * size2 * 8 * n2
        roof2k = {"workload": f"synth2048: {size2}x{size2} synthetic diagrams (BASELINE.json configs[2]), {n2} images, "
                              f"chunk {chunk2}, one stream, full path", "images_per_s": n2 * args.steps / (q0.elapsed_time(q1) / 1000.0),
                  "bound": "hbm", "achieved": alg2 / (acc2 / 1000.0) / 1e9, "unit": "GB/s",
                  "algorithmic_bytes_per_call": 10 * size2 * size2, "ms_per_step": acc2, "bad_status": bad2}
        del dev2, run2, rec2
        torch.cuda.empty_cache()

    # ---- end to end: pinned host RGB in, host records out, every step
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        engines = runner.engines
        comp = runner.streams                       # None: everything on the current stream
        S = len(engines)
        nbuf = 2 * S                                # two staging buffers per compute stream
        bufs = [torch.empty((chunk, size, size, 3), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        ready = [torch.cuda.Event() for _ in range(nbuf)]
        done = [torch.cuda.Event() for _ in range(nbuf)]
        host_rec = torch.empty((total, B.RECORD_BYTES), dtype=torch.uint8).pin_memory()

        def step_e2e():
            main = torch.cuda.current_stream()
            if comp:
                for st in comp:
                    st.wait_stream(main)
            for k, s in enumerate(range(0, per_gpu, chunk)):
                e = min(per_gpu, s + chunk)
                b = k % nbuf
                with torch.cuda.stream(copy_stream):
                    if k >= nbuf:
                        copy_stream.wait_event(done[b])        # the kernels that read this buffer have finished
                    bufs[b][:e - s].copy_(host[s:e], non_blocking=True)
                    ready[b].record(copy_stream)
                cs = comp[k % S] if comp else main
                cs.wait_event(ready[b])
                with torch.cuda.stream(cs):
                    engines[k % S].run(bufs[b][:e - s], thr, 128, n=e - s, records_out=records[s:e])
                    done[b].record(cs)
            if comp:
                for st in comp:
                    main.wait_stream(st)
            full = B.gather_records(records, total)
            host_rec.copy_(full, non_blocking=True)
            torch.cuda.synchronize()
            return host_rec

        step_e2e()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": total * args.steps / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(per_gpu * size * size * 3) * world,
               "d2h_bytes_per_step": int(total * B.RECORD_BYTES) * world}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        sections = {lib.i2s_profile_section_name(i).decode(): {"ms_per_step": ms[i] / args.steps, "launches": int(cnt[i])}
                    for i in range(nsec) if cnt[i]}
        acc_ms = sum(sections.get(k, {"ms_per_step": 0})["ms_per_step"] for k in ("edge_list", "vote"))
        calls_per_step = 8 * per_gpu                     # unique HoughCircles inputs per image (SURVEY Fact 2)
        alg_bytes = 10.0 * size * size * calls_per_step  # 10*P per call: image+edges read, int32 acc store+load
        achieved = alg_bytes / (acc_ms / 1000.0) / 1e9 if acc_ms > 0 else None
        # DRAM traffic of the two kernels from the committed ncu capture (bytes per HoughCircles call,
        # scaled to the calls one launch of this run processes); null if the capture is absent
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_hough_accum_traffic.json")))
            traffic = float(tr["dram_bytes_per_call"]) * 8 * min(chunk, per_gpu)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {size}x{size} synthetic diagrams, full path RGB->board record, "
                                   f"line threshold {thr}", "images_per_gpu": per_gpu, "global_batch": total,
                       "chunk": chunk, "streams": args.streams, "parallelism": f"image shards x{world}, all-gather of 384-byte records",
                       "l2": "inputs (3 MiB/image x batch) far larger than the 126 MB L2; no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "hough_accum (k_edge_buckets + k_vote_peaks: vote + peak find fused), 8 calls/image",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "traffic_note": "dram__bytes_read+write of both kernels per launch (8 x chunk calls), from profiles/r1_hough_accum_traffic.json",
                         "algorithmic_bytes_per_launch": 10.0 * size * size * 8 * min(chunk, per_gpu),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_call": 10 * size * size, "ms_per_step": acc_ms,
                         "share_of_step": acc_ms / prof_ms_per_step},
            "roofline_config2": (dict(roof2k, peak=peak, frac=roof2k["achieved"] / peak) if roof2k else None),
            "cpu_baseline": cpu,
            "sections": sections,
            "sections_pass": {"streams": 1, "ms_per_step": prof_ms_per_step,
                              "note": "section timers and roofline come from a second pass of the same steps on one stream"},
            "check": {"bad_status": bad_status, "boards_not_equal_truth": wrong, "images_checked": per_gpu},
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
    ap.add_argument("--cpu-images", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline2048", action="store_true")
    ap.add_argument("--images2048", type=int, default=64)
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
