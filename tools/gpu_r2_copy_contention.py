"""Does host->device DMA traffic slow the kernels down?  The device-resident step alone, the same step with
independent H2D copies of the same volume in flight (no dependencies), and run_host()."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from img2sgf_b200 import batch as B, synth

total, chunk, size = 1024, 64, 1024
imgs, _ = synth.batch("synth1024", 0, 64)
rgb = synth.to_rgb(imgs)
host = torch.from_numpy(np.ascontiguousarray(np.tile(rgb, (total // 64, 1, 1, 1)))).pin_memory()
dev = host.cuda()
r = B.BatchRunner(size, size, chunk, streams=8, copy_streams=2)
scratch = [torch.empty((chunk, size, size, 3), dtype=torch.uint8, device="cuda") for _ in range(4)]
cs = [torch.cuda.Stream() for _ in range(2)]

def timed(fn, steps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

def resident():
    r.run(dev, 150, 128); torch.cuda.synchronize()

def resident_with_copies(nbytes_frac=1.0):
    main = torch.cuda.current_stream()
    for c in cs: c.wait_stream(main)
    n = int(total / chunk * nbytes_frac)
    for k in range(n):
        with torch.cuda.stream(cs[k % 2]):
            scratch[k % 4].copy_(host[(k * chunk) % total:(k * chunk) % total + chunk], non_blocking=True)
    r.run(dev, 150, 128)
    torch.cuda.synchronize()

def d2d_with_copies():
    # same, but device->device copies of the same volume (HBM traffic without PCIe)
    main = torch.cuda.current_stream()
    for c in cs: c.wait_stream(main)
    for k in range(total // chunk):
        with torch.cuda.stream(cs[k % 2]):
            scratch[k % 4].copy_(dev[k * chunk:(k + 1) * chunk], non_blocking=True)
    r.run(dev, 150, 128)
    torch.cuda.synchronize()

out = {"resident_ms": timed(resident), "resident_plus_h2d_ms": timed(resident_with_copies),
       "resident_plus_half_h2d_ms": timed(lambda: resident_with_copies(0.5)),
       "resident_plus_d2d_ms": timed(d2d_with_copies),
       "run_host_ms": timed(lambda: r.run_host(host, 150, 128)), "resident_again_ms": timed(resident)}
print(json.dumps({k: round(v, 2) for k, v in out.items()}))
