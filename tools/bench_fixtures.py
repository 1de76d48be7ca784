"""BASELINE.json configs[1]: the 17 reference test images (contrast-enhanced arrays committed under
tests/golden/inputs), one at a time through api.process_image (the call the GUI shim makes) next to
the replay of the reference's own library calls on one host core.  Prints one JSON line.
Latency of a single ragged image, not throughput: every call allocates its workspace, copies the
image in and the results out, and synchronises."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from conftest import load_input
    from img2sgf_b200 import api
    from oracle import ref_replay as R
    names = [f"ex{i}" for i in range(1, 18)]
    imgs = [load_input(n) for n in names]
    thr = [R.choose_threshold(a.shape[1], a.shape[0]) for a in imgs]
    for a, t in zip(imgs[:3], thr[:3]):
        api.process_image(a, t)                        # warm-up: library load, first-launch costs
    torch.cuda.synchronize()
    gpu = []
    for rep in range(3):
        t0 = time.perf_counter()
        for a, t in zip(imgs, thr):
            api.process_image(a, t)
        torch.cuda.synchronize()
        gpu.append(time.perf_counter() - t0)
    import cv2
    cv2.setNumThreads(1)
    R.run(imgs[0], threshold=thr[0])
    t0 = time.perf_counter()
    for a, t in zip(imgs, thr):
        R.run(a, threshold=t)
    cpu = time.perf_counter() - t0
    mpx = sum(a.shape[0] * a.shape[1] for a in imgs) / 1e6
    print(json.dumps({"workload": "configs[1]: 17 reference test images, one api.process_image call each",
                      "megapixels": round(mpx, 2), "gpu_s_best_of_3": min(gpu), "gpu_ms_per_image": 1000 * min(gpu) / 17,
                      "cpu_replay_s_1_core": cpu, "cpu_ms_per_image": 1000 * cpu / 17, "speedup": cpu / min(gpu)}))


if __name__ == "__main__":
    main()
