"""Small invocation of every kernel of the path for compute-sanitizer (memcheck / racecheck):
one image through process_image (all stage kernels incl. the shared-memory atomics of k_vote_peaks,
k_radius, k_line_vote, k_hysteresis_list, k_circles_finish, k_median), a noisy one, a ragged batch of
three sizes (one greyscale plane), and the stage entry points on an odd-sized array."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from img2sgf_b200 import api, batch as B, synth  # noqa: E402

g, truth = synth.diagram(300, 14, 6, seed=3)
r = api.process_image(synth.to_rgb(g), 40)
print("clean:", len(r.circles), "circles, board_ready", r.board_ready)
g2, _ = synth.diagram(260, 12, 5, seed=4, noise=3.0, numbered=True)
r2 = api.process_image(synth.to_rgb(g2)[:241, :203].copy(), 40, contrast_slider=60, brightness_slider=55)
print("noisy, odd size, prologue:", len(r2.circles), "circles")
imgs = [synth.to_rgb(g)[:280, :290].copy(), np.ascontiguousarray(g2[:199, :255]), synth.to_rgb(g2)]
rec = B.RaggedRunner(streams=2, max_group=2).process_images(imgs)
print("ragged:", rec["n_circles"].tolist(), rec["status"].tolist())
img = np.ascontiguousarray(g2[:131, :77])
for b in (3, 5, 7):
    api.median_blur(img, b)
api.gaussian_blurs(img)
api.canny_grey(img)
c = api.hough_circles(np.ascontiguousarray(g[:200, :210]))
print("stage calls ok,", len(c), "circles")
