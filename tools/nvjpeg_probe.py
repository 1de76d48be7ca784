"""SURVEY 8f-3 probe: is the toolkit's nvJPEG a bit-exact stand-in for the reference's decoder?

The reference decodes with PIL (libjpeg-turbo): Image.open(path).convert('RGB') (img2sgf.py:651).  Everything
downstream is bit-exact, so a device decoder is only a drop-in if its pixels equal PIL's.  This script encodes
diagram-like images as JPEG with PIL (4:2:0 and 4:4:4 colour, greyscale), decodes them with PIL and with
nvJPEG (libnvjpeg from the CUDA toolkit, through ctypes, default and GPU-hybrid back ends) and reports how
many bytes differ.  Measurement only -- nothing in the product calls nvJPEG.
"""
import ctypes as C
import io
import json
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from img2sgf_b200 import synth  # noqa: E402


class NvjpegImage(C.Structure):
    _fields_ = [("channel", C.c_void_p * 4), ("pitch", C.c_size_t * 4)]


def main():
    lib = C.CDLL("libnvjpeg.so.12")
    out = {"library": "libnvjpeg.so.12", "cases": []}
    g, _ = synth.diagram(512, 24, 11, seed=2, noise=1.5, numbered=True)
    rgb = synth.to_rgb(g).copy()
    rgb[..., 0] = np.clip(rgb[..., 0].astype(int) + 12, 0, 255)          # a colour cast, so chroma matters
    cases = [("rgb 4:2:0 q90", Image.fromarray(rgb), dict(quality=90)),
             ("rgb 4:4:4 q95", Image.fromarray(rgb), dict(quality=95, subsampling=0)),
             ("grey q90", Image.fromarray(g), dict(quality=90))]
    for backend_name, backend in (("default", 0), ("gpu_hybrid", 2)):
        handle, state = C.c_void_p(), C.c_void_p()
        rc = lib.nvjpegCreateEx(backend, None, None, 0, C.byref(handle))
        if rc != 0:
            out["cases"].append({"backend": backend_name, "error": f"nvjpegCreateEx rc={rc}"})
            continue
        assert lib.nvjpegJpegStateCreate(handle, C.byref(state)) == 0
        for name, img, kw in cases:
            buf = io.BytesIO()
            img.save(buf, "JPEG", **kw)
            data = buf.getvalue()
            want = np.array(Image.open(io.BytesIO(data)).convert("RGB"))
            h, w = want.shape[:2]
            dev = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
            dst = NvjpegImage()
            dst.channel[0] = dev.data_ptr()
            dst.pitch[0] = 3 * w
            rc = lib.nvjpegDecode(handle, state, data, C.c_size_t(len(data)), 5, C.byref(dst), None)   # 5 = NVJPEG_OUTPUT_RGBI
            torch.cuda.synchronize()
            got = dev.cpu().numpy()
            diff = got.astype(int) - want.astype(int)
            out["cases"].append({"backend": backend_name, "image": name, "rc": rc, "bytes": int(want.size),
                                 "bytes_differing": int((diff != 0).sum()), "max_abs_diff": int(np.abs(diff).max())})
        lib.nvjpegJpegStateDestroy(state)
        lib.nvjpegDestroy(handle)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
