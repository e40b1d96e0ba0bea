"""Pack the reference's dataset PNGs into tests/golden/datasets.npz (uint8).

Run once in the build container (needs /root/reference, which does not exist on
the GPU box).  Follows load_dataset (/root/reference/src/Datasets.jl:54-65):
each `filelist.txt` line is `true.png,data.png`; the Julia loader stores the
8-bit grey value k as k/255 (1-bit PNGs as {0,1}).  We keep the raw integers
and the divisor so the tests rebuild exactly those Float64 values.
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/datasets"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "datasets.npz")


def load(path):
    im = Image.open(path)
    if im.mode == "1":
        return np.asarray(im, dtype=np.uint8), 1
    if im.mode != "L":
        raise SystemExit(f"unexpected mode {im.mode} for {path}")
    return np.asarray(im, dtype=np.uint8), 255


def main():
    out = {}
    for name in sorted(os.listdir(REF)):
        d = os.path.join(REF, name)
        pairs = [l.strip().split(",") for l in open(os.path.join(d, "filelist.txt")) if l.strip()]
        tr, da, dt, dd = [], [], [], []
        for t, n in pairs:
            a, s = load(os.path.join(d, t)); tr.append(a); dt.append(s)
            b, s = load(os.path.join(d, n)); da.append(b); dd.append(s)
        # PIL arrays are [row, col]; Julia's load() gives img[row, col] too, so the
        # (M,N,K) stack is axis-last.
        out[name + "/true"] = np.stack(tr, axis=-1)
        out[name + "/data"] = np.stack(da, axis=-1)
        out[name + "/true_div"] = np.array(dt, dtype=np.int32)
        out[name + "/data_div"] = np.array(dd, dtype=np.int32)
        print(name, out[name + "/true"].shape, dt, dd)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
