"""Throughput of the region-feature input path (the real pipeline needs >= 150 MB/s/GPU of bf16 features at
> 1 000 dialogs/s/GPU): memory-mapped bf16 shards -> O(1) lookup -> pinned batch (Prefetcher thread) -> device.

    python tools/loader_bench.py [--images 4096] [--batch 64] [--dir /tmp/gstvd_shards] [--cpu]

Prints MB/s and images/s of (a) building pinned batches on the host and (b), when a GPU is present, the same with the
host -> device copy of every batch, next to the 152 KB/image the hot path consumes."""
import argparse
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gst_visdial_b200.io import features as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=4096)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dir", default="/tmp/gstvd_shards")
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()
shutil.rmtree(a.dir, ignore_errors=True)
rng = np.random.default_rng(0)
per = 1024
ids = list(range(a.images))
dirs = []
for s in range(0, a.images, per):
    n = min(per, a.images - s)
    d = os.path.join(a.dir, f"shard{s // per}")
    F.write_shard(d, ids[s:s + n], np.maximum(rng.standard_normal((n, 37, 2048)), 0).astype(np.float32),
                  rng.uniform(0, 1, (n, 37, 5)).astype(np.float32), np.ones((n, 37), np.float32), dtype="bf16")
    dirs.append(d)
sh = F.FeatureShards(dirs)
order = rng.permutation(a.images).tolist()                      # random access, like a shuffled image list
batches = [order[i:i + a.batch] for i in range(0, a.images, a.batch)]
use_gpu = torch.cuda.is_available() and not a.cpu
bytes_per_image = 37 * 2048 * 2 + 37 * 5 * 4 + 37 * 4
for mode in (["host"] + (["host+h2d"] if use_gpu else [])):
    for rep in range(2):                                        # second pass: page cache warm
        t0 = time.perf_counter()
        n = 0
        for b in F.Prefetcher(sh, batches, depth=2, pin=use_gpu):
            if mode == "host+h2d":
                dev = {k: v.cuda(non_blocking=True) for k, v in b.items() if torch.is_tensor(v)}
            n += len(b["image_id"])
        if use_gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{mode:9s} pass {rep}: {n / dt:9.0f} images/s  {n * bytes_per_image / dt / 1e6:8.1f} MB/s  ({n} images, batch {a.batch}, {bytes_per_image / 1e3:.0f} KB/image)")
shutil.rmtree(a.dir, ignore_errors=True)
