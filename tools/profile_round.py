"""One round of the bench workload (encode -> cross-KV prefill -> beam-5 decode, batch 64, bf16) for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_round.py [--batch 64] [--beams 5] [--eager]

The profiled region (cudaProfilerStart/Stop) is ONE warm round.  Numbers printed under ncu are never bench values.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gst_visdial_b200 import _lib, synthetic as S, weights as W  # noqa: E402
from gst_visdial_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beams", type=int, default=5)
ap.add_argument("--eager", action="store_true", help="launch decode steps eagerly (no CUDA graph)")
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--hist", type=int, default=0, help="pad every history to this many tokens (0: caption only); 256 = worst case, no key tile skipped")
a = ap.parse_args()
enc_cfg = W.load_json_config(W.TINY_ENC_CONFIG if a.tiny else W.DEFAULT_ENC_CONFIG)
dec_cfg = W.load_json_config(W.TINY_DEC_CONFIG if a.tiny else W.DEFAULT_DEC_CONFIG)
eng = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=a.batch, max_beams=a.beams, flags=_lib.GSTVD_FLAG_NO_CUDA_GRAPH if a.eager else 0)
eng.load_state_dict(W.synthetic_state_dict(enc_cfg, dec_cfg, seed=0))
b = S.synthetic_batch(0, a.batch, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
if a.hist:
    g = torch.Generator().manual_seed(0)
    ids = b["enc_input_ids"]
    for i in range(a.batch):
        n = int((ids[i] != 0).sum())
        if a.hist > n:
            ids[i, n:a.hist] = torch.randint(1000, enc_cfg.vocab_size, (a.hist - n,), generator=g)
            ids[i, a.hist - 1] = 102
    b["enc_att_mask"] = (ids != 0).float()
b = {k: v.cuda() for k, v in b.items()}


def one_round():
    out = eng.encode(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"], b["enc_att_mask"], b["enc_image_mask"])
    eng.prefill_cross(a.batch, out["Le"])
    return eng.generate(a.batch, num_beams=a.beams)


for _ in range(2):
    one_round()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
ids = one_round()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"one round: {e0.elapsed_time(e1):.2f} ms (not a bench value when run under a profiler); ids[0,:6]={ids[0,:6].tolist()}")
# the decode part alone (18 steps replayed from the graph), mean of 5, plus a checksum of the ids for A/B runs of experimental flags
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
for _ in range(5):
    ids = eng.generate(a.batch, num_beams=a.beams)
g1.record()
torch.cuda.synchronize()
w = torch.arange(1, ids.numel() + 1, device=ids.device, dtype=torch.int64).view_as(ids)
print(f"generate only: {g0.elapsed_time(g1) / 5:.3f} ms per call; ids checksum {int((ids * w).sum())}")
