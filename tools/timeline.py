"""True GPU timeline of one warm round (encode -> prefill -> 18 beam-5 steps) from CUPTI through torch.profiler: kernel start/end
times as they ran (CUDA graph replay, PDL overlap, warm caches) - unlike an ncu launch list, which serialises and flushes.

    python tools/timeline.py [--hist 150] [--out gpurun_out/timeline.txt]

Prints per-kernel totals (busy time, count), the union busy time, the idle gaps, and the sequence of one decode step."""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gst_visdial_b200 import synthetic as S, weights as W  # noqa: E402
from gst_visdial_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beams", type=int, default=5)
ap.add_argument("--hist", type=int, default=150)
ap.add_argument("--out", default="")
ap.add_argument("--seq", type=int, default=160, help="kernels of the decode sequence to list")
a = ap.parse_args()
enc_cfg, dec_cfg = W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG)
eng = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=a.batch, max_beams=a.beams)
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
Lt = min(256, max(32, (a.hist + 31) // 32 * 32)) if a.hist else 256
b = {k: v.cuda() for k, v in b.items()}


def one_round():
    out = eng.encode(b["enc_input_ids"][:, :Lt].contiguous(), b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"][:, :Lt].contiguous(),
                     b["enc_att_mask"][:, :Lt].contiguous(), b["enc_image_mask"])
    eng.prefill_cross(a.batch, out["Le"])
    return eng.generate(a.batch, num_beams=a.beams)


for _ in range(3):
    one_round()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    one_round()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("gstvd::", "").replace("(anonymous namespace)::", "")
    return n[:70]


lines = []
t0, t1 = ks[0][0], max(k[1] for k in ks)
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    agg[short(n)][0] += 1
    agg[short(n)][1] += e - s
busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
for s, e, n in ks[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
lines.append(f"round: {(t1 - t0) / 1e3:.3f} ms wall on the GPU, {len(ks)} kernels, union busy {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms, "
             f"sum of kernel durations {sum(v[1] for v in agg.values()) / 1e3:.3f} ms (Lt={Lt})")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{v[1]:10.1f} us {v[0]:6d} x {v[1] / v[0]:8.2f} us  {k}")
# the decode part: from the first embed_step kernel on
first = next((i for i, k in enumerate(ks) if "embed_step" in k[2]), None)
if first is not None:
    dec = ks[first:]
    steps = [i for i, k in enumerate(dec) if "embed_step" in k[2]]
    if len(steps) > 3:
        lines.append(f"decode: {len(steps)} steps, {(dec[-1][1] - dec[0][0]) / 1e3 / len(steps):.4f} ms per step, {steps[1] - steps[0]} kernels per step")
        lo, hi = steps[2], steps[3]
        lines.append("one decode step (start offset us, duration us, gap to the previous kernel's end us):")
        prev_end = dec[lo][0]
        for s, e, n in dec[lo:min(hi, lo + a.seq)]:
            lines.append(f"  {s - dec[lo][0]:9.2f} {e - s:8.2f} {s - prev_end:8.2f}  {short(n)}")
            prev_end = max(prev_end, e)
txt = "\n".join(lines)
print(txt)
if a.out:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write(txt + "\n")
