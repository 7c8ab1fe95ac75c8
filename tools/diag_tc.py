"""Step-by-step run of the full bf16 path with a device sync after every C-ABI call (debugging aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gst_visdial_b200 import synthetic as S, weights as W  # noqa: E402
from gst_visdial_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
enc_cfg, dec_cfg = W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG)
sd = W.synthetic_state_dict(enc_cfg, dec_cfg, seed=0)
res = {}
for dtype in ("fp32", "bf16"):
    eng = Engine(enc_cfg, dec_cfg, dtype=dtype, max_batch=B)
    eng.load_state_dict(sd)
    torch.cuda.synchronize(); print(dtype, "weights ok", flush=True)
    b = S.synthetic_batch(0, B, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
    out = eng.encode(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"], b["enc_att_mask"], b["enc_image_mask"],
                     want_t=True, want_fused=True)
    torch.cuda.synchronize(); print(dtype, "encode ok", flush=True)
    eng.prefill_cross(B, out["Le"])
    torch.cuda.synchronize(); print(dtype, "prefill ok", flush=True)
    dec = torch.randint(1000, 2000, (B, 18), device="cuda")
    loss, logits = eng.score(dec, None, labels=torch.zeros_like(dec), want_logits=True)
    torch.cuda.synchronize(); print(dtype, "score ok", flush=True)
    ids = eng.generate(B, num_beams=5)
    torch.cuda.synchronize(); print(dtype, "generate ok", ids[0, :6].tolist(), flush=True)
    res[dtype] = (out["seq_t"].cpu(), out["fused"].cpu(), logits.cpu(), ids.cpu())
    eng.close()
a, c = res["fp32"], res["bf16"]
for name, x, y in (("seq_t", a[0], c[0]), ("fused", a[1], c[1]), ("logits", a[2], c[2])):
    print(name, "rel rms", float((x - y).pow(2).mean().sqrt() / x.pow(2).mean().sqrt()))
print("beam agreement", float((a[3] == c[3]).float().mean()))
