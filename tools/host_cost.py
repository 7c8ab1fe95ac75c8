"""Host-side (enqueue) cost of the C-ABI calls: the calls are asynchronous, so with a tiny batch their wall time is CPU time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gst_visdial_b200 import _lib, synthetic as S, weights as W  # noqa: E402
from gst_visdial_b200.engine import Engine  # noqa: E402

enc_cfg, dec_cfg = W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for flags, name in ((0, "graph"), (_lib.GSTVD_FLAG_NO_CUDA_GRAPH, "eager")):
    eng = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=B, max_beams=5, flags=flags)
    eng.load_state_dict(W.synthetic_state_dict(enc_cfg, dec_cfg, seed=0))
    b = {k: v.cuda() for k, v in S.synthetic_batch(0, B, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size).items()}
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = eng.encode(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"], b["enc_att_mask"], b["enc_image_mask"])
        t1 = time.perf_counter()
        eng.prefill_cross(B, out["Le"])
        t2 = time.perf_counter()
        eng.generate(B, num_beams=5)
        t3 = time.perf_counter()
        torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"{name}: encode {1e3 * (t1 - t0):.2f} ms  prefill {1e3 * (t2 - t1):.2f} ms  generate(18 steps) {1e3 * (t3 - t2):.2f} ms  drain {1e3 * (t4 - t3):.2f} ms")
    eng.close()
