"""Times the tcgen05 GEMM on the encoder shapes through the C ABI (gstvd_op_linear) with CUDA events.

Also the target of `ncu --set full --import-source on -k regex:gemm_tc_kernel`.  GSTVD_GEMM_DBG=1/2/3 and GSTVD_GEMM_BN
are measurement knobs of the kernel (epilogue without stores / epilogue skipped / MMA skipped, forced N tile).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gst_visdial_b200 import weights as W  # noqa: E402
from gst_visdial_b200.engine import Engine  # noqa: E402

shapes = [(16384, 2304, 768, 0), (16384, 3072, 768, 1), (16384, 768, 3072, 0), (16384, 768, 768, 0), (18752, 18432, 768, 0),
          (2368, 3072, 1024, 0), (320, 2304, 768, 0), (320, 768, 768, 0), (320, 3072, 768, 1), (320, 768, 3072, 0), (320, 30522, 768, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "sweep":      # text-stream shapes at trimmed history lengths (B = 64, Lt = 32 .. 256)
    shapes = [(64 * lt, n, k, act) for lt in (32, 64, 128, 192, 256) for (n, k, act) in ((2304, 768, 0), (768, 768, 0), (3072, 768, 1), (768, 3072, 0))]
    shapes += [(2368, 1024, 1024, 0), (2368, 3072, 1024, 0), (2368, 1024, 2048, 0)]
elif len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
reps = int(os.environ.get("REPS", "20"))
eng = Engine(W.load_json_config(W.TINY_ENC_CONFIG), W.load_json_config(W.TINY_DEC_CONFIG), dtype="bf16", max_batch=2)
lib, ctx = eng.lib, eng.ctx
import ctypes  # noqa: E402

for (M, N, K, act) in shapes:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    y = eng.op_linear(a, w, b, act=act, dtype="bf16")        # includes the fp32->bf16 casts; warms everything
    # time only the GEMM launches: reuse the engine's profiling events around tcgen05 launches
    eng.profile_gemm(True, 0)
    for _ in range(reps):
        eng.op_linear(a, w, b, act=act, dtype="bf16")
    r = eng.profile_read()
    eng.profile_gemm(False, 0)
    us = r["ms"] * 1e3 / max(r["launches"], 1)
    print(f"M={M:6d} N={N:6d} K={K:5d} act={act}: {us:9.2f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  ({r['launches']} launches)", flush=True)
