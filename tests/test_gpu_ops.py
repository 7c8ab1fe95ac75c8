"""GPU: single operators through the C ABI against plain PyTorch / the oracle."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import R, max_abs, rel_rms
from oracle import beam as OB

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(full_cfgs):
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = full_cfgs
    e = Engine(enc_cfg, dec_cfg, device=0, dtype="bf16", max_batch=8, max_beams=5)
    yield e
    e.close()


def _ref_linear(a, w, b, act, bf16):
    if bf16:
        a, w = a.bfloat16().float(), w.bfloat16().float()
    y = a.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if act:
        y = y * 0.5 * (1.0 + torch.erf(y / math.sqrt(2.0)))
    return y.float()


LINEAR_SHAPES = [
    (128, 128, 64), (128, 256, 64), (256, 256, 128), (37, 1024, 2048), (300, 768, 768), (2048, 2304, 768),
    (1000, 3072, 768), (777, 768, 3072), (320, 30522, 768), (5, 2, 1024), (1, 768, 768), (293, 1536, 768), (129, 40, 72),
]


@pytest.mark.parametrize("M,N,K", LINEAR_SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_linear_tcgen05_bf16(eng, M, N, K, act):
    if act and N > 4000:
        pytest.skip("gelu on the LM head shape is not a real call site")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    y = eng.op_linear(a, w, b, act=act, dtype="bf16").cpu()
    ref = _ref_linear(a, w, b, act, True)
    err = max_abs(y, ref)
    assert err < 2e-3, f"tcgen05 GEMM {M}x{N}x{K} act={act}: max abs err {err}, rel rms {rel_rms(y, ref)}"


# ---- CTA-pair GEMM (tcgen05 cta_group::2, gemm_tc2.cu): the default for the large-M throughput problems -----------------------
# ragged M / N / K on purpose: half-empty pair tiles (M % 256 in (0, 128]), N not a multiple of the tile, K tail zero-filled by TMA
PAIR_SHAPES = [(2048, 2304, 768), (16384, 768, 3072), (10240, 3072, 768), (1024, 256, 64), (1100, 768, 768), (2368, 1024, 2048),
               (4096, 1000, 72), (18752, 768, 768)]


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("bn", ["128", "256"])
def test_linear_pair_bf16(eng, M, N, K, act, bn):
    """Same products, same k order, fp32 accumulation in the tensor core: the pair kernel must be bit-identical to the single-CTA
    kernel (itself checked against fp64 above) on every shape, forced through both pair tile widths."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    os.environ["GSTVD_GEMM_2CTA"] = "0"
    try:
        y1 = eng.op_linear(a, w, b, act=act, dtype="bf16").cpu()
        os.environ["GSTVD_GEMM_2CTA"] = bn
        y2 = eng.op_linear(a, w, b, act=act, dtype="bf16").cpu()
    finally:
        os.environ.pop("GSTVD_GEMM_2CTA", None)
    if M * N * K < 4e9:                                  # the fp64 reference on the host is slow; the single-CTA kernel is the main oracle
        ref = _ref_linear(a, w, b, act, True)
        err = max_abs(y2, ref)
        assert err < 2e-3, f"pair GEMM {M}x{N}x{K} act={act} bn={bn}: max abs err {err}, rel rms {rel_rms(y2, ref)}"
    assert max_abs(y2, y1) < 1e-5, f"pair vs single-CTA kernel: {max_abs(y2, y1)}"


@pytest.mark.parametrize("M,N,K", [(37, 1024, 2048), (300, 768, 768), (64, 3072, 768), (5, 2, 1024), (129, 40, 72), (33, 1000, 128)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_simt_fp32(eng, M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    y = eng.op_linear(a, w, b, act=act, dtype="fp32").cpu()
    ref = _ref_linear(a, w, b, act, False)
    assert max_abs(y, ref) < 1e-4


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("rows,width", [(5, 768), (300, 1024), (37, 128), (64, 256), (4100, 768)])
def test_add_layernorm(eng, dtype, tol, rows, width):
    g = torch.Generator().manual_seed(rows + width)
    x, r = torch.randn(rows, width, generator=g) * 2, torch.randn(rows, width, generator=g)
    gamma, beta = torch.randn(width, generator=g), torch.randn(width, generator=g)
    y = eng.op_add_layernorm(x, r, gamma, beta, dtype=dtype).cpu()
    if dtype == "bf16":
        x, r = x.bfloat16().float(), r.bfloat16().float()
    s = x + r
    u = s.mean(-1, keepdim=True)
    v = (s - u).pow(2).mean(-1, keepdim=True)
    ref = gamma * ((s - u) / torch.sqrt(v + 1e-12)) + beta
    assert max_abs(y, ref) < tol * max(1.0, ref.abs().max().item())
    y2 = eng.op_add_layernorm(x, None, gamma, beta, dtype="fp32").cpu()
    u2 = x.mean(-1, keepdim=True)
    ref2 = gamma * ((x - u2) / torch.sqrt((x - u2).pow(2).mean(-1, keepdim=True) + 1e-12)) + beta
    assert max_abs(y2, ref2) < 2e-5 * max(1.0, ref2.abs().max().item())


def _ref_attention(q, k, v, heads, mask, neg, causal):
    B, Lq, W = q.shape
    Lk = k.shape[1]
    d = W // heads
    qh = q.view(B, Lq, heads, d).permute(0, 2, 1, 3).double()
    kh = k.view(B, Lk, heads, d).permute(0, 2, 1, 3).double()
    vh = v.view(B, Lk, heads, d).permute(0, 2, 1, 3).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    m = torch.ones(B, 1, Lq, Lk, dtype=torch.double)
    if mask is not None:
        m = m * mask[:, None, None, :].double()
    if causal:
        i = torch.arange(Lq)[:, None]; j = torch.arange(Lk)[None, :]
        m = m * (j <= i).double()
    s = s + (1.0 - m) * neg
    p = torch.softmax(s, -1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, Lq, W).float()


ATTN_CASES = [  # B, heads, Lq, Lk, D, causal, neg
    (2, 12, 256, 256, 64, False, -10000.0), (3, 8, 37, 37, 128, False, -10000.0), (2, 8, 256, 37, 128, False, -10000.0),
    (2, 8, 37, 256, 128, False, -10000.0), (3, 12, 18, 18, 64, True, -10000.0), (2, 12, 18, 293, 64, False, -1e9),
    (5, 2, 1, 293, 64, False, -1e9), (2, 2, 25, 25, 64, True, -10000.0),
]


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("B,heads,Lq,Lk,D,causal,neg", ATTN_CASES)
def test_attention(eng, dtype, tol, B, heads, Lq, Lk, D, causal, neg):
    g = torch.Generator().manual_seed(Lq * 31 + Lk)
    W = heads * D
    q, k, v = (torch.randn(B, L, W, generator=g) for L in (Lq, Lk, Lk))
    mask = (torch.rand(B, Lk, generator=g) > 0.3).float()
    mask[:, 0] = 1.0
    y = eng.op_attention(q, k, v, heads, mask, neg=neg, causal=causal, dtype=dtype).cpu()
    if dtype == "bf16":
        q, k, v = q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float()
    ref = _ref_attention(q, k, v, heads, mask, neg, causal)
    assert max_abs(y, ref) < tol, f"max abs {max_abs(y, ref)}"


@pytest.mark.parametrize("heads,Lq,Lk,D,causal", [(12, 256, 256, 64, False), (8, 37, 256, 128, False), (8, 256, 37, 128, False), (12, 25, 293, 64, False),
                                                  (12, 130, 130, 64, True),
                                                  # trimmed history lengths: query tiles of 80 / 96 / 112 rows (QW = 5, 6, 7)
                                                  (12, 160, 160, 64, False), (12, 192, 192, 64, False), (12, 224, 224, 64, False),
                                                  (12, 96, 96, 64, False), (12, 100, 100, 64, True)])
def test_attention_trailing_padding_bf16(eng, heads, Lq, Lk, D, causal):
    """History-shaped masks (valid prefix, padded tail): the tensor-core kernel skips key tiles past the last valid key."""
    g = torch.Generator().manual_seed(Lq + 7 * Lk)
    B, W = 4, heads * D
    q, k, v = (torch.randn(B, L, W, generator=g) for L in (Lq, Lk, Lk))
    lens = torch.tensor([1, min(Lk, 40), min(Lk, 130), Lk])
    mask = (torch.arange(Lk)[None, :] < lens[:, None]).float()
    y = eng.op_attention(q, k, v, heads, mask, neg=-10000.0, causal=causal, dtype="bf16").cpu()
    q, k, v = q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float()
    ref = _ref_attention(q, k, v, heads, mask, -10000.0, causal)
    assert max_abs(y, ref) < 2e-2, f"max abs {max_abs(y, ref)}"


def test_attention_fully_masked_row_is_finite(eng):
    """Additive masks: a row whose keys are all masked still attends by its raw scores (the reference's behaviour, not NaN)."""
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(1, 4, 128, generator=g) for _ in range(3))
    mask = torch.zeros(1, 4)
    y = eng.op_attention(q, k, v, 2, mask, neg=-10000.0, causal=False, dtype="fp32").cpu()
    ref = _ref_attention(q, k, v, 2, mask, -10000.0, False)
    # adding -10000 in fp32 quantises the scores to ~1e-3 (the reference's fp32 arithmetic does the same), hence the tolerance
    assert torch.isfinite(y).all() and max_abs(y, ref) < 5e-3


def test_beam_step_mass_ties_vs_oracle(eng, full_cfgs):
    """Hundreds of exactly equal top logits (and a constant row): the candidate filter of row_select overflows and the
    iterative arg-max fallback must still order by (score desc, index asc) like the oracle."""
    V, T, B, K = full_cfgs[1].vocab_size, 4, 2, 3
    g = torch.Generator().manual_seed(77)
    st = OB.BeamState(B, K, V, T)
    eng.op_beam_begin(B, K, T)
    for t in range(T):
        logits = torch.randn(B * K, V, generator=g)
        if t == 0:
            logits[:] = 0.25                                  # every token ties
        if t == 1:
            logits[:, 9000:9400] = 7.0                        # 400 tied maxima inside one 8192-entry chunk
        if t == 2:
            logits[:, 5::97] = 6.5                            # ~315 tied maxima spread over all chunks
        bi, bt, bs = st.step(logits)
        gi, gt, gs = eng.op_beam_step(logits)
        assert torch.equal(gi.cpu().long(), bi) and torch.equal(gt.cpu().long(), bt) and torch.equal(gs.cpu(), bs), f"step {t}"
    seq, _ = st.finalize()
    gseq, _ = eng.op_beam_end(T)
    assert torch.equal(gseq.cpu(), seq)


@pytest.mark.parametrize("B,K", [(3, 5), (2, 1), (4, 2)])
def test_beam_step_bit_exact_vs_oracle(eng, full_cfgs, B, K):
    """Identical fp32 logits into the CUDA beam kernels and oracle/beam.py: parents, tokens, scores and the final
    hypotheses must agree exactly (north_star: 'beam indices and top-k selections bit-exact given identical logits')."""
    V, T = full_cfgs[1].vocab_size, 18
    g = torch.Generator().manual_seed(100 + B * 10 + K)
    st = OB.BeamState(B, K, V, T)
    eng.op_beam_begin(B, K, T)
    for t in range(T):
        logits = torch.randn(B * K, V, generator=g) * 2.0
        if t >= 2:                                            # make [SEP] competitive so hypotheses finish
            logits[:, OB.EOS] += 6.0 + torch.rand(B * K, generator=g) * 3
        if t == 5:                                            # exact ties across tokens and beams
            logits[:, 2000:2004] = logits[:, 2000:2001]
        bi, bt, bs = st.step(logits)
        gi, gt, gs = eng.op_beam_step(logits)
        assert torch.equal(gi.cpu().long(), bi), f"parent beams differ at step {t}"
        assert torch.equal(gt.cpu().long(), bt), f"tokens differ at step {t}"
        assert torch.equal(gs.cpu(), bs), f"scores differ at step {t}: {(gs.cpu() - bs).abs().max()}"
    seq, sc = st.finalize()
    gseq, gsc = eng.op_beam_end(T)
    assert torch.equal(gseq.cpu(), seq)
    assert torch.allclose(gsc.cpu().double(), sc, rtol=1e-6)


def test_sample_greedy_topk_and_ngram(eng, full_cfgs):
    V = full_cfgs[1].vocab_size
    g = torch.Generator().manual_seed(9)
    rows = 6
    logits = torch.randn(rows, V, generator=g)
    tok = eng.op_sample(logits, step=0, temperature=0.7, top_k=1).cpu().long()
    assert torch.equal(tok, (logits / 0.7).argmax(-1))
    # 4-gram blocking: history holds "5000 5001 5002 5003"; prefix ends in 5000 5001 5002 -> 5003 is banned
    hist = torch.zeros(rows, 256, dtype=torch.int64)
    seg = torch.zeros(rows, 256, dtype=torch.int64)
    hist[:, 0] = 101
    hist[:, 10:14] = torch.tensor([5000, 5001, 5002, 5003])
    hist[:, 14] = 102
    hist[3, 10:14] = 0                      # row 3 has no such n-gram
    seg[4, 10:14] = 1                       # row 4: the span is an answer (segment 1) -> not part of the question history
    prefix = torch.tensor([[101, 7, 5000, 5001, 5002]] * rows)
    logits2 = logits.clone()
    logits2[:, 5003] = 50.0
    tok2 = eng.op_sample(logits2, step=4, temperature=1.0, top_k=1, ngram_blocking_size=4, hist_ids=hist, hist_segments=seg,
                         prefix=prefix).cpu().long()
    for r in range(rows):
        step = logits2[r:r + 1].clone()
        banned = R.ngram_banned_tokens((hist[r] * (seg[r] == 0).long()).tolist(), prefix[r].tolist(), 4)
        step[0, banned] = -float("inf")
        assert tok2[r].item() == step.argmax(-1).item()
    assert tok2[0].item() != 5003 and tok2[3].item() == 5003 and tok2[4].item() == 5003
    # top-k sampling stays inside the top-k set and follows the softmax of the kept logits
    lg = torch.full((8, V), -5.0)
    lg[:, 100:107] = torch.tensor([3.0, 2.5, 2.0, 1.5, 1.0, 0.5, 0.0])
    lg[:, 200] = -0.5          # 8th largest: must never be drawn with top_k = 7
    counts = torch.zeros(7)
    n = 0
    for s in range(200):
        tk = eng.op_sample(lg, step=s % 18, temperature=0.7, top_k=7, seed=1000 + s).cpu().long()
        assert ((tk >= 100) & (tk < 107)).all()
        counts += torch.bincount(tk - 100, minlength=7).float()
        n += tk.numel()
    p = torch.softmax(lg[0, 100:107] / 0.7, -1)
    assert (counts / n - p).abs().max() < 0.04, f"empirical {counts / n} vs {p}"


def test_splice_matches_reference_loop(eng):
    g = torch.Generator().manual_seed(3)
    B, Lt, Lu = 7, 64, 18
    ids = torch.zeros(B, Lt, dtype=torch.int64)
    seg = torch.zeros(B, Lt, dtype=torch.int64)
    lens = torch.tensor([5, 10, 40, 50, 60, 63, 46])
    for b in range(B):
        ids[b, : lens[b]] = torch.randint(1000, 2000, (int(lens[b]),), generator=g)
    utt = torch.zeros(B, Lu, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(3, 18, (1,), generator=g))
        utt[b, : n - 1] = torch.randint(1000, 2000, (n - 1,), generator=g)
        utt[b, n - 1] = 102
    for seg_val, strip in ((-1, False), (1, True)):
        r_ids, r_seg, r_len, abn = ids.clone(), seg.clone(), lens.clone(), set()
        u = utt.masked_fill(utt == 102, 0) if strip else utt
        r_len = r_len + R.splice(r_ids, r_seg, r_len, u, None if seg_val < 0 else seg_val, abn)
        d_ids, d_seg = ids.cuda(), seg.cuda()
        d_mask = torch.zeros(B, Lt, device="cuda")
        d_len = lens.to(torch.int32).cuda()
        d_abn = torch.zeros(B, dtype=torch.int32, device="cuda")
        eng.splice(d_ids, d_seg, d_mask, d_len, utt.cuda(), segment_value=seg_val, strip_sep=strip, abnormal=d_abn)
        assert torch.equal(d_ids.cpu(), r_ids) and torch.equal(d_seg.cpu(), r_seg)
        assert torch.equal(d_len.cpu().long(), r_len)
        assert torch.equal(d_mask.cpu(), (r_ids != 0).float())
        assert set(torch.nonzero(d_abn.cpu()).flatten().tolist()) == abn
