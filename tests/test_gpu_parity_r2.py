"""GPU: parity cases added in round 2.

  * full-size (config/bert_base_6layer_6conect_*.json) fp32 beam-5 against oracle/beam.py and a full-size 2-round dialog loop
    against oracle/dialog_loop.py: token ids identical (BASELINE.json north_star: "beam indices ... bit-exact", "greedy token
    sequences identical in fp32");
  * bf16, the benchmarked dtype: next to rel-rms the two readings of "2e-2 relative" that SURVEY.md 8d names are asserted and
    printed - max|delta| / rms(reference row) and the elementwise rtol = 2e-2, atol = 2e-2 * rms(row) pass rate - plus a floor on
    bf16-vs-fp32 token agreement;
  * nucleus filtering (utils/decoding_utils.py:23-35) and the exported _reorder_cache entry point
    (models/visual_dialog_decoder.py:177-181), both against the reference semantics.
"""
import json

import numpy as np
import pytest
import torch

from helpers import R, bf16_report, history_batch, load_golden, max_abs, rel_rms
from oracle import beam as OB
from oracle import dialog_loop as OD

pytestmark = pytest.mark.gpu

# bf16 bounds.  rel-rms is the round-1 reading of "2e-2 relative"; the max-based reading is recorded and bounded too.  The maximum
# over the 30 522 logits of a row sits ~4.5 sigma above the rms error, so its bound is looser than the rms bound by that factor.
BF16_REL_RMS = 2e-2
BF16_MAX_OVER_ROW_RMS = 9e-2
# elementwise |delta| <= 2e-2 * (rms(row) + |ref|): measured 0.973 on the logits (30 522-wide rows, rel-rms 1.3e-2), ~1.0 on the encoder
BF16_PASS_RATE = 0.96


@pytest.fixture(scope="module")
def engines(full_cfgs, full_sd):
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = full_cfgs
    e32 = Engine(enc_cfg, dec_cfg, dtype="fp32", max_batch=4)
    e32.load_state_dict(full_sd)
    e16 = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=4)
    e16.load_state_dict(full_sd)
    yield e32, e16
    e32.close(); e16.close()


def _enc(e, b, B):
    o = e.encode(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"], b["enc_att_mask"], b["enc_image_mask"])
    e.prefill_cross(B, o["Le"])
    return o


def test_full_fp32_beam5_matches_oracle(engines, full_cfgs, full_sd):
    """The headline selection mode at full size: beam 5, one image, one round, fp32 - ids identical to oracle/beam.py."""
    e32, _ = engines
    enc_cfg, dec_cfg = full_cfgs
    b = history_batch(enc_cfg, 0, 1)
    with torch.no_grad():
        ref_seq, ref_sc = OB.beam_search(full_sd, enc_cfg, dec_cfg, b, num_beams=5)
    _enc(e32, b, 1)
    seq, sc = e32.generate(1, num_beams=5, want_scores=True)
    assert torch.equal(seq.cpu(), ref_seq), f"{seq.cpu().tolist()} vs {ref_seq.tolist()}"
    assert abs(float(sc.cpu()[0]) - float(ref_sc[0])) < 1e-3


def test_full_fp32_dialog_loop_two_rounds(full_cfgs, full_sd):
    """generate.py:122-233 at full size, fp32, 2 rounds, greedy answers + perplexity pass, against oracle/dialog_loop.py."""
    from gst_visdial_b200 import synthetic as S, weights as W
    from gst_visdial_b200.dialog import generate_dialogs
    from test_gpu_model import _build_model
    enc_cfg, dec_cfg = full_cfgs
    B, rounds = 1, 2
    b = S.synthetic_batch(7, B, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
    questions = torch.stack([torch.stack([S.synthetic_utterance(7 + i, r, enc_cfg.vocab_size) for r in range(rounds)]) for i in range(B)])
    akw = dict(temperature=1.0, top_k=1, top_p=0.0, ngram_blocking_size=0)
    rq, ra, rppl, rflags, rids = OD.generate_dialogs(full_sd, enc_cfg, dec_cfg, b, questions=questions, num_rounds=rounds, a_kwargs=akw)
    model, _ = _build_model(W.DEFAULT_ENC_CONFIG, W.DEFAULT_DEC_CONFIG, full_sd, "fp32", engine_max_batch=2)
    res = generate_dialogs(model, b, questions=questions, num_rounds=rounds, a_kwargs=akw, with_ppl=True, device="cuda:0")
    assert torch.equal(res.answers.cpu(), ra), f"{res.answers.cpu().tolist()} vs {ra.tolist()}"
    assert torch.equal(res.enc_input_ids.cpu(), rids)
    assert torch.equal(res.abnormal.cpu(), rflags)
    assert np.allclose(res.answer_ppl.cpu().numpy(), rppl.numpy(), rtol=2e-3)


def test_full_bf16_tolerances_recorded(engines, full_cfgs, golden_dir, capsys):
    """bf16 against the fp32 engine (itself golden-checked at <= 1e-3) on teacher-forced logits and encoder outputs, three
    readings of the tolerance asserted; bf16-vs-fp32 greedy and beam-5 token agreement asserted, not just printed."""
    e32, e16 = engines
    g = load_golden(golden_dir, "full_b1")
    B = 3
    b = history_batch(full_cfgs[0], 0, B)
    outs = []
    for e in (e32, e16):
        o = e.encode(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"], b["enc_att_mask"], b["enc_image_mask"],
                     want_t=True, want_v=True)
        e.prefill_cross(B, o["Le"])
        dec_in = torch.cat((b["dec_input_ids"], torch.from_numpy(g["greedy_ids"]).repeat(B, 1)[:, :-1]), 1).cuda()
        _, lg = e.score(dec_in, None, labels=torch.zeros_like(dec_in), want_logits=True)
        greedy = e.generate(B, num_beams=1, top_k=1).cpu()
        beam = e.generate(B, num_beams=5).cpu()
        outs.append(dict(t=o["seq_t"].cpu(), v=o["seq_v"].cpu(), lg=lg.cpu(), greedy=greedy, beam=beam))
    valid = b["enc_att_mask"].bool()
    rep = {
        "logits": bf16_report(outs[1]["lg"].reshape(-1, outs[1]["lg"].shape[-1]), outs[0]["lg"].reshape(-1, outs[0]["lg"].shape[-1])),
        "seq_t": bf16_report(outs[1]["t"][valid], outs[0]["t"][valid]),
        "seq_v": bf16_report(outs[1]["v"].reshape(-1, outs[1]["v"].shape[-1]), outs[0]["v"].reshape(-1, outs[0]["v"].shape[-1])),
        "greedy_token_agreement": float((outs[0]["greedy"] == outs[1]["greedy"]).float().mean()),
        "beam5_token_agreement": float((outs[0]["beam"] == outs[1]["beam"]).float().mean()),
        # first position is decided by the same prefix in both dtypes: a clean single-step comparison
        "greedy_first_token_agreement": float((outs[0]["greedy"][:, 0] == outs[1]["greedy"][:, 0]).float().mean()),
    }
    with capsys.disabled():
        print("\nBF16_PARITY " + json.dumps(rep))
    for k in ("logits", "seq_t", "seq_v"):
        assert rep[k]["rel_rms"] < BF16_REL_RMS, (k, rep[k])
        assert rep[k]["max_over_row_rms"] < BF16_MAX_OVER_ROW_RMS, (k, rep[k])
        assert rep[k]["pass_rate"] >= BF16_PASS_RATE, (k, rep[k])
    assert rep["greedy_first_token_agreement"] == 1.0, rep
    # later positions: once one near-tie flips, the prefixes differ and the sequences legitimately diverge (random weights give
    # flat distributions); the floor catches a broken kernel (agreement ~ 1/30522), not rounding
    assert rep["greedy_token_agreement"] >= 0.5, rep
    assert rep["beam5_token_agreement"] >= 0.3, rep


@pytest.mark.parametrize("top_k,top_p,temperature", [(7, 0.5, 0.7), (16, 0.9, 1.0), (7, 0.05, 0.7), (12, 0.3, 1.3)])
def test_nucleus_filter_matches_reference(tiny_cfgs, tiny_sd, top_k, top_p, temperature):
    """top_p > 0 (utils/decoding_utils.py:23-35): the support of the device sampler equals the set the reference's filter keeps.
    The reference samples with torch.multinomial from the surviving logits; here every token drawn over many seeds must lie in
    that set, every token of the set must be drawn (sets of <= 16 tokens, thousands of draws), and the empirical distribution
    must match softmax of the survivors."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs
    e = Engine(enc_cfg, dec_cfg, dtype="fp32", max_batch=8)
    e.load_state_dict(tiny_sd)
    V = enc_cfg.vocab_size
    g = torch.Generator().manual_seed(top_k * 100 + int(top_p * 100))
    rows = 4
    logits = torch.randn(rows, V, generator=g) * 3.0
    logits[1, 5] = logits[1].max() + 4.0             # a peaked row: the nucleus is a single token
    ref = R.top_k_top_p_filter(logits.clone() / temperature, top_k=top_k, top_p=top_p)
    keep = [set(torch.nonzero(torch.isfinite(ref[r])).flatten().tolist()) for r in range(rows)]
    probs = torch.softmax(ref, -1)
    n_draws = 3000
    counts = torch.zeros(rows, V)
    for seed in range(n_draws):
        tok = e.op_sample(logits, step=0, temperature=temperature, top_k=top_k, top_p=top_p, seed=seed).cpu().long()
        counts[torch.arange(rows), tok] += 1
    for r in range(rows):
        drawn = set(torch.nonzero(counts[r]).flatten().tolist())
        assert drawn <= keep[r], f"row {r}: sampled tokens {sorted(drawn - keep[r])} outside the reference nucleus {sorted(keep[r])}"
        emp = counts[r] / n_draws
        likely = {t for t in keep[r] if probs[r, t] > 5e-3}
        assert likely <= drawn, f"row {r}: nucleus tokens never drawn: {sorted(likely - drawn)}"
        assert float((emp - probs[r]).abs().max()) < 0.04, f"row {r}: empirical distribution off by {float((emp - probs[r]).abs().max()):.3f}"
    assert len(keep[1]) == 1
    e.close()


@pytest.mark.parametrize("cfg", ["tiny", "full"])
def test_full_vocabulary_sampling_matches_reference(tiny_cfgs, tiny_sd, full_cfgs, full_sd, cfg):
    """top_k = 0 (utils/decoding_utils.py:17 skips the top-k cut): pure nucleus (top_p > 0) or unrestricted multinomial (top_p = 0)
    over the whole vocabulary - the branch the reference supports but its callers never take.  The support of the device sampler
    must equal the set the reference's filter keeps (sort + cumsum, :23-35) and the empirical distribution must match softmax of
    the survivors (models/visual_dialog_model.py:104-105).  Vocabulary 1000 (tiny) and 30 522 (full; one thread owns 32 tokens)."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs if cfg == "tiny" else full_cfgs
    e = Engine(enc_cfg, dec_cfg, dtype="bf16" if cfg == "full" else "fp32", max_batch=8)
    e.load_state_dict(tiny_sd if cfg == "tiny" else full_sd)
    V = enc_cfg.vocab_size
    for top_p, temperature in [(0.9, 1.0), (0.5, 0.7), (0.05, 1.0), (0.0, 1.0), (0.0, 0.6)]:
        g = torch.Generator().manual_seed(int(top_p * 100) + int(temperature * 10) + V)
        rows = 4
        logits = torch.randn(rows, V, generator=g) * 3.0
        logits[1, 5] = logits[1].max() + 14.0            # a peaked row: the nucleus is a single token
        logits[2] = torch.randn(V, generator=g) * 8.0     # a row with a small nucleus
        dev_logits = logits.cuda()
        ref = R.top_k_top_p_filter(logits.clone() / temperature, top_k=0, top_p=top_p)
        keep = [set(torch.nonzero(torch.isfinite(ref[r])).flatten().tolist()) for r in range(rows)]
        probs = torch.softmax(ref, -1)
        n_draws = 2000
        counts = torch.zeros(rows, V)
        for seed in range(n_draws):
            tok = e.op_sample(dev_logits, step=0, temperature=temperature, top_k=0, top_p=top_p, seed=seed).cpu().long()
            counts[torch.arange(rows), tok] += 1
        for r in range(rows):
            drawn = set(torch.nonzero(counts[r]).flatten().tolist())
            assert drawn <= keep[r], f"row {r}: sampled tokens {sorted(drawn - keep[r])[:8]} outside the reference nucleus ({len(keep[r])} tokens)"
            emp = counts[r] / n_draws
            likely = {t for t in keep[r] if probs[r, t] > 1e-2}
            assert likely <= drawn, f"row {r}: likely tokens never drawn: {sorted(likely - drawn)}"
            assert float((emp - probs[r]).abs().max()) < 0.045, f"row {r}: empirical distribution off by {float((emp - probs[r]).abs().max()):.3f}"
            # total-variation distance over the 64 likeliest tokens + the rest (a wrong scan order / offset shows up here)
            top = torch.topk(probs[r], min(64, V)).indices
            tv = 0.5 * ((emp[top] - probs[r, top]).abs().sum() + abs(float(emp.sum() - emp[top].sum()) - float(1.0 - probs[r, top].sum())))
            assert float(tv) < 0.15, f"row {r}: total variation {float(tv):.3f}"
        if top_p > 0:
            assert len(keep[1]) == 1
        # same seed -> same tokens
        a = e.op_sample(dev_logits, step=0, temperature=temperature, top_k=0, top_p=top_p, seed=7).cpu()
        assert torch.equal(a, e.op_sample(dev_logits, step=0, temperature=temperature, top_k=0, top_p=top_p, seed=7).cpu())
    # the n-gram ban list (utils/decoding_utils.py:38-78) is honoured by this path as well: history "... 5 6 7 8 [SEP]", decoded prefix
    # [CLS] 5 6 7 -> token 8 is banned although it carries almost all of the probability mass
    Lh = 16
    hist = torch.zeros(2, Lh, dtype=torch.int64)
    hist[:, :7] = torch.tensor([101, 9, 5, 6, 7, 8, 102])
    seg = torch.zeros(2, Lh, dtype=torch.int64)
    prefix = torch.tensor([[101, 5, 6, 7], [101, 5, 6, 9]], dtype=torch.int64)
    lg = torch.zeros(2, V)
    lg[:, 8] = 30.0
    lg[:, 11] = 28.0
    toks = torch.stack([e.op_sample(lg.cuda(), step=3, temperature=1.0, top_k=0, top_p=tp, ngram_blocking_size=4, seed=sd, hist_ids=hist,
                                    hist_segments=seg, prefix=prefix).cpu() for sd in range(50) for tp in (0.0, 0.9)])
    assert (toks[:, 0] != 8).all() and (toks[:, 0] == 11).float().mean() > 0.9, toks[:, 0]     # banned in row 0 ...
    assert (toks[:, 1] == 8).float().mean() > 0.7, toks[:, 1]                                   # ... not in row 1 (prefix ends 6 9)
    e.close()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reorder_cache_matches_index_select(tiny_cfgs, tiny_sd, dtype):
    """gstvd_reorder_cache == past_state.index_select(0, beam_idx) for every tensor of every layer
    (models/visual_dialog_decoder.py:29-31,177-181), checked directly on the cache contents: rows are (image, beam) pairs,
    beam_idx holds the parent beam inside each image (duplicated parents included - it is not a permutation), positions at or
    past ``length`` are untouched."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs
    B, K, length = 3, 5, 7
    e = Engine(enc_cfg, dec_cfg, dtype=dtype, max_batch=B, max_beams=K)
    e.load_state_dict(tiny_sd)
    T, H, layers = e.max_new_tokens, enc_cfg.hidden_size, dec_cfg.num_hidden_layers
    g = torch.Generator().manual_seed(5)
    beam_idx = torch.tensor([[2, 0, 2, 4, 4], [0, 1, 2, 3, 4], [1, 1, 1, 0, 3]], dtype=torch.int32)
    before = {}
    for layer in range(layers):
        for kv in (0, 1):
            vals = torch.randn(B, T, K, H, generator=g)
            if dtype == "bf16":
                vals = vals.bfloat16().float()              # exactly representable: the check below is bit-exact in both dtypes
            e.debug_self_cache(B, K, layer, kv, values=vals)
            before[(layer, kv)] = vals
    e.reorder_cache(beam_idx, length)
    for (layer, kv), vals in before.items():
        got = e.debug_self_cache(B, K, layer, kv).cpu()
        # the reference's tensors are [B*K, heads, len, d]: index_select(0, flat parent index) on the (image, beam) rows
        flat = vals.permute(0, 2, 1, 3).reshape(B * K, T, H)                       # [(b, k), t, H]
        parent = (torch.arange(B).unsqueeze(1) * K + beam_idx.long()).reshape(-1)  # [(b, k)] -> row of the parent
        want = flat.clone()
        want[:, :length] = flat.index_select(0, parent)[:, :length]
        want = want.reshape(B, K, T, H).permute(0, 2, 1, 3)
        assert torch.equal(got, want), (layer, kv)
    e.close()


@pytest.mark.parametrize("M,N,K1", [(320, 768, 768), (320, 768, 3072), (37, 768, 768), (5, 128, 64), (200, 1024, 256)])
def test_deferred_layernorm_chain(tiny_cfgs, tiny_sd, M, N, K1):
    """The decode step's dense -> LayerNorm(x + input) pairs without LayerNorm kernels (GemmArgs fold_* / res_* / stats_out):
    raw sums + partial row statistics out of the producing GEMM, LayerNorm folded into the consuming GEMM's weights / epilogue
    and applied on the fly to the residual, against the plain formulation in fp64 on the same bf16-rounded operands."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs
    e = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=4)
    e.load_state_dict(tiny_sd)
    g = torch.Generator().manual_seed(M * 7 + N + K1)
    bf = lambda t: t.bfloat16().double()
    a1 = torch.randn(M, K1, generator=g)
    w1 = torch.randn(N, K1, generator=g) / K1 ** 0.5
    b1 = torch.randn(N, generator=g) * 0.1
    res0 = torch.randn(M, N, generator=g) + 0.3            # a common offset: the mean term of the fold is exercised
    gamma1, beta1 = 1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
    w2 = torch.randn(N, N, generator=g) / N ** 0.5
    b2 = torch.randn(N, generator=g) * 0.1
    gamma2, beta2 = 1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
    out, x1 = e.op_deferred_ln_chain(a1, w1, b1, res0, gamma1, beta1, w2, b2, gamma2, beta2, want_x1=True)

    def ln(x, gm, bt):
        u = x.mean(-1, keepdim=True)
        s = (x - u).pow(2).mean(-1, keepdim=True)
        return gm.double() * ((x - u) / torch.sqrt(s + 1e-12)) + bt.double()

    x1_ref = bf(a1) @ bf(w1).T + b1.double() + bf(res0)
    assert max_abs(x1.cpu(), x1_ref) < 0.03, max_abs(x1.cpu(), x1_ref)           # one bf16 rounding of values of size ~3
    n1 = ln(bf(x1_ref.float()), gamma1, beta1)
    x2_ref = n1 @ w2.double().T + b2.double() + n1
    want = ln(x2_ref, gamma2, beta2)
    rep = bf16_report(out.cpu(), want.float())
    print(f"deferred LN chain M={M} N={N} K1={K1}: {rep}")
    assert rep["rel_rms"] < 1e-2 and rep["max_abs"] < 0.08, rep
    e.close()


def test_cross_attention_more_items_than_warp_slots(tiny_cfgs, tiny_sd):
    """Decode cross-attention, one warp per (image, head) item (csrc/cross_tma.cu dec_cross_warp_kernel): with more items than the
    2 x SMs x 3 resident warp slots (here 456 images x 2 heads = 912 > 888) every warp loops over several items and its box FIFO /
    mbarrier phases carry over from one item to the next.  Histories of different lengths (1 .. 4 boxes of 64 keys, ragged last
    box).  The same images decoded in chunks of 57 (every warp: at most one item) must give the same token ids, and the bf16 ids
    must agree with the fp32 parity path (generic kernels) like everywhere else."""
    from gst_visdial_b200.engine import Engine
    from gst_visdial_b200 import synthetic as S
    enc_cfg, dec_cfg = tiny_cfgs
    B, chunk = 456, 57
    b = S.synthetic_batch(0, B, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
    g = torch.Generator().manual_seed(77)
    ids = b["enc_input_ids"]
    for i in range(B):                                       # histories of 20 .. 250 tokens
        n = int((ids[i] != 0).sum())
        want = int(torch.randint(20, 250, (1,), generator=g))
        if want > n:
            ids[i, n:want] = torch.randint(104, enc_cfg.vocab_size, (want - n,), generator=g)
    b["enc_att_mask"] = (ids != 0).float()
    args = lambda lo, hi: [b[k][lo:hi] for k in ("enc_input_ids", "enc_image_feat", "enc_image_loc", "enc_segments", "enc_att_mask", "enc_image_mask")]
    outs = {}
    for dtype in ("bf16", "fp32"):
        e = Engine(enc_cfg, dec_cfg, dtype=dtype, max_batch=B, max_beams=2)
        e.load_state_dict(tiny_sd)
        o = e.encode(*args(0, B))
        e.prefill_cross(B, o["Le"])
        whole_g = e.generate(B, num_beams=1, top_k=1).cpu()
        whole_b = e.generate(B, num_beams=2).cpu()
        outs[dtype] = (whole_g, whole_b)
        if dtype == "bf16":
            parts_g, parts_b = [], []
            for lo in range(0, B, chunk):
                o = e.encode(*args(lo, lo + chunk))
                e.prefill_cross(chunk, o["Le"])
                parts_g.append(e.generate(chunk, num_beams=1, top_k=1).cpu())
                parts_b.append(e.generate(chunk, num_beams=2).cpu())
            # greedy: 456 rows whole, 57 per chunk - both run the deferred-LayerNorm decode step (M <= 512); the chunks are trimmed
            # to their own longest history, so GEMM tile shapes differ, the per-row arithmetic does not
            assert (torch.cat(parts_g) == whole_g).float().mean().item() >= 0.995
            # beam 2: 912 rows whole - above the M <= 512 limit of the deferred-LayerNorm GEMMs, so the whole batch takes the decode
            # step with LayerNorm kernels (values rounded to bf16 at other places) and near-tied beams legitimately part ways;
            # measured 0.88 - a broken item loop gives ~1 / vocab
            assert (torch.cat(parts_b) == whole_b).float().mean().item() >= 0.75
        e.close()
    agree_g = (outs["bf16"][0] == outs["fp32"][0]).float().mean().item()
    agree_first = (outs["bf16"][0][:, 0] == outs["fp32"][0][:, 0]).float().mean().item()
    print(f"cross-attention item loop: bf16 vs fp32 greedy agreement {agree_g:.3f}, first token {agree_first:.3f}")
    assert agree_first >= 0.97 and agree_g >= 0.6, (agree_first, agree_g)


def test_cross_attention_kernels_agree(full_cfgs, full_sd, monkeypatch):
    """The two decode cross-attention kernels (one warp per item - default; 8-warp CTA per item, GSTVD_CROSS_TMA=2) implement the same
    formulation and differ only in the order of the fp32 sums: full-size model, 3 images with histories of 40 / 150 / 256 tokens,
    beam 5 and greedy - the first tokens must be identical and the sequences agree (near-ties may part ways later)."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = full_cfgs
    B = 3
    b = history_batch(enc_cfg, 0, B)
    g = torch.Generator().manual_seed(5)
    ids = b["enc_input_ids"]
    for i, want in enumerate((40, 150, 256)):
        n = int((ids[i] != 0).sum())
        if want > n:
            ids[i, n:want] = torch.randint(1000, enc_cfg.vocab_size, (want - n,), generator=g)
    b["enc_att_mask"] = (ids != 0).float()
    res = {}
    for variant in ("4", "2"):
        monkeypatch.setenv("GSTVD_CROSS_TMA", variant)
        e = Engine(enc_cfg, dec_cfg, dtype="bf16", max_batch=B, max_beams=5)
        e.load_state_dict(full_sd)
        _enc(e, b, B)
        res[variant] = (e.generate(B, num_beams=1, top_k=1).cpu(), e.generate(B, num_beams=5).cpu())
        e.close()
    for k in (0, 1):
        assert torch.equal(res["4"][k][:, 0], res["2"][k][:, 0]), (res["4"][k][:, :4], res["2"][k][:, :4])
        agree = (res["4"][k] == res["2"][k]).float().mean().item()
        assert agree >= 0.6, agree


def test_sampler_is_keyed_by_global_row(tiny_cfgs, tiny_sd):
    """the draws of an image must not depend on the batch / rank split.  The sampler is keyed by
    (seed, row_offset + row, step): four rows in one call == the same rows in two calls with row offsets 0 and 2, and
    different from the same rows under another seed or offset."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs
    e = Engine(enc_cfg, dec_cfg, dtype="fp32", max_batch=8)
    e.load_state_dict(tiny_sd)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(4, enc_cfg.vocab_size, generator=g)
    logits = logits[:1].repeat(4, 1)                  # identical rows: only the RNG key distinguishes them
    kw = dict(temperature=0.7, top_k=7, top_p=0.0)
    whole = torch.stack([e.op_sample(logits, step=s, seed=99, **kw).cpu() for s in range(12)], 1)
    lo = torch.stack([e.op_sample(logits[:2], step=s, seed=99, row_offset=0, **kw).cpu() for s in range(12)], 1)
    hi = torch.stack([e.op_sample(logits[2:], step=s, seed=99, row_offset=2, **kw).cpu() for s in range(12)], 1)
    assert torch.equal(whole, torch.cat((lo, hi), 0))
    assert not torch.equal(whole[0], whole[1]), "rows of one batch must not replay the same uniforms"
    shifted = torch.stack([e.op_sample(logits[:2], step=s, seed=99, row_offset=2, **kw).cpu() for s in range(12)], 1)
    assert torch.equal(shifted, hi) and not torch.equal(shifted, lo)
    other = torch.stack([e.op_sample(logits, step=s, seed=100, **kw).cpu() for s in range(12)], 1)
    assert not torch.equal(other, whole)
    e.close()


def test_config4_config5_helpers_are_exact_under_trimming_and_chunking(tiny_cfgs, tiny_sd):
    """bench.py's config 4 / 5 paths: answer_perplexity with the host-side history bound == without it (padded positions are
    never read), and nsp_rank_items (image features expanded on the device per chunk, trimmed) == nsp_rank on the fully expanded
    rows - bit for bit, in bf16."""
    from gst_visdial_b200 import ranking as RK, synthetic as S, weights as W
    from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder
    from test_gpu_model import _build_model, _params
    enc_cfg, dec_cfg = tiny_cfgs
    vs, vf = enc_cfg.vocab_size, enc_cfg.v_feature_size
    model, _ = _build_model(W.TINY_ENC_CONFIG, W.TINY_DEC_CONFIG, tiny_sd, "bf16")
    b = S.synthetic_history_batch(0, 6, vocab_size=vs, v_feature_size=vf, rounds=[0, 1, 2, 3, 4, 5])
    ans = torch.stack([S.synthetic_utterance(i, 99, vs) for i in range(6)])
    dev = {k: v.cuda() for k, v in b.items() if k != "hist_len_bound"}
    p_full = RK.answer_perplexity(model, dev, ans.cuda(), device="cuda:0")
    p_trim = RK.answer_perplexity(model, dev, ans.cuda(), device="cuda:0", trim_history=True, hist_len_bound=int(b["hist_len_bound"][0]))
    assert torch.isfinite(p_full).all() and torch.equal(p_full, p_trim)
    params = _params(W.TINY_ENC_CONFIG, W.TINY_DEC_CONFIG, "bf16", model="enc_only_a", mode="vd_eval_val", engine_max_batch=16)
    enc = VisualDialogEncoder(params)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in tiny_sd.items() if k.startswith("encoder.")})
    enc.to("cuda:0").eval()
    items = S.synthetic_candidate_batch(0, 3, 5, vocab_size=vs, v_feature_size=vf)
    got = RK.nsp_rank_items(enc, items, chunk=4, device="cuda:0", trim_history=True)           # chunks straddle items
    flat = {"tokens": items["tokens"].reshape(15, -1), "segments": items["segments"].reshape(15, -1), "mask": items["mask"].reshape(15, -1),
            "image_feat": items["image_feat"].repeat_interleave(5, 0), "image_loc": items["image_loc"].repeat_interleave(5, 0),
            "image_mask": items["image_mask"].repeat_interleave(5, 0)}
    want = RK.nsp_rank(enc, flat, device="cuda:0").reshape(3, 5)
    assert got.shape == (3, 5) and torch.equal(got, want)
    assert ((got > 0) & (got < 1)).all()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_whole_round_graph_equals_separate_calls(tiny_cfgs, tiny_sd, dtype):
    """gstvd_round (encoder + fusion + cross-K/V prefill + all decode steps as ONE graph replay) against gstvd_encode +
    gstvd_prefill_cross + gstvd_generate: identical token ids in greedy, beam-5 and seeded sampling with 4-gram blocking, on trimmed
    and full-length inputs, on the first call (capture) and on replays; the resident state serves a following perplexity pass."""
    from gst_visdial_b200.engine import Engine
    enc_cfg, dec_cfg = tiny_cfgs
    B = 3
    b = history_batch(enc_cfg, 0, B)
    e = Engine(enc_cfg, dec_cfg, dtype=dtype, max_batch=4, max_beams=5)
    e.load_state_dict(tiny_sd)
    dev = {k: v.cuda() for k, v in b.items()}
    for Lt in (256, 96):
        ids, seg, att = dev["enc_input_ids"][:, :Lt].contiguous(), dev["enc_segments"][:, :Lt].contiguous(), dev["enc_att_mask"][:, :Lt].contiguous()
        assert int((dev["enc_input_ids"][:, Lt:] != 0).sum()) == 0
        for kw in (dict(num_beams=1, top_k=1), dict(num_beams=5), dict(num_beams=1, top_k=7, temperature=0.7, ngram_blocking_size=4, seed=5, row_offset=11)):
            o = e.encode(ids, dev["enc_image_feat"], dev["enc_image_loc"], seg, att, dev["enc_image_mask"])
            e.prefill_cross(B, o["Le"])
            want = e.generate(B, hist_ids=ids, hist_segments=seg, **kw).cpu()
            for rep in range(2):                                      # capture, then replay
                got = e.round(ids, dev["enc_image_feat"], dev["enc_image_loc"], seg, att, dev["enc_image_mask"], **kw).cpu()
                assert torch.equal(got, want), (dtype, Lt, kw, rep)
        # the round leaves encoder / cross-K/V state resident: a teacher-forced pass right after it equals one after encode + prefill
        ans = want.cuda().clone()
        l_round, _ = e.score(ans.clone(), (ans != 0).float())
        o = e.encode(ids, dev["enc_image_feat"], dev["enc_image_loc"], seg, att, dev["enc_image_mask"])
        e.prefill_cross(B, o["Le"])
        l_sep, _ = e.score(ans.clone(), (ans != 0).float())
        assert torch.equal(l_round, l_sep)
    e.close()
