"""CPU: the restatement oracle against the golden vectors produced by the reference's own modules
(oracle/gen_golden.py ran them under oracle/ref_shim.py in the development container)."""
import numpy as np
import torch

from helpers import R, history_batch, load_golden


def _check_case(golden_dir, cfgs, sd, tag, B, full_dump):
    enc_cfg, dec_cfg = cfgs
    g = load_golden(golden_dir, tag)
    b = history_batch(enc_cfg, 0, B)
    assert np.array_equal(b["enc_input_ids"].numpy().astype(np.int32), g["enc_input_ids"]), "synthetic inputs drifted"
    assert np.array_equal(b["enc_segments"].numpy().astype(np.int8), g["enc_segments"])
    with torch.no_grad():
        t, v = R.encoder(sd, enc_cfg, b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"], b["enc_segments"],
                         b["enc_att_mask"], b["enc_image_mask"])
        fused, _ = R.vlfusion(sd, t, v, b["enc_att_mask"], b["enc_image_mask"])
        if full_dump:
            assert np.abs(t.numpy() - g["seq_t"]).max() < 2e-4
            assert np.abs(v.numpy() - g["seq_v"]).max() < 2e-4
            assert np.abs(fused.numpy() - g["fused"]).max() < 2e-4
        else:
            assert np.abs(t[:, :48, :32].numpy() - g["seq_t_slice"]).max() < 2e-4
            assert np.abs(v[:, :, :32].numpy() - g["seq_v_slice"]).max() < 2e-4
            assert np.abs(fused.sum(-1).numpy() - g["fused_rowsum"]).max() < 5e-3
        assert np.abs(R.nsp_scores(sd, t, v).numpy() - g["nsp"]).max() < 2e-4
        seq, logits = R.generate_greedy_or_sample(sd, enc_cfg, dec_cfg, b, 1.0, 1, 0.0, 0, return_logits=True)
        assert np.array_equal(seq.numpy(), g["greedy_ids"])
        if full_dump:
            assert np.abs(logits.numpy() - g["greedy_logits"]).max() < 2e-3
        else:
            assert np.abs(logits[:, :, :256].numpy() - g["greedy_logits_slice"]).max() < 2e-3
            assert np.abs(torch.logsumexp(logits, -1).numpy() - g["greedy_logits_lse"]).max() < 2e-3
        loss, _, ppl = R.score_answers(sd, enc_cfg, dec_cfg, b, seq)
        assert np.abs(loss.numpy() - g["score_loss"]).max() < 2e-3
        assert np.allclose(ppl.numpy(), g["score_ppl"], rtol=2e-3)
    return b


def test_restatement_matches_reference_tiny(golden_dir, tiny_cfgs, tiny_sd):
    b = _check_case(golden_dir, tiny_cfgs, tiny_sd, "tiny_b3", 3, True)
    enc_cfg, dec_cfg = tiny_cfgs
    g = load_golden(golden_dir, "tiny_b3")
    with torch.no_grad():
        seq_ng = R.generate_greedy_or_sample(tiny_sd, enc_cfg, dec_cfg, b, 0.7, 1, 0.0, 4)
    assert np.array_equal(seq_ng.numpy(), g["greedy_ng4_ids"])


def test_restatement_matches_reference_full(golden_dir, full_cfgs, full_sd):
    _check_case(golden_dir, full_cfgs, full_sd, "full_b1", 1, False)


def test_ngram_blocking_semantics():
    """Appendix B of SURVEY.md: 4-gram blocking bans the token completing a history 4-gram; specials never recorded."""
    hist = [0] * 5 + [5, 6, 7, 8, 102] + [0] * 10
    assert R.ngram_banned_tokens(hist, [101], 4) == []
    assert R.ngram_banned_tokens(hist, [101, 5, 6], 4) == []        # key (101,5,6) contains [CLS]
    assert R.ngram_banned_tokens(hist, [101, 5, 6, 7], 4) == [8]
    assert R.ngram_banned_tokens(hist, [101, 9, 5, 6, 7], 4) == [8]
    assert R.ngram_banned_tokens(hist, [101, 6, 7, 8], 4) == []     # (6,7,8,102) holds a special id


def test_top_k_keeps_ties():
    x = torch.tensor([[0.1, 2.0, 2.0, -1.0]])
    y = R.top_k_top_p_filter(x.clone(), top_k=1)
    assert torch.isinf(y[0, 0]) and y[0, 1] == 2.0 and y[0, 2] == 2.0


def test_decoding_utils_oracle_matches_reference_goldens(golden_dir):
    """oracle/restatement.py `top_k_top_p_filter` (incl. the top_k = 0 / top_p > 0 branches) and `ngram_banned_tokens` against the
    outputs of the reference's UNMODIFIED utils/decoding_utils.py (:4-35, :38-78) on seeded inputs, committed by
    oracle/gen_golden_decoding.py - the GPU nucleus / full-vocabulary sampling tests check the kernels against these functions."""
    import os
    import numpy as np
    from oracle import gen_golden_decoding as G
    gold = np.load(os.path.join(golden_dir, "decoding_utils.npz"))
    for i, (rows, V, k, p, _) in enumerate(G.FILTER_CASES):
        x = G.filter_inputs(i)
        keep = torch.isfinite(R.top_k_top_p_filter(x.clone(), top_k=k, top_p=p))
        want = torch.from_numpy(np.unpackbits(gold[f"keep_{i}"], axis=-1)[:, :V].astype(bool))
        assert torch.equal(keep, want), f"filter case {i} (top_k {k}, top_p {p})"
        assert keep.sum(-1).tolist() == gold[f"kept_{i}"].tolist()
        if p > 0:
            assert int(keep[1].sum()) == 1                 # the peaked row: a one-token nucleus
    for i, (rows, L, n, plen) in enumerate(G.NGRAM_CASES):
        hist, prefix = G.ngram_inputs(i)
        for r in range(rows):
            got = sorted(set(R.ngram_banned_tokens(hist[r].tolist(), prefix[r].tolist(), n)))
            want = [int(t) for t in gold[f"banned_{i}"][r] if t >= 0]
            assert got == want, f"n-gram case {i} row {r}"
