"""CPU: the beam-search contract (oracle/beam.py) - internal consistency and the HF 4.16.2 behaviours it encodes."""
import torch

from oracle import beam


def test_logz_matches_torch_logsumexp():
    torch.manual_seed(0)
    x = torch.randn(7, 1000) * 3
    assert torch.allclose(beam.log_z(x).squeeze(-1), torch.logsumexp(x, -1), atol=1e-5)


def test_top_candidates_tie_order():
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    v, i = beam.top_candidates(s, 4)
    assert i.tolist() == [[1, 2, 4, 3]]


def test_first_step_only_beam0_live():
    V, K = 50, 3
    st = beam.BeamState(1, K, V, 4)
    logits = torch.zeros(K, V)
    logits[0, 7] = 5.0
    logits[1, 9] = 50.0          # would dominate if beams 1.. were live
    bi, bt, bs = st.step(logits)
    assert bi.tolist() == [[0, 0, 0]] and bt[0, 0].item() == 7


def test_eos_goes_to_hypotheses_and_finalize_appends_eos():
    V, K, T = 200, 2, 3
    st = beam.BeamState(1, K, V, T)
    l0 = torch.full((K, V), -5.0); l0[0, 11] = 3.0; l0[0, 12] = 2.0
    st.step(l0)
    l1 = torch.full((K, V), -5.0); l1[0, beam.EOS] = 6.0; l1[0, 13] = 1.0; l1[1, 14] = 1.0
    st.step(l1)
    assert len(st.hyps[0]) == 1 and st.hyps[0][0][1] == [11]
    l2 = torch.full((K, V), -5.0); l2[:, 15] = 1.0
    st.step(l2)
    seq, sc = st.finalize()
    assert seq.shape == (1, T)
    best = seq[0].tolist()
    assert best == [11, beam.EOS, 0] or best[-1] != 0


def test_beam1_equals_greedy(tiny_cfgs, tiny_sd):
    from helpers import R, history_batch
    enc_cfg, dec_cfg = tiny_cfgs
    b = history_batch(enc_cfg, 0, 2)
    with torch.no_grad():
        g = R.generate_greedy_or_sample(tiny_sd, enc_cfg, dec_cfg, b, 1.0, 1, 0.0, 0)
        s, _ = beam.beam_search(tiny_sd, enc_cfg, dec_cfg, b, num_beams=1)
    if not (g == beam.EOS).any():
        assert torch.equal(g, s)


def test_ancestry_table_equals_cache_gather():
    """The rule the bf16 beam path uses instead of `_reorder_cache` (visual_dialog_decoder.py:177-181; csrc/decode.cu
    anc_update_kernel + dec_self_attn_v2_kernel): beams never move their cache slots; position t of beam k's history is read from
    slot anc[k][t], and after a selection with parents p the rows become anc'[k][t] = anc[p[k]][t] (t < step), anc'[k][step] = p[k].
    Restated here on integers against the plain gather, including repeated parents."""
    import numpy as np
    rng = np.random.default_rng(0)
    for K, T in ((5, 18), (2, 7), (8, 32), (1, 4)):
        gathered = np.zeros((T, K), dtype=np.int64)          # [position][beam]: what index_select leaves in the cache
        slots = np.zeros((T, K), dtype=np.int64)             # [position][slot]: never moved
        anc = np.zeros((K, 32), dtype=np.int64)
        for step in range(T):
            new = rng.integers(1, 1 << 30, size=K)           # what each beam writes at this position
            gathered[step] = new
            slots[step] = new
            # what the self-attention of beam k reads at this step
            for k in range(K):
                hist = [slots[t, anc[k, t]] for t in range(step)] + [slots[step, k]]
                assert hist == list(gathered[: step + 1, k])
            parent = rng.integers(0, K, size=K)              # beam_idx of the step (not a permutation)
            gathered[: step + 1] = gathered[: step + 1][:, parent]
            new_anc = anc.copy()
            for k in range(K):
                new_anc[k, :step] = anc[parent[k], :step]
                new_anc[k, step] = parent[k]
            anc = new_anc
