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
