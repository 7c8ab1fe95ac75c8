"""TEST INFRASTRUCTURE ONLY - beam-search oracle (plain PyTorch / Python, CPU).

The reference has no beam search (utils/decoding_utils.py holds only top-k/top-p filtering and n-gram
blocking; the only beam hook is ``_reorder_cache``, models/visual_dialog_decoder.py:29-31,177-181, which is
never called).  ``BASELINE.json.north_star`` asks for beam-5, so the contract is defined here, following the
published algorithm of ``transformers==4.16.2`` (the reference's pinned dependency, requirements.txt:9):
``GenerationMixin.beam_search`` + ``BeamSearchScorer.process/finalize`` + ``BeamHypotheses.add/is_done`` with
``length_penalty=1.0``, ``early_stopping=False``, ``num_return_sequences=1``, ``max_length = 1 + max_new``.
Parity unpinned against the reference (there is nothing to pin against); the CUDA beam kernel must match
THIS file bit-exactly on indices when both are fed the same fp32 logits.

Arithmetic contract (chosen so that two implementations agree bit-for-bit):
  * logZ = fp32( m + log( sum_j exp(fp64(x_j) - fp64(m)) ) ),  m = max_j x_j   (sum and log in fp64)
  * candidate score = fp32(fp32(x - logZ) + beam_score)   (two fp32 roundings, no FMA)
  * top-2K over the K*V candidates of one image ordered by (score desc, flat index asc)
  * hypothesis score = fp64(sum_logprobs) / fp64(length incl. the start token); comparisons in fp64
  * cache reorder semantic = ``index_select(0, beam_idx)`` (models/visual_dialog_decoder.py:177-181)
"""
from __future__ import annotations

import torch

EOS, PAD = 102, 0


def log_z(logits: torch.Tensor) -> torch.Tensor:
    m = logits.max(-1, keepdim=True).values
    s = torch.exp(logits.double() - m.double()).sum(-1, keepdim=True)
    return (m.double() + torch.log(s)).float()


def candidate_scores(logits: torch.Tensor, beam_scores: torch.Tensor) -> torch.Tensor:
    """logits [B*K, V] fp32, beam_scores [B, K] fp32 -> [B, K*V] fp32."""
    B, K = beam_scores.shape
    lp = logits - log_z(logits)
    return (lp + beam_scores.reshape(B * K, 1)).reshape(B, -1)


def top_candidates(scores: torch.Tensor, n: int):
    """Top-n of each row by (score desc, index asc).  A stable descending sort defines the tie order."""
    order = torch.sort(scores, dim=-1, descending=True, stable=True).indices[:, :n]
    return scores.gather(1, order), order


class BeamState:
    """Per-batch beam bookkeeping.  ``tokens`` holds generated tokens only (the [CLS] start token is implicit and
    counts 1 towards hypothesis length)."""

    def __init__(self, B: int, K: int, V: int, max_new: int = 18):
        self.B, self.K, self.V, self.T = B, K, V, max_new
        self.t = 0
        self.beam_scores = torch.zeros(B, K, dtype=torch.float32)
        self.beam_scores[:, 1:] = -1e9
        self.tokens = torch.zeros(B, K, max_new, dtype=torch.int64)
        self.done = torch.zeros(B, dtype=torch.bool)
        self.hyps = [[] for _ in range(B)]          # list of (score fp64, tokens list, insertion order)
        self.worst = [1e9] * B
        self.last_beam_idx = None                   # [B, K] parent beam (0..K-1) of every new beam
        self.last_tokens = None                     # [B, K] token appended to every new beam

    # BeamHypotheses.add
    def _add(self, b, toks, sum_logprobs, length):
        score = float(sum_logprobs) / float(length)
        h = self.hyps[b]
        if len(h) < self.K or score > self.worst[b]:
            h.append((score, list(toks)))
            if len(h) > self.K:
                srt = sorted([(s, i) for i, (s, _) in enumerate(h)])
                del h[srt[0][1]]
                self.worst[b] = srt[1][0]
            else:
                self.worst[b] = min(score, self.worst[b])

    def step(self, logits: torch.Tensor):
        """One ``beam_search`` iteration on logits [B*K, V] of the current last position."""
        B, K, V = self.B, self.K, self.V
        cur_len = self.t + 1                         # input length incl. the start token
        sc = candidate_scores(logits.float(), self.beam_scores)
        val, idx = top_candidates(sc, 2 * K)
        nb_scores = torch.zeros(B, K, dtype=torch.float32)
        nb_tokens = torch.zeros(B, K, dtype=torch.int64)
        nb_idx = torch.zeros(B, K, dtype=torch.int64)
        for b in range(B):
            if self.done[b]:
                continue                              # padded: score 0, token PAD, parent 0
            n = 0
            for rank in range(2 * K):
                flat = int(idx[b, rank]); beam, tok = flat // V, flat % V
                s = val[b, rank]
                if tok == EOS:
                    if rank >= K:
                        continue
                    self._add(b, self.tokens[b, beam, : self.t].tolist(), s.item(), cur_len)
                else:
                    nb_scores[b, n] = s; nb_tokens[b, n] = tok; nb_idx[b, n] = beam
                    n += 1
                if n == K:
                    break
            assert n == K
            # BeamHypotheses.is_done(best_sum_logprobs = max over the 2K candidates, cur_len)
            if len(self.hyps[b]) >= K:
                cur = float(val[b].max().item()) / float(cur_len)
                if self.worst[b] >= cur:
                    self.done[b] = True
        new_tokens = self.tokens.gather(1, nb_idx[:, :, None].expand(-1, -1, self.T)).clone()
        new_tokens[:, :, self.t] = nb_tokens
        self.tokens = new_tokens
        self.beam_scores = nb_scores
        self.last_beam_idx, self.last_tokens = nb_idx, nb_tokens
        self.t += 1
        return nb_idx, nb_tokens, nb_scores

    def finalize(self):
        """BeamSearchScorer.finalize: returns (sequences [B, T] int64 zero padded, EOS appended when it fits;
        scores [B] fp64 = best hypothesis score)."""
        B, K, T = self.B, self.K, self.T
        seq = torch.zeros(B, T, dtype=torch.int64)
        best_scores = torch.zeros(B, dtype=torch.float64)
        for b in range(B):
            if not self.done[b]:
                for k in range(K):
                    self._add(b, self.tokens[b, k, : self.t].tolist(), self.beam_scores[b, k].item(), self.t + 1)
            srt = sorted(self.hyps[b], key=lambda x: x[0])
            score, toks = srt[-1]
            seq[b, : len(toks)] = torch.tensor(toks, dtype=torch.int64)
            if len(toks) < T:
                seq[b, len(toks)] = EOS
            best_scores[b] = score
        return seq, best_scores


def reorder_cache(past, beam_idx):
    """models/visual_dialog_decoder.py:177-181."""
    return tuple(tuple(p.index_select(0, beam_idx) for p in layer) for layer in past)


def beam_search(sd, enc_cfg, dec_cfg, batch, num_beams=5, max_new=18, return_trace=False, precomputed_encoder=None):
    """Full beam search over the restated reference modules (no KV cache, whole prefix every step, encoder
    states repeated per beam - the way the reference's ``use_cache=False`` decoder would have to be driven)."""
    from . import restatement as R
    ids, seg, att = batch["enc_input_ids"], batch["enc_segments"], batch["enc_att_mask"]
    B, K = ids.shape[0], num_beams
    if precomputed_encoder is not None:
        seq_t, seq_v = precomputed_encoder
    else:
        seq_t, seq_v = R.encoder(sd, enc_cfg, ids, batch["enc_image_feat"], batch["enc_image_loc"], seg, att, batch["enc_image_mask"])
    enc_h, enc_m = R.vlfusion(sd, seq_t, seq_v, att, batch["enc_image_mask"])
    enc_h = enc_h.repeat_interleave(K, 0); enc_m = enc_m.repeat_interleave(K, 0)
    st = BeamState(B, K, dec_cfg.vocab_size, max_new)
    trace = []
    for t in range(max_new):
        prefix = torch.cat((torch.full((B * K, 1), R.CLS, dtype=torch.int64), st.tokens.reshape(B * K, -1)[:, :t]), 1)
        h = R.decoder_hidden(sd, dec_cfg, prefix, None, enc_h, enc_m)
        logits = R.lm_logits(sd, h[:, -1])
        bi, bt, bs = st.step(logits)
        if return_trace:
            trace.append((logits, bi.clone(), bt.clone(), bs.clone()))
        if bool(st.done.all()):
            break
    seq, scores = st.finalize()
    return (seq, scores, trace) if return_trace else (seq, scores)
