"""Generative ranking of answer options (the scoring loop of the reference's evaluate_gen.py:62-107).

The reference expands the image features / history ``num_options`` times and runs the full encoder once per option
(evaluate_gen.py:64-70) although every option of a round shares the same (image, history).  Here the encoder and the
cross-attention K/V prefill run ONCE per (image, round); the options are teacher-forced through the decoder against that
shared state (gstvd_score_options) - a ``num_options``-fold encoder saving.  Score = sum of target log-probabilities over
non-pad positions, i.e. ``F.log_softmax(logits).gather(target)`` masked and summed (evaluate_gen.py:95-106).
"""
from __future__ import annotations

import torch


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def score_options(model, batch, option_ids: torch.Tensor, option_labels: torch.Tensor = None, device=None) -> torch.Tensor:
    """``option_ids`` int64 [B, O, L] decoder inputs ([CLS] a1 .. aN [SEP] 0 ..); ``option_labels`` [B, O, L] targets
    (default: inputs shifted left, 0 = ignore).  Returns log-likelihood scores fp32 [B, O] (higher = better)."""
    m = _unwrap(model)
    dev = torch.device(device) if device is not None else next(m.parameters()).device
    eng = m._engine(dev)
    B, O, L = option_ids.shape
    if B * O > eng.max_batch:
        raise ValueError(f"B*O = {B * O} sequences exceed engine_max_batch = {eng.max_batch}; score in chunks of images")
    enc = eng.encode(batch["enc_input_ids"], batch["enc_image_feat"], batch["enc_image_loc"], batch["enc_segments"],
                     batch["enc_att_mask"], batch["enc_image_mask"])
    eng.prefill_cross(B, enc["Le"])
    ids = option_ids.to(device=dev, dtype=torch.int64).reshape(B * O, L).contiguous().clone()
    if option_labels is None:
        labels = torch.zeros_like(ids)
        labels[:, :-1] = ids[:, 1:]
    else:
        labels = option_labels.to(device=dev, dtype=torch.int64).reshape(B * O, L).contiguous()
    mask = (ids != 0).float()
    loss, _ = eng.score(ids, mask, labels=labels, want_logits=False, options_per_image=O)
    return (-loss.sum(-1)).reshape(B, O)


def _trim(ids, seg, att, bound):
    """Leading ceil32(bound) text positions (see ``enc_valid_len`` in EncoderDecoderModel.forward): exact for every valid position."""
    if bound is None:
        return ids, seg, att
    Lt = ids.shape[1]
    Le = min(Lt, max(32, (int(bound) + 31) // 32 * 32))
    if Le >= Lt:
        return ids, seg, att
    return ids[:, :Le].contiguous(), (seg[:, :Le].contiguous() if seg is not None else None), (att[:, :Le].contiguous() if att is not None else None)


def answer_perplexity(model, batch, answer_ids: torch.Tensor, device=None, trim_history=False, hist_len_bound=None) -> torch.Tensor:
    """Teacher-forced perplexity of ``answer_ids`` int64 [B, L] (no leading [CLS], zero padded) under the model: the pass
    of generate.py:183-209, which is also BASELINE config 4 (scoring (context, answer) pairs for -select_data).
    ppl = exp(sum CE / count(ids != 0)); labels are the ids shifted left, so the first token is unscored and a generated
    [SEP] counts in the numerator but not in the length (it is replaced by [PAD] in place before the count).  fp32 [B].
    ``trim_history`` + ``hist_len_bound`` (host int: no row holds a token at or past it): the encoder runs on ceil32(bound) text
    positions, exactly like ``enc_valid_len`` of EncoderDecoderModel.forward."""
    m = _unwrap(model)
    dev = torch.device(device) if device is not None else next(m.parameters()).device
    eng = m._engine(dev)
    B = answer_ids.shape[0]
    ids_e, seg_e, att_e = _trim(batch["enc_input_ids"], batch["enc_segments"], batch["enc_att_mask"], hist_len_bound if trim_history else None)
    enc = eng.encode(ids_e, batch["enc_image_feat"], batch["enc_image_loc"], seg_e, att_e, batch["enc_image_mask"])
    eng.prefill_cross(B, enc["Le"])
    ids = answer_ids.to(device=dev, dtype=torch.int64).contiguous().clone()
    mask = (ids != 0).float()
    loss, _ = eng.score(ids, mask, labels=None, want_logits=False)      # shifts the labels and maps [SEP] -> [PAD] in ids
    return torch.exp(loss.sum(-1) / (ids != 0).sum(-1))


def select_mask(answer_ppl: torch.Tensor, threshold: float = 50) -> torch.Tensor:
    """-select_data (dataloader/dataloader_cc12m_gen.py:193-199): answers with ppl >= threshold have their labels zeroed
    (they stay in the dialog as context but are not trained on).  True = keep.  NaN ppl (an answer that is only [SEP])
    compares False with >=, so - as in the reference - it is kept."""
    return ~(answer_ppl >= threshold)


def nsp_rank(encoder, item, device=None) -> torch.Tensor:
    """Discriminative ranking step of evaluate_disc.py:79-83 (BASELINE config 5): every row of ``item`` is one
    (image, history + candidate answer) encoder input; returns softmax(nsp)[:, 0] fp32 [rows].  ``item`` uses the
    reference's chunk keys (tokens, segments, mask, image_feat, image_loc, image_mask)."""
    dev = torch.device(device) if device is not None else next(encoder.parameters()).device
    out = encoder(item["tokens"].to(dev), item["image_feat"].to(dev), item["image_loc"].to(dev), token_type_ids=item["segments"].to(dev),
                  attention_mask=item["mask"].to(dev), image_attention_mask=item["image_mask"].to(dev))
    return torch.softmax(out[3].float(), dim=1)[:, 0]


def nsp_rank_items(encoder, items, chunk: int = 500, device=None, trim_history=False) -> torch.Tensor:
    """BASELINE config 5 end to end: ``items`` holds image tensors per item ([I, 37, ...]: image_feat, image_loc, image_mask) and
    token tensors per candidate ([I, C, L]: tokens, segments, mask; optional hist_len_bound).  The reference expands the image
    features per candidate on the host (evaluate_disc.py:66-77); here they cross PCIe once per item and are expanded on the
    device per chunk of ``chunk`` encoder rows (the reference's chunk is 200, evaluate_disc.py:25).  Returns softmax(nsp)[..., 0]
    fp32 [I, C]."""
    dev = torch.device(device) if device is not None else next(encoder.parameters()).device
    nb = dict(device=dev, non_blocking=True)
    tok, seg, msk = items["tokens"].to(**nb), items["segments"].to(**nb), items["mask"].to(**nb)
    feat, loc, imask = items["image_feat"].to(**nb), items["image_loc"].to(**nb), items["image_mask"].to(**nb)
    n_items, n_cand, L = tok.shape
    bound = int(items["hist_len_bound"][0]) if (trim_history and "hist_len_bound" in items) else None
    tok, seg, msk = tok.reshape(-1, L), seg.reshape(-1, L), msk.reshape(-1, L)
    tok, seg, msk = _trim(tok, seg, msk, bound)
    owner = torch.arange(n_items, device=dev).repeat_interleave(n_cand)
    out = []
    for s0 in range(0, tok.shape[0], chunk):
        idx = owner[s0:s0 + chunk]
        part = {"tokens": tok[s0:s0 + chunk], "segments": seg[s0:s0 + chunk], "mask": msk[s0:s0 + chunk],
                "image_feat": feat.index_select(0, idx), "image_loc": loc.index_select(0, idx), "image_mask": imask.index_select(0, idx)}
        out.append(nsp_rank(encoder, part, device=dev))
    return torch.cat(out).reshape(n_items, n_cand)


def scores_to_ranks(scores: torch.Tensor) -> torch.Tensor:
    """[B, R, O] option scores -> 1-based ranks, the largest score gets rank 1 (utils/visdial_metrics.py:21-39, vectorised:
    the reference fills ``ranks[i][ranked_idx[i][j]] = j`` in two Python loops; a scatter of arange does the same)."""
    B, R, O = scores.shape
    flat = scores.reshape(-1, O)
    ranked_idx = flat.sort(1, descending=True).indices
    ranks = torch.empty_like(ranked_idx)
    ranks.scatter_(1, ranked_idx, torch.arange(O, device=flat.device).expand_as(ranked_idx))
    return (ranks + 1).reshape(B, R, O)


def sparse_metrics(gt_ranks: torch.Tensor) -> dict:
    """Recall@{1,5,10}, mean rank and mean reciprocal rank of the ground-truth option (utils/visdial_metrics.py:79-95)."""
    r = gt_ranks.reshape(-1).float()
    return {"r@1": (r <= 1).float().mean().item(), "r@5": (r <= 5).float().mean().item(), "r@10": (r <= 10).float().mean().item(),
            "mean": r.mean().item(), "mrr": r.reciprocal().mean().item()}
