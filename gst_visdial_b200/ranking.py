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
