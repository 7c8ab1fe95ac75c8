"""TEST INFRASTRUCTURE ONLY - CPU restatement of the round loop of the reference's generate.py:122-233.

Questioner -> splice -> teacher -> perplexity pass -> splice(segment 1), with the reference's own sampler made
deterministic (top_k=1) or with the beam contract of oracle/beam.py.  Single-device semantics: the
perplexity pass replaces [SEP] by [PAD] in the answer tensor in place, so ans_len and the spliced answer exclude [SEP].
Pinned indirectly: every building block is checked against the reference's modules in oracle/gen_golden.py.
"""
from __future__ import annotations

import torch

from . import beam as OB
from . import restatement as R


def generate_dialogs(sd_a, enc_cfg, dec_cfg, batch, sd_q=None, questions=None, num_rounds=10, a_kwargs=None, q_kwargs=None,
                     with_ppl=True):
    a_kwargs = dict(a_kwargs or dict(temperature=1.0, top_k=1, top_p=0.0, ngram_blocking_size=0))
    q_kwargs = dict(q_kwargs or dict(temperature=1.0, top_k=1, top_p=0.0, ngram_blocking_size=4))
    b = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    ids, seg = b["enc_input_ids"], b["enc_segments"]
    B = ids.shape[0]
    enc_len = (ids != 0).sum(-1)
    abnormal = set()
    ques_all, ans_all, ppl_all = [], [], []

    def run(sd, kw):
        kw = dict(kw)
        nb = kw.pop("num_beams", 1)
        if nb > 1:
            return OB.beam_search(sd, enc_cfg, dec_cfg, b, num_beams=nb)[0]
        return R.generate_greedy_or_sample(sd, enc_cfg, dec_cfg, b, kw["temperature"], kw["top_k"], kw["top_p"], kw["ngram_blocking_size"])

    with torch.no_grad():
        for rnd in range(num_rounds):
            ques = questions[:, rnd].clone() if sd_q is None else run(sd_q, q_kwargs)
            enc_len = enc_len + R.splice(ids, seg, enc_len, ques, None, abnormal)           # generate.py:145-160
            b["enc_att_mask"] = (ids != 0).float()
            ans = run(sd_a, a_kwargs)                                                        # generate.py:163-181
            if with_ppl:
                _, _, ppl = R.score_answers(sd_a, enc_cfg, dec_cfg, b, ans)                  # generate.py:183-209
                ppl_all.append(ppl)
            ans = ans.masked_fill(ans == R.EOS, R.PAD)                                       # in-place effect of the ppl pass (:57)
            enc_len = enc_len + R.splice(ids, seg, enc_len, ans, 1, abnormal)                # generate.py:214-228
            b["enc_att_mask"] = (ids != 0).float()
            ques_all.append(ques)
            ans_all.append(ans)
    flags = torch.zeros(B, dtype=torch.int32)
    for i in abnormal:
        flags[i] = 1
    return torch.stack(ques_all, 1), torch.stack(ans_all, 1), (torch.stack(ppl_all, 1) if with_ppl else None), flags, ids
