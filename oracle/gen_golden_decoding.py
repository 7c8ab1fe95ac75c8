"""TEST INFRASTRUCTURE ONLY - pins the token-selection helpers of the oracle on the REFERENCE'S OWN utils/decoding_utils.py.

Run in the development container only (needs /root/reference):  ``python -m oracle.gen_golden_decoding``

``utils/decoding_utils.py`` imports nothing but torch, so the UNMODIFIED ``batch_top_k_top_p_sampling`` (:4-35, including the
``top_k = 0`` and ``top_p > 0`` branches the reference's callers never take) and ``batch_ngram_blocking`` (:38-78) run as they are
on seeded inputs.  The script asserts that ``oracle/restatement.py`` (``top_k_top_p_filter``, ``ngram_banned_tokens``) reproduces
them exactly and writes the reference outputs as ``tests/golden/decoding_utils.npz``: per case the boolean keep-mask of the
filter, and per row the sorted list of tokens the n-gram blocker bans.  ``tests/test_oracle_golden.py`` regenerates the inputs
from the same seeds and re-checks the oracle everywhere (the GPU box has no /root/reference).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

# (rows, vocab, top_k, top_p, temperature)
FILTER_CASES = [(4, 1000, 7, 0.0, 0.7), (4, 1000, 1, 0.0, 1.0), (4, 1000, 0, 0.9, 1.0), (4, 1000, 0, 0.5, 0.7), (4, 1000, 0, 0.05, 1.0),
                (4, 1000, 16, 0.9, 1.0), (4, 1000, 12, 0.3, 1.3), (3, 30522, 0, 0.9, 1.0), (3, 30522, 0, 0.3, 0.6), (3, 30522, 7, 0.5, 0.7),
                (2, 1000, 0, 0.0, 1.0)]
NGRAM_CASES = [(6, 64, 4, 5), (6, 64, 3, 4), (6, 64, 2, 3), (4, 256, 4, 12), (4, 256, 4, 2)]      # (rows, hist_len, n, prefix_len)


def filter_inputs(case: int) -> torch.Tensor:
    rows, V, _, _, temperature = FILTER_CASES[case]
    g = torch.Generator().manual_seed(4000 + case)
    x = torch.randn(rows, V, generator=g) * 3.0
    x[1, 5] = x[1].max() + 14.0                  # a peaked row
    x[0, 17] = x[0, 3]                           # an exact tie
    return x / temperature


def ngram_inputs(case: int):
    """History rows over a 12-token alphabet (so n-grams repeat), specials sprinkled in, answers masked like
    models/visual_dialog_model.py:98-99 (enc_input_ids * (segments == 0)); the decoded prefix ends inside the history."""
    rows, L, n, plen = NGRAM_CASES[case]
    g = torch.Generator().manual_seed(5000 + case)
    hist = torch.randint(200, 212, (rows, L), generator=g)
    hist[:, 0] = 101
    for r in range(rows):
        for p in torch.randint(1, L, (L // 10,), generator=g).tolist():
            hist[r, p] = [0, 100, 102, 103][p % 4]
    prefix = torch.randint(200, 212, (rows, plen), generator=g)
    prefix[:, 0] = 101
    for r in range(rows):                        # copy an (n-1)-gram of the history to the end of the prefix in most rows
        if r % 3 != 2 and plen >= n:
            s = int(torch.randint(1, L - n, (1,), generator=g))
            prefix[r, plen - (n - 1):] = hist[r, s:s + n - 1]
    return hist, prefix


def main():
    spec = importlib.util.spec_from_file_location("ref_decoding_utils", os.path.join(REF, "utils", "decoding_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from oracle import restatement as R
    out = {}
    for i, (rows, V, k, p, _) in enumerate(FILTER_CASES):
        x = filter_inputs(i)
        y_ref = ref.batch_top_k_top_p_sampling(x.clone(), top_k=k, top_p=p)
        y = R.top_k_top_p_filter(x.clone(), top_k=k, top_p=p)
        keep_ref = torch.isfinite(y_ref)
        assert torch.equal(keep_ref, torch.isfinite(y)), f"filter case {i}: the restatement keeps a different set"
        assert torch.equal(y_ref[keep_ref], y[keep_ref])
        out[f"keep_{i}"] = np.packbits(keep_ref.numpy(), axis=-1)
        out[f"kept_{i}"] = keep_ref.sum(-1).numpy()
    for i, (rows, L, n, plen) in enumerate(NGRAM_CASES):
        hist, prefix = ngram_inputs(i)
        V = 400
        logits = torch.zeros(rows, V)
        y_ref = ref.batch_ngram_blocking(logits.clone(), hist, prefix, ngram_size=n)
        banned_ref = [sorted(set(torch.nonzero(torch.isinf(y_ref[r])).flatten().tolist())) for r in range(rows)]
        banned = [sorted(set(R.ngram_banned_tokens(hist[r].tolist(), prefix[r].tolist(), n))) for r in range(rows)]
        assert banned == banned_ref, f"n-gram case {i}: {banned} != {banned_ref}"
        flat = np.full((rows, 16), -1, dtype=np.int64)
        for r, b in enumerate(banned_ref):
            assert len(b) <= 16
            flat[r, :len(b)] = b
        out[f"banned_{i}"] = flat
        assert any(len(b) for b in banned_ref) or n == 4 and plen < n, f"n-gram case {i} bans nothing: weak case"
    path = os.path.join(ROOT, "tests", "golden", "decoding_utils.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("kept")}, [out[f"kept_{i}"].tolist() for i in range(len(FILTER_CASES))])


if __name__ == "__main__":
    main()
