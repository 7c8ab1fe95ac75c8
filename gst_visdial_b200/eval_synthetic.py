"""Seeded synthetic evaluation items for evaluate_gen.py (no dataset needed; nothing here touches oracle/)."""
import torch

from . import synthetic as S


def synthetic_eval_item(start, count, rnd, num_options, vocab_size, v_feature_size, max_seq_len=256, max_utt_len=25):
    """Contexts of images ``start .. start+count`` at round ``rnd`` (caption + rnd question/answer pairs + the current
    question, spliced like generate.py:145-160 / :214-228) and ``num_options`` candidate answers per image:
    int64 [count, num_options, max_utt_len] = [CLS] tokens [SEP] 0..; option ``gt[i]`` is the designated ground truth."""
    batch = S.synthetic_batch(start, count, vocab_size=vocab_size, v_feature_size=v_feature_size, max_seq_len=max_seq_len)
    ids, seg = batch["enc_input_ids"], batch["enc_segments"]
    for i in range(count):
        n = int((ids[i] != 0).sum())
        for r in range(rnd + 1):
            utts = [(S.synthetic_utterance(start + i, 2 * r, vocab_size), 0)]
            if r < rnd:
                utts.append((S.synthetic_utterance(start + i, 2 * r + 1, vocab_size), 1))
            for u, sgm in utts:
                toks = u[u != 0]
                if sgm == 1:
                    toks = toks[toks != 102]                                    # answers reach the history without [SEP]
                if n + len(toks) > max_seq_len:
                    break
                ids[i, n:n + len(toks)] = toks
                seg[i, n:n + len(toks)] = sgm
                n += len(toks)
    batch["enc_att_mask"] = (ids != 0).float()
    g = torch.Generator().manual_seed(99991 * (start + 1) + rnd)
    opts = torch.zeros(count, num_options, max_utt_len, dtype=torch.int64)
    for i in range(count):
        for o in range(num_options):
            n = int(torch.randint(2, max_utt_len - 2, (1,), generator=g))
            opts[i, o, 0] = 101
            opts[i, o, 1:1 + n] = torch.randint(min(1000, vocab_size // 2), vocab_size, (n,), generator=g)
            opts[i, o, 1 + n] = 102
    gt = torch.randint(0, num_options, (count,), generator=g)
    return batch, opts, gt
