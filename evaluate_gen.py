"""Generative evaluation: rank the answer options of every round by their log-likelihood under the teacher
(the reference's evaluate_gen.py:25-160).

    python evaluate_gen.py -mode vd_eval_val -start_path ckpt -save_name ranks.json            (reference flags)
    python evaluate_gen.py -synthetic 8 -num_options 100 -num_rounds 10                         (no dataset / checkpoint needed)

The reference flattens (image, round, option) into rows, expands the image features x (rounds x options) and pushes chunks
of 500 rows through encoder AND decoder (evaluate_gen.py:62-107) although the 100 options of a round share one
(image, history, question).  Here each (image, round) context is encoded once and its options are teacher-forced against the
shared cross-attention K/V (gst_visdial_b200.ranking.score_options -> gstvd_score_options).  Output: the reference's ranks
json (image_id, round_id, ranks) and, when ground-truth option indices are known, R@1/5/10, mean rank and MRR.
The VisDial dataset readers are host I/O outside this package: without them (-synthetic) contexts and options are seeded
random token sequences with the real shapes.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from gst_visdial_b200 import dist as D, options, synthetic as S, weights as W  # noqa: E402
from gst_visdial_b200.models.visual_dialog_decoder import VisualDialogDecoder  # noqa: E402
from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder  # noqa: E402
from gst_visdial_b200.models.visual_dialog_model import EncoderDecoderModel  # noqa: E402
from gst_visdial_b200.ranking import scores_to_ranks, score_options, sparse_metrics  # noqa: E402
from gst_visdial_b200.eval_synthetic import synthetic_eval_item  # noqa: E402


def main(argv=None):
    params = options.read_command_line(argv)
    rank, world, local = D.init_from_env("nccl")
    if world > 1:
        params['gpu_ids'] = [local]
        params['device'] = f"cuda:{local}"
    torch.cuda.set_device(params['device'])
    if params['synthetic'] <= 0:
        raise SystemExit("the VisDial readers of the reference are host I/O outside this package; run with -synthetic N or call "
                         "gst_visdial_b200.ranking.score_options() with batches from the reference's VisdialDataset")
    O, R = params['num_options'], params['num_rounds']
    per_call = max(1, params['batch_size'] // O)                      # images scored per call: rows = images x options
    p = dict(params)
    p['model'], p['mode'] = 'enc_dec_a', 'vd_eval_val'
    p['engine_max_batch'] = per_call * O                               # evaluate_gen.py:29 scores 500 rows per forward
    enc, dec = VisualDialogEncoder(p), VisualDialogDecoder(p)
    dec.decoder.bert.embeddings = enc.bert_pretrained.bert.embeddings
    model = EncoderDecoderModel(p, enc, dec)
    if params['start_path']:
        model.load_state_dict(torch.load(params['start_path'], map_location='cpu')['model_state_dict'])
    else:
        model.load_state_dict(W.synthetic_state_dict(enc.config, dec.config, seed=0))
    model.to(params['device']).eval()
    total = params['synthetic']
    start, end = D.shard_range(total, rank, world)
    ranks_json, gt_ranks = [], []
    with torch.no_grad():
        for rnd in range(R):                                           # one context per (image, round)
            for s in range(start, end, per_call):
                n = min(per_call, end - s)
                batch, opts, gt = synthetic_eval_item(s, n, rnd, O, enc.config.vocab_size, enc.config.v_feature_size, params['max_seq_len'],
                                                      params['max_utt_len'])
                scores = score_options(model, batch, opts, device=params['device'])             # [n, O]
                ranks = scores_to_ranks(scores.unsqueeze(1)).squeeze(1).cpu()
                gt_ranks.append(ranks[torch.arange(n), gt])
                for i in range(n):
                    ranks_json.append({"image_id": int(batch["image_id"][i]), "round_id": rnd + 1, "ranks": ranks[i].tolist()})
    gt_all = torch.cat(gt_ranks) if gt_ranks else torch.zeros(0)
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, (ranks_json, gt_all))
        ranks_json = [r for part in gathered for r in part[0]]
        gt_all = torch.cat([part[1] for part in gathered])
    if rank == 0:
        path = os.path.join(params['save_path'], params['save_name'] if params['save_name'] != 'generated_dialogs.json' else 'ranks.json')
        json.dump(ranks_json, open(path, "w"))
        m = sparse_metrics(gt_all)
        print(f"wrote {len(ranks_json)} rank lists to {path}")
        print(json.dumps(m))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
