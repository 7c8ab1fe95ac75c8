#!/usr/bin/env python
"""generate.py - synthetic visual-dialog generation with the questioner and teacher models (CLI of the reference's
generate.py: same single-dash flags, same output JSON layout), running on the B200 engine.

    python generate.py -mode cc12m_gen -start_path_q ckpt_q -start_path_a ckpt_a -cc12m_image_feats ... -save_name out.json
    python generate.py -synthetic 128 -batch_size 64 -num_beams 5          # seeded synthetic images + random weights

Multi-GPU: launch with torchrun (one process per GPU); images are sharded by rank and gathered once at the end.
The dataset readers (LMDB / tokenizer) of the reference are host I/O outside the hot path: without them (-synthetic)
the script uses the seeded generators of gst_visdial_b200.synthetic and writes token ids instead of decoded text.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gst_visdial_b200 import dist as D  # noqa: E402
from gst_visdial_b200 import options  # noqa: E402
from gst_visdial_b200 import synthetic as S  # noqa: E402
from gst_visdial_b200 import weights as W  # noqa: E402
from gst_visdial_b200.dialog import generate_dialogs  # noqa: E402
from gst_visdial_b200.models.visual_dialog_decoder import VisualDialogDecoder  # noqa: E402
from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder  # noqa: E402
from gst_visdial_b200.models.visual_dialog_model import EncoderDecoderModel  # noqa: E402


def build_model(params, model_name, ckpt_path, seed):
    p = dict(params)
    p['model'] = model_name
    enc, dec = VisualDialogEncoder(p), VisualDialogDecoder(p)
    dec.decoder.bert.embeddings = enc.bert_pretrained.bert.embeddings          # generate.py:65
    model = EncoderDecoderModel(p, enc, dec)
    model = torch.nn.DataParallel(model, [p['gpu_ids'][0]])
    if ckpt_path:
        ckpt = torch.load(ckpt_path, map_location='cpu')
        model.module.load_state_dict(ckpt['model_state_dict'])
        print(f"model successfully loaded from {ckpt_path}")
    else:
        model.module.load_state_dict(W.synthetic_state_dict(enc.config, dec.config, seed=seed))
        print(f"[{model_name}] no checkpoint given: seeded random weights (seed {seed})")
    model.to(p['device']).eval()
    return model


def ids_to_text(row):
    return " ".join(str(int(t)) for t in row if int(t) != 0)


def main(argv=None):
    params = options.read_command_line(argv)
    rank, world, local = D.init_from_env("nccl")
    if world > 1:
        params['gpu_ids'] = [local]
        params['device'] = f"cuda:{local}"
    torch.cuda.set_device(params['device'])
    if params['synthetic'] <= 0:
        raise SystemExit("the LMDB / tokenizer data path of the reference is host I/O outside this package; run with -synthetic N "
                         "or feed gst_visdial_b200.dialog.generate_dialogs() with batches from the reference's CC12mDataset")
    q_model = build_model(params, 'enc_dec_q', params['start_path_q'], seed=7)
    a_model = build_model(params, 'enc_dec_a', params['start_path_a'], seed=0)
    enc_cfg = a_model.module.encoder.config
    total = params['synthetic']
    start, end = D.shard_range(total, rank, world)
    a_kwargs = dict(temperature=0.7, top_k=7, top_p=0.0, ngram_blocking_size=0)
    if params['num_beams'] > 1:
        a_kwargs.update(num_beams=params['num_beams'])
    q_kwargs = dict(temperature=0.7, top_k=7, top_p=0.0, ngram_blocking_size=4)
    out = []
    with torch.no_grad():
        for s in range(start, end, params['batch_size']):
            n = min(params['batch_size'], end - s)
            batch = S.synthetic_batch(s, n, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size,
                                      max_seq_len=params['max_seq_len'])
            res = generate_dialogs(a_model, batch, q_model=q_model, num_rounds=params['num_rounds'], a_kwargs=a_kwargs, q_kwargs=q_kwargs,
                                   with_ppl=True, device=params['device'])
            q, a, ppl, abn = res.questions.cpu(), res.answers.cpu(), res.answer_ppl.cpu(), res.abnormal.cpu()
            for j in range(n):
                if abn[j]:
                    continue                                                          # generate.py:236-237
                out.append({"image_id": int(batch["image_id"][j]), "url": "", "caption": ids_to_text(batch["enc_input_ids"][j]),
                            "dialog": [{"question": ids_to_text(q[j, k]), "answer": ids_to_text(a[j, k]),
                                        "answer_ppl": float(ppl[j, k])} for k in range(params['num_rounds'])]})
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, out)
        out = [d for part in gathered for d in part]
    if rank == 0:
        path = os.path.join(params['save_path'], params['save_name'])
        json.dump(out, open(path, "w"))
        print(f"wrote {len(out)} dialogs to {path}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
