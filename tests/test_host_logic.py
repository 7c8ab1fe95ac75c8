"""CPU: host-side logic - key layout of the boundary modules, synthetic generators, sharding + gather (gloo, world 2)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_module_state_dict_layout_matches_spec(tiny_cfgs, tiny_sd):
    from gst_visdial_b200 import weights as W
    from gst_visdial_b200.models.visual_dialog_decoder import VisualDialogDecoder
    from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder
    from gst_visdial_b200.models.visual_dialog_model import EncoderDecoderModel
    params = {"model_enc_config": W.TINY_ENC_CONFIG, "model_dec_config": W.TINY_DEC_CONFIG, "gpu_ids": [0], "model": "enc_dec_a",
              "mode": "cc12m_gen"}
    enc, dec = VisualDialogEncoder(params), VisualDialogDecoder(params)
    dec.decoder.bert.embeddings = enc.bert_pretrained.bert.embeddings
    model = EncoderDecoderModel(params, enc, dec)
    spec = W.model_spec(*tiny_cfgs)
    sd = model.state_dict()
    assert list(sd.keys()) == list(spec.keys()) or set(sd.keys()) == set(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    v0 = model._version.value
    model.load_state_dict(tiny_sd, strict=True)
    assert model._version.value > v0
    w = model.encoder.bert_pretrained.bert.embeddings.word_embeddings.weight
    assert model.decoder.decoder.bert.embeddings.word_embeddings.weight is w          # generate.py:65 aliasing survives
    assert torch.equal(w, tiny_sd["encoder.bert_pretrained.bert.embeddings.word_embeddings.weight"])
    assert hasattr(model.decoder.config, "eos_token_id") and model.decoder.config.eos_token_id == 102
    assert model.decoder.decoder.config.vocab_size == tiny_cfgs[1].vocab_size
    assert "hidden_size" in model.encoder.config.to_dict()
    # the full 6-layer/6-connect layout has the 861 keys of the released checkpoints
    full = W.model_spec(W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG))
    assert len(full) == 861


def test_modules_refuse_cpu_tensors(tiny_cfgs):
    import pytest
    from gst_visdial_b200 import synthetic as S
    from gst_visdial_b200 import weights as W
    from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder
    params = {"model_enc_config": W.TINY_ENC_CONFIG, "model_dec_config": W.TINY_DEC_CONFIG, "gpu_ids": [0], "model": "enc_only_a",
              "mode": "vd_eval_val"}
    enc = VisualDialogEncoder(params)
    b = S.synthetic_batch(0, 1, vocab_size=tiny_cfgs[0].vocab_size, v_feature_size=tiny_cfgs[0].v_feature_size)
    with pytest.raises(RuntimeError):
        enc(b["enc_input_ids"], b["enc_image_feat"], b["enc_image_loc"])


def test_synthetic_inputs_are_shard_invariant():
    from gst_visdial_b200 import synthetic as S
    a = S.synthetic_batch(0, 6, vocab_size=1000, v_feature_size=64)
    b = S.synthetic_batch(4, 2, vocab_size=1000, v_feature_size=64)
    assert torch.equal(a["enc_input_ids"][4:], b["enc_input_ids"]) and torch.equal(a["enc_image_feat"][4:], b["enc_image_feat"])
    assert (a["enc_image_feat"][:, 0] - a["enc_image_feat"][:, 1:].mean(1)).abs().max() < 1e-5      # global row = mean
    assert a["enc_input_ids"][:, 0].eq(101).all() and (a["enc_att_mask"] == (a["enc_input_ids"] != 0).float()).all()


def test_shard_range_partitions():
    from gst_visdial_b200.dist import shard_range
    for total in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, e = shard_range(total, r, world)
                seen.extend(range(s, e))
            assert seen == list(range(total))


WORKER = r"""
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from gst_visdial_b200 import dist as D
rank, world, _ = D.init_from_env("gloo")
total = 7
s, e = D.shard_range(total, rank, world)
ids = torch.arange(s, e).view(-1, 1).repeat(1, 3) * 10 + rank
ppl = torch.arange(s, e).float() + 0.5
counts = [D.shard_range(total, r, world)[1] - D.shard_range(total, r, world)[0] for r in range(world)]
g_ids, g_ppl = D.gather_results([ids, ppl], counts)
assert g_ids.shape == (total, 3) and g_ppl.tolist() == [i + 0.5 for i in range(total)], (g_ids, g_ppl)
assert (g_ids[:, 0] // 10).tolist() == list(range(total))
print("rank", rank, "ok")
"""


def test_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_scores_to_ranks_matches_reference_loops():
    """utils/visdial_metrics.py:21-39 restated literally (two Python loops) vs the vectorised scatter; metrics sanity."""
    import torch
    from gst_visdial_b200.ranking import scores_to_ranks, sparse_metrics
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(3, 4, 11, generator=g)
    flat = scores.view(-1, 11)
    ranked_idx = flat.sort(1, descending=True)[1]
    ref = ranked_idx.clone().fill_(0)
    for i in range(ranked_idx.size(0)):
        for j in range(11):
            ref[i][ranked_idx[i][j]] = j
    ref = (ref + 1).view(3, 4, 11)
    assert torch.equal(scores_to_ranks(scores), ref)
    m = sparse_metrics(torch.tensor([1, 2, 6, 11]))
    assert m["r@1"] == 0.25 and m["r@5"] == 0.5 and m["r@10"] == 0.75 and abs(m["mean"] - 5.0) < 1e-6


def test_synthetic_eval_items_are_reproducible_and_well_formed():
    import torch
    from gst_visdial_b200.eval_synthetic import synthetic_eval_item
    b1, o1, g1 = synthetic_eval_item(4, 3, 2, 9, 30522, 64)
    b2, o2, g2 = synthetic_eval_item(4, 3, 2, 9, 30522, 64)
    assert torch.equal(o1, o2) and torch.equal(g1, g2) and torch.equal(b1["enc_input_ids"], b2["enc_input_ids"])
    assert o1.shape == (3, 9, 25) and (o1[:, :, 0] == 101).all() and ((o1 == 102).sum(-1) == 1).all()
    ids = b1["enc_input_ids"]
    n = (ids != 0).sum(-1)
    assert ((ids != 0).long().cumsum(-1)[torch.arange(3), n - 1] == n).all()          # no holes: tokens are a prefix
    assert (b1["enc_att_mask"] == (ids != 0).float()).all() and (n > 40).all()         # caption + two rounds + a question


def test_sampling_seeds_differ_by_role_round_and_run_seed():
    """questioner and teacher, every round, and every run seed get their own sampling stream."""
    from gst_visdial_b200.models.visual_dialog_model import derive_seed
    seeds = {derive_seed(base, role, idx) for base in (0, 1, 2**40 + 5) for role in ("q", "a", "enc_dec_q", "enc_dec_a") for idx in range(12)}
    assert len(seeds) == 3 * 4 * 12
    assert all(0 <= s < 2**64 for s in seeds)
    assert derive_seed(0, "a", 3) == derive_seed(0, "a", 3)


def test_jsonl_writer_raises_instead_of_deadlocking(tmp_path):
    """a writer thread that died (here: records that cannot be serialised) must surface as an exception on the next
    write / close even when the bounded queue is full - not as a blocked put."""
    import threading
    from gst_visdial_b200.io.output import JsonlWriter
    w = JsonlWriter(str(tmp_path / "x.jsonl"), queue_depth=1)
    w.write([{"ok": 1}])
    w.write([{"bad": object()}])                      # json.dumps fails inside the thread
    done = []

    def hammer():
        try:
            for _ in range(50):
                w.write([{"ok": 2}])
        except Exception as e:                         # noqa: BLE001
            done.append(e)

    t = threading.Thread(target=hammer, daemon=True)
    t.start()
    t.join(timeout=20)
    assert not t.is_alive(), "write() blocked on a dead writer thread"
    assert done, "the writer's failure never reached the producer"
    try:
        w.close()
    except Exception:                                  # noqa: BLE001 - close re-raises the writer's error
        pass


def test_synthetic_history_and_candidate_batches():
    """Input builders of bench.py's config 4 / 5 workloads: shapes, the reference's segment / mask conventions and the host-side
    length bound the encoder trimming relies on (no token at or past it)."""
    from gst_visdial_b200 import synthetic as S
    b = S.synthetic_history_batch(5, 6, rounds=[0, 1, 3, 9, 30, 2])
    ids, seg = b["enc_input_ids"], b["enc_segments"]
    assert ids.shape == (6, 256) and b["enc_image_feat"].shape == (6, 37, 2048)
    bound = int(b["hist_len_bound"][0])
    assert bound <= 256 and not (ids[:, bound:] != 0).any() and (ids[:, bound - 1] != 0).any()
    assert torch.equal(b["enc_att_mask"], (ids != 0).float())
    assert int((ids[0] != 0).sum()) < int((ids[3] != 0).sum())           # more rounds, longer history
    assert not (ids[:, 1:] == 101).any() and (ids[:, 0] == 101).all()
    n0 = int((ids[1] != 0).sum())
    cap = int(seg[1].sum().item())                                          # caption (segment 1) + one answer (segment 1), question segment 0
    assert 0 < cap < n0
    c = S.synthetic_candidate_batch(3, 4, 7)
    tok, sg, mk = c["tokens"], c["segments"], c["mask"]
    assert tok.shape == (4, 7, 256) and c["image_feat"].shape == (4, 37, 2048)
    assert torch.equal(mk, (tok != 0).float())
    cb = int(c["hist_len_bound"][0])
    assert not (tok[..., cb:] != 0).any()
    for i in range(4):
        assert torch.equal(tok[i, 0, :20], tok[i, 1, :20])                  # the candidates of an item share its history ...
        assert not torch.equal(tok[i, 0], tok[i, 1])                        # ... and differ in the appended answer
        last = int((tok[i, 0] != 0).sum()) - 1
        assert int(tok[i, 0, last]) == 102                                  # every utterance ends in [SEP]; the mask reaches the last one


def test_history_trim_helper():
    from gst_visdial_b200.ranking import _trim
    ids = torch.zeros(3, 256, dtype=torch.int64); ids[:, :70] = 5
    seg, att = torch.zeros_like(ids), (ids != 0).float()
    a, b, c = _trim(ids, seg, att, 70)
    assert a.shape == (3, 96) and b.shape == (3, 96) and c.shape == (3, 96) and a.is_contiguous()
    assert _trim(ids, seg, att, None)[0] is ids and _trim(ids, seg, att, 250)[0] is ids and _trim(ids, None, None, 10)[0].shape == (3, 32)
