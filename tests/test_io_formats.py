"""CPU tests of the input / output formats around the hot path."""
import base64
import json
import os

import numpy as np
import torch

from gst_visdial_b200.io import features as F
from gst_visdial_b200.io import output as O
from oracle import io_reader as RO


def _record(rng, n, w=640, h=480):
    feats = np.maximum(rng.standard_normal((n, 2048)), 0).astype(np.float32)
    x1 = rng.uniform(0, 0.7 * w, n); y1 = rng.uniform(0, 0.7 * h, n)
    boxes = np.stack([x1, y1, x1 + rng.uniform(0.1 * w, 0.3 * w, n), y1 + rng.uniform(0.1 * h, 0.3 * h, n)], 1).astype(np.float32)
    return {"image_id": 7, "image_h": h, "image_w": w, "num_boxes": n, "features": base64.b64encode(feats.tobytes()),
            "boxes": base64.b64encode(boxes.tobytes()), "cls_prob": base64.b64encode(np.zeros((n, 1601), np.float32).tobytes())}


def test_record_decoding_matches_reference_reader():
    rng = np.random.default_rng(0)
    for n in (36, 10, 50):                      # 50 > max_regions: truncated like data_utils.py:75
        item = _record(rng, n)
        f, nb, loc = F.decode_reference_record(item)
        rf, rnb, rloc = RO.read_record(item)
        assert nb == rnb and np.array_equal(f, rf) and np.array_equal(loc, rloc) and loc.dtype == rloc.dtype
        pf, ps, pm = F.pad_regions(f, nb, loc)
        qf, qs, qm = RO.encode_image_input(rf, rnb, rloc)
        assert np.array_equal(pf, qf) and np.array_equal(ps, qs) and np.array_equal(pm, qm)
        assert pf.shape == (37, 2048) and pm.sum() == min(n + 1, 37)


def test_shards_roundtrip_lookup_and_prefetch(tmp_path):
    rng = np.random.default_rng(1)
    ids = [1000 + 3 * i for i in range(23)]
    feats = rng.standard_normal((23, 37, 2048)).astype(np.float32)
    locs = rng.uniform(0, 1, (23, 37, 5)).astype(np.float32)
    masks = (rng.uniform(0, 1, (23, 37)) > 0.2).astype(np.float32)
    F.write_shard(str(tmp_path / "a"), ids[:10], feats[:10], locs[:10], masks[:10], dtype="bf16")
    F.write_shard(str(tmp_path / "b"), ids[10:], feats[10:], locs[10:], masks[10:], dtype="bf16")
    sh = F.FeatureShards([str(tmp_path / "a"), str(tmp_path / "b")])
    assert len(sh) == 23 and 1003 in sh and 5 not in sh
    pick = [ids[20], ids[0], ids[11], ids[9]]
    b = sh.batch(pick, pin=False)
    want = torch.from_numpy(feats[[20, 0, 11, 9]]).to(torch.bfloat16)
    assert b["enc_image_feat"].dtype == torch.bfloat16 and torch.equal(b["enc_image_feat"], want)
    assert torch.equal(b["enc_image_loc"], torch.from_numpy(locs[[20, 0, 11, 9]]))
    assert torch.equal(b["enc_image_mask"], torch.from_numpy(masks[[20, 0, 11, 9]]))
    F.write_shard(str(tmp_path / "c"), ids[:4], feats[:4], locs[:4], masks[:4], dtype="fp32")
    assert torch.equal(F.FeatureShards([str(tmp_path / "c")]).batch(ids[:4], pin=False)["enc_image_feat"], torch.from_numpy(feats[:4]))
    batches = [ids[i:i + 5] for i in range(0, 23, 5)]
    got = [x["image_id"].tolist() for x in F.Prefetcher(sh, batches, depth=2, pin=False)]
    assert got == batches
    try:
        list(F.Prefetcher(sh, [[1]], pin=False))
        raise AssertionError("missing image id must raise")
    except KeyError:
        pass


def test_jsonl_writer_and_reference_json(tmp_path):
    B, R = 3, 2
    q = torch.zeros(B, R, 18, dtype=torch.int64); a = torch.zeros(B, R, 18, dtype=torch.int64)
    q[:, :, :3] = torch.tensor([2001, 2002, 102]); a[:, :, :2] = torch.tensor([3001, 3002])
    ppl = torch.arange(B * R, dtype=torch.float32).reshape(B, R) + 1.5
    abn = torch.tensor([0, 1, 0], dtype=torch.int32)
    recs = O.batch_records(torch.tensor([11, 12, 13]), q, a, ppl, abn, meta={11: {"url": "u11", "caption": "c11"}})
    assert [r["image_id"] for r in recs] == [11, 13]                      # the overflowed dialog is dropped (generate.py:236-237)
    assert recs[0]["url"] == "u11" and recs[0]["dialog"][1] == {"question": [2001, 2002], "answer": [3001, 3002], "answer_ppl": 2.5}
    txt = O.batch_records([11, 12, 13], q, a, None, abn, decode=lambda ids: " ".join(map(str, ids)))
    assert txt[1]["dialog"][0] == {"question": "2001 2002", "answer": "3001 3002"}
    p = str(tmp_path / "out.jsonl")
    with O.JsonlWriter(p) as w:
        w.write(recs)
        w.write(txt)
    assert w.count == 4 and sum(1 for _ in open(p)) == 4
    n = O.jsonl_to_reference_json(p, str(tmp_path / "out.json"))
    data = json.load(open(tmp_path / "out.json"))
    assert n == 4 and data[0] == recs[0] and data[3] == txt[1]


def test_reader_matches_reference_reader_goldens(golden_dir):
    """the product reader (gst_visdial_b200/io/features.py) and its oracle (oracle/io_reader.py) against
    tests/golden/io_records.npz - outputs of the REFERENCE'S OWN ImageFeaturesH5Reader + encode_image_input
    (utils/image_features_reader.py:58-141, utils/data_utils.py:73-117) run on the same seeded records by oracle/gen_golden_io.py."""
    import hashlib
    from oracle.gen_golden_io import CASES, make_reference_record
    g = dict(np.load(os.path.join(golden_dir, "io_records.npz")))
    for i in range(len(CASES)):
        rec = make_reference_record(i)
        for reader, padder in ((F.decode_reference_record, F.pad_regions), (RO.read_record, RO.encode_image_input)):
            f, nb, loc = reader(rec)
            assert nb == int(g[f"c{i}_num_boxes"]) == CASES[i][0] + 1
            assert np.array_equal(f[0], g[f"c{i}_global_row"]) and np.array_equal(np.asarray(loc), g[f"c{i}_loc"])
            pf, ps, pm = padder(f, nb, loc)
            assert np.array_equal(ps, g[f"c{i}_pad_loc"]) and np.array_equal(pm, g[f"c{i}_pad_mask"])
            sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pf).tobytes()).digest(), dtype=np.uint8)
            assert np.array_equal(sha, g[f"c{i}_pad_feat_sha256"]), f"case {i}: padded feature block differs from the reference reader's"
