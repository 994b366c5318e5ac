"""Packed region-feature shards (vlbert_b200.region_shards).

  * format round trip: what a shard returns is byte for byte what base64 + np.frombuffer give on the reference-format JSON record
    (the decode at pretrain/data/datasets/conceptual_captions.py:99-118), including records without features and with no boxes;
  * `collate_boxes` equals the reference BatchCollator's boxes column (clip_pad_boxes with pad -2, then stack);
  * where /root/reference exists: the UNMODIFIED `ConceptualCaptionsDataset.__getitem__` returns identical items when it reads
    through an attached shard instead of JSON files (same `random` seed), with and without precomputed features.
"""
import base64
import json
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))


@pytest.fixture(scope="module")
def RS():
    import vlbert_b200
    return vlbert_b200.region_shards


def synth_record(rng, n, feat_dim=16, n_cls=11, with_feat=True):
    """one frcnn record in the reference's JSON schema (base64 text of float32 arrays)"""
    boxes = np.concatenate([rng.uniform(0, 200, (n, 2)), rng.uniform(220, 400, (n, 2))], 1).astype(np.float32)
    classes = rng.dirichlet(np.ones(n_cls), n).astype(np.float32) if n else np.zeros((0, n_cls), np.float32)
    rec = {"num_boxes": n, "image_w": 640, "image_h": 480,
           "boxes": base64.encodebytes(boxes.tobytes()).decode(), "classes": base64.encodebytes(classes.tobytes()).decode()}
    if with_feat:
        rec["features"] = base64.encodebytes(rng.standard_normal((n, feat_dim)).astype(np.float32).tobytes()).decode()
    return rec


def test_shard_round_trip_is_byte_exact(RS, tmp_path):
    rng = np.random.default_rng(3)
    recs = [synth_record(rng, n, with_feat=(i % 3 != 2)) for i, n in enumerate([5, 1, 9, 0, 36])]
    keys = ["frcnn/%05d.json" % i for i in range(len(recs))]
    path = RS.write_shard(str(tmp_path / "regions.vlbrs"), recs, keys)
    sh = RS.RegionShard(path)
    assert len(sh) == len(recs) and [sh.key(i) for i in range(len(sh))] == keys
    for i, r in enumerate(recs):
        got = sh.record_by_key(keys[i])
        assert got["num_boxes"] == r["num_boxes"] and got["image_w"] == 640 and got["image_h"] == 480
        for name in ("boxes", "classes", "features"):
            if name not in r:
                assert name not in got
                continue
            want = base64.decodebytes(r[name].encode())
            assert bytes(got[name]) == want, name
            # the reference's own decode expression on both sides
            a = np.frombuffer(got[name], dtype=np.float32).reshape((r["num_boxes"], -1)) if r["num_boxes"] else None
            b = np.frombuffer(want, dtype=np.float32).reshape((r["num_boxes"], -1)) if r["num_boxes"] else None
            assert (a is None and b is None) or np.array_equal(a, b)
        boxes, classes, feats = sh.arrays(i)
        assert boxes.shape == (r["num_boxes"], 4 if r["num_boxes"] else 0)
    sh.close()
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.bin"
        bad.write_bytes(b"0" * 64)
        RS.RegionShard(str(bad))


def test_vqa_style_records(RS, tmp_path):
    """records without class scores and with the optional image_box_feature row (vqa/data/datasets/vqa.py:188-215)"""
    rng = np.random.default_rng(5)
    recs = []
    for n, extra in ((4, True), (7, False)):
        r = synth_record(rng, n)
        del r["classes"], r["image_w"], r["image_h"]
        if extra:
            r["image_box_feature"] = base64.encodebytes(rng.standard_normal((1, 16)).astype(np.float32).tobytes()).decode()
        recs.append(r)
    sh = RS.RegionShard(RS.write_shard(str(tmp_path / "vqa.vlbrs"), recs, ["a.json", "b.json"]))
    for k, r in zip(("a.json", "b.json"), recs):
        got = sh.record_by_key(k)
        assert set(got) == set(r) | {"image_w", "image_h"}
        for name in ("boxes", "features", "image_box_feature"):
            if name in r:
                assert bytes(got[name]) == base64.decodebytes(r[name].encode()), name
        # the VQA decode expressions on the shard record
        feats = np.frombuffer(got["features"], dtype=np.float32).reshape((got["num_boxes"], -1))
        assert feats.shape == (r["num_boxes"], 16)
        if "image_box_feature" in r:
            assert np.frombuffer(got["image_box_feature"], dtype=np.float32).reshape((1, -1)).shape == (1, 16)


def test_collate_boxes_equals_the_reference_collator(RS):
    g = torch.Generator().manual_seed(0)
    items = [torch.randn(n, 4 + 8, generator=g) for n in (3, 7, 1, 7)]
    K = max(t.shape[0] for t in items)

    def clip_pad_boxes(t, K, pad):                       # common/utils/clip_pad.py:23-38
        out = torch.zeros((K, t.shape[1]), dtype=t.dtype) + pad
        out[:min(t.shape[0], K)] = t[:min(t.shape[0], K)]
        return out

    ref = torch.stack([clip_pad_boxes(t, K, -2) for t in items], 0)
    assert torch.equal(RS.collate_boxes(items), ref)
    buf = torch.empty(8, 16, 12)
    assert torch.equal(RS.collate_boxes(items, out=buf), ref)


class _Tok(object):
    """the three tokenizer entry points __getitem__ uses when the MLM task is off"""
    def tokenize(self, text):
        return text.split()

    def convert_tokens_to_ids(self, toks):
        return [sum(t.encode()) % 30000 + 1 for t in toks]


@pytest.mark.parametrize("with_feat", [True, False])
def test_reference_dataset_reads_identical_items_through_a_shard(RS, tmp_path, with_feat):
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_shim.install()
    import importlib
    for name in ("pycocotools", "pycocotools.coco"):           # imported by a sibling dataset module; not used here
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = getattr(sys.modules["pycocotools.coco"], "COCO", object)
    cc = importlib.import_module("pretrain.data.datasets.conceptual_captions")
    rng = np.random.default_rng(11)
    n_items = 6
    os.makedirs(tmp_path / "frcnn", exist_ok=True)
    recs, db = [], []
    for i in range(n_items):
        r = synth_record(rng, int(rng.integers(2, 12)), with_feat=with_feat)
        recs.append(r)
        with open(tmp_path / "frcnn" / ("%05d.json" % i), "w") as f:
            json.dump(r, f)
        db.append({"caption": ["a", "photo", "of", "thing", str(i), "on", "table"], "image": "img/%05d.jpg" % i, "frcnn": "frcnn/%05d.json" % i})
    RS.write_shard(str(tmp_path / "regions.vlbrs"), recs, ["frcnn/%05d.json" % i for i in range(n_items)])

    def dataset():
        ds = object.__new__(cc.ConceptualCaptionsDataset)     # the constructor needs the annotation files and a BERT vocabulary
        ds.database, ds.data_path, ds.transform, ds.seq_len = db, str(tmp_path), None, 14
        ds.with_precomputed_visual_feat, ds.add_image_as_a_box = with_feat, True
        ds.with_rel_task, ds.with_mlm_task, ds.with_mvrc_task, ds.mask_raw_pixels = True, False, True, False
        ds.tokenizer = _Tok()
        ds.zipreader = None
        return ds

    ds_json = dataset()
    if not with_feat:      # the image branch: no image files here -> the reference falls back to its zero image
        ds_json._load_image = types.MethodType(lambda self, path: (_ for _ in ()).throw(IOError("no image")), ds_json)
    ds_shard = RS.attach(dataset(), RS.RegionShard(str(tmp_path / "regions.vlbrs")),
                         key_of=lambda p: os.path.relpath(p, str(tmp_path)))
    if not with_feat:
        ds_shard._load_image = ds_json._load_image
    for i in range(n_items):
        random.seed(100 + i)
        a = ds_json[i]
        random.seed(100 + i)
        b = ds_shard[i]
        assert len(a) == len(b) == 8
        for x, y in zip(a, b):
            if torch.is_tensor(x):
                assert x.dtype == y.dtype and torch.equal(x, y)
            elif isinstance(x, np.ndarray):
                assert np.array_equal(x, y)
            else:
                assert x == y


@pytest.mark.parametrize("image_box_feature,image_as_box", [(True, True), (False, True), (False, False)])
def test_reference_vqa_dataset_reads_identical_items_through_a_shard(RS, tmp_path, image_box_feature, image_as_box):
    """the same seam in vqa/data/datasets/vqa.py:184-262 (records without class scores, optional image_box_feature row)"""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_shim.install()
    import importlib
    vqa = importlib.import_module("vqa.data.datasets.vqa")
    rng = np.random.default_rng(21)
    recs, db = [], []
    for i in range(5):
        r = synth_record(rng, int(rng.integers(3, 10)))
        del r["classes"]
        if image_box_feature:
            r["image_box_feature"] = base64.encodebytes(rng.standard_normal((1, 16)).astype(np.float32).tobytes()).decode()
        recs.append(r)
        fn = str(tmp_path / ("box_%03d.json" % i))
        with open(fn, "w") as f:
            json.dump(r, f)
        db.append({"box_fn": fn, "image_fn": "none.jpg", "width": 640, "height": 480, "question": "what is left of thing %d" % i})
    RS.write_shard(str(tmp_path / "vqa.vlbrs"), recs, [d["box_fn"] for d in db])

    def dataset():
        ds = object.__new__(vqa.VQA)
        ds.database, ds.transform, ds.test_mode, ds.use_imdb = db, None, True, False
        ds.with_precomputed_visual_feat, ds.add_image_as_a_box = True, image_as_box   # False: the in-place clamp hits the decoded buffer
        ds.tokenizer = _Tok()
        return ds

    ds_json = dataset()
    ds_shard = RS.attach(dataset(), RS.RegionShard(str(tmp_path / "vqa.vlbrs")))
    for i in range(len(db)):
        a, b = ds_json[i], ds_shard[i]
        assert a[0] is None and b[0] is None
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3]
    # a second pass over the shard still sees the file's bytes where the first pass clamped in place on private pages only
    fresh = RS.RegionShard(str(tmp_path / "vqa.vlbrs"))
    for k, r in zip([d["box_fn"] for d in db], recs):
        assert bytes(fresh.record_by_key(k)["boxes"]) == base64.decodebytes(r["boxes"].encode())
