"""Packed region-feature shards: the on-disk format in front of the hot path.

The reference stores the precomputed Faster R-CNN output of every image as one JSON file whose `boxes`, `classes` and `features`
entries are base64 text of float32 arrays (`pretrain/data/datasets/conceptual_captions.py:99-118`, `vqa/data/datasets/vqa.py`):
per sample `json.load` of ~0.6 MB of text, three `base64.decodebytes`, and only then `np.frombuffer`.  A shard keeps the same
records as raw little-endian float32 behind a fixed-size index, memory-mapped (private, copy-on-write): a record costs two index reads and
three zero-copy `memoryview`s, and everything downstream of `np.frombuffer` in the reference's `__getitem__` (confidence sort,
whole-image box, masking tasks, truncation) runs unchanged on identical bytes -- parity is exact by construction and checked
item by item in `tests/test_region_shards.py`.

    write_shard(path, records)                 records: dicts in the reference's JSON schema (base64 strings or float32 arrays)
    shard = RegionShard(path)                  mmap; len(shard), shard.key(i), shard.record(i) / shard.record_by_key(key)
    attach(dataset, shard)                     the reference dataset then reads shard records through its own __getitem__
    collate_boxes(list of [n_i, D] tensors)    == torch.stack([clip_pad_boxes(b, max n, pad=-2)]) (pretrain/data/collate_batch.py:22,38-40),
                                               one preallocated (optionally pinned) [B, max n, D] buffer

File layout (little endian): magic "VLBRS001" | uint64 count | count x index entry | key bytes | payload.
Index entry = 8 x uint64: key offset, key length | (image-box feature dim << 32), payload offset, num_boxes, box dim, class dim,
feature dim, (image_w << 32 | image_h); bit 63 of the class / feature dim says that the record has that entry at all (the VQA
records carry no class scores, `vqa/data/datasets/vqa.py:188-215`, and may carry one extra `image_box_feature` row).
Payload of a record = boxes [n, box dim] | classes [n, class dim] | features [n, feature dim] | image_box_feature [1, dim], float32,
64-byte aligned.
"""
import base64
import mmap
import os
import struct

import numpy as np
import torch

MAGIC = b"VLBRS001"
_ENTRY = struct.Struct("<8Q")
_ALIGN = 64


def _as_f32(x, n):
    """base64 text (the reference's encoding) or an array -> contiguous float32 [n, -1]"""
    if isinstance(x, str):
        x = np.frombuffer(base64.decodebytes(x.encode()), dtype=np.float32)
    elif isinstance(x, (bytes, bytearray, memoryview)):
        x = np.frombuffer(x, dtype=np.float32)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    return x.reshape(n, -1) if n > 0 else x.reshape(0, 0)


def write_shard(path, records, keys=None):
    """records: iterable of dicts with the reference's keys `boxes`, `classes`, `features` (optional), `num_boxes`, `image_w`,
    `image_h`; keys: the strings the dataset will ask for (the `frcnn` path of the annotation line); default "0", "1", ..."""
    records = list(records)
    keys = [str(i) for i in range(len(records))] if keys is None else [str(k) for k in keys]
    assert len(keys) == len(records) and len(set(keys)) == len(keys), "one unique key per record"
    count = len(records)
    key_blob = b"".join(k.encode() for k in keys)
    pos = len(MAGIC) + 8 + count * _ENTRY.size + len(key_blob)
    pos = (pos + _ALIGN - 1) // _ALIGN * _ALIGN
    entries, payloads = [], []
    koff = len(MAGIC) + 8 + count * _ENTRY.size
    for k, r in zip(keys, records):
        n = int(r["num_boxes"])
        boxes = _as_f32(r["boxes"], n)
        has_cls, has_feat = r.get("classes") is not None, r.get("features") is not None
        classes = _as_f32(r["classes"], n) if has_cls else np.zeros((n, 0), np.float32)
        feats = _as_f32(r["features"], n) if has_feat else np.zeros((n, 0), np.float32)
        ibf = _as_f32(r["image_box_feature"], 1) if r.get("image_box_feature") is not None else np.zeros((1, 0), np.float32)
        blob = boxes.tobytes() + classes.tobytes() + feats.tobytes() + ibf.tobytes()
        pad = (-len(blob)) % _ALIGN
        entries.append((koff, len(k.encode()) | (ibf.shape[1] << 32), pos, n, boxes.shape[1] if n else 0,
                        (classes.shape[1] if n else 0) | ((1 << 63) if has_cls else 0),
                        (feats.shape[1] if n else 0) | ((1 << 63) if has_feat else 0),
                        (int(r.get("image_w", 0)) << 32) | int(r.get("image_h", 0))))
        payloads.append(blob + b"\0" * pad)
        koff += len(k.encode())
        pos += len(blob) + pad
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", count))
        for e in entries:
            f.write(_ENTRY.pack(*e))
        f.write(key_blob)
        f.write(b"\0" * ((-f.tell()) % _ALIGN))
        for p in payloads:
            f.write(p)
    os.replace(tmp, path)
    return path


class _Record(dict):
    """What `json.load` of a reference frcnn file returns, with the three arrays as zero-copy views of the mapped shard."""


class RegionShard(object):
    def __init__(self, path):
        self.path = path
        self._f = open(path, "rb")
        # private copy-on-write mapping: the reference datasets clamp `boxes` IN PLACE on the tensor that shares memory with the
        # decoded buffer when no image box is prepended (vqa/data/datasets/vqa.py:199-226) -- harmless on their private bytes
        # object, a fault on a read-only mapping; with ACCESS_COPY such a write touches a private page and never the file
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_COPY)
        if self._mm[:8] != MAGIC:
            raise ValueError("%s is not a vlbert_b200 region shard" % path)
        (self.count,) = struct.unpack_from("<Q", self._mm, 8)
        self._index = np.frombuffer(self._mm, dtype="<u8", count=self.count * 8, offset=16).reshape(self.count, 8)
        self._view = memoryview(self._mm)
        self._keys = None

    def __len__(self):
        return int(self.count)

    def key(self, i):
        e = self._index[i]
        return bytes(self._view[int(e[0]):int(e[0]) + (int(e[1]) & 0xFFFFFFFF)]).decode()

    def _key_table(self):
        if self._keys is None:
            self._keys = {self.key(i): i for i in range(self.count)}
        return self._keys

    def record(self, i):
        e = [int(v) for v in self._index[i]]
        lo = (1 << 63) - 1
        n, bd, cd, fd, has_cls, has_feat, ibd = e[3], e[4], e[5] & lo, e[6] & lo, bool(e[5] >> 63), bool(e[6] >> 63), e[1] >> 32
        o = e[2]
        r = _Record(num_boxes=n, image_w=e[7] >> 32, image_h=e[7] & 0xFFFFFFFF)
        r["boxes"] = self._view[o:o + 4 * n * bd]
        o += 4 * n * bd
        if has_cls:
            r["classes"] = self._view[o:o + 4 * n * cd]
        o += 4 * n * cd
        if has_feat:
            r["features"] = self._view[o:o + 4 * n * fd]
        o += 4 * n * fd
        if ibd:
            r["image_box_feature"] = self._view[o:o + 4 * ibd]
        return r

    def record_by_key(self, key):
        return self.record(self._key_table()[str(key)])

    def arrays(self, i):
        """(boxes [n, 4+], class scores [n, C], features [n, F] or None) as float32 arrays viewing the shard"""
        r = self.record(i)
        n = r["num_boxes"]
        def arr(raw):
            a = np.frombuffer(raw, np.float32)
            return a.reshape(n, -1) if n else a.reshape(0, 0)
        return arr(r["boxes"]), (arr(r["classes"]) if "classes" in r else None), (arr(r["features"]) if "features" in r else None)

    def close(self):
        """Unmap if no record views are alive any more; otherwise the mapping is released with the last of them."""
        self._index = None
        try:
            self._view.release()
            self._mm.close()
        except BufferError:
            pass
        self._f.close()


def attach(dataset, shard, key_of=None):
    """Make a reference dataset object (`ConceptualCaptionsDataset`, `VQA`, ...: anything that loads region records through
    `self._load_json(path)` and decodes them with `self.b64_decode`) read from `shard` instead.  The two methods are the only
    seam: the rest of its `__getitem__` runs as written, on the same float32 bytes.  key_of: maps the path the dataset asks for
    to the shard key (default: the path itself, then its basename)."""
    table = shard._key_table()

    def load(path):
        k = key_of(path) if key_of is not None else (path if path in table else os.path.basename(path))
        return shard.record_by_key(k)

    dataset._load_json = load
    dataset.b64_decode = lambda raw: raw        # np.frombuffer(memoryview) in the reference's own code: zero copy
    return dataset


def collate_boxes(boxes_list, pad=-2.0, out=None, pin_memory=False):
    """The `boxes` column of the reference's BatchCollator: every [n_i, D] tensor padded with `pad` to the longest and stacked
    (pretrain/data/collate_batch.py:22,38-40 with common/utils/clip_pad.py: clip_pad_boxes), written into one buffer."""
    B = len(boxes_list)
    n_max = max(int(b.shape[0]) for b in boxes_list)
    D = int(boxes_list[0].shape[1])
    if out is None:
        out = torch.empty((B, n_max, D), dtype=boxes_list[0].dtype, pin_memory=pin_memory)
    else:
        out = out[:B, :n_max]
    out.fill_(pad)
    for i, b in enumerate(boxes_list):
        out[i, :b.shape[0]] = b
    return out
