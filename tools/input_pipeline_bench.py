"""Host-side cost of getting one sample's precomputed region record in front of the hot path, on one core:

  json    the reference's decode: json.load of the per-image file + base64.decodebytes x 3 + np.frombuffer
          (pretrain/data/datasets/conceptual_captions.py:99-118)
  shard   vlbert_b200.region_shards: index lookup in the memory-mapped shard + np.frombuffer on zero-copy views

Synthetic records in the reference's schema with its real sizes (36 boxes, 1601 class scores, 2048-d features), files on local
disk, page cache warm (both arms).  CPU only.   python tools/input_pipeline_bench.py [--records 64] [--reps 5]
"""
import argparse
import base64
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import importlib.util
    spec = importlib.util.spec_from_file_location("region_shards", os.path.join(os.path.dirname(__file__), "..", "vl-bert_b200", "region_shards.py"))
    RS = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RS)
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="vlb_regions_")
    recs, keys = [], []
    for i in range(a.records):
        n = 36
        rec = {"num_boxes": n, "image_w": 1000, "image_h": 600}
        for name, d in (("boxes", 4), ("classes", 1601), ("features", 2048)):
            rec[name] = base64.encodebytes(rng.standard_normal((n, d)).astype(np.float32).tobytes()).decode()
        recs.append(rec)
        keys.append("%06d.json" % i)
        with open(os.path.join(tmp, keys[-1]), "w") as f:
            json.dump(rec, f)
    shard_path = RS.write_shard(os.path.join(tmp, "regions.vlbrs"), recs, keys)
    shard = RS.RegionShard(shard_path)

    def decode(frcnn_data, b64):          # the three expressions of conceptual_captions.py:103-118
        n = frcnn_data["num_boxes"]
        return [np.frombuffer(b64(frcnn_data[k]), dtype=np.float32).reshape((n, -1)) for k in ("boxes", "classes", "features")]

    def arm_json():
        s = 0.0
        for k in keys:
            with open(os.path.join(tmp, k), "r") as f:
                d = json.load(f)
            s += float(decode(d, lambda t: base64.decodebytes(t.encode()))[2][0, 0])
        return s

    def arm_shard():
        s = 0.0
        for k in keys:
            s += float(decode(shard.record_by_key(k), lambda raw: raw)[2][0, 0])
        return s

    assert arm_json() == arm_shard()
    res = {}
    for name, fn in (("json", arm_json), ("shard", arm_shard)):
        fn()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        dt = (time.perf_counter() - t0) / (a.reps * a.records)
        res[name] = dt
        print("%-6s %9.1f us per record  %9.0f records/s per core" % (name, dt * 1e6, 1.0 / dt))
    sz_json = sum(os.path.getsize(os.path.join(tmp, k)) for k in keys) / a.records
    print("bytes per record: json %.0f, shard %.0f  |  speed-up %.0fx  |  cores needed for 10.6 k samples/s (one B200 at config 2): json %.1f, shard %.2f"
          % (sz_json, os.path.getsize(shard_path) / a.records, res["json"] / res["shard"], 10600 * res["json"], 10600 * res["shard"]))


if __name__ == "__main__":
    main()
