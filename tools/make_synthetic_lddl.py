#!/usr/bin/env python
"""Write a synthetic, sequence-binned pretraining dataset in the parquet layout shims/thirdparty/lddl/torch reads (the on-disk edge
of the hot path): <out>/meta.json + <out>/bin_<k>/shard_<i>.parquet.

    python tools/make_synthetic_lddl.py --out /tmp/lddl_synth --samples 8192 --seq-len 512 --bin-size 64 --max-pred 80 --shards 8
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shims", "thirdparty"))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--seq-len", type=int, default=512)
    ap.add_argument("--bin-size", type=int, default=64)
    ap.add_argument("--max-pred", type=int, default=80)
    ap.add_argument("--vocab", type=int, default=30522)
    ap.add_argument("--shards", type=int, default=8, help="shards per bin (ranks read shards i == rank mod world)")
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args(argv)
    import pyarrow as pa
    import pyarrow.parquet as pq
    from lddl.torch import _synth_samples
    samples = _synth_samples(a.samples, a.seq_len, a.max_pred, a.vocab, a.seed)
    bins = {}
    for s in samples:
        L = len(s["a_ids"]) + len(s["b_ids"])
        bins.setdefault(0 if a.bin_size <= 0 else (L - 1) // a.bin_size, []).append(s)
    os.makedirs(a.out, exist_ok=True)
    json.dump(dict(seq_len=a.seq_len, bin_size=a.bin_size, max_pred=a.max_pred, vocab=a.vocab, samples=a.samples,
                   bin_counts={str(k): len(v) for k, v in bins.items()}), open(os.path.join(a.out, "meta.json"), "w"))
    schema = pa.schema([("a_ids", pa.list_(pa.int32())), ("b_ids", pa.list_(pa.int32())), ("masked_lm_positions", pa.list_(pa.int32())),
                        ("masked_lm_ids", pa.list_(pa.int32())), ("is_random_next", pa.bool_())])
    for k, rows in sorted(bins.items()):
        d = os.path.join(a.out, f"bin_{k}")
        os.makedirs(d, exist_ok=True)
        for i in range(a.shards):
            part = rows[i::a.shards]
            if part:
                pq.write_table(pa.Table.from_pylist(part, schema=schema), os.path.join(d, f"shard_{i}.parquet"))
    print(json.dumps({"out": a.out, "bins": {k: len(v) for k, v in sorted(bins.items())}}))


if __name__ == "__main__":
    main()
