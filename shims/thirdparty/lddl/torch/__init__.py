"""lddl.torch stand-in: `get_bert_pretrain_data_loader` with the signature and batch format the reference driver relies on
(PyTorch/LanguageModeling/BERT/run_pretraining.py:557-570 call site; batch consumed at :520-521,603-609,663-665):

    an iterable with len() whose items are dicts of five int64 CPU tensors (pinned when pin_memory is set)
        input_ids [B,S]  token_type_ids [B,S]  attention_mask [B,S]  labels [B,S] (-1 = not masked)  next_sentence_labels [B]

Two sources, both sharded by rank and SEQUENCE-BINNED the way LDDL's phase-2 data is (scripts/run_pretraining.sh:41,54: bin size 64):
every batch comes from ONE length bin and is padded to that bin's upper edge, and all ranks draw the same bin at the same step (the bin
sequence is a function of (base_seed, epoch) only), so DDP steps see identical shapes on every rank.

  * a directory written by tools/make_synthetic_lddl.py: parquet shards `bin_<k>/shard_<i>.parquet` of pre-tokenised samples
    (columns a_ids, b_ids: list<int32> incl. [CLS]/[SEP]; masked_lm_positions: list<int32>; masked_lm_ids: list<int32>;
    is_random_next: bool) plus `meta.json` {seq_len, bin_size, max_pred, vocab}.  (The real LDDL stores tokens as text and needs the
    vocab file + network-installed package; tokenisation is outside the hot path)
  * the string "synthetic[?key=value&...]" (keys: seq_len, max_pred, samples, bin_size, vocab; defaults 512/80/4096/0/30522): the
    same samples generated in memory.  bin_size=0 pads everything to seq_len (the benchmark's worst case).
"""
import json
import logging
import os

import torch


def _rank_world(local_rank):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", max(local_rank, 0))), int(os.environ.get("WORLD_SIZE", 1))


def _synth_samples(n, seq_len, max_pred, vocab, seed, len_seed=None):
    """n samples with lengths ~ U{seq_len/4 .. seq_len}: dicts of python lists in the parquet schema.  The LENGTHS come from `len_seed`
    (rank-independent: every rank then holds equally populated bins, as LDDL's balanced shards do), the tokens from `seed`."""
    gl = torch.Generator().manual_seed(seed if len_seed is None else len_seed)
    lens = torch.randint(max(seq_len // 4, 8), seq_len + 1, (n,), generator=gl).tolist()
    g = torch.Generator().manual_seed(seed)
    out = []
    for L in lens:
        la = max(3, L // 2)
        ids = torch.randint(1000, min(vocab, 30522), (L,), generator=g, dtype=torch.int64)
        ids[0], ids[la - 1], ids[L - 1] = 101, 102, 102
        n_pred = max(1, min(max_pred, int(round(0.15 * L))))
        cand = torch.tensor([i for i in range(1, L - 1) if i != la - 1])
        pos = cand[torch.randperm(len(cand), generator=g)[:n_pred]].sort().values
        out.append(dict(a_ids=ids[:la].tolist(), b_ids=ids[la:].tolist(), masked_lm_positions=pos.tolist(),
                        masked_lm_ids=torch.randint(1000, min(vocab, 30522), (len(pos),), generator=g).tolist(),
                        is_random_next=bool(torch.randint(0, 2, (1,), generator=g).item())))
    return out


def _collate(samples, pad_to, pin):
    B = len(samples)
    ids = torch.zeros(B, pad_to, dtype=torch.int64)
    tt = torch.zeros(B, pad_to, dtype=torch.int64)
    am = torch.zeros(B, pad_to, dtype=torch.int64)
    lab = torch.full((B, pad_to), -1, dtype=torch.int64)
    nsl = torch.zeros(B, dtype=torch.int64)
    for i, s in enumerate(samples):
        a, b = s["a_ids"], s["b_ids"]
        L = len(a) + len(b)
        ids[i, :L] = torch.tensor(list(a) + list(b), dtype=torch.int64)
        tt[i, len(a):L] = 1
        am[i, :L] = 1
        if len(s["masked_lm_positions"]):
            lab[i, torch.tensor(list(s["masked_lm_positions"]), dtype=torch.int64)] = torch.tensor(list(s["masked_lm_ids"]), dtype=torch.int64)
        nsl[i] = 1 if s["is_random_next"] else 0
    batch = {"input_ids": ids, "token_type_ids": tt, "attention_mask": am, "labels": lab, "next_sentence_labels": nsl}
    if pin and torch.cuda.is_available():
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


class BertPretrainBinnedLoader:
    def __init__(self, bins, seq_len, bin_size, batch_size, base_seed, start_epoch, pin, bin_weights=None):
        """bins: {bin index -> list of this rank's samples}; bin_weights: {bin index -> GLOBAL population} (identical on every rank)."""
        self.bins = {k: v for k, v in bins.items() if len(v) > 0}
        self.bin_weights = bin_weights
        self.seq_len, self.bin_size, self.batch_size = seq_len, bin_size, batch_size
        self.base_seed, self.epoch, self.pin = base_seed, start_epoch, pin
        total = sum(len(v) for v in self.bins.values())
        self._len = max(1, total // batch_size)
        self._cache = {}

    def __len__(self):
        return self._len

    def _pad_to(self, k):
        return self.seq_len if self.bin_size <= 0 else min(self.seq_len, (k + 1) * self.bin_size)

    def __iter__(self):
        keys = sorted(self.bins)
        weights = torch.tensor([float((self.bin_weights or {}).get(k, len(self.bins[k]))) for k in keys])
        g = torch.Generator().manual_seed(self.base_seed * 1000003 + self.epoch)       # same on every rank: same bin sequence
        cursor = {k: 0 for k in keys}
        for step in range(self._len):
            k = keys[int(torch.multinomial(weights, 1, generator=g).item())]
            pool = self.bins[k]
            start = cursor[k]
            sel = [pool[(start + i) % len(pool)] for i in range(self.batch_size)]
            cursor[k] = (start + self.batch_size) % len(pool)
            key = (k, start)
            if key not in self._cache:
                if len(self._cache) > 64:
                    self._cache.clear()
                self._cache[key] = _collate(sel, self._pad_to(k), self.pin)
            yield self._cache[key]
        self.epoch += 1


def _parse_spec(spec):
    opts = dict(seq_len=512, max_pred=80, samples=4096, bin_size=0, vocab=30522)
    if "?" in spec:
        for kv in spec.split("?", 1)[1].split("&"):
            if "=" in kv:
                k, v = kv.split("=", 1)
                opts[k] = int(v)
    for k in list(opts):                                  # environment overrides (the reference CLI has no such flags)
        env = os.environ.get("LDDL_SYNTH_" + k.upper())
        if env:
            opts[k] = int(env)
    return opts


def get_bert_pretrain_data_loader(path, local_rank=0, shuffle_buffer_size=16384, shuffle_buffer_warmup_factor=16, vocab_file=None,
                                  data_loader_kwargs=None, mlm_probability=0.15, base_seed=12345, log_dir=None, log_level=logging.INFO,
                                  return_raw_samples=False, start_epoch=0, sequence_length_alignment=8, ignore_index=-1, **unused):
    kw = dict(data_loader_kwargs or {})
    batch_size = int(kw.get("batch_size", 32))
    pin = bool(kw.get("pin_memory", False))
    rank, world = _rank_world(local_rank)
    bins, weights = {}, None
    if path is not None and os.path.isdir(str(path)) and os.path.exists(os.path.join(str(path), "meta.json")):
        import pyarrow.parquet as pq
        meta = json.load(open(os.path.join(path, "meta.json")))
        seq_len, bin_size = int(meta["seq_len"]), int(meta.get("bin_size", 0))
        weights = {int(k): float(v) for k, v in meta.get("bin_counts", {}).items()} or None
        for d in sorted(os.listdir(path)):
            if not d.startswith("bin_"):
                continue
            k = int(d.split("_")[1])
            shards = sorted(f for f in os.listdir(os.path.join(path, d)) if f.endswith(".parquet"))
            mine = [f for i, f in enumerate(shards) if i % world == rank] or shards[rank % max(len(shards), 1):][:1]
            rows = []
            for f in mine:
                rows += pq.read_table(os.path.join(path, d, f)).to_pylist()
            bins[k] = rows
    else:
        o = _parse_spec(str(path or "synthetic"))
        seq_len, bin_size = o["seq_len"], o["bin_size"]
        per_rank = max(batch_size, o["samples"] // world)
        for s in _synth_samples(per_rank, seq_len, o["max_pred"], o["vocab"], seed=base_seed + 7919 * rank + 1, len_seed=base_seed):
            L = len(s["a_ids"]) + len(s["b_ids"])
            k = 0 if bin_size <= 0 else (L - 1) // bin_size
            bins.setdefault(k, []).append(s)
    return BertPretrainBinnedLoader(bins, seq_len, bin_size, batch_size, int(base_seed), int(start_epoch), pin, bin_weights=weights)
