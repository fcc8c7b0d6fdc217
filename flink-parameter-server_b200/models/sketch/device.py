"""Device tier of the streaming sketches: packed sketches in HBM shards, updates as one-sided
reductions over NVLink (ops/csrc/fps_sketch.cu), queries as popcount / int-dot scans + top-K merge.

Keys (word hashes, arbitrary ints or strings) are interned to dense slot ids on the host
(:class:`KeyInterner`); slot ``s`` lives on shard ``s % G`` like every other table.
``hash64`` is shared with the host tier so a device-built sketch equals the host-built one.
"""
from __future__ import annotations

from typing import Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ...ops import native
from ...store.sharded_table import ShardedTable
from .hashing import hash64, java_string_hash
from .utils import merge_topk


class KeyInterner:
    """Opaque key -> dense int slot (string / opaque ids)."""

    def __init__(self):
        self.slot: Dict[Hashable, int] = {}
        self.keys: List[Hashable] = []

    def __call__(self, key: Hashable) -> int:
        s = self.slot.get(key)
        if s is None:
            s = len(self.keys)
            self.slot[key] = s
            self.keys.append(key)
        return s

    def __len__(self):
        return len(self.keys)


class DeviceSketch:
    """kind in {"bloom", "tow", "minhash"}; ``capacity`` = max number of distinct keys.

    ``time_aware=True``: records carry a time slot and the model is keyed ``(wordHash, slot)``
    (TimeAwareBloomFilter.scala:31-92, TimeAwareTugOfWar.scala:17-60); a query then gets one answer per
    slot of the word and is scored against same-slot keys only.  A second one-column table counts the
    occurrences of every key (the keyed word count of MinHashPredict.scala:19-37) with the same
    one-sided ``red.add``."""

    def __init__(self, kind: str, capacity: int, num_hashes: int, array_size: int = 0, group=None,
                 device: Optional[int] = None, time_aware: bool = False):
        self.kind, self.num_hashes, self.array_size = kind, int(num_hashes), int(array_size)
        self.time_aware = bool(time_aware)
        if kind == "bloom":
            self.words = (self.array_size + 31) // 32
        elif kind == "tow":
            self.words = self.num_hashes
        elif kind == "minhash":
            self.words = 2 * self.num_hashes
        else:
            raise ValueError(kind)
        self.table = ShardedTable(capacity, self.words, group=group, device=device, init="zeros")
        self.freq = ShardedTable(capacity, 1, group=group, device=device, init="zeros")
        self.dev = self.table.cuda_device
        self.rows_i32 = self.table.local.view(torch.int32)
        if kind == "minhash":
            self.rows_i32.fill_(-1)  # all-ones u64 = +inf for the min reduction
            self.table.barrier()
        self.interner = KeyInterner()
        self._word_slot: Dict[Hashable, int] = {}     # (word[, time slot]) -> key slot, one hash per distinct word
        self._slots_of_word: Dict[int, List[int]] = {}
        self._key_slot_dev = None
        self._key_slot_n = -1
        self._gather = None

    # -- keys -----------------------------------------------------------------------------------
    def _key(self, word: str, slot=None) -> int:
        k = (word, slot) if self.time_aware else word
        s = self._word_slot.get(k)
        if s is None:
            h = java_string_hash(word)
            s = self.interner((h, int(slot)) if self.time_aware else h)
            self._word_slot[k] = s
            if self.time_aware:
                lst = self._slots_of_word.setdefault(h, [])
                if int(slot) not in lst:
                    lst.append(int(slot))
        return s

    def _local_key_slots(self) -> torch.Tensor:
        """Time slot of every local row (-1: not a key), uploaded when the dictionary has grown."""
        if self._key_slot_n != len(self.interner):
            n_local = self.rows_i32.shape[0]
            ids = self.table.local_ids().cpu().numpy()
            slots = np.full(n_local, -1, dtype=np.int32)
            keys = self.interner.keys
            ok = ids < len(keys)
            slots[ok] = [keys[i][1] for i in ids[ok]]
            self._key_slot_dev = torch.from_numpy(slots).to(self.dev)
            self._key_slot_n = len(keys)
        return self._key_slot_dev

    # -- train: (tweetId, [words][, timeSlot]) records -----------------------------------------
    def update(self, records: Iterable[Tuple[object, Sequence[str]]]) -> None:
        keys, tweets = [], []
        key = self._key
        for rec in records:
            tid = int(rec[0])
            slot = rec[2] if self.time_aware else None
            ks = [key(w, slot) for w in rec[1]]
            keys.extend(ks)
            tweets.extend([tid] * len(ks))
        if not keys:
            return
        k = torch.tensor(keys, dtype=torch.int32).to(self.dev, non_blocking=True)
        t = torch.tensor(tweets, dtype=torch.int64).to(self.dev, non_blocking=True)
        self.update_ids(k, t)

    def update_ids(self, keys: torch.Tensor, tweets: torch.Tensor) -> None:
        """Tensor fast path: ``keys`` int32 dense key slots (already interned, < capacity) and
        ``tweets`` int64 tweet ids of the (word, tweet) occurrences, on the device or in pinned host
        memory.  One kernel, ``num_hashes`` one-sided reductions per occurrence, plus the keyed
        occurrence count."""
        k = keys.to(self.dev, non_blocking=True)
        t = tweets.to(self.dev, non_blocking=True)
        native.sketch_update(self.table.table_c, self.kind, k, t, self.num_hashes, self.array_size)
        if self._ones is None or self._ones.shape[0] < k.numel():
            self._ones = torch.ones((max(k.numel(), 1 << 16), 1), dtype=torch.float32, device=self.dev)
        native.push_add(self.freq.table_c, k, self._ones[: k.numel()])

    _ones = None

    # -- export (the close() dump of the *PSLogic classes) -------------------------------------
    def model(self) -> List[Tuple[int, object]]:
        torch.cuda.synchronize()
        self.table.barrier()
        ids = self.table.local_ids().cpu().tolist()
        rows = self.rows_i32.cpu().numpy()
        out = []
        for slot, id_ in enumerate(ids):
            if id_ >= len(self.interner):
                continue
            key = self.interner.keys[id_]
            r = rows[slot, : self.words]
            if self.kind == "bloom":
                bits = np.unpackbits(r.view(np.uint8), bitorder="little")[: self.array_size]
                out.append((key, frozenset(np.nonzero(bits)[0].tolist())))
            elif self.kind == "tow":
                out.append((key, r.astype(np.int64).tolist()))
            else:
                packed = r.view(np.uint64)
                out.append((key, [int(p & np.uint64(0xFFFFFFFF)) for p in packed]))
        return out

    def load_model(self, entries: Iterable[Tuple[object, object]], counts: Optional[Dict[object, int]] = None) -> None:
        """Model load of the predict jobs (``transformWithModelLoad``, FPS:715-908): ``entries`` are the
        train jobs' dumps -- ``(wordHash | (wordHash, slot), bitset | counters | arg-min tweet ids)``; they
        are written into the owning shards with one-sided stores.  ``counts``: keyed word counts."""
        keys, rows = [], []
        for key, val in entries:
            if self.time_aware:
                h, slot = key
                lst = self._slots_of_word.setdefault(h, [])
                if int(slot) not in lst:
                    lst.append(int(slot))
            keys.append(self.interner(key))
            r = np.zeros(self.words, dtype=np.int32)
            if self.kind == "bloom":
                bits = np.zeros(self.words * 32, dtype=np.uint8)
                bits[list(val)] = 1
                r[:] = np.packbits(bits, bitorder="little").view(np.int32)
            elif self.kind == "tow":
                r[:] = np.asarray(val, dtype=np.int64).astype(np.int32)
            else:       # arg-min tweet ids only: the hash half is not part of the dump (and not scored)
                r[:] = np.asarray(val, dtype=np.uint64).astype(np.uint64).view(np.int32)
            rows.append(r)
        if not keys:
            return
        ids = torch.tensor(keys, dtype=torch.int64, device=self.dev)
        vals = torch.from_numpy(np.stack(rows)).to(self.dev).view(torch.float32)
        native.push_assign(self.table.table_c, ids, vals.contiguous())
        if counts:
            c = torch.tensor([[float(counts.get(self.interner.keys[k], 0))] for k in keys],
                             dtype=torch.float32, device=self.dev)
            native.push_assign(self.freq.table_c, ids, c)
        self.table.barrier()

    def word_count(self) -> Dict[Hashable, int]:
        """Occurrences per key of the local shard (MinHashPredict.scala:19-37)."""
        self.freq.barrier()
        ids = self.freq.local_ids().cpu().tolist()
        f = self.freq.local[:, 0].cpu().tolist()
        return {self.interner.keys[i]: int(c) for i, c in zip(ids, f) if i < len(self.interner) and c > 0}

    # -- predict: co-occurrence top-K of a query word against the local shard -----------------
    def slots_of(self, word: str) -> List[int]:
        """Time slots in which ``word`` was seen (a time-aware pull has one answer per slot)."""
        return sorted(self._slots_of_word.get(java_string_hash(word), []))

    def _local_scores(self, word: str, num_means: int, slot, cooccurrence: bool):
        key = (java_string_hash(word), int(slot)) if self.time_aware else java_string_hash(word)
        kslot = self.interner.slot.get(key)
        if kslot is None:
            return None
        kid = torch.tensor([kslot], device=self.dev)
        q = self.table.pull(kid).view(torch.int32)[0].contiguous()        # the PULL of the query sketch
        n_local = self.rows_i32.shape[0]
        est = torch.empty(n_local, dtype=torch.float32, device=self.dev)
        ks = self._local_key_slots() if self.time_aware else None
        qs = int(slot) if self.time_aware else 0
        if self.kind == "bloom":
            native.bloom_query(self.rows_i32, self.words, q, float(self.array_size), float(self.num_hashes),
                               est, key_slot=ks, query_slot=qs)
        elif self.kind == "tow":
            native.sketch_query("tow", self.rows_i32, self.words, self.num_hashes, q, est,
                                num_means=num_means, key_slot=ks, query_slot=qs)
        else:
            fq = float(self.freq.pull(kid)[0, 0]) if cooccurrence else 0.0
            native.sketch_query("minhash", self.rows_i32, self.words, self.num_hashes, q, est, key_slot=ks,
                                query_slot=qs, freq=self.freq.local[:, 0].contiguous() if cooccurrence else None,
                                query_freq=fq)
        ids = self.table.local_ids()
        return torch.where(ids < len(self.interner), est, torch.full_like(est, -3.0e38))

    def _local_topk(self, est: torch.Tensor, K: int):
        """(scores[K], key ids[K]) of the local shard: radix select + bitonic sort (fps_row_topk)."""
        n = est.numel()
        K = max(1, min(K, 2048))
        idx = torch.arange(n, dtype=torch.int32, device=self.dev)
        row = 32768
        if n > 2 * row:     # two levels: K per 32K-slice, then K of the survivors
            pad = (-n) % row
            e2 = torch.nn.functional.pad(est, (0, pad), value=-3.0e38).view(-1, row)
            i2 = torch.nn.functional.pad(idx, (0, pad), value=-1).view(-1, row)
            s, i = native.row_topk(e2.contiguous(), i2.contiguous(), min(K, row))
            est, idx = s.reshape(1, -1).contiguous(), i.reshape(1, -1).contiguous()
        else:
            est, idx = est.view(1, -1), idx.view(1, -1)
        s, i = native.row_topk(est, idx, K)
        ids = self.table.local_ids()
        gid = torch.where(i[0] >= 0, ids[i[0].clamp(min=0).long()], torch.full_like(i[0], -1, dtype=torch.int64))
        return s[0], gid

    def query_local(self, word: str, K: int, num_means: int = 1, slot=None,
                    cooccurrence: bool = False) -> List[Tuple[float, object]]:
        est = self._local_scores(word, num_means, slot, cooccurrence)
        if est is None:
            return []
        s, gid = self._local_topk(est, K)
        out = []
        for sc, g in zip(s.cpu().tolist(), gid.cpu().tolist()):
            if g >= 0 and sc > -1.0e38:
                out.append((sc, self.interner.keys[g]))
        return out

    def query(self, word: str, K: int, num_means: int = 1, slot=None,
              cooccurrence: bool = False) -> List[Tuple[float, object]]:
        """Scatter the query to every shard, gather the local top-K lists and merge them (E8 + E9: the
        predict jobs' push-broadcast and parallelism-1 merge sink).  The partial lists travel as
        one-sided stores into the merging rank's buffers (:class:`P2PGather`), not through NCCL; the
        merge is ``fps_row_topk``.  Every rank must call it with the same word and needs the same
        interning (feed all ranks the same key stream or share the dictionary)."""
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.table.group) > 1
        if not multi:
            return self.query_local(word, K, num_means, slot, cooccurrence)
        est = self._local_scores(word, num_means, slot, cooccurrence)
        Kc = max(1, min(K, 2048))
        if est is None:
            s = torch.full((Kc,), -3.0e38, device=self.dev)
            gid = torch.full((Kc,), -1, dtype=torch.int64, device=self.dev)
        else:
            s, gid = self._local_topk(est, Kc)
        if self._gather is None:
            from ...parallel.fabric import P2PGather

            self._gather = P2PGather(2048 * 12, group=self.table.group, device=self.table.device)
        packed = torch.cat([s.view(torch.int32), gid.to(torch.int32)])          # [2K] one message per rank
        parts = self._gather.gather(packed, dst=None)
        sc = torch.cat([p[:Kc].view(torch.float32) for p in parts]).view(1, -1).contiguous()
        ids = torch.cat([p[Kc:] for p in parts]).view(1, -1).contiguous()
        ms, mi = native.row_topk(sc, ids, Kc)
        out = []
        for a, g in zip(ms[0].cpu().tolist(), mi[0].cpu().tolist()):
            if g >= 0 and a > -1.0e38:
                out.append((a, self.interner.keys[g]))
        return out

    def close(self):
        if self._gather is not None:
            self._gather.close()
        self.freq.close()
        self.table.close()


# host oracles (same hash family as the kernels) ----------------------------------------------
def bloom_positions64(tweet_id: int, num_hashes: int, array_size: int) -> List[int]:
    return [hash64(tweet_id, i) % array_size for i in range(num_hashes)]


def tow_bits64(tweet_id: int, num_hashes: int) -> List[int]:
    return [1 if (hash64(tweet_id, j >> 6) >> (j & 63)) & 1 else -1 for j in range(num_hashes)]


def minhash_packed64(tweet_id: int, num_hashes: int) -> List[int]:
    return [((hash64(tweet_id, j) >> 32) << 32) | (tweet_id & 0xFFFFFFFF) for j in range(num_hashes)]
