"""Integer index work of the hot path, restated in numpy (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference itself only has the sorted-unique vocabulary map with OOV -> vocab_size
(rec_pangu/dataset/base_dataset.py:57-61,92) and the V+1-row table convention
(rec_pangu/models/layers/embedding.py:32).  The hashed-id encoder and the row-shard routing are the
B200 build's own integer work for BASELINE.json config 5; they are restated here
so the CUDA kernels can be checked bit-exactly.
"""
import numpy as np

__all__ = ["splitmix64", "hash_to_row", "shard_route", "bucket_by_owner", "bounds_check"]

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 lanes (wrap-around arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over='ignore'):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        x = x ^ (x >> np.uint64(31))
    return x


def hash_to_row(raw: np.ndarray, vocab_size: int) -> np.ndarray:
    """Hashed-id encoder: row = splitmix64(raw) mod V, int64 in [0, V).  Row V stays the OOV slot."""
    return (splitmix64(raw.view(np.uint64) if raw.dtype == np.int64 else raw) % np.uint64(vocab_size)).astype(np.int64)


def shard_route(idx: np.ndarray, world: int):
    """Row-shard routing for tables split over ``world`` GPUs: owner = idx mod G, local row = idx div G."""
    idx = idx.astype(np.int64)
    return (idx % world).astype(np.int32), idx // world


def bucket_by_owner(idx: np.ndarray, world: int):
    """Stable bucketing of one rank's indices by owner rank.

    Returns (perm, counts, local_rows): ``perm`` lists positions grouped by owner (stable within an owner),
    ``counts[g]`` is how many go to rank g, ``local_rows`` are the local row ids in send order.
    """
    owner, local = shard_route(idx, world)
    perm = np.argsort(owner, kind='stable').astype(np.int64)
    counts = np.bincount(owner, minlength=world).astype(np.int64)
    return perm, counts, local[perm]


def bounds_check(idx: np.ndarray, rows: int) -> bool:
    """True iff every index addresses a valid row of a [rows, D] table (rows = vocab_size + 1)."""
    return bool(((idx >= 0) & (idx < rows)).all())
