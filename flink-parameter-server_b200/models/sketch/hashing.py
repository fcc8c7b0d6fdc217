"""Hash functions used by the sketches.

* :func:`java_string_hash` -- ``String.hashCode`` (the sketches key every word by it, e.g.
  BloomFilter.scala:48), so model files written here use the same keys as the reference's.
* :func:`murmur3_32` -- MurmurHash3 x86_32 with a seed (role of ``MurmurHash3.stringHash(id, i)``).
* :func:`hash64` -- seeded 64-bit mixer (role of ``LongHashFunction.xx(seed).hashLong(x)``; the
  device kernels use the same mixer so host and GPU sketches agree bit for bit).
"""
from __future__ import annotations

MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


def java_string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & MASK32
    return h - (1 << 32) if h & 0x80000000 else h


def _rotl32(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & MASK32


def murmur3_32(data, seed: int = 0) -> int:
    """MurmurHash3 x86_32 of a str (utf-8) / bytes; returns a signed 32-bit int like the JVM.
    Integers are hashed through their decimal string (tweet ids arrive as strings in the reference's
    readers; numeric ids hash to the same value as their string form)."""
    if isinstance(data, int):
        data = str(data)
    if isinstance(data, str):
        data = data.encode("utf-8")
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & MASK32
    n = len(data)
    for i in range(0, n - n % 4, 4):
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * c1) & MASK32
        k = _rotl32(k, 15)
        k = (k * c2) & MASK32
        h ^= k
        h = _rotl32(h, 13)
        h = (h * 5 + 0xE6546B64) & MASK32
    tail = data[n - n % 4:]
    k = 0
    if len(tail) >= 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * c1) & MASK32
        k = _rotl32(k, 15)
        k = (k * c2) & MASK32
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & MASK32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & MASK32
    h ^= h >> 16
    return h - (1 << 32) if h & 0x80000000 else h


def hash64(x: int, seed: int = 0) -> int:
    """splitmix64-style seeded mixer; unsigned 64-bit result (same constants as fps_sketch.cu)."""
    z = (x + 0x9E3779B97F4A7C15 * (seed + 1)) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def hash64_signed(x: int, seed: int = 0) -> int:
    h = hash64(x, seed)
    return h - (1 << 64) if h & (1 << 63) else h


def floor_mod(a: int, m: int) -> int:
    return a % m  # Python's % is already floorMod for positive m
