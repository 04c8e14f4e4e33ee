"""Deterministic synthetic inputs shared by the oracle and the GPU engine:
synth(tag, i, n) = SHAKE256(tag || LE64(i)).digest(n)."""
import hashlib
import struct

import numpy as np


def synth(tag, i, n):
    return hashlib.shake_256(tag.encode() + struct.pack("<Q", i)).digest(n)


def synth_rows(tag, count, n, start=0):
    return np.frombuffer(b"".join(synth(tag, start + i, n) for i in range(count)), dtype=np.uint8).reshape(count, n)


def synth_block(tag, count, n):
    """count x n pseudo-random bytes from ONE SHAKE256 stream (fast for large counts)."""
    return np.frombuffer(hashlib.shake_256(tag.encode()).digest(count * n), dtype=np.uint8).reshape(count, n)
