"""Shared helpers for the tests: seeded synthetic workloads built with the ORACLE (checker side)."""
import numpy as np

import cref
import pyref as o


def rand_fp(rng, n):
    v = rng.integers(0, 2**64, n, dtype=np.uint64)
    return np.where(v >= np.uint64(o.P), v - np.uint64(o.P), v).astype(np.uint64)


def rand_fp6(rng):
    return rand_fp(rng, 6)


def rand_scalars(rng, n):
    """n x 32 little-endian bytes, each < 2^254 < q (canonical without reduction)."""
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s[:, 31] &= 0x3F
    return s


def pack_msgs(msgs):
    lens = [len(m) for m in msgs]
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if sum(lens) else np.zeros(0, dtype=np.uint8)
    return blob, off


def make_workload(seed, n, msg_len=8, lens=None, nthreads=None):
    """Keys, messages and valid signatures from the C oracle. Returns dict of numpy arrays."""
    rng = np.random.default_rng(seed)
    nt = nthreads or cref.default_threads()
    sk = rand_scalars(rng, n)
    nonce = rand_scalars(rng, n)
    if lens is None:
        lens = [msg_len] * n
    msgs = [bytes(rng.integers(0, 256, L, dtype=np.uint8)) for L in lens]
    blob, off = pack_msgs(msgs)
    pk, inf = cref.keygen(sk, nt)
    sigs = cref.sign_many(sk, pk, inf, blob, off, nonce, nt)
    rand = rand_scalars(rng, n)
    return dict(sk=sk, nonce=nonce, msgs=msgs, blob=blob, off=off, pk=pk, inf=inf, sigs=sigs, rand=rand)


def int_le(b):
    return int.from_bytes(bytes(b), "little")


def pt_from96(b, inf=0):
    if inf:
        return o.INF
    b = bytes(b)
    return (tuple(int.from_bytes(b[8 * i:8 * i + 8], "little") for i in range(6)),
            tuple(int.from_bytes(b[48 + 8 * i:56 + 8 * i], "little") for i in range(6)))


def pt_to96(pt):
    if pt is o.INF:
        return np.zeros(96, dtype=np.uint8)
    return np.frombuffer(o.f6_to_bytes(pt[0]) + o.f6_to_bytes(pt[1]), dtype=np.uint8).copy()


KAT96 = pt_to96((o.KAT_X, o.KAT_Y))
