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


# ---- lazily reduced building blocks (csrc/debug_ops.cuh): expected values and input sets -----------------------
def lazy_ops_inputs(rng, n_random=200):
    """(a, b) pairs for debug_lazy_ops: a holds ARBITRARY 64-bit limbs (incl. non-canonical ones), b canonical limbs."""
    P = o.P
    nc_edge = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**64 - 2, 2**64 - 2**32, 2**64 - 2**32 + 1, P + 2**32 - 2, 2**63,
               2**32 - 1, 2**32, 0xFFFFFFFF00000000, 0xFFFFFFFFFFFF0000]
    c_edge = [0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, P - 2**32, (P - 1) // 2, 2**63]
    a, b = [], []
    for i, x in enumerate(nc_edge):
        for j, y in enumerate(c_edge):
            a.append([(x if (k + i) % 3 else nc_edge[(i + k) % len(nc_edge)]) for k in range(6)])
            b.append([(y if (k + j) % 2 else c_edge[(j + k) % len(c_edge)]) for k in range(6)])
    a += [[2**64 - 1] * 6, [P] * 6, [0] * 6]
    b += [[P - 1] * 6, [P - 1] * 6, [1, 0, 0, 0, 0, 0]]
    for _ in range(n_random):
        a.append([int(v) for v in rng.integers(0, 2**64, 6, dtype=np.uint64)])
        b.append([int(v) for v in rand_fp(rng, 6)])
    return np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64)


def check_lazy_ops(a, b, out):
    """out[i] = the 8 results of debug_lazy_ops for (a[i], b[i]); asserts them against the big-int oracle."""
    P = o.P
    for i in range(len(a)):
        ta = tuple(int(v) % P for v in a[i])
        tb = tuple(int(v) for v in b[i])
        ar = ta[1:] + ta[:1]
        br = tb[5:] + tb[:5]
        got = [tuple(int(v) for v in out[i, j]) for j in range(8)]
        assert got[0] == o.f6_mul(ta, ar), i
        assert got[1] == o.f6_sub(o.f6_sub(o.f6_sqr(ta), tb), tb), i
        assert got[2] == o.f6_sub(o.f6_sub(o.f6_sqr(ta), tb), o.f6_scalar(br, ta[3])), i
        assert got[3] == o.f6_sub(o.f6_mul(ta, ar), o.f6_scalar(tb, ta[5])), i
        n = got[5][0]
        assert got[5][1:] == (0, 0, 0, 0, 0) and (n == 0) == (not any(tb)), i
        assert o.f6_mul(tb, got[4]) == (n, 0, 0, 0, 0, 0), i
        want = o.f6_sub(o.f6_scalar(tb, ta[0]), o.f6_scalar(br, ta[1]))
        assert got[6] == want and got[7] == want, i
