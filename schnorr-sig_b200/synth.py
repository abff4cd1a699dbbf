"""Seeded synthetic workloads.  Keys and signatures are produced by the engine's own
device signer (KeyPair::new / KeyPair::sign restated as kernels); the signer is parity-tested against
the oracle in tests/.  Everything derives from one integer seed through numpy's counter-based Philox
generator, so any shard of a workload can be regenerated independently (multi-GPU: shard = rank)."""
import numpy as np

DEFAULT_SEED = 0x5343484E4F525231   # "SCHNORR1"


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[seed & (2**64 - 1), stream]))


def scalars(seed, stream, n):
    """n x 32 LE bytes, uniformly below 2^254 (< q: canonical without reduction)."""
    s = _rng(seed, stream).integers(0, 256, (n, 32), dtype=np.uint8)
    s[:, 31] &= 0x3F
    s[:, 0] |= 1   # never zero
    return s


def messages(seed, stream, n, length):
    blob = _rng(seed, stream).integers(0, 256, n * length, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length)).astype(np.uint64)
    return blob, off


def host_inputs(seed, n, msg_len=8, shard=0):
    """Secret keys, nonces, messages, batch randomisers for shard `shard` (host arrays)."""
    base = 16 * shard
    sk = scalars(seed, base + 1, n)
    nonce = scalars(seed, base + 2, n)
    blob, off = messages(seed, base + 3, n, msg_len)
    rand = scalars(seed, base + 4, n)
    return dict(sk=sk, nonce=nonce, blob=blob, off=off, rand=rand, n=n, msg_len=msg_len)


def signed_workload(engine, seed, n, msg_len=8, shard=0):
    """host_inputs + public keys and valid signatures from the device signer (host arrays)."""
    w = host_inputs(seed, n, msg_len, shard)
    pk, inf = engine.keygen(w["sk"])
    w["pk"], w["inf"] = pk, inf
    w["sigs"] = engine.sign_many(w["sk"], pk, inf, w["blob"], w["off"], w["nonce"])
    return w


# the off-subgroup point of the reference's tests (src/signature.rs:387-404): expects InvalidPublicKey
KAT_OFF_SUBGROUP = np.frombuffer(b"".join(int(c).to_bytes(8, "little") for c in (
    0x9BFCD3244AFCB637, 0x39005E478830B187, 0x7046F1C03B42C6CC, 0xB5EEAC99193711E5, 0x7FD272E724307B98, 0xCC371DD6DD5D8625,
    0x9D03FDC216DFAAE8, 0xBF4ADE2A7665D9B8, 0xF08B022D5B3262B7, 0x2EAF583A3CF15C6F, 0xA92531E4B1338285, 0x5B8157814141A7A7)),
    dtype=np.uint8).copy()


def inject_faults(w, every=1024):
    """Corrupts every `every`-th item round-robin and returns the expected verdicts:
    0 flip message byte -> 2; 1 swap P_i with P_{i+1} -> 2 (both); 2 e := 0 -> 2;
    3 x := identity encoding -> 2; 4 P_i := off-subgroup KAT point -> 1."""
    n = w["n"]
    expect = np.zeros(n, dtype=np.uint8)
    sigs, pk, blob = w["sigs"].copy(), w["pk"].copy(), w["blob"].copy()
    kind = 0
    for i in range(every // 2, n - 1, every):
        if kind == 0 and w["msg_len"] > 0:
            blob[int(w["off"][i])] ^= 0x01
            expect[i] = 2
        elif kind == 1:
            pk[[i, i + 1]] = pk[[i + 1, i]]
            expect[i] = expect[i + 1] = 2
        elif kind == 2:
            sigs[i, 49:] = 0
            expect[i] = 2
        elif kind == 3:
            sigs[i, :48] = 0
            sigs[i, 48] = 0x80
            expect[i] = 2
        elif kind == 4:
            pk[i] = KAT_OFF_SUBGROUP
            expect[i] = 1
        kind = (kind + 1) % 5
    out = dict(w)
    out.update(sigs=sigs, pk=pk, blob=blob, expect=expect)
    return out
