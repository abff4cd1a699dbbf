"""ctypes loader for oracle/libcref.so (oracle #2, the C restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/cref.c.  Array conventions (numpy, C-contiguous):
  sigs  uint8 [n,81]   x (48 LE limb bytes) | flag byte | e (32 LE bytes)   src/signature.rs:208-214
  pks   uint8 [n,96]   affine x (48) | y (48), the in-memory PublicKey(AffinePoint) src/public.rs:24
  pk_inf uint8 [n]     1 = identity
  msgs  uint8 blob + uint64 offsets [n+1]
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcref.so")
    src = os.path.join(_HERE, "cref.c")
    hdr = os.path.join(_HERE, "..", "include", "cheetah_params.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcref.so"], stdout=subprocess.DEVNULL)
    return so


_FAST = None


def fast_lib():
    """oracle/libcfast.so: the OPTIMISED CPU port (oracle/cfast.c) -- bench baseline, cross-checked against cref.c."""
    global _FAST
    if _FAST is None:
        so = os.path.join(_HERE, "libcfast.so")
        src = os.path.join(_HERE, "cfast.c")
        hdr = os.path.join(_HERE, "..", "include", "cheetah_params.h")
        if (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr)):
            subprocess.check_call(["make", "-C", _HERE, "-B", "libcfast.so"], stdout=subprocess.DEVNULL)
        _FAST = C.CDLL(so)
    return _FAST


def verify_many_fast(sigs, pks, pk_inf, msgs, off, nthreads=1):
    """Signature::verify for n items through the optimised port; same verdict codes as verify_many."""
    sigs, pks, msgs = _u8(sigs), _u8(pks), _u8(msgs)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = sigs.shape[0]
    inf = None if pk_inf is None else _u8(pk_inf)
    out = np.zeros(n, dtype=np.uint8)
    fast_lib().cfast_verify_many(C.c_uint64(n), _p(sigs), _p(pks), _p(inf), _p(msgs), _p(off), _p(out), C.c_int(nthreads))
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def default_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def verify_many(sigs, pks, pk_inf, msgs, off, nthreads=1):
    sigs, pks, msgs = _u8(sigs), _u8(pks), _u8(msgs)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = sigs.shape[0]
    inf = None if pk_inf is None else _u8(pk_inf)
    out = np.zeros(n, dtype=np.uint8)
    lib().cref_verify_many(C.c_uint64(n), _p(sigs), _p(pks), _p(inf), _p(msgs), _p(off), _p(out), C.c_int(nthreads))
    return out


def hash_messages(rx48, pks, msgs, off, nthreads=1):
    rx48, pks, msgs = _u8(rx48), _u8(pks), _u8(msgs)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = rx48.shape[0]
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().cref_hash_messages(C.c_uint64(n), _p(rx48), _p(pks), _p(msgs), _p(off), _p(out), C.c_int(nthreads))
    return out


def keygen(sk32, nthreads=1):
    sk32 = _u8(sk32)
    n = sk32.shape[0]
    pk = np.zeros((n, 96), dtype=np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    lib().cref_keygen(C.c_uint64(n), _p(sk32), _p(pk), _p(inf), C.c_int(nthreads))
    return pk, inf


def sign_many(sk32, pks, pk_inf, msgs, off, nonce32, nthreads=1):
    sk32, pks, msgs, nonce32 = _u8(sk32), _u8(pks), _u8(msgs), _u8(nonce32)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = sk32.shape[0]
    inf = None if pk_inf is None else _u8(pk_inf)
    out = np.zeros((n, 81), dtype=np.uint8)
    lib().cref_sign_many(C.c_uint64(n), _p(sk32), _p(pks), _p(inf), _p(msgs), _p(off), _p(nonce32), _p(out),
                         C.c_int(nthreads))
    return out


def verify_batch(sigs, pks, pk_inf, msgs, off, rand32, nthreads=1):
    """-> (verdict, lhs97, rhs97)"""
    sigs, pks, msgs, rand32 = _u8(sigs), _u8(pks), _u8(msgs), _u8(rand32)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = sigs.shape[0]
    inf = None if pk_inf is None else _u8(pk_inf)
    verdict = C.c_int(-1)
    lhs = np.zeros(97, dtype=np.uint8)
    rhs = np.zeros(97, dtype=np.uint8)
    lib().cref_verify_batch(C.c_uint64(n), _p(sigs), _p(pks), _p(inf), _p(msgs), _p(off), _p(rand32),
                            C.byref(verdict), _p(lhs), _p(rhs), C.c_int(nthreads))
    return verdict.value, lhs, rhs


# ---- low-level probes -------------------------------------------------------------------
def _f6(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def fp6_mul(a, b):
    r = np.zeros(6, dtype=np.uint64)
    lib().cref_fp6_mul(_p(_f6(a)), _p(_f6(b)), _p(r))
    return r


def fp6_inv(a):
    r = np.zeros(6, dtype=np.uint64)
    lib().cref_fp6_inv(_p(_f6(a)), _p(r))
    return r


def fp6_sqrt(a):
    r = np.zeros(6, dtype=np.uint64)
    ok = lib().cref_fp6_sqrt(_p(_f6(a)), _p(r))
    return bool(ok), r


def rescue_permutation(state):
    s = np.array(state, dtype=np.uint64)
    lib().cref_rescue_permutation(_p(s))
    return s


def pt_mul(pt96, inf, k32):
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    lib().cref_pt_mul(_p(_u8(pt96)), C.c_int(int(inf)), _p(_u8(k32)), _p(out), C.byref(oi))
    return out, oi.value


def pt_add(a96, ainf, b96, binf):
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    lib().cref_pt_add(_p(_u8(a96)), C.c_int(int(ainf)), _p(_u8(b96)), C.c_int(int(binf)), _p(out), C.byref(oi))
    return out, oi.value


def is_torsion_free(pt96, inf=0):
    return bool(lib().cref_is_torsion_free(_p(_u8(pt96)), C.c_int(int(inf))))


def compress(pt96, inf=0):
    out = np.zeros(49, dtype=np.uint8)
    lib().cref_compress(_p(_u8(pt96)), C.c_int(int(inf)), _p(out))
    return out


def decompress(in49):
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    ok = lib().cref_decompress(_p(_u8(in49)), _p(out), C.byref(oi))
    return bool(ok), out, oi.value


def scalar_mul(a32, b32):
    r = np.zeros(32, dtype=np.uint8)
    lib().cref_scalar_mul(_p(_u8(a32)), _p(_u8(b32)), _p(r))
    return r


def scalar_reduce(a32):
    r = np.zeros(32, dtype=np.uint8)
    lib().cref_scalar_reduce(_p(_u8(a32)), _p(r))
    return r


def workload(seed, n, msg_len, nthreads=1):
    """Seeded keys / messages / valid signatures produced entirely by this oracle (used by the CPU
    baseline legs of bench.py and by tests)."""
    rng = np.random.default_rng(seed)
    sk = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sk[:, 31] &= 0x3F
    nonce = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    nonce[:, 31] &= 0x3F
    blob = rng.integers(0, 256, n * msg_len, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(msg_len)
    pk, inf = keygen(sk, nthreads)
    sigs = sign_many(sk, pk, inf, blob, off, nonce, nthreads)
    return dict(sk=sk, nonce=nonce, blob=blob, off=off, pk=pk, inf=inf, sigs=sigs)
