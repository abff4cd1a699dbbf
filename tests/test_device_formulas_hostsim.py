"""CPU checks of the DEVICE headers' formulas (schnorr-sig_b200/csrc/*.cuh compiled for the host by
tests/hostsim, PTX replaced by portable C) against the oracle.  This is host-logic coverage: the real
parity tests run the CUDA kernels through the C ABI (tests/test_gpu_*.py, -m gpu)."""
import ctypes as C

import numpy as np
import pytest

import cref
import pyref as o
from util import KAT96, make_workload, pt_from96, pt_to96, rand_fp, rand_fp6, rand_scalars, int_le


def p(a):
    return a.ctypes.data_as(C.c_void_p)


EDGE = [0, 1, 2, 7, 2**32 - 1, 2**32, 2**32 + 1, 2**63, o.P - 1, o.P - 2, o.P - 2**32, (o.P - 1) // 2, (o.P + 1) // 2]


def test_fp_ops(hostsim):
    rng = np.random.default_rng(1)
    vals = EDGE + [int(x) for x in rand_fp(rng, 200)]
    for a in vals[:40]:
        for b in vals[:40]:
            assert hostsim.hs_fp_mul(C.c_uint64(a), C.c_uint64(b)) == a * b % o.P
            assert hostsim.hs_fp_add(C.c_uint64(a), C.c_uint64(b)) == (a + b) % o.P
            assert hostsim.hs_fp_sub(C.c_uint64(a), C.c_uint64(b)) == (a - b) % o.P
    for a in vals:
        assert hostsim.hs_fp_sqr(C.c_uint64(a)) == a * a % o.P
        assert hostsim.hs_fp_mul_small(C.c_uint64(a), C.c_uint32(7)) == 7 * a % o.P
        assert hostsim.hs_fp_mul_small(C.c_uint64(a), C.c_uint32(2**32 - 1)) == (2**32 - 1) * a % o.P
        if a:
            assert hostsim.hs_fp_inv(C.c_uint64(a)) == pow(a, o.P - 2, o.P)
        r = C.c_uint64(0)
        ok = hostsim.hs_fp_sqrt(C.c_uint64(a * a % o.P), C.byref(r))
        assert ok and r.value in (a, (o.P - a) % o.P)
        assert hostsim.hs_rescue_inv_sbox(C.c_uint64(a)) == pow(a, o.INV_ALPHA, o.P)
    assert not hostsim.hs_fp_sqrt(C.c_uint64(7), C.byref(C.c_uint64(0)))       # 7 generates Fp^*


def test_fp_reduce160_full_range(hostsim):
    rng = np.random.default_rng(2)
    for _ in range(2000):
        w = [int(x) for x in rng.integers(0, 2**32, 4, dtype=np.uint64)] + [int(rng.integers(0, 512))]
        x = sum(v << (32 * i) for i, v in enumerate(w))
        assert hostsim.hs_fp_reduce160(*[C.c_uint32(v) for v in w]) == x % o.P
    top = [2**32 - 1] * 4 + [511]
    assert hostsim.hs_fp_reduce160(*[C.c_uint32(v) for v in top]) == sum(v << (32 * i) for i, v in enumerate(top)) % o.P


def test_fp6_mul_sqr_inv_sqrt(hostsim):
    rng = np.random.default_rng(3)
    cases = [rand_fp6(rng) for _ in range(60)]
    cases += [np.array([o.P - 1] * 6, dtype=np.uint64), np.zeros(6, dtype=np.uint64),
              np.array([1, 0, 0, 0, 0, 0], dtype=np.uint64), np.array([0, 0, 0, 0, 0, o.P - 1], dtype=np.uint64),
              np.array([5, 0, 9, 0, 11, 0], dtype=np.uint64), np.array([0, 3, 0, 0, 0, 0], dtype=np.uint64)]
    r = np.zeros(6, dtype=np.uint64)
    for i, a in enumerate(cases):
        b = cases[(i * 7 + 3) % len(cases)]
        ta, tb = tuple(int(x) for x in a), tuple(int(x) for x in b)
        hostsim.hs_fp6_mul(p(a), p(b), p(r))
        assert tuple(int(x) for x in r) == o.f6_mul(ta, tb)
        hostsim.hs_fp6_sqr(p(a), p(r))
        assert tuple(int(x) for x in r) == o.f6_sqr(ta)
        if any(ta):
            hostsim.hs_fp6_inv(p(a), p(r))
            assert tuple(int(x) for x in r) == o.f6_inv(ta)
        sq = np.array(o.f6_sqr(ta), dtype=np.uint64)
        ok = hostsim.hs_fp6_sqrt(p(sq), p(r))
        assert ok and tuple(int(x) for x in r) in (ta, o.f6_neg(ta))
        assert bool(hostsim.hs_fp6_lex_largest(p(a))) == o.f6_lex_largest(ta)
    # non-squares are rejected (curve constant B) and agree with the oracle on random inputs
    assert not hostsim.hs_fp6_sqrt(p(np.array(o.CURVE_B, dtype=np.uint64)), p(r))
    for a in cases[:30]:
        ok = hostsim.hs_fp6_sqrt(p(a), p(r))
        assert bool(ok) == cref.fp6_sqrt(a)[0]


def test_rescue_and_hash_message(hostsim):
    rng = np.random.default_rng(4)
    for _ in range(4):
        s = rand_fp(rng, 12)
        want = o.rescue_permutation([int(x) for x in s])
        hostsim.hs_rescue_permutation(p(s))
        assert [int(x) for x in s] == want
    lens = [0, 1, 6, 7, 8, 13, 14, 15, 21, 49, 50, 80, 160, 163]
    w = make_workload(5, len(lens), lens=lens)
    want = cref.hash_messages(w["sigs"][:, :48], w["pk"], w["blob"], w["off"])
    out = np.zeros(32, dtype=np.uint8)
    for i, m in enumerate(w["msgs"]):
        rx = np.frombuffer(bytes(w["sigs"][i, :48]), dtype=np.uint64).copy()
        px = np.frombuffer(bytes(w["pk"][i, :48]), dtype=np.uint64).copy()
        py0 = int(np.frombuffer(bytes(w["pk"][i, 48:56]), dtype=np.uint64)[0])
        mb = np.frombuffer(m, dtype=np.uint8).copy() if m else np.zeros(1, dtype=np.uint8)
        hostsim.hs_hash_message(p(rx), p(px), C.c_uint64(py0), p(mb), C.c_uint64(len(m)), p(out))
        assert bytes(out) == bytes(want[i])


def test_scalar_field(hostsim):
    rng = np.random.default_rng(6)
    r = np.zeros(32, dtype=np.uint8)
    edge = [0, 1, o.Q - 1, o.Q - 2, 2**254, (o.Q - 1) // 2]
    vals = edge + [int_le(s) for s in rand_scalars(rng, 30)]
    for a in vals:
        for b in vals[:12]:
            A = np.frombuffer(a.to_bytes(32, "little"), dtype=np.uint8).copy()
            B = np.frombuffer(b.to_bytes(32, "little"), dtype=np.uint8).copy()
            hostsim.hs_sc_mul(p(A), p(B), p(r)); assert int_le(r) == a * b % o.Q
            hostsim.hs_sc_add(p(A), p(B), p(r)); assert int_le(r) == (a + b) % o.Q
            hostsim.hs_sc_sub(p(A), p(B), p(r)); assert int_le(r) == (a - b) % o.Q
    for v in [0, o.Q - 1, o.Q, o.Q + 1, 2 * o.Q, 2 * o.Q + 5, 2**256 - 1] + [int_le(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(20)]:
        V = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8).copy()
        hostsim.hs_sc_from_u256(p(V), p(r)); assert int_le(r) == v % o.Q
        assert bool(hostsim.hs_sc_geq_q(p(V))) == (v >= o.Q)


def test_point_formulas(hostsim):
    rng = np.random.default_rng(8)
    G = o.generator()
    g96 = pt_to96(G)
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    pts = [o.pt_mul(G, int_le(s)) for s in rand_scalars(rng, 4)] + [G, (o.KAT_X, o.KAT_Y)]
    for a in pts:
        for b in pts:
            A, B = pt_to96(a), pt_to96(b)
            want = o.pt_add(a, b)
            hostsim.hs_pt_add(p(A), 0, p(B), 0, p(out), C.byref(oi))
            assert (o.INF if oi.value else pt_from96(out)) == want
            hostsim.hs_pt_madd(p(A), 0, p(B), 0, p(out), C.byref(oi))
            assert (o.INF if oi.value else pt_from96(out)) == want
        A = pt_to96(a)
        hostsim.hs_pt_dbl(p(A), 0, p(out), C.byref(oi))
        assert pt_from96(out) == o.pt_add(a, a)
        # exceptional cases: P + (-P), identity operands
        N = pt_to96(o.pt_neg(a))
        hostsim.hs_pt_add(p(A), 0, p(N), 0, p(out), C.byref(oi)); assert oi.value == 1
        hostsim.hs_pt_madd(p(A), 0, p(N), 0, p(out), C.byref(oi)); assert oi.value == 1
        hostsim.hs_pt_add(p(A), 1, p(A), 0, p(out), C.byref(oi)); assert pt_from96(out) == a and not oi.value
        hostsim.hs_pt_add(p(A), 0, p(A), 1, p(out), C.byref(oi)); assert pt_from96(out) == a and not oi.value
        hostsim.hs_pt_madd(p(A), 1, p(A), 0, p(out), C.byref(oi)); assert pt_from96(out) == a and not oi.value
        hostsim.hs_pt_madd(p(A), 0, p(A), 1, p(out), C.byref(oi)); assert pt_from96(out) == a and not oi.value
    hostsim.hs_pt_dbl(p(g96), 1, p(out), C.byref(oi)); assert oi.value == 1
    # a point of order 2 (y = 0) doubles to the identity: find one from the cofactor structure
    kat = (o.KAT_X, o.KAT_Y)
    t2 = o.pt_mul(kat, o.COFACTOR // 2 * o.Q)
    assert t2 is not o.INF and t2[1] == o.F6_ZERO
    T2 = pt_to96(t2)
    hostsim.hs_pt_dbl(p(T2), 0, p(out), C.byref(oi)); assert oi.value == 1
    hostsim.hs_pt_add(p(T2), 0, p(T2), 0, p(out), C.byref(oi)); assert oi.value == 1


def test_torsion_check_including_small_order_points(hostsim):
    kat = (o.KAT_X, o.KAT_Y)
    assert hostsim.hs_torsion_free(p(pt_to96(o.generator())), 0) == 1
    assert hostsim.hs_torsion_free(p(KAT96), 0) == 0
    assert hostsim.hs_torsion_free(p(KAT96), 1) == 1            # identity
    n = o.COFACTOR * o.Q
    # points of small order 2, 5, 10, 29 and of order q*2: adversarial keys that hit P+P / P-P inside the chain
    for f in (2, 5, 10, 29, 2 * 5 * 29 * 181):
        pt = o.pt_mul(kat, n // f)
        if pt is o.INF:
            continue
        assert hostsim.hs_torsion_free(p(pt_to96(pt)), 0) == int(o.is_torsion_free(pt)) == 0
    mixed = o.pt_add(o.generator(), o.pt_mul(kat, n // 2))       # order 2q
    assert hostsim.hs_torsion_free(p(pt_to96(mixed)), 0) == 0
    rng = np.random.default_rng(9)
    for s in rand_scalars(rng, 3):
        pt = o.pt_mul(o.generator(), int_le(s))
        assert hostsim.hs_torsion_free(p(pt_to96(pt)), 0) == 1


def test_fixed_and_double_base(hostsim):
    rng = np.random.default_rng(10)
    G = o.generator()
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    ks = [0, 1, 255, 256, o.Q - 1, 2**248] + [int_le(s) for s in rand_scalars(rng, 4)]
    for k in ks:
        K = np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint8).copy()
        hostsim.hs_fixed_base_mul(p(K), p(out), C.byref(oi))
        assert (o.INF if oi.value else pt_from96(out)) == o.pt_mul(G, k % o.Q)
    P = o.pt_mul(G, 0x1234567)
    P96 = pt_to96(P)
    hs = [0, 1, 2, 3, o.Q - 1, o.Q - 2] + [int_le(s) for s in rand_scalars(rng, 4)]
    for h in hs:
        for e in (0, 5, int_le(rand_scalars(rng, 1)[0])):
            H = np.frombuffer(h.to_bytes(32, "little"), dtype=np.uint8).copy()
            E = np.frombuffer(e.to_bytes(32, "little"), dtype=np.uint8).copy()
            hostsim.hs_double_base(p(P96), 0, p(H), p(E), p(out), C.byref(oi))
            assert (o.INF if oi.value else pt_from96(out)) == o.pt_mul2(P, h, G, e)
    # identity public key: result is e*G
    E = np.frombuffer((77).to_bytes(32, "little"), dtype=np.uint8).copy()
    hostsim.hs_double_base(p(P96), 1, p(E), p(E), p(out), C.byref(oi))
    assert pt_from96(out) == o.pt_mul(G, 77)


def test_decompress(hostsim):
    w = make_workload(12, 6)
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    for i in range(6):
        rec = w["sigs"][i, :49].copy()
        ok, want, winf = cref.decompress(rec)
        assert hostsim.hs_decompress(p(rec), p(out), C.byref(oi)) == 1 and ok
        assert bytes(out) == bytes(want)
        rec[48] ^= 0x40
        assert hostsim.hs_decompress(p(rec), p(out), C.byref(oi)) == 1
        assert bytes(out[:48]) == bytes(want[:48]) and bytes(out[48:]) != bytes(want[48:])
    for bad in (np.zeros(49, np.uint8), np.full(49, 255, np.uint8)):
        assert hostsim.hs_decompress(p(bad), p(out), C.byref(oi)) == 0
    ident = np.zeros(49, np.uint8); ident[48] = 0x80
    assert hostsim.hs_decompress(p(ident), p(out), C.byref(oi)) == 1 and oi.value == 1


def test_verify_one_matches_oracle_verdicts(hostsim):
    lens = [8, 0, 7, 80, 160, 3]
    w = make_workload(13, len(lens), lens=lens)
    sigs, pk = w["sigs"].copy(), w["pk"].copy()
    sigs[1, 49:] = 0                      # e := 0
    pk[2] = KAT96                         # off-subgroup key
    sigs[3, :48] = 0; sigs[3, 48] = 0x80  # x := identity encoding
    sigs[4, 8:16] = 0xFF                  # non-canonical limb
    want = cref.verify_many(sigs, pk, w["inf"], w["blob"], w["off"])
    assert list(want) == [0, 2, 1, 2, 3, 0]
    for i, m in enumerate(w["msgs"]):
        mb = np.frombuffer(m, dtype=np.uint8).copy() if m else np.zeros(1, dtype=np.uint8)
        got = hostsim.hs_verify_one(p(sigs[i].copy()), p(pk[i].copy()), 0, p(mb), C.c_uint64(len(m)))
        assert got == want[i]


def test_shared_doubling_core(hostsim):
    """torsion_check_and_mul: [q]P == O and h*P from one doubling chain, incl. adversarial keys."""
    rng = np.random.default_rng(21)
    d = np.zeros(64, dtype=np.int8)
    for k in [0, 1, 8, 9, 15, 16, o.Q - 1, 2**255 - 1] + [int_le(s) for s in rand_scalars(rng, 40)]:
        K = np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint8).copy()
        hostsim.hs_recode_signed_w4(p(K), p(d))
        assert all(abs(int(x)) <= 8 for x in d) and sum(int(x) << (4 * i) for i, x in enumerate(d)) == k
    G = o.generator()
    kat = (o.KAT_X, o.KAT_Y)
    n = o.COFACTOR * o.Q
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    pts = [G, o.pt_mul(G, 12345), kat, o.pt_mul(kat, n // 2), o.pt_mul(kat, n // 5), o.pt_mul(kat, n // 10),
           o.pt_add(G, o.pt_mul(kat, n // 2)), o.pt_mul(kat, o.Q)]
    hs = [0, 1, 2, o.Q - 1, 0x8888888888888888, int_le(rand_scalars(rng, 1)[0])]
    pts = [pt for pt in pts if pt is not o.INF]      # n/5 * KAT is the identity: KAT's order has no factor 5
    assert len(pts) >= 6
    for pt in pts:
        for h in hs:
            H = np.frombuffer(h.to_bytes(32, "little"), dtype=np.uint8).copy()
            tf = hostsim.hs_torsion_check_and_mul(p(pt_to96(pt)), 0, p(H), p(out), C.byref(oi))
            assert bool(tf) == o.is_torsion_free(pt)
            assert (o.INF if oi.value else pt_from96(out)) == o.pt_mul(pt, h)
    H = np.frombuffer((77).to_bytes(32, "little"), dtype=np.uint8).copy()
    assert hostsim.hs_torsion_check_and_mul(p(KAT96), 1, p(H), p(out), C.byref(oi)) == 1 and oi.value == 1


# ---- fast path in (X, Y, w) coordinates (schnorr-sig_b200/csrc/affine.cuh) ------------------------

def test_verify_core_fast(hostsim):
    """[q]P == O and h*P + e*G from the (X, Y, w) chain; exceptional keys are reported, never mis-evaluated."""
    rng = np.random.default_rng(33)
    G = o.generator()
    kat = (o.KAT_X, o.KAT_Y)
    n = o.COFACTOR * o.Q
    out = np.zeros(96, dtype=np.uint8)
    TF, NTF, EXC = 0, 1, 2

    def run(pt, h, e):
        H = np.frombuffer(h.to_bytes(32, "little"), dtype=np.uint8).copy()
        E = np.frombuffer(e.to_bytes(32, "little"), dtype=np.uint8).copy()
        return hostsim.hs_verify_core_fast(p(pt_to96(pt)), p(H), p(E), p(out))

    good = [o.pt_mul(G, 12345), o.pt_mul(G, int_le(rand_scalars(rng, 1)[0]))]
    hs = [0, 1, 2, 8, 0x8888888888888888, o.Q - 1, 2**255 - 1] + [int_le(s) for s in rand_scalars(rng, 3)]
    es = [0, 1, 4096, 4097, o.Q - 1] + [int_le(s) for s in rand_scalars(rng, 2)]
    for pt in good:
        for h in hs:
            for e in es[:3] if h not in hs[-3:] else es:
                fr = run(pt, h, e)
                want = o.pt_mul2(pt, h % o.Q if h < o.Q else h, G, e)
                if want is o.INF:
                    assert fr == EXC
                else:   # incl. tiny / highly regular scalars whose bucket aggregation needs the O == R doubling
                    assert fr == TF, (h, e)
                    assert pt_from96(out) == want
    # h*P + e*G == O  (e = -h*k for P = k*G): the identity result is left to the exact routine
    assert run(o.pt_mul(G, 5), 3, o.Q - 15) == EXC
    # off-subgroup key of large order: decided by the fast path
    fr = run(kat, hs[-1], es[-1])
    assert fr == NTF and pt_from96(out) == o.pt_mul2(kat, hs[-1], G, es[-1])
    # small-order / mixed-order adversarial keys: exceptional or a correct "not torsion free"
    for pt in (o.pt_mul(kat, n // 2), o.pt_mul(kat, n // 10), o.pt_mul(kat, n // 29), o.pt_add(G, o.pt_mul(kat, n // 2)),
               o.pt_mul(kat, o.Q)):
        if pt is o.INF:
            continue
        fr = run(pt, hs[-2], es[-2])
        assert fr in (NTF, EXC)
        if fr == NTF:
            assert pt_from96(out) == o.pt_mul2(pt, hs[-2], G, es[-2])


def test_verify_one_fast_matches_oracle_verdicts(hostsim):
    lens = [8, 0, 7, 80, 160, 3, 8, 8]
    w = make_workload(13, len(lens), lens=lens)
    sigs, pk, inf = w["sigs"].copy(), w["pk"].copy(), w["inf"].copy()
    sigs[1, 49:] = 0                      # e := 0
    pk[2] = KAT96                         # off-subgroup key
    sigs[3, :48] = 0; sigs[3, 48] = 0x80  # x := identity encoding
    sigs[4, 8:16] = 0xFF                  # non-canonical limb
    n = o.COFACTOR * o.Q
    pk[6] = pt_to96(o.pt_mul((o.KAT_X, o.KAT_Y), n // 2))   # key of order 2
    inf[7] = 1                            # identity key
    want = cref.verify_many(sigs, pk, inf, w["blob"], w["off"])
    assert list(want[:6]) == [0, 2, 1, 2, 3, 0] and want[6] == 1
    used = C.c_int(0)
    exact_used = []
    for i, m in enumerate(w["msgs"]):
        mb = np.frombuffer(m, dtype=np.uint8).copy() if m else np.zeros(1, dtype=np.uint8)
        got = hostsim.hs_verify_one_fast(p(sigs[i].copy()), p(pk[i].copy()), int(inf[i]), p(mb), C.c_uint64(len(m)), C.byref(used))
        assert got == want[i], i
        exact_used.append(used.value)
    assert exact_used[:6] == [0, 0, 0, 0, 0, 0] and exact_used[6:] == [1, 1]


def test_fp_denominator_jacobian_operations(hostsim):
    """(X, Y, w) coordinates with w in Fp (affine.cuh: jf_dbl / jf_add / fp6_cofactor_norm)."""
    hostsim.hs_fp6_cofactor_norm.restype = C.c_uint64
    rng = np.random.default_rng(41)
    c = np.zeros(6, dtype=np.uint64)
    for d in [rand_fp6(rng) for _ in range(20)] + [np.array([1, 0, 0, 0, 0, 0], dtype=np.uint64),
                                                   np.array([0, 0, 0, 0, 0, o.P - 1], dtype=np.uint64),
                                                   np.array([0, 3, 0, 0, 0, 0], dtype=np.uint64)]:
        n = hostsim.hs_fp6_cofactor_norm(p(d), p(c))
        td = tuple(int(x) for x in d)
        assert n != 0 and o.f6_mul(td, tuple(int(x) for x in c)) == (n, 0, 0, 0, 0, 0)
    assert hostsim.hs_fp6_cofactor_norm(p(np.zeros(6, dtype=np.uint64)), p(c)) == 0
    G = o.generator()
    kat = (o.KAT_X, o.KAT_Y)
    pts = [o.pt_mul(G, int_le(s)) for s in rand_scalars(rng, 5)] + [kat, G]
    ws = [1, 2, o.P - 1, 0x123456789abcdef] + [int(x) for x in rand_fp(rng, 3)]
    out = np.zeros(96, dtype=np.uint8)
    NOP, ADD, SUB, SET, SETNEG = 0, 1, 2, 3, 4
    for jf_add, jf_dbl in ((hostsim.hs_jf_add, hostsim.hs_jf_dbl), (hostsim.hs_jf_add_fused, hostsim.hs_jf_dbl_fused)):
        for i, a in enumerate(pts):
            wa, wb = ws[i % len(ws)], ws[(3 * i + 1) % len(ws)]
            b = pts[(i + 2) % len(pts)]
            A, B = pt_to96(a), pt_to96(b)
            assert jf_dbl(p(A), C.c_uint64(wa), p(out)) == 0 and pt_from96(out) == o.pt_add(a, a)
            for mode, want in ((ADD, o.pt_add(a, b)), (SUB, o.pt_add(a, o.pt_neg(b))), (SET, b), (SETNEG, o.pt_neg(b)), (NOP, a)):
                assert jf_add(p(A), C.c_uint64(wa), p(B), C.c_uint64(wb), mode, p(out)) == 0
                assert pt_from96(out) == want, (i, mode)
            # exceptional inputs are reported when active, ignored when masked
            assert jf_add(p(A), C.c_uint64(wa), p(A), C.c_uint64(wb), ADD, p(out)) == 1
            assert jf_add(p(A), C.c_uint64(wa), p(A), C.c_uint64(wb), SUB, p(out)) == 1
            assert jf_add(p(A), C.c_uint64(wa), p(A), C.c_uint64(wb), NOP, p(out)) == 0 and pt_from96(out) == a
    t2 = o.pt_mul(kat, o.COFACTOR // 2 * o.Q)
    assert hostsim.hs_jf_dbl(p(pt_to96(t2)), C.c_uint64(5), p(out)) == 1
    assert hostsim.hs_jf_dbl_fused(p(pt_to96(t2)), C.c_uint64(5), p(out)) == 1


def test_executed_multiply_counts_match_design_doc(hostsim):
    """The executed 32x32->64 multiply counts quoted in DESIGN.md §4 (test-only counter in the host build)."""
    hostsim.hs_wide_count_reset.restype = C.c_ulonglong
    w = make_workload(91, 2, lens=[8, 8])
    used = C.c_int(0)
    counts = {}
    for name, fn in (("fast", lambda i, mb, m: hostsim.hs_verify_one_fast(p(w["sigs"][i].copy()), p(w["pk"][i].copy()), 0, p(mb),
                                                                         C.c_uint64(len(m)), C.byref(used))),
                     ("exact", lambda i, mb, m: hostsim.hs_verify_one(p(w["sigs"][i].copy()), p(w["pk"][i].copy()), 0, p(mb),
                                                                      C.c_uint64(len(m))))):
        hostsim.hs_gtab()
        hostsim.hs_wide_count_reset()
        mb = np.frombuffer(w["msgs"][0], dtype=np.uint8).copy()
        assert fn(0, mb, w["msgs"][0]) == 0
        counts[name] = hostsim.hs_wide_count_reset()
    import cost_model
    # the executed-multiply figures behind roofline.frac_executed (cost_model.py is their single source)
    assert counts["fast"] == cost_model.W_EXECUTED_FAST_L8, counts
    assert counts["exact"] == cost_model.W_EXECUTED_EXACT_L8, counts
    assert cost_model.w_per_verify(8) == 787338 and cost_model.w_per_verify(80) == 843618
    assert cost_model.w_per_hash(8) == 56280 and cost_model.permutations_for(160) == 5
    print(counts)


def test_exact_bucket_accumulation_step(hostsim):
    """jf_madd_exact (MSM bucket accumulation): identity accumulator, P + P, P - P, 2-torsion, generic."""
    rng = np.random.default_rng(43)
    G = o.generator()
    kat = (o.KAT_X, o.KAT_Y)
    pts = [o.pt_mul(G, int_le(s)) for s in rand_scalars(rng, 3)] + [kat]
    t2 = o.pt_mul(kat, o.COFACTOR // 2 * o.Q)
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)

    def run(a, b, neg, w=0x1234567):
        A = pt_to96(a) if a is not o.INF else np.zeros(96, dtype=np.uint8)
        hostsim.hs_jf_madd_exact(p(A), int(a is o.INF), C.c_uint64(w), p(pt_to96(b)), int(neg), p(out), C.byref(oi))
        return o.INF if oi.value else pt_from96(out)

    for a in pts:
        for b in pts:
            for neg in (False, True):
                want = o.pt_add(a, o.pt_neg(b) if neg else b)
                assert run(a, b, neg) == want
        assert run(o.INF, a, False) == a and run(o.INF, a, True) == o.pt_neg(a)
        assert run(a, a, False, w=1) == o.pt_add(a, a) and run(a, a, True, w=o.P - 1) is o.INF
    assert run(t2, t2, False) is o.INF and run(t2, t2, True) is o.INF


def test_exact_general_addition_in_fp_denominator_coordinates(hostsim):
    """jf_add_exact (MSM reduction stages): identity operands, P + P, P - P, generic, for scrambled denominators."""
    rng = np.random.default_rng(44)
    G = o.generator()
    kat = (o.KAT_X, o.KAT_Y)
    pts = [o.pt_mul(G, int_le(s)) for s in rand_scalars(rng, 3)] + [kat, o.INF]
    t2 = o.pt_mul(kat, o.COFACTOR // 2 * o.Q)
    out = np.zeros(96, dtype=np.uint8)
    oi = C.c_int(0)
    z96 = np.zeros(96, dtype=np.uint8)

    def run(a, b, wa=0x1234567, wb=0xfedcba987):
        A = pt_to96(a) if a is not o.INF else z96
        B = pt_to96(b) if b is not o.INF else z96
        hostsim.hs_jf_add_exact(p(A), int(a is o.INF), C.c_uint64(wa), p(B), int(b is o.INF), C.c_uint64(wb), p(out), C.byref(oi))
        return o.INF if oi.value else pt_from96(out)

    for a in pts:
        for b in pts:
            assert run(a, b) == o.pt_add(a, b)
        if a is not o.INF:
            assert run(a, o.pt_neg(a)) is o.INF and run(a, a, 3, 5) == o.pt_add(a, a)
    assert run(t2, t2) is o.INF


def test_windowed_tonelli_shanks_in_fp(hostsim):
    """fp_sqrt_or_none (windowed discrete log, include/fp_sqrt_tables.h): squares, non-residues, roots of unity."""
    rng = np.random.default_rng(51)
    g = pow(7, (o.P - 1) >> 32, o.P)                      # generator of the 2-power torsion
    vals = [1, 4, o.P - 1, 2**32, 2**32 - 1, g, pow(g, 2, o.P), pow(g, 3, o.P), pow(g, 1 << 31, o.P), pow(g, (1 << 24) + 2, o.P),
            pow(7, 12345, o.P), pow(7, 2 * 999983, o.P)] + [int(x) for x in rand_fp(rng, 300) if x]
    r = C.c_uint64(0)
    n_sq = 0
    for a in vals:
        is_sq = pow(a, (o.P - 1) // 2, o.P) == 1
        ok = hostsim.hs_fp_sqrt(C.c_uint64(a), C.byref(r))
        assert bool(ok) == is_sq, a
        if is_sq:
            assert r.value < o.P and r.value * r.value % o.P == a
            n_sq += 1
    assert 100 < n_sq < len(vals) - 100


def test_lazily_reduced_building_blocks_accept_non_canonical_operands(hostsim):
    """fp6_mul_nc / fp6_sqr_sub2 / fp6_sqr_sub_scaled / fp6_mul_sub_scaled / fused cofactor (csrc/debug_ops.cuh) with
    operands in [p, 2^64) wherever the fast path hands them non-canonical values."""
    from util import check_lazy_ops, lazy_ops_inputs
    a, b = lazy_ops_inputs(np.random.default_rng(61))
    out = np.zeros((len(a), 8, 6), dtype=np.uint64)
    for i in range(len(a)):
        hostsim.hs_debug_lazy_ops(p(a[i].copy()), p(b[i].copy()), p(out[i]))
    check_lazy_ops(a, b, out)


def test_sha512_hmac_and_derivation_steps(hostsim):
    """derive.cuh (f4): SHA-512 / HMAC-SHA512 against hashlib, master key and private children against the oracle's
    restatement of src/derivation.rs:66-153."""
    import hashlib
    import hmac
    rng = np.random.default_rng(71)
    out = np.zeros(64, dtype=np.uint8)
    for L in (0, 1, 55, 56, 64, 100, 111):
        data = rng.integers(0, 256, max(L, 1), dtype=np.uint8)
        hostsim.hs_sha512_short(p(data), L, p(out))
        assert bytes(out) == hashlib.sha512(bytes(data[:L])).digest()
    for kl, dl in ((32, 53), (34, 32), (1, 0), (128, 111), (64, 64)):
        key = rng.integers(0, 256, kl, dtype=np.uint8)
        data = rng.integers(0, 256, max(dl, 1), dtype=np.uint8)
        hostsim.hs_hmac_sha512(p(key), kl, p(data), dl, p(out))
        assert bytes(out) == hmac.new(bytes(key), bytes(data[:dl]), hashlib.sha512).digest()
    seed = rng.integers(0, 256, 32, dtype=np.uint8)
    x = np.zeros(64, dtype=np.uint8)
    assert hostsim.hs_derive_master(p(seed), p(x)) == 1
    sk, chain, ok = o.hd_master_key(bytes(seed))
    assert ok and bytes(x) == sk.to_bytes(32, "little") + chain
    pk49 = np.frombuffer(bytes(o.compress(o.pt_mul(o.generator(), sk))), dtype=np.uint8).copy()
    child = np.zeros(64, dtype=np.uint8)
    for index in (0, 1, 2**31 - 1, 2**31, 2**32 - 1, 123456789):
        assert hostsim.hs_derive_private_child(p(x), p(pk49), C.c_uint32(index), p(child)) == 1
        ck, cc, cok = o.hd_derive_private(sk, chain, index)
        assert cok and bytes(child) == ck.to_bytes(32, "little") + cc
