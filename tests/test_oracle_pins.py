"""Pins the oracles (oracle/pyref.py, oracle/cref.c) against every fact the reference holds for the
hot path and against each other.  CPU only."""
import json
import os

import numpy as np
import pytest

import cref
import pyref as o
from util import KAT96, make_workload, pack_msgs, pt_from96, pt_to96, rand_fp6, rand_scalars, int_le

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- facts read from the reference ------------------------------------------------------------
def test_field_and_curve_constants():
    assert o.P == 2**64 - 2**32 + 1                                   # README.md:4
    assert o.f6_mul((0, 1, 0, 0, 0, 0), (0, 0, 0, 0, 0, 1)) == (7, 0, 0, 0, 0, 0)   # u * u^5 = 7, README.md:8
    assert o.CURVE_B == (395, 1, 0, 0, 0, 0)                           # README.md:4-5


def test_kat_point_is_on_curve_and_off_subgroup():
    """src/signature.rs:385-406 / src/error.rs:47-70: raw limbs are canonical, c0-first, on the curve,
    and the point is NOT in the prime-order subgroup."""
    kat = (o.KAT_X, o.KAT_Y)
    assert all(c < o.P for c in o.KAT_X + o.KAT_Y)
    assert o.on_curve(kat)
    assert not o.is_torsion_free(kat)
    assert not cref.is_torsion_free(KAT96)
    # group order: #E = cofactor * q kills every point
    assert o.pt_mul(kat, o.COFACTOR * o.Q) is o.INF


def test_derived_generator_kat():
    """[cofactor] * KAT point, a point of exact order q (placeholder generator)."""
    G = o.generator()
    assert [hex(c) for c in G[0]] == ['0x88d96bf5ecf8bece', '0x310124185de3845f', '0x4e0641872249a1c5',
                                      '0xb3bbecb8d7bf04b2', '0xc071ab909eece75e', '0x470e1fc36b9e011f']
    assert [hex(c) for c in G[1]] == ['0x129ceb48ab1c38f7', '0x914e968c33511a8d', '0xd3944c84ce90db8f',
                                      '0x68802a73e3c97ee0', '0xd0b0a9463ce1de24', '0xed278ad637f63d69']
    assert o.on_curve(G) and o.is_torsion_free(G) and cref.is_torsion_free(pt_to96(G))


def test_subgroup_order_facts():
    # q prime-ish sanity (Fermat, several bases) and Hasse interval
    for a in (2, 3, 5, 7, 11):
        assert pow(a, o.Q - 1, o.Q) == 1
    n = o.COFACTOR * o.Q
    assert abs(n - (o.P**6 + 1)) <= 2 * o.P**3
    assert (o.INV_ALPHA * 7) % (o.P - 1) == 1


def test_byte_lengths_and_identity_encoding():
    """src/constants.rs:12-40, src/public.rs:95-101."""
    assert o.compress(o.INF) == bytes(48) + b"\x80"
    assert bytes(cref.compress(np.zeros(96, np.uint8), 1)) == bytes(48) + b"\x80"
    w = make_workload(3, 2)
    assert w["sigs"].shape == (2, 81) and w["pk"].shape == (2, 96)
    assert len(o.compress(pt_from96(w["pk"][0]))) == 49


def test_invalid_public_key_encodings():
    """src/public.rs:114-129,150-151: 49 zero bytes, all-0xff and flag byte 255 do not decode."""
    assert o.decompress(bytes(49))[0] is False                        # x = 0 is not on the curve
    assert not cref.decompress(np.zeros(49, np.uint8))[0]
    assert o.decompress(b"\xff" * 49)[0] is False
    assert not cref.decompress(np.full(49, 255, np.uint8))[0]
    G = o.compress(o.generator())
    assert o.decompress(G[:48] + b"\xff")[0] is False
    ok, pt = o.decompress(G)
    assert ok and pt == o.generator()
    ok, pt96, inf = cref.decompress(np.frombuffer(G, np.uint8))
    assert ok and not inf and bytes(pt96) == bytes(pt_to96(o.generator()))
    assert o.decompress(bytes(48) + b"\x80") == (True, o.INF)


def test_scalar_encodings():
    """src/private.rs:119-133,154-160: scalars are 32-byte LE; all-0xff and top byte 127 are >= q."""
    assert o.Q.to_bytes(32, "little")[31] == 0x7A
    assert int_le(cref.scalar_reduce(np.full(32, 255, np.uint8))) == (2**256 - 1) % o.Q
    one = np.zeros(32, np.uint8); one[0] = 1
    assert int_le(cref.scalar_reduce(one)) == 1


# ---- oracle #2 (C) against oracle #1 (Python big-int) ------------------------------------------
def test_c_oracle_field_ops_match_bigint():
    rng = np.random.default_rng(11)
    for _ in range(40):
        a, b = rand_fp6(rng), rand_fp6(rng)
        ta, tb = tuple(int(x) for x in a), tuple(int(x) for x in b)
        assert tuple(int(x) for x in cref.fp6_mul(a, b)) == o.f6_mul(ta, tb)
        assert tuple(int(x) for x in cref.fp6_inv(a)) == o.f6_inv(ta) == o.f6_inv_fermat(ta)
        ok, r = cref.fp6_sqrt(np.array(o.f6_sqr(ta), dtype=np.uint64))
        assert ok and tuple(int(x) for x in r) in (ta, o.f6_neg(ta))
    # edge values: 0, 1, p-1
    edge = np.array([0, 1, o.P - 1, o.P - 1, 0, 1], dtype=np.uint64)
    te = tuple(int(x) for x in edge)
    assert tuple(int(x) for x in cref.fp6_mul(edge, edge)) == o.f6_mul(te, te)
    assert not cref.fp6_sqrt(np.array(o.CURVE_B, dtype=np.uint64))[0]          # B is a non-square (App. A)


def test_c_oracle_rescue_matches_bigint():
    rng = np.random.default_rng(12)
    for _ in range(5):
        s = [int(x) % o.P for x in rng.integers(0, 2**64, 12, dtype=np.uint64)]
        assert [int(x) for x in cref.rescue_permutation(s)] == o.rescue_permutation(s)
    assert [int(x) for x in cref.rescue_permutation([0] * 12)] == o.rescue_permutation([0] * 12)


def test_c_oracle_points_and_scalars_match_bigint():
    rng = np.random.default_rng(13)
    G = o.generator()
    for _ in range(4):
        k = rand_scalars(rng, 1)[0]
        out, inf = cref.pt_mul(pt_to96(G), 0, k)
        assert not inf and pt_from96(out) == o.pt_mul(G, int_le(k))
        a, b = rand_scalars(rng, 1)[0], rng.integers(0, 256, 32, dtype=np.uint8)
        assert int_le(cref.scalar_mul(a, b)) == (int_le(a) * (int_le(b) % o.Q)) % o.Q
    # P + (-P), P + P, identity operands
    g96 = pt_to96(G)
    assert cref.pt_add(g96, 0, pt_to96(o.pt_neg(G)), 0)[1] == 1
    out, inf = cref.pt_add(g96, 0, g96, 0)
    assert pt_from96(out) == o.pt_mul(G, 2)
    out, inf = cref.pt_add(g96, 1, g96, 0)
    assert pt_from96(out) == G


@pytest.mark.parametrize("lens", [[0, 1, 6, 7, 8, 13, 14, 24, 48, 80, 160]])
def test_hash_message_padding_rule_and_parity(lens):
    """src/signature.rs:284-301: 7-byte chunks, 0x01 marker on a short tail, NO padding element when
    len % 7 == 0 -- so m||0x01 (7k bytes) collides with m (7k-1 bytes)."""
    w = make_workload(21, len(lens), lens=lens)
    d = cref.hash_messages(w["sigs"][:, :48], w["pk"], w["blob"], w["off"])
    for i, m in enumerate(w["msgs"]):
        rx = o.f6_from_bytes(bytes(w["sigs"][i, :48]))
        assert bytes(d[i]) == o.hash_message(rx, pt_from96(w["pk"][i]), m)
    m6 = b"abcdef"
    rx = o.f6_from_bytes(bytes(w["sigs"][0, :48])); pk = pt_from96(w["pk"][0])
    assert o.hash_message(rx, pk, m6) == o.hash_message(rx, pk, m6 + b"\x01")
    assert o.message_to_felts(b"") == [] and o.message_to_felts(b"\x05") == [0x0105]


# ---- the reference's behavioural tests replayed on the oracles --------------------------------
def test_reference_behaviour_single_verify():
    """src/signature.rs:333-426, src/error.rs:66-82."""
    w = make_workload(31, 4, lens=[160, 8, 8, 32])
    sigs, pk, inf, blob, off = w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]
    assert (cref.verify_many(sigs, pk, inf, blob, off) == 0).all()
    for i in range(4):
        assert o.verify(bytes(sigs[i, :49]), int_le(sigs[i, 49:]), w["msgs"][i], pt_from96(pk[i])) == o.OK
    # wrong message
    msgs = list(w["msgs"]); msgs[0] = bytes([42]) + msgs[0][1:]
    b2, o2 = pack_msgs(msgs)
    assert list(cref.verify_many(sigs, pk, inf, b2, o2)) == [2, 0, 0, 0]
    assert o.verify(bytes(sigs[0, :49]), int_le(sigs[0, 49:]), msgs[0], pt_from96(pk[0])) == o.INVALID_SIGNATURE
    # wrong public key = generator -> InvalidSignature ; off-subgroup key -> InvalidPublicKey
    pk2 = pk.copy(); pk2[0] = pt_to96(o.generator()); pk2[1] = KAT96
    assert list(cref.verify_many(sigs, pk2, inf, blob, off)) == [2, 1, 0, 0]
    assert o.verify(bytes(sigs[1, :49]), int_le(sigs[1, 49:]), w["msgs"][1], (o.KAT_X, o.KAT_Y)) == o.INVALID_PUBLIC_KEY
    # x := identity encoding ; e := 0
    s2 = sigs.copy(); s2[0, :48] = 0; s2[0, 48] = 0x80; s2[1, 49:] = 0
    assert list(cref.verify_many(s2, pk, inf, blob, off)) == [2, 2, 0, 0]
    assert o.verify(bytes(s2[0, :49]), int_le(s2[0, 49:]), w["msgs"][0], pt_from96(pk[0])) == o.INVALID_SIGNATURE
    # states the reference cannot hold / panics on
    s3 = sigs.copy(); s3[0, 49:] = 0xFF; s3[1, :8] = 0xFF
    assert list(cref.verify_many(s3, pk, inf, blob, off)) == [3, 3, 0, 0]
    # the flag byte of sig.x is ignored by single verify (src/signature.rs:186)
    s4 = sigs.copy(); s4[:, 48] ^= 0x40
    assert (cref.verify_many(s4, pk, inf, blob, off) == 0).all()


def test_reference_behaviour_batch():
    """src/batch.rs:138-179, tests/schnorr.rs:148-183: batches of 1, 3, 5 with a repeated signer verify;
    swapping two public keys fails."""
    for n, lens in ((1, [8]), (3, [24, 24, 48]), (5, [8] * 5)):
        w = make_workload(40 + n, n, lens=lens)
        if n == 5:   # signers 3 and 4 reuse key 0 (src/batch.rs:167-169)
            for j in (3, 4):
                w["sk"][j] = w["sk"][0]
            w["pk"], w["inf"] = cref.keygen(w["sk"])
            w["sigs"] = cref.sign_many(w["sk"], w["pk"], w["inf"], w["blob"], w["off"], w["nonce"])
        v, lhs, rhs = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], 2)
        assert v == 0 and bytes(lhs[:48]) == bytes(rhs[:48])
        pv, pl, pr = o.verify_batch([(bytes(s[:49]), int_le(s[49:])) for s in w["sigs"]],
                                    [pt_from96(k) for k in w["pk"]], w["msgs"], [int_le(r) for r in w["rand"]])
        assert pv == 0 and bytes(lhs[:96]) == bytes(pt_to96(pl)) and bytes(rhs[:96]) == bytes(pt_to96(pr))
        if n == 5:
            pk2 = w["pk"].copy(); pk2[[1, 2]] = pk2[[2, 1]]
            assert cref.verify_batch(w["sigs"], pk2, w["inf"], w["blob"], w["off"], w["rand"], 3)[0] == 2
    # the y-sign flag matters in the batch path (src/batch.rs:104): flipping it breaks the batch
    s2 = w["sigs"].copy(); s2[0, 48] ^= 0x40
    assert cref.verify_batch(s2, w["pk"], w["inf"], w["blob"], w["off"], w["rand"], 1)[0] == 2


def test_params_json_matches_oracle():
    with open(os.path.join(ROOT, "params", "params.json")) as f:
        pj = json.load(f)
    assert int(pj["p"]["value"], 16) == o.P and int(pj["q"]["value"], 16) == o.Q
    assert [int(c, 16) for c in pj["generator"]["x"]] == list(o.generator()[0])
    assert [[int(c, 16) for c in row] for row in pj["rescue"]["ark"]] == o.rescue_round_constants()
    assert pj["rescue"]["rounds"] == o.RESCUE_ROUNDS


def test_generated_subgroup_check_digits_represent_q():
    """The kernels' subgroup check consumes q through the generated digit table (include/cheetah_params.h:
    CHEETAH_Q_WNAF5): it must represent q exactly, with odd digits below 16 (eight buckets), and stop below bit 252 --
    the shared doubling chain of the verification ends with the last challenge window (tools/gen_params.py: fold_top)."""
    import re
    text = open(os.path.join(ROOT, "include", "cheetah_params.h")).read()
    body = re.search(r"CHEETAH_Q_WNAF5\[(\d+)\] = \{(.*?)\};", text, re.S)
    digits = [int(t) for t in body.group(2).replace("\n", " ").split(",") if t.strip()]
    assert len(digits) == int(body.group(1)) == 256
    assert sum(d << i for i, d in enumerate(digits)) == o.Q
    assert all(d == 0 or (d % 2 == 1 and abs(d) < 16) for d in digits)
    top = max(i for i, d in enumerate(digits) if d)
    assert top <= 251 and int(re.search(r"#define CHEETAH_Q_WNAF5_LEN (\d+)", text).group(1)) == top + 1
    assert sum(1 for d in digits if d) == 44


def test_upstream_dump_known_answers_when_present():
    """rust/dump_params (run on a machine with cargo + network) writes params/upstream_dump.json from the REAL cheetah /
    hash crates.  Absent here -> the oracle stays "parity unpinned" at value level and this test SKIPS;
    present -> every dumped known answer must be reproduced: fixed-base multiples of G, the Rescue digests that separate
    the padding rules, the y-sign flag rule, and the seeded key pairs / signatures."""
    import pytest
    d = o.upstream_dump()
    if d is None:
        pytest.skip("params/upstream_dump.json absent: no Rust toolchain / network on this image (parity unpinned)")
    G = o.generator()
    for name, k in (("2G", 2), ("3G", 3), ("8192G", 8192)):
        assert o.pt_mul(G, k) == (o._limbs_from_hex(d[name]["x"]), o._limbs_from_hex(d[name]["y"]))
    assert bytes(o.compress(G)) == bytes.fromhex(d["generator"]["compressed"])
    for case in d["rescue"]:
        if case["input_len"] in (1, 7, 8, 9):
            assert o.digest_to_bytes(o.rescue_hash_field([1] * case["input_len"])) == bytes.fromhex(case["digest"])
    for s_ in d["signatures"]:
        sk = int.from_bytes(bytes.fromhex(s_["private_key"]), "little")
        pk = o.pt_mul(G, sk)
        assert pk == (o._limbs_from_hex(s_["public_key"]["x"]), o._limbs_from_hex(s_["public_key"]["y"]))
        sig = bytes.fromhex(s_["signature"])
        assert o.verify(sig[:49], int.from_bytes(sig[49:], "little"), bytes.fromhex(s_["msg"]), pk) == 0


def test_optimised_cpu_port_agrees_with_the_textbook_oracle():
    """oracle/cfast.c (windowed NAFs, Straus-Shamir with a base-point table, lazy reduction: the CPU arm of bench.py)
    returns the verdicts of oracle/cref.c on valid, faulty, malformed and adversarial inputs."""
    import cref
    from util import KAT96, make_workload, pt_to96
    n = 260
    w = make_workload(303, n, lens=[int(x) for x in np.random.default_rng(303).integers(0, 170, n)])
    sigs, pk, inf = w["sigs"].copy(), w["pk"].copy(), w["inf"].copy()
    order = o.COFACTOR * o.Q
    kat = (o.KAT_X, o.KAT_Y)
    sigs[5, 49] ^= 1                                   # wrong e
    pk[9] = KAT96                                      # off-subgroup key
    sigs[11, :48] = 0; sigs[11, 48] = 0x80             # x := identity encoding
    sigs[13, 8:16] = 0xFF                              # non-canonical limb
    sigs[17, 49 + 31] = 0xFF                           # e >= q
    inf[19] = 1                                        # identity key
    sigs[23, 49:] = 0                                  # e := 0
    pk[29] = pt_to96(o.pt_mul(kat, order // 2))        # order 2
    pk[31] = pt_to96(o.pt_mul(kat, order // 29))       # order 29
    pk[37] = pt_to96(o.pt_add(o.generator(), o.pt_mul(kat, order // 2)))   # order 2q
    pk[41, 8:16] = 0xFF                                # non-canonical key limb
    pk[43], pk[44] = w["pk"][44].copy(), w["pk"][43].copy()
    nt = cref.default_threads()
    a = cref.verify_many(sigs, pk, inf, w["blob"], w["off"], nt)
    b = cref.verify_many_fast(sigs, pk, inf, w["blob"], w["off"], nt)
    assert np.array_equal(a, b)
    assert {0, 1, 2, 3} == set(int(v) for v in a)
