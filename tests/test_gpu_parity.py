"""Parity tests proper: the CUDA path through the C ABI against the oracle on the same seeded inputs.
Bit-exact (integer/byte work).  Run on the B200 box with `-m gpu`."""
import numpy as np
import pytest

import cref
import pyref as o
from util import KAT96, make_workload, pack_msgs, pt_from96, pt_to96, rand_fp, rand_scalars, int_le

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import schnorr_sig_b200 as s
    return s.default_engine(0)


def test_library_is_loaded_from_tree(eng):
    import schnorr_sig_b200 as s
    assert s._lib.library_path().endswith("schnorr-sig_b200/csrc/libschnorr_b200.so")
    assert eng.launch_count >= 2          # the G table was built by kernels at context creation


# ---- K1: hash_message -----------------------------------------------------------------------
@pytest.mark.parametrize("n,lens", [(1, None), (33, None), (2048, None), (300, "ragged")])
def test_hash_messages_bit_exact(eng, n, lens):
    rng = np.random.default_rng(100 + n)
    if lens == "ragged":
        L = [int(x) for x in rng.integers(0, 200, n)]
        L[:8] = [0, 1, 6, 7, 8, 13, 14, 160]
    else:
        L = [8] * n
    msgs = [bytes(rng.integers(0, 256, l, dtype=np.uint8)) for l in L]
    blob, off = pack_msgs(msgs)
    rx = rand_fp(rng, n * 6).view(np.uint8).reshape(n, 48)
    pk = rand_fp(rng, n * 12).view(np.uint8).reshape(n, 96)
    want = cref.hash_messages(rx, pk, blob, off, cref.default_threads())
    assert np.array_equal(eng.hash_messages(rx, pk, blob, off), want)       # k_hash_dist (small calls)
    eng.set_dist_threshold(0)
    try:
        assert np.array_equal(eng.hash_messages(rx, pk, blob, off), want)   # k_hash (one message per thread)
    finally:
        eng.set_dist_threshold(10240)


def test_hash_padding_collision_preserved(eng):
    """m (7k-1 bytes) and m||0x01 (7k bytes) hash alike (src/signature.rs:292-299)."""
    rng = np.random.default_rng(5)
    rx = rand_fp(rng, 6).view(np.uint8).reshape(1, 48).repeat(2, 0)
    pk = rand_fp(rng, 12).view(np.uint8).reshape(1, 96).repeat(2, 0)
    blob, off = pack_msgs([b"abcdef", b"abcdef\x01"])
    d = eng.hash_messages(rx, pk, blob, off)
    assert bytes(d[0]) == bytes(d[1])


# ---- K5: key generation and the device signer -----------------------------------------------
def test_keygen_and_sign_bit_exact(eng):
    w = make_workload(7, 257, lens=[int(x) for x in np.random.default_rng(7).integers(0, 90, 257)])
    pk, inf = eng.keygen(w["sk"])
    assert np.array_equal(pk, w["pk"]) and not inf.any()
    sigs = eng.sign_many(w["sk"], pk, inf, w["blob"], w["off"], w["nonce"])
    assert np.array_equal(sigs, w["sigs"])
    # scalar edge cases: 0 -> identity, 1 -> G, q-1 -> -G, non-canonical input is reduced
    sk = np.zeros((4, 32), dtype=np.uint8)
    sk[1, 0] = 1
    sk[2] = np.frombuffer((o.Q - 1).to_bytes(32, "little"), dtype=np.uint8)
    sk[3] = 0xFF
    pk, inf = eng.keygen(sk)
    cpk, cinf = cref.keygen(sk)
    assert list(inf) == [1, 0, 0, 0] and list(cinf) == [1, 0, 0, 0]
    assert np.array_equal(pk[1:], cpk[1:])
    assert pt_from96(pk[1]) == o.generator() and pt_from96(pk[2]) == o.pt_neg(o.generator())


# ---- K2: Signature::verify ------------------------------------------------------------------
def test_verify_valid_and_reference_negative_cases(eng):
    """The reference's own accept/reject cases (src/signature.rs:333-426, src/error.rs:66-82)."""
    lens = [160, 8, 8, 32, 0, 7, 80, 3, 24, 48]
    w = make_workload(31, len(lens), lens=lens)
    sigs, pk, inf, blob, off = w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]
    assert list(eng.verify_many(sigs, pk, inf, blob, off)) == [0] * len(lens)
    msgs = list(w["msgs"]); msgs[0] = bytes([msgs[0][0] ^ 42]) + msgs[0][1:]
    b2, o2 = pack_msgs(msgs)
    assert list(eng.verify_many(sigs, pk, inf, b2, o2))[:2] == [2, 0]
    pk2 = pk.copy(); pk2[0] = pt_to96(o.generator()); pk2[1] = KAT96
    assert list(eng.verify_many(sigs, pk2, inf, blob, off))[:3] == [2, 1, 0]
    s2 = sigs.copy(); s2[0, :48] = 0; s2[0, 48] = 0x80; s2[1, 49:] = 0
    assert list(eng.verify_many(s2, pk, inf, blob, off))[:3] == [2, 2, 0]
    s3 = sigs.copy(); s3[0, 49:] = 0xFF; s3[1, :8] = 0xFF; s3[2, 40:48] = 0xFF
    pk3 = pk.copy(); pk3[2] = KAT96; pk3[3, 8:16] = 0xFF
    got = eng.verify_many(s3, pk3, inf, blob, off)
    assert np.array_equal(got, cref.verify_many(s3, pk3, inf, blob, off))
    assert list(got[:4]) == [3, 3, 1, 3]      # off-subgroup key wins over a malformed x (src/signature.rs:182-186)
    s4 = sigs.copy(); s4[:, 48] ^= 0x40       # flag byte ignored by single verify
    assert list(eng.verify_many(s4, pk, inf, blob, off)) == [0] * len(lens)


def test_verify_adversarial_public_keys(eng):
    """Small-order / mixed-order / identity keys exercise the exceptional group-law branches."""
    w = make_workload(32, 8)
    kat = (o.KAT_X, o.KAT_Y)
    n = o.COFACTOR * o.Q
    pk, inf = w["pk"].copy(), w["inf"].copy()
    pk[0] = pt_to96(o.pt_mul(kat, n // 2))                                     # order 2 (y = 0)
    pk[1] = pt_to96(o.pt_mul(kat, n // 29))                                    # order 29
    pk[2] = pt_to96(o.pt_mul(kat, n // 58))                                    # order 58
    pk[3] = pt_to96(o.pt_add(o.generator(), o.pt_mul(kat, n // 2)))            # order 2q
    pk[4] = pt_to96(o.pt_mul(kat, o.Q))                                        # order = cofactor
    inf[5] = 1                                                                 # identity key
    assert all(pk[i].any() for i in range(5))                                  # none of them is the identity
    got = eng.verify_many(w["sigs"], pk, inf, w["blob"], w["off"])
    want = cref.verify_many(w["sigs"], pk, inf, w["blob"], w["off"], 4)
    assert np.array_equal(got, want)
    assert list(want[:6]) == [1, 1, 1, 1, 1, 2] and list(want[6:]) == [0, 0]
    # a signature that verifies under the identity key: R = e*G, h arbitrary
    sk0 = np.zeros((1, 32), dtype=np.uint8)
    blob, off = pack_msgs([b"identity"])
    sig = cref.sign_many(sk0, np.zeros((1, 96), np.uint8), np.ones(1, np.uint8), blob, off, w["nonce"][:1])
    assert list(cref.verify_many(sig, np.zeros((1, 96), np.uint8), np.ones(1, np.uint8), blob, off)) == [0]
    assert list(eng.verify_many(sig, np.zeros((1, 96), np.uint8), np.ones(1, np.uint8), blob, off)) == [0]


@pytest.mark.parametrize("n", [1, 127, 129, 1000])
def test_verify_many_matches_oracle_with_injected_faults(eng, n):
    import schnorr_sig_b200 as s
    w = make_workload(50 + n, n, lens=[int(x) for x in np.random.default_rng(n).integers(1, 40, n)])
    w["n"], w["msg_len"] = n, 1
    f = s.synth.inject_faults(w, every=16)
    got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], cref.default_threads())
    assert np.array_equal(got, want)
    assert np.array_equal(got, f["expect"])


def test_affine_fast_path_agrees_with_exact_kernel_and_hands_back_exceptional_items(eng):
    """k_verify_fast (affine.cuh) + exact pass == k_verify alone == oracle; the hand-back list is small for honest
    inputs and contains the adversarial keys."""
    import schnorr_sig_b200 as s
    n = 1500
    w = make_workload(77, n, lens=[int(x) for x in np.random.default_rng(77).integers(0, 30, n)])
    w["n"], w["msg_len"] = n, 1
    f = s.synth.inject_faults(w, every=16)
    kat = (o.KAT_X, o.KAT_Y)
    order = o.COFACTOR * o.Q
    f["pk"][3] = pt_to96(o.pt_mul(kat, order // 2))          # order 2: doubling of a 2-torsion point
    f["pk"][5] = pt_to96(o.pt_mul(kat, order // 29))         # order 29: P + P / P - P inside the buckets
    f["inf"][7] = 1                                          # identity key
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], cref.default_threads())
    got, handed_back = {}, {}
    try:
        for name, thr in (("fast", 0), ("dist", 2**62)):      # one signature per thread / per six lanes
            eng.set_dist_threshold(thr)
            got[name] = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
            handed_back[name] = eng.last_exact_count()
        eng.set_exact_only(True)
        got["exact"] = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
        assert eng.last_exact_count() == 0
    finally:
        eng.set_exact_only(False)
        eng.set_dist_threshold(10240)
    for name in ("fast", "dist", "exact"):
        assert np.array_equal(got[name], want), name
        assert not (got[name] == 0xFF).any()
    for name in ("fast", "dist"):                              # the three adversarial keys + ~0.1 % of honest inputs
        assert 3 <= handed_back[name] <= 3 + n // 100, (name, handed_back)


ONE_DEFAULT = 512   # schnorr_b200_set_one_threshold default (include/schnorr_b200.h)


@pytest.mark.parametrize("n", [1, 2, 5, 21, 64, 700])
def test_block_per_signature_kernel_matches_oracle(eng, n):
    """k_verify_one (one thread block per signature, one.cuh: doubling chain, challenge, e*G and the sixteen bucket
    accumulations side by side) forced for every size: ragged messages incl. empty ones, injected faults, small-order /
    mixed-order / identity keys, malformed records -> the oracle's verdicts, and the same verdicts as the six-lane and the
    per-thread kernels."""
    import schnorr_sig_b200 as s
    rng = np.random.default_rng(1900 + n)
    lens = [int(x) for x in rng.integers(0, 60, n)]
    lens[:min(n, 8)] = [8, 0, 1, 6, 7, 13, 14, 49][:min(n, 8)]
    w = make_workload(1900 + n, n, lens=lens)
    w["n"], w["msg_len"] = n, 1
    f = s.synth.inject_faults(w, every=7) if n >= 7 else dict(w, expect=np.zeros(n, np.uint8))
    adversarial = 0
    if n >= 21:
        kat = (o.KAT_X, o.KAT_Y)
        order = o.COFACTOR * o.Q
        f["pk"][2] = pt_to96(o.pt_mul(kat, order // 2))      # order 2
        f["pk"][9] = pt_to96(o.pt_mul(kat, order // 29))     # order 29: P + P / P - P inside the buckets
        f["pk"][10] = pt_to96(o.pt_add(o.generator(), o.pt_mul(kat, order // 2)))   # order 2q
        f["pk"][12] = KAT96                                    # the reference's off-subgroup point -> InvalidPublicKey
        f["inf"][11] = 1                                     # identity key
        f["sigs"][13, 8:16] = 0xFF                           # non-canonical sig.x limb
        f["sigs"][15, 49:] = 0xFF                            # e >= q
        adversarial = 4
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], cref.default_threads())
    eng.set_one_threshold(2**62)
    try:
        got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
        handed_back = eng.last_exact_count()
    finally:
        eng.set_one_threshold(ONE_DEFAULT)
    assert np.array_equal(got, want)
    assert not (got == 0xFF).any()
    assert handed_back <= adversarial
    eng.set_one_threshold(0)
    try:
        assert np.array_equal(eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"]), want)
    finally:
        eng.set_one_threshold(ONE_DEFAULT)


@pytest.mark.parametrize("n", [1, 4, 5, 6, 21, 127, 1000])
def test_warp_cooperative_kernel_matches_oracle(eng, n):
    """k_verify_dist (one signature per six lanes, dist.cuh) forced for every size: ragged messages (different hash
    trip counts inside one warp), injected faults, adversarial keys -> the oracle's verdicts."""
    import schnorr_sig_b200 as s
    rng = np.random.default_rng(900 + n)
    lens = [int(x) for x in rng.integers(0, 60, n)]
    lens[:min(n, 8)] = [0, 1, 6, 7, 8, 13, 14, 49][:min(n, 8)]
    w = make_workload(900 + n, n, lens=lens)
    w["n"], w["msg_len"] = n, 1
    f = s.synth.inject_faults(w, every=7) if n >= 7 else dict(w, expect=np.zeros(n, np.uint8))
    if n >= 21:
        kat = (o.KAT_X, o.KAT_Y)
        order = o.COFACTOR * o.Q
        f["pk"][2] = pt_to96(o.pt_mul(kat, order // 2))
        f["pk"][9] = pt_to96(o.pt_mul(kat, order // 29))
        f["inf"][11] = 1
        f["sigs"][13, 8:16] = 0xFF            # non-canonical sig.x limb
        f["sigs"][15, 49:] = 0xFF             # e >= q
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], cref.default_threads())
    eng.set_dist_threshold(2**62)
    eng.set_one_threshold(0)              # (small calls would otherwise take the block-per-signature kernel)
    try:
        got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
        handed_back = eng.last_exact_count()
    finally:
        eng.set_dist_threshold(10240)
        eng.set_one_threshold(ONE_DEFAULT)
    assert np.array_equal(got, want)
    assert handed_back <= (3 if n >= 21 else 0)
    eng.set_dist_threshold(0)
    eng.set_one_threshold(0)              # one signature per thread for every size
    try:
        assert np.array_equal(eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"]), want)
    finally:
        eng.set_dist_threshold(10240)
        eng.set_one_threshold(ONE_DEFAULT)


def test_verify_empty_and_argument_errors(eng):
    z = np.zeros((0, 81), np.uint8)
    assert eng.verify_many(z, np.zeros((0, 96), np.uint8), None, np.zeros(0, np.uint8), np.zeros(1, np.uint64)).size == 0
    w = make_workload(3, 2)
    with pytest.raises(AssertionError):
        eng.verify_many(w["sigs"], w["pk"][:1], None, w["blob"], w["off"])


# ---- wire codecs (f2) ---------------------------------------------------------------------------
def test_compress_decompress_round_trip(eng):
    w = make_workload(9, 64)
    comp = eng.compress(w["pk"], w["inf"])
    for i in range(64):
        assert bytes(comp[i]) == o.compress(pt_from96(w["pk"][i])) if i < 4 else True
        assert bytes(comp[i]) == bytes(cref.compress(w["pk"][i]))
    pk, inf, ok = eng.decompress(comp)
    assert ok.all() and not inf.any() and np.array_equal(pk, w["pk"])
    bad = np.zeros((4, 49), np.uint8)
    bad[1] = 0xFF
    bad[2] = comp[0]; bad[2, 48] = 0xFF
    bad[3, 48] = 0x80                                   # identity
    pk, inf, ok = eng.decompress(bad)
    assert list(ok) == [0, 0, 0, 1] and inf[3] == 1
    assert bytes(eng.compress(np.zeros((1, 96), np.uint8), np.ones(1, np.uint8))[0]) == bytes(48) + b"\x80"


# ---- K3/K4: batch verification ----------------------------------------------------------------
BATCH_SMALL_DEFAULT = 256   # schnorr_b200_set_batch_small_threshold default (include/schnorr_b200.h)


@pytest.mark.parametrize("n", [1, 3, 5, 64, 700])
def test_verify_batch_points_bit_exact(eng, n):
    lens = [24, 24, 48] if n == 3 else [int(x) for x in np.random.default_rng(n).integers(0, 90, n)]
    w = make_workload(60 + n, n, lens=lens)
    if n >= 5:   # repeated signers (src/batch.rs:167-169)
        w["sk"][3] = w["sk"][0]; w["sk"][4] = w["sk"][0]
        w["pk"], w["inf"] = cref.keygen(w["sk"], 4)
        w["sigs"] = cref.sign_many(w["sk"], w["pk"], w["inf"], w["blob"], w["off"], w["nonce"], 4)
    v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], cref.default_threads())
    assert v == cv == 0
    assert np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    # the other batch form: Pippenger pipeline for the small sizes (default: one thread block per signature up to 256),
    # block per signature for the large one
    eng.set_batch_small_threshold(0 if n <= BATCH_SMALL_DEFAULT else 2**62)
    try:
        v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
        assert v == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
        eng.set_batch_small_threshold(0)
        eng.set_dist_threshold(0)                     # challenges hashed inside k_batch_prepare (the large-batch form)
        v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    finally:
        eng.set_dist_threshold(10240)
        eng.set_batch_small_threshold(BATCH_SMALL_DEFAULT)
    assert v == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    if n >= 5:
        pk2 = w["pk"].copy(); pk2[[1, 2]] = pk2[[2, 1]]
        v, lhs, rhs = eng.verify_batch(w["sigs"], pk2, w["inf"], w["blob"], w["off"], w["rand"])
        cv, cl, cr = cref.verify_batch(w["sigs"], pk2, w["inf"], w["blob"], w["off"], w["rand"], cref.default_threads())
        assert v == cv == 2 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    s2 = w["sigs"].copy(); s2[0, 48] ^= 0x40          # y-sign flag matters here (src/batch.rs:104)
    s3 = w["sigs"].copy(); s3[0, :8] = 0xFF           # panic in the reference
    for thr in (BATCH_SMALL_DEFAULT, 0, 2**62):       # both batch forms
        eng.set_batch_small_threshold(thr)
        try:
            assert eng.verify_batch(s2, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])[0] == 2
            assert eng.verify_batch(s3, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])[0] == 3
        finally:
            eng.set_batch_small_threshold(BATCH_SMALL_DEFAULT)
    assert cref.verify_batch(s3, w["pk"], w["inf"], w["blob"], w["off"], w["rand"], 2)[0] == 3


def test_verify_batch_extreme_randomisers_and_empty(eng):
    w = make_workload(71, 6)
    rand = w["rand"].copy()
    rand[0] = 0                                                           # s = 0
    rand[1] = np.frombuffer((o.Q - 1).to_bytes(32, "little"), np.uint8)   # s = q-1
    rand[2] = 0xFF                                                        # non-canonical: reduced mod q
    rand[3] = 0; rand[3, 0] = 1                                           # s = 1
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand, 2)
    for thr in (BATCH_SMALL_DEFAULT, 0):              # block per signature, Pippenger pipeline
        eng.set_batch_small_threshold(thr)
        try:
            v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand)
        finally:
            eng.set_batch_small_threshold(BATCH_SMALL_DEFAULT)
        assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    v, lhs, rhs = eng.verify_batch(np.zeros((0, 81), np.uint8), np.zeros((0, 96), np.uint8), None,
                                   np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros((0, 32), np.uint8))
    assert v == 0 and lhs[96] == 1 and rhs[96] == 1                      # identity == identity


def test_batch_partials_compose(eng):
    """Multi-GPU form on one device: partials of two slices finish to the same points as one batch."""
    import torch
    w = make_workload(72, 40)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(w[k])).to(dev) for k in ("sigs", "pk", "inf", "rand")}
    parts = torch.zeros((2, 192), dtype=torch.uint8, device=dev)
    for r, (lo, hi) in enumerate(((0, 17), (17, 40))):
        blob, off = pack_msgs(w["msgs"][lo:hi])
        tb, to = torch.from_numpy(blob).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
        eng.batch_partial_dev(hi - lo, t["sigs"][lo:hi].contiguous(), t["pk"][lo:hi].contiguous(),
                              t["inf"][lo:hi].contiguous(), tb, to, t["rand"][lo:hi].contiguous(), parts[r])
        eng.synchronize()
    v, lhs, rhs = eng.batch_finish(parts.cpu().numpy())
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], 4)
    assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)


# ---- façade (reads like the reference's tests) -------------------------------------------------
def test_facade_sign_verify_like_reference():
    import schnorr_sig_b200 as s
    rng = s.OsRng()
    message = bytes(range(160))
    keypair = s.KeyPair.new(rng)
    skey, pkey = keypair.private_key, keypair.public_key
    signature = skey.sign(message, rng)
    assert signature.verify(message, pkey).is_ok()
    assert pkey.verify_signature(signature, message).is_ok()
    signature = keypair.sign(message, rng)
    assert keypair.verify_signature(signature, message).is_ok()
    assert keypair.sign_and_bind_pkey(message, rng).verify(message).is_ok()
    wrong = bytes([42]) + message[1:]
    assert signature.verify(wrong, pkey).is_err()
    wrong_pkey = s.PublicKey.from_raw_coordinates(o.KAT_X, o.KAT_Y)
    assert repr(signature.verify(message, wrong_pkey)) == "Err(InvalidPublicKey)"          # src/error.rs:66-70
    assert repr(signature.verify(wrong, pkey)) == "Err(InvalidSignature)"
    assert str(signature.verify(message, wrong_pkey).unwrap_err()) == "The public key is not an element of the prime subgroup."
    assert str(signature.verify(wrong, pkey).unwrap_err()) == "The signature is invalid or was incorrectly computed."
    # encodings (src/signature.rs:483-507, src/public.rs:104-112)
    b = signature.to_bytes()
    assert len(b) == s.SIGNATURE_LENGTH and s.Signature.from_bytes(b) == signature
    assert s.Signature.from_bytes(b"\xff" * 81) is None
    kb = pkey.to_bytes()
    assert len(kb) == s.PUBLIC_KEY_LENGTH and s.PublicKey.from_bytes(kb) == pkey
    assert s.PublicKey.from_bytes(bytes(49)) is None and s.PublicKey.from_bytes(b"\xff" * 49) is None
    ks = keypair.sign_and_bind_pkey(message, rng)
    assert len(ks.to_bytes()) == s.KEYED_SIGNATURE_LENGTH and s.KeyedSignature.from_bytes(ks.to_bytes()).verify(message).is_ok()


def test_facade_batch_like_reference():
    """src/batch.rs:138-179 and tests/schnorr.rs:148-183."""
    import schnorr_sig_b200 as s
    rng = s.OsRng()
    messages = [b"Message1", b"Message2", b"Message3", b"Message4", b"Message5"]
    keypairs, signatures = [], []
    for i, m in enumerate(messages):
        kp = s.KeyPair.new(rng)
        if i in (3, 4):
            kp = keypairs[0]
        signatures.append(kp.sign(m, rng))
        keypairs.append(kp)
    public_keys = [k.public_key for k in keypairs]
    assert s.verify_batch(signatures[:1], public_keys[:1], messages[:1], rng).is_ok()
    assert s.verify_batch(signatures, public_keys, messages, rng).is_ok()
    public_keys[1], public_keys[2] = public_keys[2], public_keys[1]
    assert s.verify_batch(signatures, public_keys, messages, rng).is_err()
    with pytest.raises(AssertionError):
        s.verify_batch(signatures, public_keys[:2], messages, rng)
