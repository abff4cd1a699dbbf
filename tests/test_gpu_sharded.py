"""ShardedVerifier with the real engine (single process = world 1) and the full-size size-independent
properties: at 2^20 signatures the oracle is too slow, so the check is the injected-fault pattern
(every valid signature accepts, every corrupted one is rejected with the right code)."""
import numpy as np
import pytest

import cref
from util import make_workload

pytestmark = pytest.mark.gpu


def test_sharded_verifier_single_process_matches_oracle():
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    w = make_workload(5, 200, lens=[int(x) for x in np.random.default_rng(5).integers(0, 50, 200)])
    sv = s.ShardedVerifier(eng)
    got = sv.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])
    assert np.array_equal(got, cref.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], cref.default_threads()))
    v, lhs, rhs = sv.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], cref.default_threads())
    assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)


@pytest.mark.parametrize("log2n,msg_len", [(20, 8), (16, 80)])
def test_full_size_fault_pattern(log2n, msg_len):
    """BASELINE configs[2] size: 2^20 device-signed signatures, 1/1024 corrupted in five ways."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 1 << log2n
    w = s.synth.signed_workload(eng, s.synth.DEFAULT_SEED, n, msg_len=msg_len)
    # the device signer itself is spot-checked against the oracle (SURVEY.md §7 "hard parts")
    idx = np.random.default_rng(1).choice(n, 256, replace=False)
    off = (np.arange(257, dtype=np.uint64) * np.uint64(msg_len))
    blob = w["blob"].reshape(n, msg_len)[idx].reshape(-1)
    cpk, cinf = cref.keygen(w["sk"][idx], cref.default_threads())
    assert np.array_equal(cpk, w["pk"][idx])
    csig = cref.sign_many(w["sk"][idx], cpk, cinf, blob, off, w["nonce"][idx], cref.default_threads())
    assert np.array_equal(csig, w["sigs"][idx])
    f = s.synth.inject_faults(w, every=1024)
    got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
    assert np.array_equal(got, f["expect"])
    assert int((f["expect"] == 2).sum()) > 0 and int((f["expect"] == 1).sum()) > 0
    # batch at 2^16 (BASELINE configs[3]): all valid -> Ok, one corrupted -> Err, points equal the oracle's x
    nb = 1 << 16
    sl = slice(0, nb)
    boff = w["off"][:nb + 1]
    v, lhs, rhs = eng.verify_batch(w["sigs"][sl], w["pk"][sl], w["inf"][sl], w["blob"][:nb * msg_len], boff, w["rand"][sl])
    assert v == 0 and np.array_equal(lhs[:48], rhs[:48])
    bad = w["sigs"][sl].copy()
    bad[nb // 2, 49] ^= 1
    assert eng.verify_batch(bad, w["pk"][sl], w["inf"][sl], w["blob"][:nb * msg_len], boff, w["rand"][sl])[0] == 2


def test_cpp_facade_example_runs():
    """The C++ host facade (include/schnorr_b200.hpp) over the C ABI: sign, verify, batch, codecs."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "verify_example")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def test_keyed_signature_wire_records_and_localisation():
    """f2: 130-byte KeyedSignature records verified straight from the wire (key decompressed on device);
    f3: failed-batch localisation."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    w = make_workload(8, 50, lens=[int(x) for x in np.random.default_rng(8).integers(0, 60, 50)])
    comp = np.stack([cref.compress(w["pk"][i]) for i in range(50)])
    keyed = np.concatenate([comp, w["sigs"]], axis=1)
    assert keyed.shape == (50, 130)
    assert (eng.verify_keyed_many(keyed, w["blob"], w["off"]) == 0).all()
    bad = keyed.copy()
    bad[3, 48] = 0xFF            # undecodable key -> 3
    bad[4, 49 + 49:] = 0         # e := 0 -> 2
    bad[5, 48] ^= 0x40           # other root: -P, a different (valid) key -> 2
    got = eng.verify_keyed_many(bad, w["blob"], w["off"])
    assert list(got[3:6]) == [3, 2, 2] and (np.delete(got, [3, 4, 5]) == 0).all()
    # localisation through the facade
    sigs = [s.Signature(bytes(x[:49]), bytes(x[49:])) for x in w["sigs"]]
    pks = [s.PublicKey(bytes(k)) for k in w["pk"]]
    assert s.verify_batch(sigs, pks, w["msgs"]).is_ok() and s.locate_invalid(sigs, pks, w["msgs"]) == []
    pks[7], pks[9] = pks[9], pks[7]
    assert s.verify_batch(sigs, pks, w["msgs"]).is_err()
    loc = s.locate_invalid(sigs, pks, w["msgs"])
    assert [i for i, _ in loc] == [7, 9] and all(e == s.SignatureError(s.SignatureError.InvalidSignature) for _, e in loc)


def test_batch_with_skewed_randomisers_exercises_bucket_splitting():
    """All randomisers equal (and tiny): every s_i R_i term lands in the same bucket of every window, so
    single buckets hold thousands of points and are split across many segments (k_msm_segment_fixup)."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 3000
    w = make_workload(17, n, msg_len=8, nthreads=cref.default_threads())
    for val in (1, 0x0123456789ABCDEF):
        rand = np.zeros((n, 32), dtype=np.uint8)
        rand[:] = np.frombuffer(int(val).to_bytes(32, "little"), dtype=np.uint8)
        v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand)
        cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand, cref.default_threads())
        assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    # repeated signer with the SAME message and nonce: identical points with identical scalars (P + P inside a bucket)
    w2 = {k: (np.repeat(v[:1], 64, axis=0) if isinstance(v, np.ndarray) and v.ndim == 2 else v) for k, v in w.items()}
    blob = np.tile(w["blob"][:8], 64)
    off = np.arange(65, dtype=np.uint64) * np.uint64(8)
    rand = np.repeat(w["rand"][:1], 64, axis=0)
    cv, cl, cr = cref.verify_batch(w2["sigs"], w2["pk"], np.zeros(64, np.uint8), blob, off, rand, 4)
    for thr in (0, 256):      # Pippenger pipeline (P + P inside a bucket), block per signature (P + P in the final sum)
        eng.set_batch_small_threshold(thr)
        try:
            v, lhs, rhs = eng.verify_batch(w2["sigs"], w2["pk"], np.zeros(64, np.uint8), blob, off, rand)
        finally:
            eng.set_batch_small_threshold(256)
        assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)


def test_batch_with_narrow_top_window_uses_the_block_fixup():
    """Window widths whose top window has only 3 bits (c = 6, 12, 14: a handful of buckets hold an eighth of all
    points each and span hundreds of segments -> k_msm_fixup_long) give the same points as the oracle."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 6000
    w = make_workload(19, n, msg_len=8, nthreads=cref.default_threads())
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], cref.default_threads())
    assert cv == 0
    try:
        for c, t in ((6, 0), (12, 0), (14, 8), (9, 64)):
            eng.set_msm_geometry(c, t)        # test hook of the host planner (schnorr_b200_set_msm_geometry)
            v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
            assert v == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr), c
    finally:
        eng.set_msm_geometry(0, 0)


def test_mid_size_random_faults_against_oracle():
    """2^13 signatures with ragged messages and 1/16 injected faults: verdict vector equals the oracle's."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 1 << 13
    lens = [int(x) for x in np.random.default_rng(3).integers(1, 120, n)]
    w = make_workload(23, n, lens=lens, nthreads=cref.default_threads())
    w["n"], w["msg_len"] = n, 1
    f = s.synth.inject_faults(w, every=16)
    got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], cref.default_threads())
    assert np.array_equal(got, want) and np.array_equal(got, f["expect"])
    d = eng.hash_messages(f["sigs"][:, :48].copy(), f["pk"], f["blob"], f["off"])
    ok = f["sigs"][:, :48].view(np.uint64).max(axis=1) < np.uint64(0xFFFFFFFF00000001)
    cd = cref.hash_messages(f["sigs"][:, :48].copy(), f["pk"], f["blob"], f["off"], cref.default_threads())
    assert np.array_equal(d[ok], cd[ok])


def test_three_verification_kernels_agree_at_2_17():
    """k_verify_fast (one signature per thread), k_verify_dist (six lanes per signature) and the exact Jacobian
    k_verify return the same verdict vector on 2^17 device-signed signatures with ragged messages and 1/64 faults."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 1 << 17
    w = s.synth.signed_workload(eng, 0xC0FFEE, n, msg_len=8)
    f = s.synth.inject_faults(w, every=64)
    got = {}
    try:
        eng.set_dist_threshold(0)
        got["fast"] = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
        assert eng.last_exact_count() == 0
        eng.set_dist_threshold(2**62)
        got["dist"] = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
        assert eng.last_exact_count() == 0
        eng.set_exact_only(True)
        got["exact"] = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
    finally:
        eng.set_exact_only(False)
        eng.set_dist_threshold(10240)
    assert np.array_equal(got["fast"], f["expect"])
    assert np.array_equal(got["dist"], f["expect"])
    assert np.array_equal(got["exact"], f["expect"])
