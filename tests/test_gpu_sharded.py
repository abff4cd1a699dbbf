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
