"""f3: failed-batch localisation on the device (schnorr_b200_locate_invalid), in BATCH semantics.

The reference returns one Err for a bad batch (src/batch.rs:125-129).  An item is a culprit when its own term of the
batch equation fails: R_i decompressed with its flag byte (src/batch.rs:104), no subgroup check on the key
(src/batch.rs:102-106).  The oracle for "item i is bad" is cref.verify_batch on the one-item batch {i}."""
import numpy as np
import pytest

import cref
from util import KAT96, make_workload

pytestmark = pytest.mark.gpu


def _oracle_item(w, sigs, pk, i):
    off = np.array([0, int(w["off"][i + 1] - w["off"][i])], dtype=np.uint64)
    blob = w["blob"][int(w["off"][i]):int(w["off"][i + 1])]
    return cref.verify_batch(sigs[i:i + 1], pk[i:i + 1], w["inf"][i:i + 1], blob, off, w["rand"][i:i + 1], 1)[0]


def test_flipped_y_flag_is_a_batch_only_failure_and_is_located():
    """flip bit 6 of sig.x[48] -- src/batch.rs:104 uses it, src/signature.rs:186 ignores it: every single
    verification accepts, the batch fails, and the localisation names exactly that item."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 6000
    w = make_workload(31, n, lens=[int(x) for x in np.random.default_rng(31).integers(0, 30, n)])
    assert eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])[0] == 0
    assert not eng.locate_invalid(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"]).any()
    bad = w["sigs"].copy()
    bad[4321, 48] ^= 0x40
    assert (eng.verify_many(bad, w["pk"], w["inf"], w["blob"], w["off"]) == 0).all()          # x-only: still accepted
    assert eng.verify_batch(bad, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])[0] == 2
    flags = eng.locate_invalid(bad, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    assert list(np.nonzero(flags)[0]) == [4321] and flags[4321] == 2
    assert _oracle_item(w, bad, w["pk"], 4321) == 2 and _oracle_item(w, bad, w["pk"], 4320) == 0
    # facade
    sigs = [s.Signature(bytes(x[:49]), bytes(x[49:])) for x in bad[4300:4400]]
    pks = [s.PublicKey(bytes(k)) for k in w["pk"][4300:4400]]
    msgs = w["msgs"][4300:4400]
    assert s.verify_batch(sigs, pks, msgs).is_err()
    assert all(sig.verify(m, k).is_ok() for sig, k, m in list(zip(sigs, pks, msgs))[20:23])
    assert [i for i, _ in s.locate_invalid(sigs, pks, msgs)] == [21]


def test_scattered_culprits_of_every_kind_match_the_oracle_item_by_item():
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 40000
    w = make_workload(37, n, msg_len=8)
    sigs, pk = w["sigs"].copy(), w["pk"].copy()
    culprits = {}
    rng = np.random.default_rng(5)
    for k, i in enumerate(sorted(int(v) for v in rng.choice(n - 1, 60, replace=False))):
        kind = k % 5
        if kind == 0:
            sigs[i, 49 + 3] ^= 0x10                  # e changed
        elif kind == 1:
            sigs[i, 48] ^= 0x40                      # y-sign flag flipped
        elif kind == 2:
            pk[i] = KAT96                            # off-subgroup key: no subgroup check here, the equation just fails
        elif kind == 3:
            pk[i] = w["pk"][(i + 1) % n]             # somebody else's key
        else:
            sigs[i, 48] = 0x07                       # reserved flag bits set: from_compressed fails -> malformed
        culprits[i] = 3 if kind == 4 else 2
    flags = eng.locate_invalid(sigs, pk, w["inf"], w["blob"], w["off"], w["rand"])
    assert {int(i): int(flags[i]) for i in np.nonzero(flags)[0]} == culprits
    for i in list(culprits)[:15]:
        assert _oracle_item(w, sigs, pk, i) == culprits[i]
    for i in (0, 1, n - 1, 12345):
        if i not in culprits:
            assert _oracle_item(w, sigs, pk, i) == 0


@pytest.mark.parametrize("log2n,density", [(17, 0), (17, 1), (15, 40)])
def test_bisection_at_scale_and_degradation_to_the_per_item_pass(log2n, density):
    """One culprit among 2^17 (deep bisection), none at all, and a batch riddled with them (1 in 40: the suspects stop
    shrinking and everything goes through the per-item check)."""
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    n = 1 << log2n
    w = s.synth.signed_workload(eng, 0x10CA7E + log2n, n, msg_len=8)
    sigs = w["sigs"].copy()
    want = []
    if density == 1:
        want = [n // 3 + 17]
    elif density > 1:
        want = list(range(7, n, density))
    for i in want:
        sigs[i, 49] ^= 1
    flags = eng.locate_invalid(sigs, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    assert list(np.nonzero(flags)[0]) == want and (flags[want] == 2).all()
