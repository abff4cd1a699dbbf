"""Full-size parity against the ORACLE itself (oracle/cref.c), not against a rule-built expectation:

 * BASELINE configs[2]: every one of the 2^20 verdicts and challenge digests of the bench workload;
 * BASELINE configs[1]: hash digests on random samples at 2^16 / 2^20 / 2^22 messages of 8 / 80 / 160 bytes;
 * BASELINE configs[3]: both points of a 2^16-signature batch;
 * the pipelined host path (several chunks) with adversarial keys on the chunk boundaries;
 * BASELINE configs[4] in miniature: the multi-device context (schnorr_b200_create_multi) against the single-device
   engine and the oracle, injected invalid signature included.

Reference behaviour restated by the oracle: src/signature.rs:181-205,274-306, src/batch.rs:84-130."""
import os
import subprocess

import numpy as np
import pytest

import cref
import pyref as o
from util import KAT96, make_workload, pt_to96, rand_fp

pytestmark = pytest.mark.gpu

P = np.uint64(0xFFFFFFFF00000001)


def _engine():
    import schnorr_sig_b200 as s
    return s, s.default_engine(0)


def test_bench_workload_2_20_every_verdict_and_digest_equals_the_oracle():
    """north_star: "bit-exact with the reference on 2^20 synthetic signatures" -- the bench workload (device-signed,
    8-byte messages, 1/1024 corrupted in five ways) through the pipelined HOST entry point (k_verify_fast in chunks),
    all 2^20 verdicts against cref.verify_many, all 2^20 challenge digests against cref.hash_messages."""
    s, eng = _engine()
    n = 1 << 20
    w = s.synth.signed_workload(eng, s.synth.DEFAULT_SEED, n, msg_len=8)
    f = s.synth.inject_faults(w, every=1024)
    got = eng.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"])
    assert eng.last_exact_count() == 0
    nt = cref.default_threads()
    want = cref.verify_many(f["sigs"], f["pk"], f["inf"], f["blob"], f["off"], nt)
    assert np.array_equal(got, want)
    assert np.array_equal(want, f["expect"])
    assert int((want == 2).sum()) > 600 and int((want == 1).sum()) > 150
    # digests: items whose sig.x was replaced by the identity encoding still hash (x = 0 is canonical)
    rx = np.ascontiguousarray(f["sigs"][:, :48])
    d = eng.hash_messages(rx, f["pk"], f["blob"], f["off"])
    cd = cref.hash_messages(rx, f["pk"], f["blob"], f["off"], nt)
    assert np.array_equal(d, cd)


@pytest.mark.parametrize("log2n", [16, 20, 22])
def test_hash_sweep_digests_on_random_samples(log2n):
    """BASELINE configs[1] (hash_message sweep): uniform field elements for R.x / P, messages of 8, 80 and 160 bytes;
    the device hashes all 2^log2n messages, 4096 random indices are compared with the oracle."""
    import torch
    s, eng = _engine()
    n = 1 << log2n
    rng = np.random.default_rng(100 + log2n)
    dev = torch.device("cuda", 0)
    rx = rand_fp(rng, 6 * n).view(np.uint8).reshape(n, 48)
    pk = rand_fp(rng, 12 * n).view(np.uint8).reshape(n, 96)
    d_rx, d_pk = torch.from_numpy(rx).to(dev), torch.from_numpy(pk).to(dev)
    idx = np.sort(rng.choice(n, 4096, replace=False))
    for L in (8, 80, 160):
        blob = rng.integers(0, 256, n * L, dtype=np.uint8)
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
        d_blob, d_off = torch.from_numpy(blob).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
        d_out = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
        eng.hash_messages_dev(n, d_rx, d_pk, d_blob, d_off, d_out)
        eng.synchronize()
        got = d_out.cpu().numpy()[idx]
        sub_blob = blob.reshape(n, L)[idx].reshape(-1)
        sub_off = np.arange(len(idx) + 1, dtype=np.uint64) * np.uint64(L)
        want = cref.hash_messages(rx[idx].copy(), pk[idx].copy(), sub_blob, sub_off, cref.default_threads())
        assert np.array_equal(got, want), (log2n, L)
        del d_blob, d_off, d_out


def test_batch_2_16_points_equal_the_oracle():
    """BASELINE configs[3]: 2^16 signatures, caller-supplied randomisers: lhs = sum s_i R_i - sum s_i h_i P_i and
    rhs = (sum s_i e_i) G as affine points against cref.verify_batch; one corrupted signature -> Err."""
    s, eng = _engine()
    n = 1 << 16
    w = s.synth.signed_workload(eng, 0xBA7C4, n, msg_len=80)
    nt = cref.default_threads()
    v, lhs, rhs = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], nt)
    assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
    bad = w["sigs"].copy()
    bad[n // 3, 49] ^= 1
    v, lhs, rhs = eng.verify_batch(bad, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    cv, cl, cr = cref.verify_batch(bad, w["pk"], w["inf"], w["blob"], w["off"], w["rand"], nt)
    assert v == cv == 2 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)


def _adversarial_keys():
    kat = (o.KAT_X, o.KAT_Y)
    n = o.COFACTOR * o.Q
    G = o.generator()
    pts = [o.pt_mul(kat, n // 2), o.pt_mul(kat, n // 10), o.pt_mul(kat, n // 29), o.pt_add(G, o.pt_mul(kat, n // 2)),
           o.pt_mul(kat, o.Q), kat]
    return [pt_to96(p) for p in pts if p is not o.INF]


def test_pipelined_chunks_with_adversarial_keys_on_the_boundaries():
    """The host entry point cuts a call of more than three kernel waves into 1 + 3 + rest waves; small-order,
    mixed-order, off-subgroup and identity keys and malformed records sit on both sides of every chunk boundary
    (they take the hand-back path to the exact kernel, per chunk work lists)."""
    s, eng = _engine()
    wave = 148 * 2 * 128
    n = 4 * wave + 1000
    w = s.synth.signed_workload(eng, 0xC4A1, n, msg_len=8)
    sigs, pk, inf = w["sigs"].copy(), w["pk"].copy(), w["inf"].copy()
    adv = _adversarial_keys()
    spots = []
    for b in (wave, 4 * wave, n - 1, 0, 2 * wave):
        for d in range(-3, 4):
            if 0 <= b + d < n:
                spots.append(b + d)
    for k, i in enumerate(spots):
        kind = k % (len(adv) + 3)
        if kind < len(adv):
            pk[i] = adv[kind]
        elif kind == len(adv):
            inf[i] = 1                     # identity key
        elif kind == len(adv) + 1:
            sigs[i, 8:16] = 0xFF           # non-canonical limb of sig.x -> 3
        else:
            sigs[i, 49 + 31] = 0xFF        # e >= q -> 3
    got = eng.verify_many(sigs, pk, inf, w["blob"], w["off"])
    assert eng.last_exact_count() >= 5
    want = cref.verify_many(sigs, pk, inf, w["blob"], w["off"], cref.default_threads())
    assert np.array_equal(got, want)
    assert {0, 1, 3} <= set(int(v) for v in np.unique(want))


@pytest.mark.parametrize("devices", [None, 2])
def test_offset_table_that_goes_bad_in_a_late_chunk_is_refused(devices):
    """The C ABI validates the message offsets chunk by chunk inside the pipeline (behind the kernels of the chunks
    before): a table that decreases, or runs past its last entry, anywhere -- first chunk, last chunk, a shard of a
    multi-device context -- ends the call with EARG, never with an out-of-range read, and the context stays usable."""
    s, eng0 = _engine()
    eng = eng0 if devices is None else s.Engine(_device_list(devices))
    wave = 148 * 2 * 128
    n = 4 * wave + 1000
    w = s.synth.signed_workload(eng0, 0xBAD0, n, msg_len=8)
    sigs, pk, inf, blob = w["sigs"], w["pk"], w["inf"], w["blob"]
    good = np.ascontiguousarray(w["off"], dtype=np.uint64)
    out = np.full(n, 255, dtype=np.uint8)
    for where in (5, wave + 7, n - 3):
        for kind in ("decreasing", "beyond_the_end"):
            off = good.copy()
            if kind == "decreasing":
                off[where] = off[where - 1] - 1 if off[where - 1] else off[where + 1] + 1
            else:
                off[where] = off[-1] + 64   # non-decreasing up to here, then back down: also caught as a decrease
            with pytest.raises(s.EngineError, match="non-decreasing"):
                eng.verify_many_raw(n, sigs, pk, inf, blob, off, out)
    off = good.copy()
    off[0] = 1
    with pytest.raises(s.EngineError, match="non-decreasing"):
        eng.verify_many_raw(n, sigs, pk, inf, blob, off, out)
    eng.verify_many_raw(n, sigs, pk, inf, blob, good, out)       # and the context still works
    assert int(out.max()) == 0


def _device_list(want):
    import torch
    c = torch.cuda.device_count()
    return [k % c for k in range(want)]


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_multi_device_context_equals_single_device_and_oracle(shards):
    """schnorr_b200_create_multi: verdicts, digests and both batch points of the sharded host entry points equal the
    single-device engine and the oracle; an invalid signature injected into the LAST shard turns the batch into Err.
    On a box with fewer GPUs than shards the device list wraps around (independent contexts on the same GPU)."""
    s, eng = _engine()
    multi = s.Engine(_device_list(shards))
    try:
        assert multi.device_count == shards
        n = 4096 * shards + 777                       # every shard gets a slice (multi.cuh: >= 4096 per shard)
        lens = [int(x) for x in np.random.default_rng(shards).integers(0, 40, n)]
        w = s.synth.host_inputs(0x5EED + shards, n, 8)
        rng = np.random.default_rng(7 + shards)
        msgs = [bytes(rng.integers(0, 256, L, dtype=np.uint8)) for L in lens]
        from util import pack_msgs
        blob, off = pack_msgs(msgs)
        pk, inf = multi.keygen(w["sk"])
        sigs = multi.sign_many(w["sk"], pk, inf, blob, off, w["nonce"])
        pk1, inf1 = eng.keygen(w["sk"])
        assert np.array_equal(pk, pk1) and np.array_equal(inf, inf1)
        assert np.array_equal(sigs, eng.sign_many(w["sk"], pk, inf, blob, off, w["nonce"]))
        bad = sigs.copy()
        bad[::997, 49] ^= 1
        bad[n - 5, 3] ^= 2
        pkb = pk.copy()
        pkb[4095:4098] = KAT96                        # off-subgroup keys across the first slice boundary
        got = multi.verify_many(bad, pkb, inf, blob, off)
        assert np.array_equal(got, eng.verify_many(bad, pkb, inf, blob, off))
        sub = np.concatenate([np.arange(0, 600), np.arange(3900, 4300), np.arange(n - 400, n)])
        sblob, soff = pack_msgs([msgs[i] for i in sub])
        assert np.array_equal(got[sub], cref.verify_many(bad[sub], pkb[sub], inf[sub], sblob, soff, cref.default_threads()))
        rx = np.ascontiguousarray(sigs[:, :48])
        assert np.array_equal(multi.hash_messages(rx, pk, blob, off), eng.hash_messages(rx, pk, blob, off))
        # batch: n <= 2^12 against the oracle, the full call against the single-device engine
        m = 4096
        mblob, moff = pack_msgs(msgs[:m])
        for e in (multi, eng):
            e.set_dist_threshold(0)                   # the per-thread prepare path on every shard
        try:
            v, lhs, rhs = multi.verify_batch(sigs, pk, inf, blob, off, w["rand"])
            v1, l1, r1 = eng.verify_batch(sigs, pk, inf, blob, off, w["rand"])
            assert v == v1 == 0 and np.array_equal(lhs, l1) and np.array_equal(rhs, r1)
            last = sigs.copy()
            last[n - 3, 49] ^= 1                      # invalid signature on the last shard
            assert multi.verify_batch(last, pk, inf, blob, off, w["rand"])[0] == 2
        finally:
            for e in (multi, eng):
                e.set_dist_threshold(10240)
        small = s.Engine(_device_list(2))             # 2 shards x 2048 would stay on one shard: force the split below
        try:
            v, lhs, rhs = multi.verify_batch(sigs[:m], pk[:m], inf[:m], mblob, moff, w["rand"][:m])
            cv, cl, cr = cref.verify_batch(sigs[:m], pk[:m], inf[:m], mblob, moff, w["rand"][:m], cref.default_threads())
            assert v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
            v2, l2, r2 = small.verify_batch(sigs[:2 * m], pk[:2 * m], inf[:2 * m], *pack_msgs(msgs[:2 * m]), w["rand"][:2 * m])
            c2 = cref.verify_batch(sigs[:2 * m], pk[:2 * m], inf[:2 * m], *pack_msgs(msgs[:2 * m]), w["rand"][:2 * m],
                                   cref.default_threads())
            assert v2 == c2[0] == 0 and np.array_equal(l2, c2[1]) and np.array_equal(r2, c2[2])
        finally:
            small.close()
    finally:
        multi.close()


def test_multi_device_cpp_example_runs():
    """examples/multi_device_example.cpp drives two (and five) device contexts through the plain C ABI."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "multi_device_example")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    for args in (["20000", "2"], ["33000", "5"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
