"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes drive
schnorr_sig_b200.distributed.ShardedVerifier with an oracle-backed worker (no GPU needed)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    import schnorr_sig_b200 as s
    for n in (0, 1, 2, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = s.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi)) if n <= 1000 else []
                assert hi - lo in (n // world, n // world + 1)
            if n <= 1000:
                assert seen == list(range(n))
            assert s.shard_bounds(n, world - 1, world)[1] == n
    with pytest.raises(ValueError):
        s.shard_bounds(4, 2, 2)


def test_synth_shards_are_distinct_and_reproducible():
    import schnorr_sig_b200 as s
    a = s.synth.host_inputs(7, 64, 8, shard=0)
    b = s.synth.host_inputs(7, 64, 8, shard=1)
    a2 = s.synth.host_inputs(7, 64, 8, shard=0)
    assert not np.array_equal(a["sk"], b["sk"]) and np.array_equal(a["sk"], a2["sk"])
    assert (a["sk"][:, 31] < 0x40).all() and a["off"][-1] == 64 * 8


class OracleWorker:
    """Stands in for the Engine in the CPU test: same methods, computed by the oracle."""

    def verify_many(self, sigs, pk, inf, blob, off):
        import cref
        return cref.verify_many(sigs, pk, inf, blob, off, 1)

    def batch_partial(self, sigs, pk, inf, blob, off, rand):
        import cref
        import pyref as o
        from util import int_le, pt_from96
        n = sigs.shape[0]
        lin = 0
        acc = o.INF
        bad = 0
        for i in range(n):
            x49 = bytes(sigs[i, :49]); e = int_le(sigs[i, 49:]); s = int_le(rand[i]) % o.Q
            msg = bytes(blob[int(off[i]):int(off[i + 1])])
            P = pt_from96(pk[i], inf[i] if inf is not None else 0)
            ok, R = o.decompress(x49)
            if not ok or e >= o.Q:
                bad = 1
                break
            h = o.scalar_from_digest(o.hash_message(R[0] if R is not o.INF else o.F6_ZERO, P, msg))
            lin = (lin + s * e) % o.Q
            acc = o.pt_add(acc, o.pt_mul(R, s))
            acc = o.pt_add(acc, o.pt_mul(o.pt_neg(P), h * s % o.Q))
        out = np.zeros(24, dtype=np.uint64)
        if acc is o.INF:
            out[0] = 1; out[6] = 1
        else:
            out[0:6] = acc[0]; out[6:12] = acc[1]; out[12] = 1
        for k in range(4):
            out[18 + k] = (lin >> (64 * k)) & (2**64 - 1)
        out[22] = bad
        return out.view(np.uint8)

    def batch_finish(self, partials):
        import pyref as o
        from util import pt_to96
        acc, lin, bad = o.INF, 0, 0
        for p in partials:
            w = np.ascontiguousarray(p).view(np.uint64)
            if any(w[12:18]):
                assert list(w[12:18]) == [1, 0, 0, 0, 0, 0]          # the oracle worker emits Z = 1
                acc = o.pt_add(acc, (tuple(int(c) for c in w[0:6]), tuple(int(c) for c in w[6:12])))
            lin = (lin + sum(int(w[18 + k]) << (64 * k) for k in range(4))) % o.Q
            bad |= int(w[22])
        rhs = o.pt_mul(o.generator(), lin)
        lx = acc[0] if acc is not o.INF else o.F6_ZERO
        rx = rhs[0] if rhs is not o.INF else o.F6_ZERO
        v = 3 if bad else (0 if lx == rx else 2)
        return v, np.append(pt_to96(acc), 1 if acc is o.INF else 0).astype(np.uint8), \
            np.append(pt_to96(rhs), 1 if rhs is o.INF else 0).astype(np.uint8)


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import cref
    import schnorr_sig_b200 as s
    from util import make_workload
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11                                   # ragged: 6 + 5
        w = make_workload(99, n, lens=[0, 3, 7, 8, 9, 14, 20, 1, 2, 5, 30], nthreads=1)
        sigs = w["sigs"].copy()
        sigs[7, 49:] = 0                         # one invalid signature, lands in rank 1's slice
        sv = s.ShardedVerifier(OracleWorker(), dist)
        got = sv.verify_many(sigs, w["pk"], w["inf"], w["blob"], w["off"])
        want = cref.verify_many(sigs, w["pk"], w["inf"], w["blob"], w["off"], 1)
        ok1 = np.array_equal(got, want) and list(want).count(2) == 1
        v, lhs, rhs = sv.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
        cv, cl, cr = cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"], 1)
        ok2 = v == cv == 0 and np.array_equal(lhs, cl) and np.array_equal(rhs, cr)
        v3, _, _ = sv.verify_batch(sigs, w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
        ok3 = v3 == 2                            # the corrupted signature on the non-root shard fails the batch
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    finally:
        dist.destroy_process_group()


def test_sharded_verifier_world2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True, True), (1, True, True, True)]
