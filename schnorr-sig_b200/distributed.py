"""Multi-GPU sharding of the verification path: one process per GPU.

* independent verification: contiguous slices, NO data-path collective; the verdict bytes are
  all-gathered only if the caller wants the full vector on every rank.
* batch verification: every rank reduces its slice to one 192-byte partial (Jacobian point of its
  partial MSM + partial sum s_i e_i + malformed flag); ONE small all_gather moves the partials and
  the finish kernel adds them, multiplies G and compares x-coordinates.  Randomisers are indexed by
  global signature position, so the result does not depend on the sharding.

`worker` is anything with the Engine's host-array methods (verify_many / batch_partial /
batch_finish); the CPU tests drive the same plumbing over gloo with an oracle-backed worker.
"""
import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced partition of range(n): the first n % world shards get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def slice_messages(blob, off, lo, hi):
    """Sub-blob and rebased offsets of messages lo..hi."""
    off = np.asarray(off, dtype=np.uint64)
    b0, b1 = int(off[lo]), int(off[hi])
    return np.ascontiguousarray(blob[b0:b1]), (off[lo:hi + 1] - off[lo]).astype(np.uint64)


class ShardedVerifier:
    def __init__(self, worker, dist=None, device=None):
        """dist: an initialised torch.distributed module (or None for a single process)."""
        self.worker = worker
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def _all_gather_bytes(self, local: np.ndarray, sizes):
        """all_gather of per-rank uint8 vectors of (known) different sizes."""
        if self.dist is None:
            return [local]
        import torch
        m = max(sizes) if sizes else 0
        dev = self.device if self.device is not None else "cpu"
        buf = torch.zeros(max(m, 1), dtype=torch.uint8, device=dev)
        if local.size:
            buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
        out = [torch.zeros_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf)
        return [o.cpu().numpy()[:s] for o, s in zip(out, sizes)]

    def verify_many(self, sigs81, pk96, pk_inf, blob, off, gather=True):
        """Signature::verify over all n items; this rank computes its slice.  Returns the full verdict
        vector (gather=True, one all_gather of n/world bytes per rank) or just the local slice."""
        n = sigs81.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        b, o = slice_messages(blob, off, lo, hi)
        inf = None if pk_inf is None else pk_inf[lo:hi]
        local = self.worker.verify_many(sigs81[lo:hi], pk96[lo:hi], inf, b, o)
        if not gather:
            return local
        sizes = [shard_bounds(n, r, self.world)[1] - shard_bounds(n, r, self.world)[0] for r in range(self.world)]
        return np.concatenate(self._all_gather_bytes(local, sizes)) if n else local

    def verify_batch(self, sigs81, pk96, pk_inf, blob, off, rand32):
        """verify_batch over all n items sharded across ranks -> (verdict, lhs97, rhs97) on every rank."""
        n = sigs81.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        b, o = slice_messages(blob, off, lo, hi)
        inf = None if pk_inf is None else pk_inf[lo:hi]
        part = self.worker.batch_partial(sigs81[lo:hi], pk96[lo:hi], inf, b, o, rand32[lo:hi])
        parts = self._all_gather_bytes(np.asarray(part, dtype=np.uint8).reshape(-1), [192] * self.world)
        return self.worker.batch_finish(np.stack(parts))
