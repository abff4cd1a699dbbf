"""Device field arithmetic (the PTX carry-chain code that a CPU box cannot execute) against the big-int
oracle on edge values and random inputs, through the schnorr_b200_debug_field_ops test hook."""
import numpy as np
import pytest

import pyref as o
from util import rand_fp

pytestmark = pytest.mark.gpu

EDGE = [0, 1, 2, 7, 2**32 - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**33 - 1, 2**63, 2**63 - 1, o.P - 1, o.P - 2, o.P - 2**32,
        o.P - 2**32 + 1, o.P - 2**32 - 1, (o.P - 1) // 2, (o.P + 1) // 2, 0xFFFFFFFE00000002, 0xFFFFFFFF00000000,
        0x00000000FFFFFFFF, 0xFFFFFFFE00000001, 0x8000000080000000, 0x7FFFFFFF7FFFFFFF]


def test_field_ops_edge_and_random():
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    rng = np.random.default_rng(77)
    a, b = [], []
    for x in EDGE:                       # every edge value against every other, spread over the six limbs
        for k in range(0, len(EDGE), 6):
            a.append([x] * 6)
            b.append((EDGE[k:k + 6] + EDGE)[:6])
    for _ in range(3000):
        a.append([int(v) for v in rand_fp(rng, 6)])
        b.append([int(v) for v in rand_fp(rng, 6)])
    # sparse / structured Fp6 elements
    a += [[o.P - 1] * 6, [0] * 6, [1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, o.P - 1], [0, 1, 0, 0, 0, 0]]
    b += [[o.P - 1] * 6, [o.P - 1] * 6, [o.P - 1] * 6, [0, 0, 0, 0, 0, o.P - 1], [0, 0, 0, 0, 0, 1]]
    A = np.array(a, dtype=np.uint64)
    B = np.array(b, dtype=np.uint64)
    out = eng.debug_field_ops(A, B)
    for i in range(len(a)):
        ta, tb = tuple(a[i]), tuple(b[i])
        got = [tuple(int(v) for v in out[i, j]) for j in range(8)]
        assert got[0] == o.f6_mul(ta, tb), i
        assert got[1] == o.f6_sqr(ta), i
        assert got[2] == o.f6_add(ta, tb), i
        assert got[3] == o.f6_sub(ta, tb), i
        assert got[4] == tuple(x * y % o.P for x, y in zip(ta, tb)), i
        assert got[5] == tuple(x * x % o.P for x in ta), i
        if any(ta):
            assert got[6] == o.f6_inv(ta), i
        if i % 16 == 0:
            assert got[7] == tuple(pow(x, o.INV_ALPHA, o.P) for x in ta), i


def test_lazily_reduced_building_blocks_on_device():
    """The lazily reduced forms k_verify_fast is built from (fp6_mul_nc, fp6_sqr_sub2, fp6_sqr_sub_scaled,
    fp6_mul_sub_scaled, the fused cofactor/norm, non-canonical scalings) on the device, with non-canonical operands
    wherever the kernel may hand them one."""
    import schnorr_sig_b200 as s
    from util import check_lazy_ops, lazy_ops_inputs
    eng = s.default_engine(0)
    a, b = lazy_ops_inputs(np.random.default_rng(62), n_random=1500)
    check_lazy_ops(a, b, eng.debug_lazy_ops(a, b))
