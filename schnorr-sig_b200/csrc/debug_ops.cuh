// TEST HOOK (schnorr_b200_debug_lazy_ops, tests/hostsim): the lazily reduced building blocks of the fast path
// evaluated on raw 64-bit inputs, so that the tests can feed NON-canonical representatives (values in [p, 2^64))
// to every operand that claims to accept them.  a: any 64-bit limbs; b: canonical limbs.
#pragma once
#include "affine.cuh"

namespace sb {

SB_DEV fp6 fp6_canon(const fp6& a) {
    fp6 r;
#pragma unroll
    for (int i = 0; i < 6; i++) r.c[i] = fp_canon(a.c[i]);
    return r;
}
SB_DEV void debug_lazy_ops(const fp6& a, const fp6& b, fp6* out /*8*/) {
    fp6 ar = fp6{{a.c[1], a.c[2], a.c[3], a.c[4], a.c[5], a.c[0]}};
    fp6 br = fp6{{b.c[5], b.c[0], b.c[1], b.c[2], b.c[3], b.c[4]}};
    out[0] = fp6_canon(fp6_mul_nc(a, ar));
    out[1] = fp6_sqr_sub2(a, b);
    out[2] = fp6_sqr_sub_scaled(a, b, br, a.c[3]);
    out[3] = fp6_mul_sub_scaled(a, ar, b, a.c[5]);
    fp6 c;
    fp_t n;
    fp6_cofactor_norm(&b, &c, &n);
    out[4] = fp6_canon(c);
    out[5] = fp6{{n, 0, 0, 0, 0, 0}};
    fp6 nbr;
#pragma unroll
    for (int i = 0; i < 6; i++) nbr.c[i] = FP_P - br.c[i];
    out[6] = fp6_canon(fp6_scale_diff_nc(b, a.c[0], nbr, a.c[1]));
    out[7] = fp6_scale_diff(b, a.c[0], br, a.c[1]);
}

}  // namespace sb
