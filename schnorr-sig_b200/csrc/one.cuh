// Lowest-latency form of Signature::verify (src/signature.rs:181-205) for calls of a FEW signatures -- the reference's own
// Criterion case is one verification (benches/schnorr.rs:22-66): one thread block per signature.
//
// The six-lane kernel (dist.cuh) walks one dependency chain per signature: challenge hash, 252 doublings interleaved with
// ~100 bucket additions, 27 additions of bucket aggregation, 20 table additions -- 0.93 ms however idle the GPU is.  Most
// of that chain is not a true dependency: only the doublings D_j = 2^j P depend on each other.  Here the four warps of a
// block split the work of ONE signature in three barrier-separated phases:
//   A  warp 0: the doubling chain in Jacobian coordinates on 24 lanes (jac_dbl_dist of batch.cuh: the independent products
//      of dbl-2007-bl on four six-lane groups side by side, 4 product rounds = 1.3 us per doubling), every D_j stored to
//      shared memory (253 x 144 B);   warp 1: the challenge hash on six lanes and its signed digits;   warp 2: e*G from the
//      fixed-base table on six lanes  -- side by side
//   B  the sixteen bucket accumulations (eight of the subgroup check, eight of the challenge), four per warp, each over the
//      compacted list of chain steps of ITS bucket (jac_add_dist: 5 product rounds on 24 lanes)
//   C  warp 0: aggregation of the subgroup-check buckets and the [q]P == O test;   warp 1: aggregation of the challenge
//      buckets, + e*G, x-only comparison
// Latency = the doubling chain + ~45 additions instead of chain + ~150 additions + hash: 0.50 ms for one verification
// through the host API.  Every addition that meets an exceptional operand (identity, equal x) falls back to the complete
// per-thread routine of curve.cuh, so the kernel is EXACT: only the identity key is handed to k_verify.  Additions are
// merely re-associated with respect to torsion_check_and_mul / verify_points: same group elements, same verdicts.
// (A first version ran the (X, Y, w) formulas of dist.cuh on six lanes with a helper warp for X^2: 0.58 ms,
// profiles/r2_variants.md.)
#pragma once
#include "dist.cuh"

namespace sb {

static constexpr int ONE_THREADS = 128;
static constexpr int ONE_SLOTS = 16;  // bucket accumulations: 0..7 subgroup check (odd digits of q), 8..15 challenge (|d| = 1..8)

struct one_shared {
    fp_t cx[SB_CHAIN_STEPS][6], cy[SB_CHAIN_STEPS][6], cz[SB_CHAIN_STEPS][6];  // the chain D_j (Jacobian)
    fp_t bx[ONE_SLOTS][6], by[ONE_SLOTS][6], bz[ONE_SLOTS][6];                 // bucket sums
    fp_t ex[6], ey[6], ew;                                                     // e*G as (X, Y, w): a Jacobian point with Z = w
    uint32_t mds2[24];
    int8_t hd[64];
    uint8_t lst[ONE_SLOTS][64];
    uint8_t neg[ONE_SLOTS][64];
    int cnt[ONE_SLOTS];
    int e_empty, torsion_free;
};

// MINB = resident blocks per SM the register allocation must allow: 2 (248 registers, no spills) while every block of the
// call is resident anyway (n <= 2 x SMs), 4 (128 registers; the rarely taken exact routines spill) beyond
template <int MINB>
__global__ void __launch_bounds__(ONE_THREADS, MINB) k_verify_one(soa_batch in, const uint8_t* __restrict__ msgs,
                                                             const uint64_t* __restrict__ msg_off,
                                                             const uint64_t* __restrict__ gtab, uint8_t* __restrict__ verdicts,
                                                             uint32_t* __restrict__ work_list, uint32_t* __restrict__ work_count) {
    __shared__ one_shared S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool in24 = lane < 24;             // the 24 lanes of the Jacobian toolkit: four groups x six coefficients
    const int g = in24 ? lane / 6 : 0, k = lane % 6;
    const size_t i = blockIdx.x, n = in.n;
    const uint8_t fl = in.flags[i];          // block-uniform from here on
    if (fl & (FL_MALFORMED | FL_PK_INF)) {
        if (tid == 0) {
            uint8_t v = (fl & FL_MALFORMED) ? VERDICT_MALFORMED : VERDICT_NEEDS_EXACT;
            verdicts[i] = v;
            if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
        }
        return;
    }
    if (tid < 24) S.mds2[tid] = c_mds_row[tid % 12];
    if (tid < ONE_SLOTS) S.cnt[tid] = 0;
    if (tid == 0) {
        S.e_empty = 1;
        S.torsion_free = 0;
    }
    __syncthreads();
    const bool x_ok = !(fl & FL_X_BAD);
    const uint64_t* pl = reinterpret_cast<const uint64_t*>(in.planes);
    fp_t sx = pl[((size_t)(0 + (k >> 1)) * n + i) * 2 + (k & 1)];
    fp_t px = pl[((size_t)(5 + (k >> 1)) * n + i) * 2 + (k & 1)];
    fp_t py = pl[((size_t)(8 + (k >> 1)) * n + i) * 2 + (k & 1)];

    // ---- phase A: doubling chain | challenge | e*G ------------------------------------------------------------------
    if (warp == 0) {
        if (in24) {
            fp_t x = px, y = py, z = k == 0 ? 1 : 0;
#pragma unroll 1
            for (int j = 0; j < SB_CHAIN_STEPS; j++) {
                if (g == 0) {
                    S.cx[j][k] = x;
                    S.cy[j][k] = y;
                    S.cz[j][k] = z;
                }
                if (j < SB_CHAIN_STEPS - 1) jac_dbl_dist(x, y, z, g, k);
            }
        }
    } else if (warp == 1) {
        if (lane < 6) {
            scalar h = sc_zero();
            uint64_t off = msg_off[i];
            if (x_ok) h = dchallenge_scalar(0x3fu, sx, px, py, msgs + off, msg_off[i + 1] - off, k, 0, S.mds2);
            if (k == 0) recode_signed_w4(h, S.hd);
        }
    } else if (warp == 2) {
        if (lane < 6) {
            scalar e = load_scalar_planes(in.planes, 3, n, i);
            dpt R{0, 0, 1}, T;
            T.w = 1;
            bool e_acc = true, exc = false;
            int carry = 0;
#pragma unroll 1
            for (int w = 0; w < GTAB_WINDOWS; w++) {
                int raw = (int)sc_bits(e, GTAB_W * w, GTAB_W) + carry;
                bool neg = raw > (1 << (GTAB_W - 1));
                carry = neg ? 1 : 0;
                int dg = neg ? (1 << GTAB_W) - raw : raw;
                const uint64_t* ent = gtab + ((size_t)w * GTAB_ENTRIES + (dg ? dg : 1)) * GTAB_ENTRY_U64;
                T.X = ent[k];
                T.Y = ent[6 + k];
                exc |= djf_add(0x3fu, &R, &T, jf_add_mode(e_acc, dg == 0, neg), k, 0);
                e_acc = e_acc && dg == 0;
            }
            S.ex[k] = R.X;
            S.ey[k] = R.Y;
            if (k == 0) {
                S.ew = R.w;
                S.e_empty = exc ? 2 : (e_acc ? 1 : 0);   // 2: the table points collided (never for a reduced scalar)
            }
        }
    }
    __syncthreads();

    // ---- phase B: sixteen bucket accumulations, four per warp ------------------------------------------------------
    if (in24) {
#pragma unroll 1
        for (int q4 = 0; q4 < 4; q4++) {
            const int slot = warp * 4 + q4;
            if (lane == 0) {
                int c = 0;
                if (slot < 8) {
#pragma unroll 1
                    for (int j = 0; j < CHEETAH_Q_WNAF5_LEN; j++) {
                        int dq = SB_QWNAF(j);
                        if (dq != 0 && ((dq < 0 ? -dq : dq) >> 1) == slot) {
                            S.lst[slot][c] = (uint8_t)j;
                            S.neg[slot][c] = dq < 0;
                            c++;
                        }
                    }
                } else {
#pragma unroll 1
                    for (int w = 0; w < 64; w++) {
                        int dh = S.hd[w];
                        if (dh != 0 && (dh < 0 ? -dh : dh) == slot - 7) {
                            S.lst[slot][c] = (uint8_t)(4 * w);
                            S.neg[slot][c] = dh < 0;
                            c++;
                        }
                    }
                }
                S.cnt[slot] = c;
            }
            __syncwarp(HORNER_MASK);
            const int cnt = S.cnt[slot];
            fp_t x = 1, y = 1, z = 0;            // identity (curve.cuh: jac_identity; only z matters)
#pragma unroll 1
            for (int t = 0; t < cnt; t++) {
                int j = S.lst[slot][t];
                fp_t x2 = S.cx[j][k], y2 = S.cy[j][k], z2 = S.cz[j][k];
                if (S.neg[slot][t]) y2 = fp_neg(y2);
                if (t == 0) {
                    x = x2;
                    y = y2;
                    z = z2;
                } else {
                    jac_add_dist_exact(x, y, z, x2, y2, z2, g, k);
                }
            }
            if (g == 0) {
                S.bx[slot][k] = x;
                S.by[slot][k] = y;
                S.bz[slot][k] = z;
            }
        }
    }
    __syncthreads();

    // ---- phase C: aggregation ---------------------------------------------------------------------------------------
    if (warp == 0 && in24) {       // [q]P = 2 O_1 + R_0 (odd digits 2k+1)
        fp_t rx = S.bx[7][k], ry = S.by[7][k], rz = S.bz[7][k];
        fp_t ox = rx, oy = ry, oz = rz;
#pragma unroll 1
        for (int b = 6; b >= 0; b--) {
            jac_add_dist_exact(rx, ry, rz, S.bx[b][k], S.by[b][k], S.bz[b][k], g, k);
            if (b >= 1) jac_add_dist_exact(ox, oy, oz, rx, ry, rz, g, k);
        }
        jac_dbl_dist(ox, oy, oz, g, k);
        jac_add_dist_exact(ox, oy, oz, rx, ry, rz, g, k);
        bool ident = __ballot_sync(HORNER_MASK, oz == 0) == HORNER_MASK;
        if (lane == 0) S.torsion_free = ident;
    }
    uint8_t v = VERDICT_INVALID_SIGNATURE;
    bool have_v = false;
    if (warp == 1 && in24) {       // h*P = O_0 (digits m = k+1), + e*G
        fp_t rx = S.bx[15][k], ry = S.by[15][k], rz = S.bz[15][k];
        fp_t ox = rx, oy = ry, oz = rz;
#pragma unroll 1
        for (int b = 6; b >= 0; b--) {
            jac_add_dist_exact(rx, ry, rz, S.bx[8 + b][k], S.by[8 + b][k], S.bz[8 + b][k], g, k);
            jac_add_dist_exact(ox, oy, oz, rx, ry, rz, g, k);
        }
        fp_t ez = (S.e_empty == 0 && k == 0) ? S.ew : 0;
        jac_add_dist_exact(ox, oy, oz, S.ex[k], S.ey[k], ez, g, k);
        jac_pt R;
        R.X = gather_fp6(ox);
        R.Y = gather_fp6(oy);
        R.Z = gather_fp6(oz);
        fp6 sxx = gather_fp6(sx);
        v = jac_x_equals(R, sxx) ? VERDICT_OK : VERDICT_INVALID_SIGNATURE;   // x-only comparison, src/signature.rs:200
        have_v = lane == 0;
    }
    __syncthreads();
    if (have_v) {
        if (S.e_empty == 2) v = VERDICT_NEEDS_EXACT;
        else if (!S.torsion_free) v = VERDICT_INVALID_PUBLIC_KEY;
        else if (!x_ok) v = VERDICT_MALFORMED;
        verdicts[i] = v;
        if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
    }
}

}  // namespace sb
