// schnorr_b200 -- kernels, host orchestration and the C ABI (include/schnorr_b200.h).
//
// Kernel inventory ("kernels the new build must create"):
//   k_ingest          AoS wire records -> SoA limb planes (coalesced 128-bit stores), input checks
//   k_hash            K1  hash_message per thread
//   k_verify          K2  Signature::verify per thread (challenge hash + subgroup check + h*P + e*G)
//   k_keygen / k_sign K5  fixed-base multiplication; device signer for synthetic inputs
//   k_gtab_*              builds the fixed-base table of G at context creation
//   k_imad_peak       K6  integer-multiply roofline calibration
//   batch kernels     K3/K4  in batch.cuh (Pippenger MSM)
// No CPU fallback exists: every entry point needs a live CUDA context.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cheetah_params.h"
#include "../../include/schnorr_b200.h"
#include "verify.cuh"
#include "dist.cuh"
#include "debug_ops.cuh"
#include "derive.cuh"

using namespace sb;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static constexpr size_t GTAB_U64 = (size_t)GTAB_WINDOWS * GTAB_ENTRIES * 12;

enum scratch_slot { SL_A = 0, SL_B, SL_C, SL_D, SL_E, SL_F, SL_G, SL_H, SL_I, SL_J, SL_K, SL_L, SL_M, SL_COUNT };

struct schnorr_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    uint64_t* gtab = nullptr;
    void* scratch[SL_COUNT] = {};
    size_t scratch_cap[SL_COUNT] = {};
    uint64_t launches = 0;
    int sm_count = 148;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // bracket the dominant kernel of the last call
    size_t verify_wave = 148 * 256;                 // signatures resident at once in k_verify (filled at creation)
    cudaStream_t copy_stream = nullptr;             // host->device staging of the pipelined host entry points
    cudaStream_t aux_stream = nullptr;              // second compute stream: consecutive pipeline chunks alternate between
    cudaEvent_t ev_aux = nullptr;                   // the two, so a chunk's blocks fill the SMs the previous chunk drains
    bool exact_only = false;                        // schnorr_b200_set_exact_only: skip the fast path (A/B measurements, tests)
    int msm_c_override = 0;                         // schnorr_b200_set_msm_geometry (tests): forced window width / segment length
    uint32_t msm_t_override = 0;
    int last_msm_c = 0, last_msm_K = 0;             // Pippenger geometry of the last batch call (schnorr_b200_last_batch_plan)
    uint32_t last_msm_T = 0;
    size_t dist_max = 10240;                        // calls up to this many signatures use the six-lanes-per-signature kernel
    size_t one_max = 512;                           // ... and up to this many the block-per-signature kernel (one.cuh)
    size_t batch_dist_max = (size_t)1 << 18;        // batches up to this size hash their challenges on six lanes per signature
    size_t batch_small_max = 256;                   // batches up to this size run one thread block per signature (k_batch_small)
    int exact_counters_used = 0;                    // work-list counters written by the last verify call
    static constexpr int MAX_CHUNKS = 16;
    cudaEvent_t ev_chunk[MAX_CHUNKS] = {};
    std::string err;
    // multi-device context (schnorr_b200_create_multi): one single-device context per listed device; empty otherwise
    std::vector<schnorr_b200_ctx*> shards;
};

#define CUDA_TRY(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                     \
            return SCHNORR_B200_ECUDA;                                                           \
        }                                                                                        \
    } while (0)

static int ensure_scratch(schnorr_b200_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_cap[slot] < bytes) {
        if (ctx->scratch[slot]) {
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            CUDA_TRY(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_cap[slot] = 0;
        }
        size_t cap = bytes + bytes / 8;
        CUDA_TRY(ctx, cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_cap[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return 0;
}

// ------------------------------------------------------------------------------------------------
// SoA device layout: plane-major ulonglong2, plane p of an Fp6 holds coefficients (2p, 2p+1).
//   sx[3][n] sig.x   se[2][n] sig.e   px[3][n] py[3][n] public key   fl[n] flags
// A warp reads 32 x 16 B = 512 contiguous bytes per plane.
// ------------------------------------------------------------------------------------------------
static constexpr uint8_t FL_PK_INF = 1;     // public key is the identity
static constexpr uint8_t FL_X_BAD = 2;      // sig.x has a non-canonical limb
static constexpr uint8_t FL_MALFORMED = 4;  // e >= q or non-canonical public-key limb
static constexpr int SOA_PLANES = 11;

struct soa_batch {
    ulonglong2* planes;  // [11][n]: 0-2 sx, 3-4 se, 5-7 px, 8-10 py
    uint8_t* flags;      // [n]
    uint8_t* sig_flag;   // [n] byte 48 of the compressed point (used by batch verification only)
    size_t n;
};

// The planes are read exactly once per launch: streaming loads (evict-first) keep them from displacing the thread-local
// bucket lines that the L2 is really caching in these kernels.
__device__ __forceinline__ ulonglong2 load_plane(const ulonglong2* p) {
    return __ldcs(p);
}
__device__ __forceinline__ fp6 load_fp6_planes(const ulonglong2* planes, int first_plane, size_t n, size_t i) {
    ulonglong2 a = load_plane(&planes[(size_t)(first_plane + 0) * n + i]);
    ulonglong2 b = load_plane(&planes[(size_t)(first_plane + 1) * n + i]);
    ulonglong2 c = load_plane(&planes[(size_t)(first_plane + 2) * n + i]);
    return fp6{{a.x, a.y, b.x, b.y, c.x, c.y}};
}
__device__ __forceinline__ scalar load_scalar_planes(const ulonglong2* planes, int first_plane, size_t n, size_t i) {
    ulonglong2 a = load_plane(&planes[(size_t)(first_plane + 0) * n + i]);
    ulonglong2 b = load_plane(&planes[(size_t)(first_plane + 1) * n + i]);
    return sc_from_u64x4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint64_t load_u64_le(const uint8_t* p) {
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v |= (uint64_t)p[k] << (8 * k);
    return v;
}

// Block-wide agreement on the challenge hash.  The permutation takes a block barrier per half-round when every hashing
// thread of the block runs the same number of Rescue permutations (rescue.cuh: the warps then share instruction-cache
// lines), and barriers need EVERY thread of the block: threads without work of their own (past the end of the batch,
// malformed items, identity keys) therefore hash along on the message of the block's first hashing thread and drop the
// result.  Must be called by all threads of the block.
//   np         this thread's permutation count, -1 if it has nothing to hash
//   off / len  in: this thread's message; out: the message to hash (its own, or the first hasher's)
//   returns    sync  = all hashing threads agree on the count (take the barriers)
//              any   = at least one thread of the block hashes
struct hash_vote {
    bool sync, any;
};
__device__ __forceinline__ hash_vote block_hash_vote(int np, uint64_t& off, uint64_t& len) {
    __shared__ int s_np_min, s_np_max, s_first;
    __shared__ uint64_t s_off, s_len;
    if (threadIdx.x == 0) {
        s_np_min = 1 << 30;
        s_np_max = -1;
        s_first = 1 << 30;
    }
    __syncthreads();
    if (np >= 0) {
        atomicMin(&s_np_min, np);
        atomicMax(&s_np_max, np);
        atomicMin(&s_first, (int)threadIdx.x);
    }
    __syncthreads();
    if ((int)threadIdx.x == s_first) {
        s_off = off;
        s_len = len;
    }
    __syncthreads();
    hash_vote v;
    v.any = s_np_max >= 0;
    v.sync = v.any && s_np_min >= s_np_max;
    if (np < 0 && v.any) {
        off = s_off;
        len = s_len;
    }
    return v;
}

// AoS -> SoA.  The 81-byte signature records of a block are staged through shared memory with
// coalesced 16-byte loads; public keys (96 B, 8-byte aligned) are read directly.
static constexpr int INGEST_THREADS = 128;
__global__ void __launch_bounds__(INGEST_THREADS) k_ingest(size_t n, const uint8_t* __restrict__ sigs81,
                                                           const uint8_t* __restrict__ pk96,
                                                           const uint8_t* __restrict__ pk_inf, soa_batch out) {
    __shared__ __align__(16) uint8_t s_sig[INGEST_THREADS * 81];
    size_t base = (size_t)blockIdx.x * INGEST_THREADS;
    size_t cnt = n - base < (size_t)INGEST_THREADS ? n - base : (size_t)INGEST_THREADS;
    size_t bytes = cnt * 81;
    const uint8_t* src = sigs81 + base * 81;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        size_t nv = bytes / 16;
        for (size_t k = threadIdx.x; k < nv; k += INGEST_THREADS)
            reinterpret_cast<uint4*>(s_sig)[k] = reinterpret_cast<const uint4*>(src)[k];
        for (size_t k = nv * 16 + threadIdx.x; k < bytes; k += INGEST_THREADS) s_sig[k] = src[k];
    } else {
        for (size_t k = threadIdx.x; k < bytes; k += INGEST_THREADS) s_sig[k] = src[k];
    }
    __syncthreads();
    size_t i = base + threadIdx.x;
    if (i >= n) return;
    const uint8_t* rec = s_sig + threadIdx.x * 81;
    uint64_t x[6], e[4];
#pragma unroll
    for (int k = 0; k < 6; k++) x[k] = load_u64_le(rec + 8 * k);
#pragma unroll
    for (int k = 0; k < 4; k++) e[k] = load_u64_le(rec + 49 + 8 * k);
    const uint64_t* pk = reinterpret_cast<const uint64_t*>(pk96 + i * 96);
    uint64_t p[12];
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = pk[k];
    uint8_t fl = 0;
    bool inf = pk_inf != nullptr && pk_inf[i] != 0;
    if (inf) fl |= FL_PK_INF;
    bool x_ok = true, pk_ok = true;
#pragma unroll
    for (int k = 0; k < 6; k++) x_ok &= x[k] < FP_P;
#pragma unroll
    for (int k = 0; k < 12; k++) pk_ok &= p[k] < FP_P;
    if (!x_ok) fl |= FL_X_BAD;
    if ((!pk_ok && !inf) || sc_geq_q(sc_from_u64x4(e[0], e[1], e[2], e[3]))) fl |= FL_MALFORMED;
    size_t N = out.n;
    ulonglong2* pl = out.planes;
    pl[0 * N + i] = make_ulonglong2(x[0], x[1]);
    pl[1 * N + i] = make_ulonglong2(x[2], x[3]);
    pl[2 * N + i] = make_ulonglong2(x[4], x[5]);
    pl[3 * N + i] = make_ulonglong2(e[0], e[1]);
    pl[4 * N + i] = make_ulonglong2(e[2], e[3]);
#pragma unroll
    for (int k = 0; k < 6; k++) pl[(size_t)(5 + k) * N + i] = make_ulonglong2(p[2 * k], p[2 * k + 1]);
    out.flags[i] = fl;
    out.sig_flag[i] = rec[48];
}

// ------------------------------------------------------------------------------------------------
// K2: Signature::verify, one signature per thread
// ------------------------------------------------------------------------------------------------
#ifndef VERIFY_THREADS_N
#define VERIFY_THREADS_N 128
#endif
static constexpr int VERIFY_THREADS = VERIFY_THREADS_N;
#ifndef VERIFY_MIN_BLOCKS
#define VERIFY_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(VERIFY_THREADS, VERIFY_MIN_BLOCKS) k_verify(soa_batch in, const uint8_t* __restrict__ msgs,
                                                           const uint64_t* __restrict__ msg_off,
                                                           const uint64_t* __restrict__ gtab,
                                                           uint8_t* __restrict__ verdicts,
                                                           const uint32_t* __restrict__ work_list,
                                                           const uint32_t* __restrict__ work_count) {
    // running point D_j of each thread: shared memory, padded to 152 B per thread (2-way bank conflicts at most)
    struct d_slot {
        jac_pt p;
        uint64_t pad;
    };
    __shared__ d_slot s_d[VERIFY_THREADS];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (work_list) {  // exact pass over the items the fast path handed back
        if (i >= *work_count) return;
        i = work_list[i];
    }
    if (i >= in.n) return;
    uint8_t fl = in.flags[i];
    if (fl & FL_MALFORMED) {
        verdicts[i] = VERDICT_MALFORMED;
        return;
    }
    fp6 sx = load_fp6_planes(in.planes, 0, in.n, i);
    scalar e = load_scalar_planes(in.planes, 3, in.n, i);
    fp6 px = load_fp6_planes(in.planes, 5, in.n, i);
    fp6 py = load_fp6_planes(in.planes, 8, in.n, i);
    bool pk_inf = fl & FL_PK_INF;
    bool x_ok = !(fl & FL_X_BAD);
    uint64_t off = msg_off[i];
    scalar h = sc_zero();
    if (x_ok) h = challenge_scalar(sx, px, py, pk_inf, msgs + off, msg_off[i + 1] - off);
    verdicts[i] = verify_points(sx, x_ok, e, px, py, pk_inf, h, gtab, &s_d[threadIdx.x].p);
}

// K2 fast path: the same verdicts through point arithmetic in (X, Y, w) coordinates, w in Fp (affine.cuh).
// Items that meet an exceptional case of the chord-and-tangent formulas (identity / small-order keys, colliding
// partial sums: ~0.1 % of honest inputs) are appended to `work_list` and finished by k_verify.
#ifndef VERIFY_FAST_MIN_BLOCKS
#define VERIFY_FAST_MIN_BLOCKS 2
#endif
static constexpr size_t VERIFY_FAST_SMEM = (size_t)(1 + FAST_BH_SHARED) * VERIFY_THREADS * sizeof(jf_pt);  // 106 496 B: two blocks per SM
__global__ void __launch_bounds__(VERIFY_THREADS, VERIFY_FAST_MIN_BLOCKS) k_verify_fast(soa_batch in, const uint8_t* __restrict__ msgs,
                                                           const uint64_t* __restrict__ msg_off,
                                                           const uint64_t* __restrict__ gtab,
                                                           uint8_t* __restrict__ verdicts,
                                                           uint32_t* __restrict__ work_list,
                                                           uint32_t* __restrict__ work_count) {
    // dynamic shared memory (VERIFY_FAST_SMEM bytes): the running point D_j of every thread, then FAST_BH_SHARED
    // challenge buckets per thread, bucket-major.  A slot is 104 bytes: 64-bit accesses of a half-warp fall into sixteen
    // distinct even banks whatever bucket each thread addresses.
    extern __shared__ __align__(16) unsigned char s_fast[];
    jf_pt* s_d = reinterpret_cast<jf_pt*>(s_fast);
    jf_pt* s_bh = s_d + VERIFY_THREADS;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < in.n;
    uint8_t fl = live ? in.flags[i] : FL_MALFORMED;
    // `work`: this thread has a verification of its own.  Every other thread of the block (past the end of the batch,
    // malformed record, identity key) runs the SAME instruction stream on stand-in inputs -- the generator as the key,
    // zero scalars, the first hashing thread's message -- so that all 128 threads reach every barrier of the hash and of
    // the point loops; its result is dropped.
    bool work = live && !(fl & (FL_MALFORMED | FL_PK_INF));
    uint64_t off = work ? msg_off[i] : 0, len = work ? msg_off[i + 1] - off : 0;
    hash_vote hv = block_hash_vote(work ? hash_message_permutations(len) : -1, off, len);
    fp6 sx = fp6_zero(), px, py;
    scalar e = sc_zero();
    if (work) {
        sx = load_fp6_planes(in.planes, 0, in.n, i);
        e = load_scalar_planes(in.planes, 3, in.n, i);
        px = load_fp6_planes(in.planes, 5, in.n, i);
        py = load_fp6_planes(in.planes, 8, in.n, i);
    } else {  // 1 * G, the first entry of the fixed-base table
        const uint64_t* g = gtab + (size_t)1 * GTAB_ENTRY_U64;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            px.c[k] = g[k];
            py.c[k] = g[6 + k];
        }
    }
    bool x_ok = !(fl & FL_X_BAD);
    scalar h = sc_zero();
    if (hv.any) {  // block-uniform.  An item with a malformed x hashes too (result unused)
        h = challenge_scalar(sx, px, py, false, msgs + off, len, hv.sync);
        if (!x_ok || !work) h = sc_zero();
    }
    uint8_t v = verify_points_fast(sx, x_ok, e, px, py, false, h, gtab, s_d + threadIdx.x, s_bh + threadIdx.x, VERIFY_THREADS);
    if (!live) return;
    if (fl & FL_MALFORMED) v = VERDICT_MALFORMED;
    else if (fl & FL_PK_INF) v = VERDICT_NEEDS_EXACT;   // the identity key is the exact kernel's business
    verdicts[i] = v;
    if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
}

// K2 for small calls: one signature per group of six lanes (dist.cuh), five signatures per warp.  Same verdicts and
// the same hand-back list as k_verify_fast.
static constexpr int DIST_THREADS = 128;
static constexpr int DIST_SIGS_PER_BLOCK = (DIST_THREADS / 32) * 5;
__global__ void __launch_bounds__(DIST_THREADS) k_verify_dist(soa_batch in, const uint8_t* __restrict__ msgs,
                                                              const uint64_t* __restrict__ msg_off,
                                                              const uint64_t* __restrict__ gtab,
                                                              uint8_t* __restrict__ verdicts,
                                                              uint32_t* __restrict__ work_list,
                                                              uint32_t* __restrict__ work_count) {
    __shared__ uint32_t s_mds2[24];  // the circulant MDS row twice: M[i][j] = s_mds2[j - i + 12]
    if (threadIdx.x < 24) s_mds2[threadIdx.x] = c_mds_row[threadIdx.x % 12];
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int g = lane / 6, k = lane % 6;
    if (g >= 5) return;  // lanes 30, 31 idle
    size_t i = ((size_t)blockIdx.x * (DIST_THREADS / 32) + warp) * 5 + g;
    if (i >= in.n) return;
    int gbase = 6 * g;
    unsigned mask = 0x3fu << gbase;
    uint8_t fl = in.flags[i];
    if (fl & FL_MALFORMED) {
        if (k == 0) verdicts[i] = VERDICT_MALFORMED;
        return;
    }
    uint8_t v;
    if (fl & FL_PK_INF) {
        v = VERDICT_NEEDS_EXACT;
    } else {
        bool x_ok = !(fl & FL_X_BAD);
        // the lane's coefficient k of an Fp6 held in planes p0..p0+2 (ulonglong2 = two coefficients)
        const uint64_t* pl = reinterpret_cast<const uint64_t*>(in.planes);
        size_t n = in.n;
        fp_t sx = pl[((size_t)(0 + (k >> 1)) * n + i) * 2 + (k & 1)];
        fp_t px = pl[((size_t)(5 + (k >> 1)) * n + i) * 2 + (k & 1)];
        fp_t py = pl[((size_t)(8 + (k >> 1)) * n + i) * 2 + (k & 1)];
        scalar e = load_scalar_planes(in.planes, 3, n, i);
        uint64_t off = msg_off[i];
        scalar h = sc_zero();
        if (x_ok) h = dchallenge_scalar(mask, sx, px, py, msgs + off, msg_off[i + 1] - off, k, gbase, s_mds2);
        dpt r;
        int fr = dverify_core(mask, px, py, h, e, gtab, &r, k, gbase);
        bool eq = dall(mask, gbase, r.X == fp_mul(sx, fp_sqr_nc(r.w)));  // x(R) == sig.x  <=>  X == sig.x w^2
        v = fr == FAST_EXCEPTIONAL ? VERDICT_NEEDS_EXACT
            : (fr == FAST_NOT_TORSION_FREE ? VERDICT_INVALID_PUBLIC_KEY
               : (!x_ok ? VERDICT_MALFORMED : (eq ? VERDICT_OK : VERDICT_INVALID_SIGNATURE)));
    }
    if (k == 0) {
        verdicts[i] = v;
        if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
    }
}


// ------------------------------------------------------------------------------------------------
// K1: hash_message, one message per thread (AoS inputs: 8-byte aligned records)
// ------------------------------------------------------------------------------------------------
static constexpr int HASH_THREADS = 128;
__global__ void __launch_bounds__(HASH_THREADS) k_hash(size_t n, const uint8_t* __restrict__ rx48,
                                                       const uint8_t* __restrict__ pk96,
                                                       const uint8_t* __restrict__ msgs,
                                                       const uint64_t* __restrict__ msg_off,
                                                       uint8_t* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    uint64_t off = live ? msg_off[i] : 0, len = live ? msg_off[i + 1] - off : 0;
    // threads past the end of the batch hash the first live thread's message along (every thread reaches the barriers)
    hash_vote hv = block_hash_vote(live ? hash_message_permutations(len) : -1, off, len);
    size_t src = live ? i : (size_t)blockIdx.x * blockDim.x;   // the block's first item always exists
    const uint64_t* r = reinterpret_cast<const uint64_t*>(rx48 + src * 48);
    const uint64_t* p = reinterpret_cast<const uint64_t*>(pk96 + src * 96);
    fp6 rx, px;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        rx.c[k] = r[k];
        px.c[k] = p[k];
    }
    fp_t py0 = p[6];
    fp_t d[4];
    hash_message(rx, px, py0, msgs + off, len, d, hv.sync);
    if (!live) return;
    ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + i * 32);
    o[0] = make_ulonglong2(d[0], d[1]);
    o[1] = make_ulonglong2(d[2], d[3]);
}

// K1 for small calls: one message per group of six lanes (dist.cuh), five messages per warp
__global__ void __launch_bounds__(DIST_THREADS) k_hash_dist(size_t n, const uint8_t* __restrict__ rx48,
                                                            const uint8_t* __restrict__ pk96,
                                                            const uint8_t* __restrict__ msgs,
                                                            const uint64_t* __restrict__ msg_off,
                                                            uint8_t* __restrict__ digests) {
    __shared__ uint32_t s_mds2[24];
    if (threadIdx.x < 24) s_mds2[threadIdx.x] = c_mds_row[threadIdx.x % 12];
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int g = lane / 6, k = lane % 6;
    if (g >= 5) return;
    size_t i = ((size_t)blockIdx.x * (DIST_THREADS / 32) + warp) * 5 + g;
    if (i >= n) return;
    int gbase = 6 * g;
    unsigned mask = 0x3fu << gbase;
    const uint64_t* r = reinterpret_cast<const uint64_t*>(rx48 + i * 48);
    const uint64_t* p = reinterpret_cast<const uint64_t*>(pk96 + i * 96);
    uint64_t off = msg_off[i];
    fp_t d[4];
    dhash_message(mask, r[k], p[k], p[6 + k], msgs + off, msg_off[i + 1] - off, k, gbase, s_mds2, d);
    if (k < 4) reinterpret_cast<uint64_t*>(digests + i * 32)[k] = d[k];
}

// ------------------------------------------------------------------------------------------------
// K5: fixed-base multiplication -> key generation and the device signer
// ------------------------------------------------------------------------------------------------
static constexpr int SIGN_THREADS = 128;
__device__ __forceinline__ void store_point96(uint8_t* dst, const fp6& x, const fp6& y) {
    uint64_t* o = reinterpret_cast<uint64_t*>(dst);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        o[k] = x.c[k];
        o[6 + k] = y.c[k];
    }
}
__global__ void __launch_bounds__(SIGN_THREADS) k_keygen(size_t n, const uint8_t* __restrict__ sk32,
                                                         const uint64_t* __restrict__ gtab, uint8_t* __restrict__ pk96,
                                                         uint8_t* __restrict__ pk_inf) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    scalar k = sc_from_u256(sc_load_le(sk32 + 32 * i));
    jac_pt P = fixed_base_mul(k, gtab);
    fp6 x, y;
    bool inf;
    jac_to_affine(P, x, y, inf);
    store_point96(pk96 + 96 * i, x, y);
    if (pk_inf) pk_inf[i] = inf ? 1 : 0;
}
__global__ void __launch_bounds__(SIGN_THREADS) k_sign(size_t n, const uint8_t* __restrict__ sk32,
                                                       const uint8_t* __restrict__ pk96,
                                                       const uint8_t* __restrict__ pk_inf,
                                                       const uint8_t* __restrict__ msgs,
                                                       const uint64_t* __restrict__ msg_off,
                                                       const uint8_t* __restrict__ nonce32,
                                                       const uint64_t* __restrict__ gtab, uint8_t* __restrict__ sigs81) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    scalar sk = sc_from_u256(sc_load_le(sk32 + 32 * i));
    scalar r = sc_from_u256(sc_load_le(nonce32 + 32 * i));
    const uint64_t* p = reinterpret_cast<const uint64_t*>(pk96 + i * 96);
    fp6 px, py;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        px.c[k] = p[k];
        py.c[k] = p[6 + k];
    }
    bool inf = pk_inf != nullptr && pk_inf[i] != 0;
    jac_pt R = fixed_base_mul(r, gtab);
    fp6 x, y;
    bool r_inf;
    jac_to_affine(R, x, y, r_inf);  // identity -> x = y = 0
    uint64_t off = msg_off[i];
    scalar h = challenge_scalar(x, px, py, inf, msgs + off, msg_off[i + 1] - off);
    scalar e = sc_sub(r, sc_mul(sk, h));
    uint8_t* o = sigs81 + 81 * i;
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int b = 0; b < 8; b++) o[8 * k + b] = (uint8_t)(x.c[k] >> (8 * b));
    o[48] = compress_flags(y, r_inf);
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int b = 0; b < 4; b++) o[49 + 4 * k + b] = (uint8_t)(e.l[k] >> (8 * b));
}

// ------------------------------------------------------------------------------------------------
// compressed point codecs (PublicKey::from_bytes / to_bytes, src/public.rs:49-56)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_decompress(size_t n, const uint8_t* __restrict__ in49, uint8_t* __restrict__ pk96,
                                                    uint8_t* __restrict__ pk_inf, uint8_t* __restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* rec = in49 + 49 * i;
    fp6 x;
#pragma unroll
    for (int k = 0; k < 6; k++) x.c[k] = load_u64_le(rec + 8 * k);
    fp6 ox, oy;
    bool inf;
    bool good = decompress_point(x, rec[48], ox, oy, inf);
    store_point96(pk96 + 96 * i, ox, oy);
    pk_inf[i] = inf ? 1 : 0;
    ok[i] = good ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_compress(size_t n, const uint8_t* __restrict__ pk96,
                                                  const uint8_t* __restrict__ pk_inf, uint8_t* __restrict__ out49) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* p = reinterpret_cast<const uint64_t*>(pk96 + i * 96);
    bool inf = pk_inf != nullptr && pk_inf[i] != 0;
    fp6 y;
#pragma unroll
    for (int k = 0; k < 6; k++) y.c[k] = p[6 + k];
    uint8_t* o = out49 + 49 * i;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        uint64_t v = inf ? 0 : p[k];
#pragma unroll
        for (int b = 0; b < 8; b++) o[8 * k + b] = (uint8_t)(v >> (8 * b));
    }
    o[48] = compress_flags(y, inf);
}

// KeyedSignature wire records (src/signature.rs:237-271): 49-byte compressed public key || 81-byte
// signature.  The key is decompressed on the device; records whose key does not decode get verdict 3
// (KeyedSignature::from_bytes returns None for them).
__global__ void __launch_bounds__(128) k_split_keyed(size_t n, const uint8_t* __restrict__ keyed130, uint8_t* __restrict__ pk96,
                                                     uint8_t* __restrict__ pk_inf, uint8_t* __restrict__ ok,
                                                     uint8_t* __restrict__ sigs81) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* rec = keyed130 + 130 * i;
    fp6 x;
#pragma unroll
    for (int k = 0; k < 6; k++) x.c[k] = load_u64_le(rec + 8 * k);
    fp6 ox, oy;
    bool inf;
    bool good = decompress_point(x, rec[48], ox, oy, inf);
    store_point96(pk96 + 96 * i, ox, oy);
    pk_inf[i] = inf ? 1 : 0;
    ok[i] = good ? 1 : 0;
    for (int k = 0; k < 81; k++) sigs81[81 * i + k] = rec[49 + k];
}
__global__ void k_mask_verdicts(size_t n, const uint8_t* __restrict__ ok, uint8_t* __restrict__ verdicts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !ok[i]) verdicts[i] = VERDICT_MALFORMED;
}

// ------------------------------------------------------------------------------------------------
// f4: hierarchical deterministic derivation (src/derivation.rs:66-277), one child per thread (derive.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_derive_master(size_t n, const uint8_t* __restrict__ seeds32, uint8_t* __restrict__ xsk64,
                                                       uint8_t* __restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ok[i] = derive_master(seeds32 + 32 * i, xsk64 + 64 * i) ? 1 : 0;
}
// parent49 = PublicKey::from(parent sk).to_bytes(), computed once by k_parent_public below
__global__ void __launch_bounds__(128) k_derive_private(size_t n, const uint8_t* __restrict__ parent_xsk64,
                                                        const uint8_t* __restrict__ parent49, const uint32_t* __restrict__ indices,
                                                        uint8_t* __restrict__ children64, uint8_t* __restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ok[i] = derive_private_child(parent_xsk64, parent49, indices[i], children64 + 64 * i) ? 1 : 0;
}
__global__ void k_parent_public(const uint8_t* __restrict__ parent_xsk64, const uint64_t* __restrict__ gtab, uint8_t* __restrict__ out49) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    scalar k = sc_from_u256(sc_load_le(parent_xsk64));
    jac_pt P = fixed_base_mul(k, gtab);
    fp6 x, y;
    bool inf;
    jac_to_affine(P, x, y, inf);
    for (int c = 0; c < 6; c++)
        for (int b = 0; b < 8; b++) out49[8 * c + b] = (uint8_t)((inf ? 0 : x.c[c]) >> (8 * b));
    out49[48] = compress_flags(y, inf);
}
// children81 = 49-byte compressed child key || 32-byte chain code; parent96 / parent_inf = the decompressed parent key
__global__ void __launch_bounds__(SIGN_THREADS) k_derive_public(size_t n, const uint8_t* __restrict__ parent_xpk81,
                                                                const uint8_t* __restrict__ parent96,
                                                                const uint8_t* __restrict__ parent_inf,
                                                                const uint32_t* __restrict__ indices,
                                                                const uint64_t* __restrict__ gtab, uint8_t* __restrict__ children81,
                                                                uint8_t* __restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t index = indices[i];
    uint8_t* o = children81 + 81 * i;
    scalar t = derive_public_tweak(parent_xpk81, parent_xpk81 + 49, index, o + 49);
    const uint64_t* p = reinterpret_cast<const uint64_t*>(parent96);
    fp6 px, py;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        px.c[k] = p[k];
        py.c[k] = p[6 + k];
    }
    jac_pt acc = fixed_base_mul(t, gtab);                       // I_L G   (src/derivation.rs:263)
    bool tweak_inf = jac_is_identity(acc);
    acc = jac_madd(acc, px, py, parent_inf[0] != 0);            // + parent key (:264)
    fp6 x, y;
    bool inf;
    jac_to_affine(acc, x, y, inf);
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int b = 0; b < 8; b++) o[8 * k + b] = (uint8_t)((inf ? 0 : x.c[k]) >> (8 * b));
    o[48] = compress_flags(y, inf);
    ok[i] = (!tweak_inf && (index >> 31) == 0) ? 1 : 0;         // :270-275
}

// ------------------------------------------------------------------------------------------------
// fixed-base table of G:  gtab[i][d] = d * 2^(13 i) * G  (affine), i < 20, 1 <= d <= 4096 (curve.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void k_gtab_bases(jac_pt* bases, fp6 gx, fp6 gy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GTAB_WINDOWS) return;
    jac_pt p = jac_from_affine(gx, gy, false);
#pragma unroll 1
    for (int k = 0; k < GTAB_W * i; k++) jac_dbl_mem(&p);
    bases[i] = p;
}
__global__ void __launch_bounds__(128) k_gtab_fill(const jac_pt* __restrict__ bases, uint64_t* __restrict__ gtab) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= GTAB_WINDOWS * GTAB_ENTRIES) return;
    int i = t / GTAB_ENTRIES, b = t % GTAB_ENTRIES;
    jac_pt base = bases[i];
    jac_pt acc = jac_identity();
#pragma unroll 1
    for (int bit = GTAB_W - 1; bit >= 0; bit--) {
        jac_dbl_mem(&acc);
        if ((b >> bit) & 1) jac_add_mem(&acc, &base, false);
    }
    fp6 x, y;
    bool inf;
    jac_to_affine(acc, x, y, inf);  // slot 0 (the identity) is written as zeros and never read
    uint64_t* o = gtab + (size_t)t * 12;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        o[k] = x.c[k];
        o[6 + k] = y.c[k];
    }
}

// ------------------------------------------------------------------------------------------------
// test hook: raw field operations on the device (the PTX paths cannot be exercised on a CPU box)
//   out[i] = { fp6_mul(a,b) (6) | fp6_sqr(a) (6) | a+b (6) | a-b (6) | limb-wise fp_mul (6) | limb-wise fp_sqr(a) (6) |
//              fp6_inv(a) (6) | limb-wise rescue x^(1/7) of a (6) }
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_debug_field(size_t n, const uint64_t* __restrict__ a6, const uint64_t* __restrict__ b6,
                                                     uint64_t* __restrict__ out48) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp6 a, b;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        a.c[k] = a6[i * 6 + k];
        b.c[k] = b6[i * 6 + k];
    }
    fp6 m = fp6_mul(a, b), q = fp6_sqr(a), ad = fp6_add(a, b), su = fp6_sub(a, b), iv = fp6_inv(a);
    uint64_t* o = out48 + i * 48;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        o[k] = m.c[k];
        o[6 + k] = q.c[k];
        o[12 + k] = ad.c[k];
        o[18 + k] = su.c[k];
        o[24 + k] = fp_mul(a.c[k], b.c[k]);
        o[30 + k] = fp_sqr(a.c[k]);
        o[36 + k] = iv.c[k];
        o[42 + k] = rescue_inv_sbox(a.c[k]);
    }
}

// test hook for the lazily reduced forms of the fast path (debug_ops.cuh): out[i] = 8 x 6 limbs
__global__ void __launch_bounds__(128) k_debug_lazy(size_t n, const uint64_t* __restrict__ a6, const uint64_t* __restrict__ b6,
                                                    uint64_t* __restrict__ out48) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp6 a, b, r[8];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        a.c[k] = a6[i * 6 + k];
        b.c[k] = b6[i * 6 + k];
    }
    debug_lazy_ops(a, b, r);
    for (int j = 0; j < 8; j++)
        for (int k = 0; k < 6; k++) out48[i * 48 + 6 * j + k] = r[j].c[k];
}

// ------------------------------------------------------------------------------------------------
// K6: integer-multiply roofline calibration.  8 independent accumulator chains per thread of
// 32x32+64 -> 64 multiply-adds (IMAD.WIDE.U32), no memory traffic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_imad_peak(int iters, uint64_t* sink) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
        }
        b += 3;
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x123456789abcdefULL) sink[0] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline unsigned grid_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static int alloc_soa(schnorr_b200_ctx* ctx, size_t n, soa_batch* b) {
    void *p, *f, *sf;
    if (int rc = ensure_scratch(ctx, SL_A, sizeof(ulonglong2) * SOA_PLANES * n, &p)) return rc;
    if (int rc = ensure_scratch(ctx, SL_B, n, &f)) return rc;
    if (int rc = ensure_scratch(ctx, SL_C, n, &sf)) return rc;
    b->planes = (ulonglong2*)p;
    b->flags = (uint8_t*)f;
    b->sig_flag = (uint8_t*)sf;
    b->n = n;
    return 0;
}

// The offset table of a HOST call must start at 0 and be non-decreasing (the reference's `&[&[u8]]` cannot express
// anything else); a bad table would otherwise turn into out-of-bounds device reads.  O(n) on the host.
static bool msg_off_valid(size_t n, const uint64_t* msg_off) {
    if (msg_off[0] != 0) return false;
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return false;
    return true;
}
#define CHECK_MSG_OFF(ctx, n, msg_off)                                                            \
    do {                                                                                          \
        if (!msg_off_valid((n), (msg_off))) {                                                     \
            (ctx)->err = "message offsets must start at 0 and be non-decreasing";                 \
            return SCHNORR_B200_EARG;                                                             \
        }                                                                                         \
    } while (0)

// host -> device staging helper
static int stage_in(schnorr_b200_ctx* ctx, int slot, const void* host, size_t bytes, void** dev) {
    if (int rc = ensure_scratch(ctx, slot, bytes, dev)) return rc;
    if (host && bytes) CUDA_TRY(ctx, cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

#define MULTI_DISPATCH(ctx, call)                                  \
    do {                                                           \
        if ((ctx) && !(ctx)->shards.empty()) return call;          \
    } while (0)
#define NOT_ON_MULTI(ctx)                                                                                          \
    do {                                                                                                           \
        if ((ctx) && !(ctx)->shards.empty()) {                                                                     \
            (ctx)->err = "device-pointer entry points need a single-device context (buffers live on one device)"; \
            return SCHNORR_B200_EARG;                                                                              \
        }                                                                                                          \
    } while (0)

static int verify_many_host(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sigs81, const uint8_t* pk96, const uint8_t* pk_inf,
                            const uint8_t* msgs, const uint64_t* msg_off, uint8_t* verdicts, bool validate_offsets);
#include "batch.cuh"
#include "multi.cuh"
// K2 for the smallest calls: one thread block per signature (uses the six-lane primitives of dist.cuh and the 24-lane
// Jacobian toolkit of batch.cuh)
#include "one.cuh"


// Signature::verify over an ingested SoA batch: fast path (per-thread or warp-cooperative by call size), then the exact kernel over the handful of
// items it handed back.  `list_base` = first element of this batch in the per-call work list (pipelined chunks
// use disjoint regions), `counter` = index of its counter.
static int launch_verify(schnorr_b200_ctx* ctx, const soa_batch& soa, const uint8_t* msgs, const uint64_t* msg_off,
                         uint8_t* verdicts, size_t list_base, int counter, size_t list_total, cudaStream_t st) {
    unsigned grid = grid_for(soa.n, VERIFY_THREADS);
    if (counter == 0) ctx->exact_counters_used = 0;
    if (ctx->exact_only) {
        k_verify<<<grid, VERIFY_THREADS, 0, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, nullptr, nullptr);
        ctx->launches += 1;
        return 0;
    }
    void* wl;
    if (int rc = ensure_scratch(ctx, SL_M, 4 * (list_total + schnorr_b200_ctx::MAX_CHUNKS), &wl)) return rc;
    uint32_t* counters = (uint32_t*)wl;
    uint32_t* list = counters + schnorr_b200_ctx::MAX_CHUNKS + list_base;
    CUDA_TRY(ctx, cudaMemsetAsync(counters + counter, 0, 4, st));
    if (counter + 1 > ctx->exact_counters_used) ctx->exact_counters_used = counter + 1;
    if (soa.n <= ctx->one_max)
        if (soa.n <= (size_t)2 * ctx->sm_count)   // every block resident at two per SM: the full register budget
            k_verify_one<2><<<(unsigned)soa.n, ONE_THREADS, 0, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, list, counters + counter);
        else
            k_verify_one<4><<<(unsigned)soa.n, ONE_THREADS, 0, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, list, counters + counter);
    else if (soa.n <= ctx->dist_max)
        k_verify_dist<<<grid_for(soa.n, DIST_SIGS_PER_BLOCK), DIST_THREADS, 0, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, list,
                                                                                   counters + counter);
    else
        k_verify_fast<<<grid, VERIFY_THREADS, VERIFY_FAST_SMEM, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, list, counters + counter);
    // the exact kernel sizes itself from the device-side counter: blocks beyond it exit at once
    k_verify<<<grid, VERIFY_THREADS, 0, st>>>(soa, msgs, msg_off, ctx->gtab, verdicts, list, counters + counter);
    ctx->launches += 2;
    return 0;
}

extern "C" {

int schnorr_b200_params_pinned(void) { return CHEETAH_PARAMS_PINNED; }
const char* schnorr_b200_params_provenance(void) { return CHEETAH_PARAMS_PROVENANCE; }

int schnorr_b200_create(int device, schnorr_b200_ctx** out) {
    if (!out) return SCHNORR_B200_EARG;
    *out = nullptr;
#if !CHEETAH_PARAMS_PINNED
    {   // once per process: results are bit-exact against the RESTATED oracle only
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "schnorr_b200: curve / hash parameters are NOT pinned to the upstream cheetah / hash crates (%s). "
                            "Keys and signatures interoperate with the real schnorr-sig crate only after "
                            "rust/dump_params has been run and include/cheetah_params.h regenerated.\n",
                    CHEETAH_PARAMS_PROVENANCE);
        }
    }
#endif
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return SCHNORR_B200_ENODEV;
    }
    schnorr_b200_ctx* ctx = new schnorr_b200_ctx();
    ctx->device = device;
    auto fail = [&](int rc) {
        fprintf(stderr, "schnorr_b200_create: %s\n", ctx->err.c_str());
        schnorr_b200_destroy(ctx);
        return rc;
    };
#define CREATE_TRY(expr)                                                          \
    do {                                                                          \
        cudaError_t e_ = (expr);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e_);        \
            return fail(SCHNORR_B200_ECUDA);                                      \
        }                                                                         \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    {
        int per_sm = 0;
        CREATE_TRY(cudaFuncSetAttribute(k_verify_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VERIFY_FAST_SMEM));
        CREATE_TRY(cudaFuncSetAttribute(k_verify_fast, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_verify_fast, VERIFY_THREADS, VERIFY_FAST_SMEM));
        if (per_sm < 1) per_sm = 1;
        ctx->verify_wave = (size_t)ctx->sm_count * per_sm * VERIFY_THREADS;
    }
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    for (int c = 0; c < schnorr_b200_ctx::MAX_CHUNKS; c++) CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_chunk[c], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreate(&ctx->ev_k0));
    CREATE_TRY(cudaEventCreate(&ctx->ev_k1));
    CREATE_TRY(cudaMemcpyToSymbol(c_ark, RESCUE_ARK, sizeof(uint64_t) * 2 * RESCUE_ROUNDS * 12));
    CREATE_TRY(cudaMemcpyToSymbol(c_q_wnaf5, CHEETAH_Q_WNAF5, sizeof(CHEETAH_Q_WNAF5)));  // the tail of the symbol stays zero
    CREATE_TRY(cudaMemcpyToSymbol(c_q_wnaf4, CHEETAH_Q_WNAF4, 256));
    CREATE_TRY(cudaMalloc(&ctx->gtab, GTAB_U64 * sizeof(uint64_t)));
    jac_pt* bases = nullptr;
    CREATE_TRY(cudaMalloc(&bases, sizeof(jac_pt) * GTAB_WINDOWS));
    fp6 gx, gy;
    memcpy(gx.c, CHEETAH_GX, 48);
    memcpy(gy.c, CHEETAH_GY, 48);
    k_gtab_bases<<<1, 32, 0, ctx->stream>>>(bases, gx, gy);
    k_gtab_fill<<<(GTAB_WINDOWS * GTAB_ENTRIES + 127) / 128, 128, 0, ctx->stream>>>(bases, ctx->gtab);
    ctx->launches += 2;
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaStreamSynchronize(ctx->stream));
    cudaFree(bases);
#undef CREATE_TRY
    *out = ctx;
    return SCHNORR_B200_OK;
}

int schnorr_b200_create_multi(const int* devices, int n_devices, schnorr_b200_ctx** out) {
    if (!out) return SCHNORR_B200_EARG;
    *out = nullptr;
    if (!devices || n_devices <= 0 || n_devices > 64) return SCHNORR_B200_EARG;
    if (n_devices == 1) return schnorr_b200_create(devices[0], out);
    schnorr_b200_ctx* ctx = new schnorr_b200_ctx();
    ctx->device = devices[0];
    for (int k = 0; k < n_devices; k++) {
        schnorr_b200_ctx* sh = nullptr;
        int rc = schnorr_b200_create(devices[k], &sh);
        if (rc != SCHNORR_B200_OK) {
            schnorr_b200_destroy(ctx);
            return rc;
        }
        ctx->shards.push_back(sh);
    }
    ctx->sm_count = ctx->shards[0]->sm_count;
    ctx->verify_wave = ctx->shards[0]->verify_wave;
    *out = ctx;
    return SCHNORR_B200_OK;
}
int schnorr_b200_device_count(const schnorr_b200_ctx* ctx) { return !ctx ? 0 : (ctx->shards.empty() ? 1 : (int)ctx->shards.size()); }

void schnorr_b200_destroy(schnorr_b200_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->shards.empty()) {
        for (schnorr_b200_ctx* sh : ctx->shards) schnorr_b200_destroy(sh);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    for (int s = 0; s < SL_COUNT; s++)
        if (ctx->scratch[s]) cudaFree(ctx->scratch[s]);
    if (ctx->gtab) cudaFree(ctx->gtab);
    for (int c = 0; c < schnorr_b200_ctx::MAX_CHUNKS; c++)
        if (ctx->ev_chunk[c]) cudaEventDestroy(ctx->ev_chunk[c]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
    if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* schnorr_b200_last_error(const schnorr_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int schnorr_b200_set_stream(schnorr_b200_ctx* ctx, void* stream) {
    if (!ctx) return SCHNORR_B200_EARG;
    NOT_ON_MULTI(ctx);
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return SCHNORR_B200_OK;
}
int schnorr_b200_synchronize(schnorr_b200_ctx* ctx) {
    if (!ctx) return SCHNORR_B200_EARG;
    for (schnorr_b200_ctx* sh : ctx->shards)
        if (int rc = schnorr_b200_synchronize(sh)) return rc;
    if (!ctx->shards.empty()) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}
uint64_t schnorr_b200_launch_count(const schnorr_b200_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t total = ctx->launches;
    for (const schnorr_b200_ctx* sh : ctx->shards) total += sh->launches;
    return total;
}
int schnorr_b200_last_kernel_ms(schnorr_b200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return SCHNORR_B200_EARG;
    if (!ctx->shards.empty()) {  // the slowest device
        *ms = 0;
        for (schnorr_b200_ctx* sh : ctx->shards) {
            float m = 0;
            if (int rc = schnorr_b200_last_kernel_ms(sh, &m)) return rc;
            if (m > *ms) *ms = m;
        }
        return SCHNORR_B200_OK;
    }
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_k1));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
    return SCHNORR_B200_OK;
}

int schnorr_b200_set_exact_only(schnorr_b200_ctx* ctx, int exact_only) {
    if (!ctx) return SCHNORR_B200_EARG;
    for (schnorr_b200_ctx* sh : ctx->shards) sh->exact_only = exact_only != 0;
    ctx->exact_only = exact_only != 0;
    return SCHNORR_B200_OK;
}
int schnorr_b200_set_dist_threshold(schnorr_b200_ctx* ctx, size_t max_signatures) {
    if (!ctx) return SCHNORR_B200_EARG;
    for (schnorr_b200_ctx* sh : ctx->shards) sh->dist_max = max_signatures;
    ctx->dist_max = max_signatures;
    return SCHNORR_B200_OK;
}
int schnorr_b200_set_one_threshold(schnorr_b200_ctx* ctx, size_t max_signatures) {
    if (!ctx) return SCHNORR_B200_EARG;
    for (schnorr_b200_ctx* sh : ctx->shards) sh->one_max = max_signatures;
    ctx->one_max = max_signatures;
    return SCHNORR_B200_OK;
}
int schnorr_b200_set_batch_small_threshold(schnorr_b200_ctx* ctx, size_t max_signatures) {
    if (!ctx) return SCHNORR_B200_EARG;
    for (schnorr_b200_ctx* sh : ctx->shards) sh->batch_small_max = max_signatures;
    ctx->batch_small_max = max_signatures;
    return SCHNORR_B200_OK;
}
int schnorr_b200_set_batch_dist_threshold(schnorr_b200_ctx* ctx, size_t max_signatures) {
    if (!ctx) return SCHNORR_B200_EARG;
    ctx->batch_dist_max = max_signatures;
    for (schnorr_b200_ctx* sh : ctx->shards) sh->batch_dist_max = max_signatures;
    return SCHNORR_B200_OK;
}
int schnorr_b200_set_msm_geometry(schnorr_b200_ctx* ctx, int window_bits, unsigned segment_len) {
    if (!ctx || (window_bits != 0 && (window_bits < 4 || window_bits > 16)) || (segment_len != 0 && segment_len < 8))
        return SCHNORR_B200_EARG;
    ctx->msm_c_override = window_bits;
    ctx->msm_t_override = segment_len;
    for (schnorr_b200_ctx* sh : ctx->shards) {
        sh->msm_c_override = window_bits;
        sh->msm_t_override = segment_len;
    }
    return SCHNORR_B200_OK;
}
int schnorr_b200_last_batch_plan(const schnorr_b200_ctx* ctx, int* window_bits, int* windows, unsigned* segment_len) {
    if (!ctx) return SCHNORR_B200_EARG;
    const schnorr_b200_ctx* c = ctx->shards.empty() ? ctx : ctx->shards[0];
    if (window_bits) *window_bits = c->last_msm_c;
    if (windows) *windows = c->last_msm_K;
    if (segment_len) *segment_len = c->last_msm_T;
    return SCHNORR_B200_OK;
}
int schnorr_b200_last_exact_count(schnorr_b200_ctx* ctx, uint64_t* count) {
    if (!ctx || !count) return SCHNORR_B200_EARG;
    *count = 0;
    if (!ctx->shards.empty()) {
        for (schnorr_b200_ctx* sh : ctx->shards) {
            uint64_t c = 0;
            if (int rc = schnorr_b200_last_exact_count(sh, &c)) return rc;
            *count += c;
        }
        return SCHNORR_B200_OK;
    }
    if (ctx->exact_counters_used == 0 || !ctx->scratch[SL_M]) return SCHNORR_B200_OK;
    uint32_t c[schnorr_b200_ctx::MAX_CHUNKS] = {};
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(c, ctx->scratch[SL_M], 4 * ctx->exact_counters_used, cudaMemcpyDeviceToHost));
    for (int i = 0; i < ctx->exact_counters_used; i++) *count += c[i];
    return SCHNORR_B200_OK;
}

// ---- hash_messages ---------------------------------------------------------------------------
int schnorr_b200_hash_messages_dev(schnorr_b200_ctx* ctx, size_t n, const uint8_t* rx48, const uint8_t* pk96,
                                   const uint8_t* msgs, const uint64_t* msg_off, uint8_t* digests) {
    NOT_ON_MULTI(ctx);
    if (!ctx || (n && (!rx48 || !pk96 || !msg_off || !digests))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    if (n <= ctx->dist_max)
        k_hash_dist<<<grid_for(n, DIST_SIGS_PER_BLOCK), DIST_THREADS, 0, ctx->stream>>>(n, rx48, pk96, msgs, msg_off, digests);
    else
        k_hash<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, ctx->stream>>>(n, rx48, pk96, msgs, msg_off, digests);
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SCHNORR_B200_OK;
}

int schnorr_b200_hash_messages(schnorr_b200_ctx* ctx, size_t n, const uint8_t* rx48, const uint8_t* pk96,
                               const uint8_t* msgs, const uint64_t* msg_off, uint8_t* digests) {
    MULTI_DISPATCH(ctx, multi_hash_messages(ctx, n, rx48, pk96, msgs, msg_off, digests));
    if (!ctx || (n && (!rx48 || !pk96 || !msg_off || !digests))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CHECK_MSG_OFF(ctx, n, msg_off);
    size_t mb = msg_off[n];
    if (mb && !msgs) return SCHNORR_B200_EARG;
    void *d_rx, *d_pk, *d_m, *d_off, *d_out;
    if (int rc = stage_in(ctx, SL_D, rx48, n * 48, &d_rx)) return rc;
    if (int rc = stage_in(ctx, SL_E, pk96, n * 96, &d_pk)) return rc;
    if (int rc = stage_in(ctx, SL_F, msgs, mb, &d_m)) return rc;
    if (int rc = stage_in(ctx, SL_G, msg_off, (n + 1) * 8, &d_off)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n * 32, &d_out)) return rc;
    if (int rc = schnorr_b200_hash_messages_dev(ctx, n, (uint8_t*)d_rx, (uint8_t*)d_pk, (uint8_t*)d_m, (uint64_t*)d_off,
                                                (uint8_t*)d_out))
        return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(digests, d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

// ---- verify_many -----------------------------------------------------------------------------
int schnorr_b200_verify_many_dev(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sigs81, const uint8_t* pk96,
                                 const uint8_t* pk_inf, const uint8_t* msgs, const uint64_t* msg_off,
                                 uint8_t* verdicts) {
    NOT_ON_MULTI(ctx);
    if (!ctx || (n && (!sigs81 || !pk96 || !msg_off || !verdicts))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    soa_batch soa;
    if (int rc = alloc_soa(ctx, n, &soa)) return rc;
    k_ingest<<<grid_for(n, INGEST_THREADS), INGEST_THREADS, 0, ctx->stream>>>(n, sigs81, pk96, pk_inf, soa);
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    if (int rc = launch_verify(ctx, soa, msgs, msg_off, verdicts, 0, 0, n, ctx->stream)) return rc;
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SCHNORR_B200_OK;
}

// Host entry point.  Large calls are pipelined: the batch is cut into up to 8 chunks; chunk c+1 is copied
// host->device on a second stream while chunk c is being verified, and each chunk's verdicts are copied
// back as soon as its kernel ends, so the PCIe time hides behind the kernels.
int schnorr_b200_verify_many(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sigs81, const uint8_t* pk96,
                             const uint8_t* pk_inf, const uint8_t* msgs, const uint64_t* msg_off, uint8_t* verdicts) {
    MULTI_DISPATCH(ctx, multi_verify_many(ctx, n, sigs81, pk96, pk_inf, msgs, msg_off, verdicts));
    if (!ctx || (n && (!sigs81 || !pk96 || !msg_off || !verdicts))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    if (msg_off[0] != 0 || msg_off[n] < msg_off[0]) {
        ctx->err = "message offsets must start at 0 and be non-decreasing";
        return SCHNORR_B200_EARG;
    }
    // the rest of the table is validated chunk by chunk inside the pipeline, behind the kernels of the chunks before
    // (one pass over 2^20 offsets on the calling thread, in front of the first copy, was 1 % of the call)
    return verify_many_host(ctx, n, sigs81, pk96, pk_inf, msgs, msg_off, verdicts, true);
}
} // extern "C" (the internal host pipeline below has C++ linkage; it is declared before multi.cuh)

// The pipelined host path on ONE device.  `msg_off` is a table of n + 1 offsets into `msgs` whose first entry may be
// non-zero (the multi-device layer hands every shard its slice of the caller's table as it is) and whose END POINTS the
// caller has checked (msg_off[0] <= msg_off[n] <= size of the blob).  With `validate_offsets` every chunk's part of the
// table is checked right before that chunk is enqueued -- non-decreasing and not beyond msg_off[n], so that no copy and
// no kernel of the chunk leaves the blob -- while the GPU is busy with the chunks before; a bad table ends the call with
// EARG after the work already enqueued has drained.
static int verify_many_host(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sigs81, const uint8_t* pk96,
                            const uint8_t* pk_inf, const uint8_t* msgs, const uint64_t* msg_off, uint8_t* verdicts,
                            bool validate_offsets) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t base = msg_off[0];
    size_t mb = msg_off[n] - base;
    if (mb && !msgs) return SCHNORR_B200_EARG;
    void *d_sig, *d_pk, *d_inf = nullptr, *d_m, *d_off, *d_out;
    if (int rc = ensure_scratch(ctx, SL_D, n * 81, &d_sig)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 96, &d_pk)) return rc;
    if (int rc = ensure_scratch(ctx, SL_F, mb, &d_m)) return rc;
    if (int rc = ensure_scratch(ctx, SL_G, (n + 1) * 8, &d_off)) return rc;
    if (pk_inf)
        if (int rc = ensure_scratch(ctx, SL_I, n, &d_inf)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n, &d_out)) return rc;
    soa_batch soa;
    if (int rc = alloc_soa(ctx, n, &soa)) return rc;
    // chunk plan in whole kernel waves (no partial-wave tail between chunks): 1, 3, 8 waves, then the rest.
    // Every chunk's copy hides behind the kernels of the chunks before it as long as the host link sustains
    // ~6 GB/s (a wave of 37 888 signatures = 7 MB of input takes ~3.5 ms to verify).
    size_t bounds[5] = {0, n, n, n, n};
    int chunks = 1;
    {
        const size_t w = ctx->verify_wave;
        const size_t cum[3] = {w, 4 * w, 12 * w};
        int k = 0;
        while (k < 3 && n > cum[k] + 2 * w) {
            bounds[k + 1] = cum[k];
            k++;
        }
        bounds[k + 1] = n;
        chunks = k + 1;
    }
    cudaStream_t cs = ctx->copy_stream, ks = ctx->stream, ks2 = ctx->aux_stream;
    // neither the copy stream nor the second compute stream may overtake work of a previous call that still reads the
    // scratch buffers
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_chunk[0], ks));
    CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->ev_chunk[0], 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ks2, ctx->ev_chunk[0], 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_off, msg_off, (n + 1) * 8, cudaMemcpyHostToDevice, cs));
    for (int c = 0; c < chunks; c++) {
        size_t lo = bounds[c], hi = bounds[c + 1], cn = hi - lo;
        if (validate_offsets) {
            uint64_t bad = msg_off[hi] > msg_off[n];
            for (size_t i = lo; i < hi; i++) bad |= msg_off[i + 1] < msg_off[i];   // branch-free: vectorises
            if (bad) {
                cudaStreamSynchronize(cs);
                cudaStreamSynchronize(ks2);
                cudaStreamSynchronize(ks);
                ctx->err = "message offsets must start at 0 and be non-decreasing";
                return SCHNORR_B200_EARG;
            }
        }
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)d_sig + 81 * lo, sigs81 + 81 * lo, cn * 81, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)d_pk + 96 * lo, pk96 + 96 * lo, cn * 96, cudaMemcpyHostToDevice, cs));
        if (pk_inf) CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)d_inf + lo, pk_inf + lo, cn, cudaMemcpyHostToDevice, cs));
        size_t b0 = msg_off[lo], b1 = msg_off[hi];
        if (b1 > b0) CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)d_m + (b0 - base), msgs + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_chunk[c], cs));
        // consecutive chunks run on alternating compute streams: the blocks of chunk c + 1 start on the SMs that chunk c
        // is draining instead of waiting for its last block (chunks touch disjoint regions of every buffer)
        cudaStream_t st = (c & 1) ? ks2 : ks;
        CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
        soa_batch sc;
        sc.planes = soa.planes + (size_t)SOA_PLANES * lo;   // a private [11][cn] region per chunk
        sc.flags = soa.flags + lo;
        sc.sig_flag = soa.sig_flag + lo;
        sc.n = cn;
        k_ingest<<<grid_for(cn, INGEST_THREADS), INGEST_THREADS, 0, st>>>(cn, (uint8_t*)d_sig + 81 * lo, (uint8_t*)d_pk + 96 * lo,
                                                                         pk_inf ? (uint8_t*)d_inf + lo : nullptr, sc);
        if (st == ks) cudaEventRecord(ctx->ev_k0, ks);
        // the kernels index the message blob with the caller's absolute offsets: shift the device base accordingly
        if (int rc = launch_verify(ctx, sc, (uint8_t*)d_m - base, (uint64_t*)d_off + lo, (uint8_t*)d_out + lo, lo, c, n, st)) return rc;
        if (st == ks) cudaEventRecord(ctx->ev_k1, ks);
        ctx->launches += 1;
    }
    if (chunks > 1) {  // the main stream continues after the second one
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_aux, ks2));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ks, ctx->ev_aux, 0));
    }
    // verdicts come back after EVERY chunk has been enqueued: a device-to-host copy into pageable memory blocks the
    // host until the kernels before it have finished, which would serialise the pipeline above (1 byte per signature:
    // nothing to overlap anyway)
    CUDA_TRY(ctx, cudaMemcpyAsync(verdicts, d_out, n, cudaMemcpyDeviceToHost, ks));
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ks));
    return SCHNORR_B200_OK;
}
extern "C" {

// ---- KeyedSignature::verify over wire records -------------------------------------------------
int schnorr_b200_verify_keyed_many_dev(schnorr_b200_ctx* ctx, size_t n, const uint8_t* keyed130, const uint8_t* msgs,
                                       const uint64_t* msg_off, uint8_t* verdicts) {
    NOT_ON_MULTI(ctx);
    if (!ctx || (n && (!keyed130 || !msg_off || !verdicts))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_pk, *d_inf, *d_ok, *d_sig;
    if (int rc = ensure_scratch(ctx, SL_J, n * 96, &d_pk)) return rc;
    if (int rc = ensure_scratch(ctx, SL_K, n * 2, &d_inf)) return rc;
    if (int rc = ensure_scratch(ctx, SL_L, n * 81 + 256, &d_sig)) return rc;
    d_ok = (uint8_t*)d_inf + n;
    k_split_keyed<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, keyed130, (uint8_t*)d_pk, (uint8_t*)d_inf, (uint8_t*)d_ok,
                                                           (uint8_t*)d_sig);
    ctx->launches += 1;
    if (int rc = schnorr_b200_verify_many_dev(ctx, n, (uint8_t*)d_sig, (uint8_t*)d_pk, (uint8_t*)d_inf, msgs, msg_off, verdicts))
        return rc;
    k_mask_verdicts<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, (uint8_t*)d_ok, verdicts);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SCHNORR_B200_OK;
}
int schnorr_b200_verify_keyed_many(schnorr_b200_ctx* ctx, size_t n, const uint8_t* keyed130, const uint8_t* msgs,
                                   const uint64_t* msg_off, uint8_t* verdicts) {
    MULTI_DISPATCH(ctx, multi_verify_keyed_many(ctx, n, keyed130, msgs, msg_off, verdicts));
    if (!ctx || (n && (!keyed130 || !msg_off || !verdicts))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CHECK_MSG_OFF(ctx, n, msg_off);
    size_t mb = msg_off[n];
    if (mb && !msgs) return SCHNORR_B200_EARG;
    void *d_rec, *d_m, *d_off, *d_out;
    if (int rc = stage_in(ctx, SL_D, keyed130, n * 130, &d_rec)) return rc;
    if (int rc = stage_in(ctx, SL_F, msgs, mb, &d_m)) return rc;
    if (int rc = stage_in(ctx, SL_G, msg_off, (n + 1) * 8, &d_off)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n, &d_out)) return rc;
    if (int rc = schnorr_b200_verify_keyed_many_dev(ctx, n, (uint8_t*)d_rec, (uint8_t*)d_m, (uint64_t*)d_off, (uint8_t*)d_out))
        return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(verdicts, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

// ---- keygen / sign ---------------------------------------------------------------------------
int schnorr_b200_keygen_dev(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sk32, uint8_t* pk96, uint8_t* pk_inf) {
    NOT_ON_MULTI(ctx);
    if (!ctx || (n && (!sk32 || !pk96))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    k_keygen<<<grid_for(n, SIGN_THREADS), SIGN_THREADS, 0, ctx->stream>>>(n, sk32, ctx->gtab, pk96, pk_inf);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SCHNORR_B200_OK;
}
int schnorr_b200_keygen(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sk32, uint8_t* pk96, uint8_t* pk_inf) {
    MULTI_DISPATCH(ctx, multi_keygen(ctx, n, sk32, pk96, pk_inf));
    if (!ctx || (n && (!sk32 || !pk96))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_sk, *d_pk, *d_inf;
    if (int rc = stage_in(ctx, SL_D, sk32, n * 32, &d_sk)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 96, &d_pk)) return rc;
    if (int rc = ensure_scratch(ctx, SL_I, n, &d_inf)) return rc;
    if (int rc = schnorr_b200_keygen_dev(ctx, n, (uint8_t*)d_sk, (uint8_t*)d_pk, (uint8_t*)d_inf)) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(d_sk, 0, n * 32, ctx->stream));  // the staging arena is reused: do not keep secret keys in it
    CUDA_TRY(ctx, cudaMemcpyAsync(pk96, d_pk, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    if (pk_inf) CUDA_TRY(ctx, cudaMemcpyAsync(pk_inf, d_inf, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}
int schnorr_b200_sign_many_dev(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sk32, const uint8_t* pk96,
                               const uint8_t* pk_inf, const uint8_t* msgs, const uint64_t* msg_off,
                               const uint8_t* nonce32, uint8_t* sigs81) {
    NOT_ON_MULTI(ctx);
    if (!ctx || (n && (!sk32 || !pk96 || !msg_off || !nonce32 || !sigs81))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    k_sign<<<grid_for(n, SIGN_THREADS), SIGN_THREADS, 0, ctx->stream>>>(n, sk32, pk96, pk_inf, msgs, msg_off, nonce32,
                                                                        ctx->gtab, sigs81);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SCHNORR_B200_OK;
}
int schnorr_b200_sign_many(schnorr_b200_ctx* ctx, size_t n, const uint8_t* sk32, const uint8_t* pk96,
                           const uint8_t* pk_inf, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* nonce32,
                           uint8_t* sigs81) {
    MULTI_DISPATCH(ctx, multi_sign_many(ctx, n, sk32, pk96, pk_inf, msgs, msg_off, nonce32, sigs81));
    if (!ctx || (n && (!sk32 || !pk96 || !msg_off || !nonce32 || !sigs81))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CHECK_MSG_OFF(ctx, n, msg_off);
    size_t mb = msg_off[n];
    if (mb && !msgs) return SCHNORR_B200_EARG;
    void *d_sk, *d_pk, *d_inf = nullptr, *d_m, *d_off, *d_nonce, *d_out;
    if (int rc = stage_in(ctx, SL_D, sk32, n * 32, &d_sk)) return rc;
    if (int rc = stage_in(ctx, SL_E, pk96, n * 96, &d_pk)) return rc;
    if (int rc = stage_in(ctx, SL_F, msgs, mb, &d_m)) return rc;
    if (int rc = stage_in(ctx, SL_G, msg_off, (n + 1) * 8, &d_off)) return rc;
    if (pk_inf)
        if (int rc = stage_in(ctx, SL_I, pk_inf, n, &d_inf)) return rc;
    if (int rc = stage_in(ctx, SL_J, nonce32, n * 32, &d_nonce)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n * 81, &d_out)) return rc;
    if (int rc = schnorr_b200_sign_many_dev(ctx, n, (uint8_t*)d_sk, (uint8_t*)d_pk, (uint8_t*)d_inf, (uint8_t*)d_m,
                                            (uint64_t*)d_off, (uint8_t*)d_nonce, (uint8_t*)d_out))
        return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(d_sk, 0, n * 32, ctx->stream));      // secret keys and nonces do not stay in the
    CUDA_TRY(ctx, cudaMemsetAsync(d_nonce, 0, n * 32, ctx->stream));   // reusable staging arena
    CUDA_TRY(ctx, cudaMemcpyAsync(sigs81, d_out, n * 81, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

// ---- f4: hierarchical deterministic derivation ---------------------------------------------------------------
int schnorr_b200_derive_master_keys(schnorr_b200_ctx* ctx, size_t n, const uint8_t* seeds32, uint8_t* xsk64, uint8_t* ok) {
    if (!ctx || (n && (!seeds32 || !xsk64 || !ok))) return SCHNORR_B200_EARG;
    MULTI_DISPATCH(ctx, schnorr_b200_derive_master_keys(ctx->shards[0], n, seeds32, xsk64, ok));
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_in, *d_out, *d_ok;
    if (int rc = stage_in(ctx, SL_D, seeds32, n * 32, &d_in)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 64, &d_out)) return rc;
    if (int rc = ensure_scratch(ctx, SL_I, n, &d_ok)) return rc;
    k_derive_master<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint8_t*)d_in, (uint8_t*)d_out, (uint8_t*)d_ok);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemsetAsync(d_in, 0, n * 32, ctx->stream));   // seeds are secrets
    CUDA_TRY(ctx, cudaMemcpyAsync(xsk64, d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, n * 64, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}
int schnorr_b200_derive_private_children(schnorr_b200_ctx* ctx, size_t n, const uint8_t* parent_xsk64, const uint32_t* indices,
                                         uint8_t* children_xsk64, uint8_t* ok) {
    if (!ctx || !parent_xsk64 || (n && (!indices || !children_xsk64 || !ok))) return SCHNORR_B200_EARG;
    MULTI_DISPATCH(ctx, schnorr_b200_derive_private_children(ctx->shards[0], n, parent_xsk64, indices, children_xsk64, ok));
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_par, *d_idx, *d_out, *d_ok;
    if (int rc = ensure_scratch(ctx, SL_D, 128, &d_par)) return rc;          // [0,64) parent, [64,128) its public key bytes
    CUDA_TRY(ctx, cudaMemcpyAsync(d_par, parent_xsk64, 64, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = stage_in(ctx, SL_F, indices, n * 4, &d_idx)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 64, &d_out)) return rc;
    if (int rc = ensure_scratch(ctx, SL_I, n, &d_ok)) return rc;
    uint8_t* d_pk49 = (uint8_t*)d_par + 64;
    k_parent_public<<<1, 32, 0, ctx->stream>>>((uint8_t*)d_par, ctx->gtab, d_pk49);
    k_derive_private<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint8_t*)d_par, d_pk49, (uint32_t*)d_idx, (uint8_t*)d_out,
                                                               (uint8_t*)d_ok);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(children_xsk64, d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_par, 0, 128, ctx->stream));              // secrets do not stay in the arena
    CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, n * 64, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}
int schnorr_b200_derive_public_children(schnorr_b200_ctx* ctx, size_t n, const uint8_t* parent_xpk81, const uint32_t* indices,
                                        uint8_t* children_xpk81, uint8_t* ok) {
    if (!ctx || !parent_xpk81 || (n && (!indices || !children_xpk81 || !ok))) return SCHNORR_B200_EARG;
    MULTI_DISPATCH(ctx, schnorr_b200_derive_public_children(ctx->shards[0], n, parent_xpk81, indices, children_xpk81, ok));
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_par, *d_idx, *d_out, *d_ok;
    // parent record (81) | decompressed key (96) | identity flag (1) | decode flag (1), 16-byte aligned pieces
    if (int rc = ensure_scratch(ctx, SL_D, 96 + 96 + 16 + 16, &d_par)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_par, parent_xpk81, 81, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = stage_in(ctx, SL_F, indices, n * 4, &d_idx)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 81, &d_out)) return rc;
    if (int rc = ensure_scratch(ctx, SL_I, n, &d_ok)) return rc;
    uint8_t *d_pk96 = (uint8_t*)d_par + 96, *d_inf = d_pk96 + 96, *d_dec = d_inf + 16;
    k_decompress<<<1, 128, 0, ctx->stream>>>(1, (uint8_t*)d_par, d_pk96, d_inf, d_dec);   // ExtendedPublicKey::from_bytes (:295-311)
    k_derive_public<<<grid_for(n, SIGN_THREADS), SIGN_THREADS, 0, ctx->stream>>>(n, (uint8_t*)d_par, d_pk96, d_inf, (uint32_t*)d_idx,
                                                                                ctx->gtab, (uint8_t*)d_out, (uint8_t*)d_ok);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    uint8_t dec = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&dec, d_dec, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(children_xpk81, d_out, n * 81, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!dec) {  // the parent key does not decode: no child exists
        ctx->err = "parent extended public key does not decode";
        return SCHNORR_B200_EARG;
    }
    return SCHNORR_B200_OK;
}

// ---- codecs ----------------------------------------------------------------------------------
int schnorr_b200_decompress(schnorr_b200_ctx* ctx, size_t n, const uint8_t* in49, uint8_t* pk96, uint8_t* pk_inf,
                            uint8_t* ok) {
    MULTI_DISPATCH(ctx, schnorr_b200_decompress(ctx->shards[0], n, in49, pk96, pk_inf, ok));
    if (!ctx || (n && (!in49 || !pk96 || !pk_inf || !ok))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_in, *d_pk, *d_inf, *d_ok;
    if (int rc = stage_in(ctx, SL_D, in49, n * 49, &d_in)) return rc;
    if (int rc = ensure_scratch(ctx, SL_E, n * 96, &d_pk)) return rc;
    if (int rc = ensure_scratch(ctx, SL_I, n, &d_inf)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n, &d_ok)) return rc;
    k_decompress<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint8_t*)d_in, (uint8_t*)d_pk, (uint8_t*)d_inf,
                                                            (uint8_t*)d_ok);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(pk96, d_pk, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(pk_inf, d_inf, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}
int schnorr_b200_compress(schnorr_b200_ctx* ctx, size_t n, const uint8_t* pk96, const uint8_t* pk_inf, uint8_t* out49) {
    MULTI_DISPATCH(ctx, schnorr_b200_compress(ctx->shards[0], n, pk96, pk_inf, out49));
    if (!ctx || (n && (!pk96 || !out49))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_pk, *d_inf = nullptr, *d_out;
    if (int rc = stage_in(ctx, SL_E, pk96, n * 96, &d_pk)) return rc;
    if (pk_inf)
        if (int rc = stage_in(ctx, SL_I, pk_inf, n, &d_inf)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n * 49, &d_out)) return rc;
    k_compress<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint8_t*)d_pk, (uint8_t*)d_inf, (uint8_t*)d_out);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out49, d_out, n * 49, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

// ---- test hook -------------------------------------------------------------------------------
int schnorr_b200_debug_field_ops(schnorr_b200_ctx* ctx, size_t n, const uint64_t* a6, const uint64_t* b6, uint64_t* out48) {
    MULTI_DISPATCH(ctx, schnorr_b200_debug_field_ops(ctx->shards[0], n, a6, b6, out48));
    if (!ctx || (n && (!a6 || !b6 || !out48))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_a, *d_b, *d_o;
    if (int rc = stage_in(ctx, SL_D, a6, n * 48, &d_a)) return rc;
    if (int rc = stage_in(ctx, SL_E, b6, n * 48, &d_b)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n * 48 * 8, &d_o)) return rc;
    k_debug_field<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint64_t*)d_a, (uint64_t*)d_b, (uint64_t*)d_o);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out48, d_o, n * 48 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

int schnorr_b200_debug_lazy_ops(schnorr_b200_ctx* ctx, size_t n, const uint64_t* a6, const uint64_t* b6, uint64_t* out48) {
    MULTI_DISPATCH(ctx, schnorr_b200_debug_lazy_ops(ctx->shards[0], n, a6, b6, out48));
    if (!ctx || (n && (!a6 || !b6 || !out48))) return SCHNORR_B200_EARG;
    if (n == 0) return SCHNORR_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void *d_a, *d_b, *d_o;
    if (int rc = stage_in(ctx, SL_D, a6, n * 48, &d_a)) return rc;
    if (int rc = stage_in(ctx, SL_E, b6, n * 48, &d_b)) return rc;
    if (int rc = ensure_scratch(ctx, SL_H, n * 48 * 8, &d_o)) return rc;
    k_debug_lazy<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, (uint64_t*)d_a, (uint64_t*)d_b, (uint64_t*)d_o);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out48, d_o, n * 48 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SCHNORR_B200_OK;
}

// ---- roofline calibration --------------------------------------------------------------------
int schnorr_b200_imad_peak(schnorr_b200_ctx* ctx, int iters, double* wide_mul_per_s, double* elapsed_ms) {
    MULTI_DISPATCH(ctx, schnorr_b200_imad_peak(ctx->shards[0], iters, wide_mul_per_s, elapsed_ms));
    if (!ctx || iters <= 0 || !wide_mul_per_s) return SCHNORR_B200_EARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* sink;
    if (int rc = ensure_scratch(ctx, SL_H, 64, &sink)) return rc;
    int blocks = ctx->sm_count * 8;
    cudaEvent_t a, b;
    CUDA_TRY(ctx, cudaEventCreate(&a));
    CUDA_TRY(ctx, cudaEventCreate(&b));
    k_imad_peak<<<blocks, 256, 0, ctx->stream>>>(iters / 8 + 1, (uint64_t*)sink);  // warm-up
    CUDA_TRY(ctx, cudaEventRecord(a, ctx->stream));
    k_imad_peak<<<blocks, 256, 0, ctx->stream>>>(iters, (uint64_t*)sink);
    CUDA_TRY(ctx, cudaEventRecord(b, ctx->stream));
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    double total = (double)blocks * 256.0 * (double)iters * 8.0;
    *wide_mul_per_s = total / (ms * 1e-3);
    if (elapsed_ms) *elapsed_ms = ms;
    return SCHNORR_B200_OK;
}

}  // extern "C"
