// Multi-device context through the plain C ABI (include/schnorr_b200.h: schnorr_b200_create_multi): the host entry
// points shard a call over every listed device; the results must equal those of a single-device context bit for bit
// (verdicts of Signature::verify, src/signature.rs:181-205; both points of verify_batch, src/batch.rs:84-130).
//   g++ -std=c++17 -Iinclude examples/multi_device_example.cpp -Lschnorr-sig_b200/csrc -lschnorr_b200 -lcudart
// usage: multi_device_example [n_signatures] [n_devices]   (devices beyond the box's count wrap around: the same GPU
// listed twice gives two independent shards, which is how a single-GPU box exercises the sharded paths)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime_api.h>

#include "schnorr_b200.h"

#define CHECK(expr)                                                                                    \
    do {                                                                                               \
        int rc_ = (expr);                                                                              \
        if (rc_ != 0) {                                                                                \
            std::printf("FAILED %s -> %d (%s / %s)\n", #expr, rc_, schnorr_b200_last_error(multi),     \
                        schnorr_b200_last_error(single));                                              \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

int main(int argc, char** argv) {
    size_t n = argc > 1 ? (size_t)std::atoll(argv[1]) : 20000;
    int want = argc > 2 ? std::atoi(argv[2]) : 2;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        std::printf("no CUDA device\n");
        return 2;
    }
    std::vector<int> devices;
    for (int k = 0; k < want; k++) devices.push_back(k % count);
    schnorr_b200_ctx *multi = nullptr, *single = nullptr;
    CHECK(schnorr_b200_create(0, &single));
    CHECK(schnorr_b200_create_multi(devices.data(), (int)devices.size(), &multi));
    if (schnorr_b200_device_count(multi) != want || schnorr_b200_device_count(single) != 1) return 1;

    std::mt19937_64 gen(7);
    std::vector<uint8_t> sk(n * 32), nonce(n * 32), rnd(n * 32), pk(n * 96), inf(n), sigs(n * 81), msgs;
    std::vector<uint64_t> off(n + 1, 0);
    for (auto& b : sk) b = (uint8_t)gen();
    for (auto& b : nonce) b = (uint8_t)gen();
    for (auto& b : rnd) b = (uint8_t)gen();
    for (size_t i = 0; i < n; i++) {
        sk[32 * i + 31] &= 0x3f;  // < 2^254 < q
        nonce[32 * i + 31] &= 0x3f;
        rnd[32 * i + 31] &= 0x3f;
        size_t len = gen() % 40;  // ragged messages, including empty ones
        off[i + 1] = off[i] + len;
        for (size_t k = 0; k < len; k++) msgs.push_back((uint8_t)gen());
    }
    if (msgs.empty()) msgs.push_back(0);
    // sharded key generation and signing must equal the single-device results
    std::vector<uint8_t> pk1(n * 96), inf1(n), sigs1(n * 81);
    CHECK(schnorr_b200_keygen(multi, n, sk.data(), pk.data(), inf.data()));
    CHECK(schnorr_b200_keygen(single, n, sk.data(), pk1.data(), inf1.data()));
    CHECK(schnorr_b200_sign_many(multi, n, sk.data(), pk.data(), inf.data(), msgs.data(), off.data(), nonce.data(), sigs.data()));
    CHECK(schnorr_b200_sign_many(single, n, sk.data(), pk.data(), inf.data(), msgs.data(), off.data(), nonce.data(), sigs1.data()));
    bool ok = pk == pk1 && inf == inf1 && sigs == sigs1;
    if (!ok) std::printf("keygen / sign differ between the multi- and the single-device context\n");
    // corrupt a few signatures on both sides of every slice boundary
    std::vector<uint8_t> bad = sigs;
    for (size_t i = 0; i < n; i += n / 13 + 1) bad[81 * i + 49] ^= 1;
    for (int k = 1; k < want; k++) {
        size_t b = n * k / want;
        if (b > 0 && b < n) { bad[81 * (b - 1) + 50] ^= 4; bad[81 * b + 51] ^= 8; }  // e only: R must stay decompressible
    }
    std::vector<uint8_t> vm(n, 255), vs(n, 254), dm(n * 32), ds(n * 32, 1);
    CHECK(schnorr_b200_verify_many(multi, n, bad.data(), pk.data(), inf.data(), msgs.data(), off.data(), vm.data()));
    CHECK(schnorr_b200_verify_many(single, n, bad.data(), pk.data(), inf.data(), msgs.data(), off.data(), vs.data()));
    size_t rejected = 0;
    for (size_t i = 0; i < n; i++) rejected += vm[i] != 0;
    if (!(vm == vs && rejected > 0 && rejected < n)) {
        std::printf("verify_many: verdicts differ (rejected %zu of %zu)\n", rejected, n);
        for (size_t i = 0, shown = 0; i < n && shown < 8; i++)
            if (vm[i] != vs[i]) std::printf("  item %zu: multi %d single %d\n", i, vm[i], vs[i]), shown++;
        ok = false;
    }
    std::vector<uint8_t> rx(n * 48);
    for (size_t i = 0; i < n; i++) std::memcpy(&rx[48 * i], &sigs[81 * i], 48);
    CHECK(schnorr_b200_hash_messages(multi, n, rx.data(), pk.data(), msgs.data(), off.data(), dm.data()));
    CHECK(schnorr_b200_hash_messages(single, n, rx.data(), pk.data(), msgs.data(), off.data(), ds.data()));
    if (dm != ds) std::printf("hash_messages differ\n"), ok = false;
    // batch: both points identical, valid batch accepted, corrupted batch rejected
    int v_m = -1, v_s = -1, v_bad = -1;
    uint8_t lm[97], rm[97], ls[97], rs[97];
    CHECK(schnorr_b200_verify_batch(multi, n, sigs.data(), pk.data(), inf.data(), msgs.data(), off.data(), rnd.data(), &v_m, lm, rm));
    CHECK(schnorr_b200_verify_batch(single, n, sigs.data(), pk.data(), inf.data(), msgs.data(), off.data(), rnd.data(), &v_s, ls, rs));
    CHECK(schnorr_b200_verify_batch(multi, n, bad.data(), pk.data(), inf.data(), msgs.data(), off.data(), rnd.data(), &v_bad, nullptr, nullptr));
    if (!(v_m == 0 && v_s == 0 && v_bad == 2 && !std::memcmp(lm, ls, 97) && !std::memcmp(rm, rs, 97) && !std::memcmp(lm, rm, 48))) {
        std::printf("verify_batch: verdicts multi %d single %d corrupted %d, lhs equal %d, rhs equal %d\n", v_m, v_s, v_bad,
                    !std::memcmp(lm, ls, 97), !std::memcmp(rm, rs, 97));
        ok = false;
    }
    // the device-pointer entry points refuse a multi-device context; a bad offset table is an argument error
    uint8_t dummy[192];
    if (schnorr_b200_batch_partial_dev(multi, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dummy) != SCHNORR_B200_EARG)
        std::printf("a *_dev entry point accepted a multi-device context\n"), ok = false;
    std::vector<uint64_t> bad_off = off;
    if (n > 2) bad_off[2] = bad_off[1] + (uint64_t(1) << 40), bad_off[3 < n ? 3 : n] = 0;
    if (schnorr_b200_verify_many(multi, n, sigs.data(), pk.data(), inf.data(), msgs.data(), bad_off.data(), vm.data()) != SCHNORR_B200_EARG ||
        schnorr_b200_verify_many(single, n, sigs.data(), pk.data(), inf.data(), msgs.data(), bad_off.data(), vs.data()) != SCHNORR_B200_EARG)
        std::printf("a decreasing offset table was not rejected\n"), ok = false;
    std::printf("%s\n", ok ? "ok" : "FAILED");
    schnorr_b200_destroy(multi);
    schnorr_b200_destroy(single);
    return ok ? 0 : 1;
}
