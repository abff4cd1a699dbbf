// Example / compile check of the C++ facade (include/schnorr_b200.hpp): reads like the reference's
// own tests (src/signature.rs:333-360, src/batch.rs:138-150).  Needs a CUDA device at run time.
//   g++ -std=c++17 -Iinclude examples/verify_example.cpp -Lschnorr-sig_b200/csrc -lschnorr_b200 -o verify_example
#include <cstdio>
#include <random>

#include "schnorr_b200.hpp"

using namespace schnorr_sig;

int main() {
    Engine& eng = Engine::instance();
    std::mt19937_64 gen(1);
    Rng rng = [&](uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) p[i] = (uint8_t)gen(); };

    // key pair + signature from the device signer (KeyPair::new / KeyPair::sign)
    uint8_t sk[32], nonce[32], inf = 0;
    rng(sk, 32); rng(nonce, 32); sk[31] &= 0x3f; nonce[31] &= 0x3f;
    PublicKey pkey;
    eng.check(schnorr_b200_keygen(eng.get(), 1, sk, pkey.xy.data(), &inf), "keygen");
    std::vector<uint8_t> message = {'M', 'e', 's', 's', 'a', 'g', 'e', '1'};
    uint64_t off[2] = {0, message.size()};
    std::array<uint8_t, 81> raw{};
    eng.check(schnorr_b200_sign_many(eng.get(), 1, sk, pkey.xy.data(), &inf, message.data(), off, nonce, raw.data()), "sign");
    Signature signature = Signature::from_raw(raw);

    bool ok = !signature.verify(message, pkey).has_value();                       // assert!(signature.verify(..).is_ok())
    ok &= !pkey.verify_signature(signature, message).has_value();
    std::vector<uint8_t> wrong = message; wrong[0] = 42;
    Result r = signature.verify(wrong, pkey);
    ok &= r.has_value() && *r == SignatureError::InvalidSignature;
    ok &= !verify_batch({signature}, {pkey}, {message}, rng).has_value();
    auto pk2 = PublicKey::from_bytes(pkey.to_bytes());
    ok &= pk2.has_value() && pk2->xy == pkey.xy;
    std::printf("%s\n", ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
}
