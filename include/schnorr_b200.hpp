// schnorr_b200.hpp -- header-only C++ host facade over the C ABI (schnorr_b200.h).
//
// The reference is compiled code (Rust); no Rust toolchain exists on this image, so the host side
// above the C ABI is written in C++ and mirrors the reference's public interface for the hot path:
// same names, argument meaning and error behaviour.
//
//   schnorr_sig::Signature::verify(message, pkey)            src/signature.rs:181-205
//   schnorr_sig::PublicKey::verify_signature(sig, message)   src/signature.rs:170-176
//   schnorr_sig::KeyedSignature::verify(message)             src/signature.rs:232-234
//   schnorr_sig::verify_batch(signatures, public_keys, messages, rng)   src/batch.rs:31-50
//   schnorr_sig::SignatureError {InvalidPublicKey, InvalidSignature}     src/error.rs:13-31
//   to_bytes()/from_bytes() of Signature / PublicKey          src/signature.rs:208-227, src/public.rs:49-56
//
// Result<(), SignatureError> is std::optional<SignatureError> (nullopt = Ok).  Inputs on which the
// reference panics (unwrap of a non-canonical encoding, length-mismatch asserts) throw std::logic_error.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "schnorr_b200.h"

namespace schnorr_sig {

constexpr size_t SCALAR_LENGTH = 32, BASEFIELD_LENGTH = 48, PUBLIC_KEY_LENGTH = 49, SIGNATURE_LENGTH = 81,
                 KEYED_SIGNATURE_LENGTH = 130;  // src/constants.rs:12-33

enum class SignatureError { InvalidPublicKey = 1, InvalidSignature = 2 };
inline const char* to_string(SignatureError e) {
    return e == SignatureError::InvalidPublicKey ? "The public key is not an element of the prime subgroup."
                                                 : "The signature is invalid or was incorrectly computed.";
}
using Result = std::optional<SignatureError>;  // nullopt == Ok(())

// One engine (CUDA context, tables, scratch) per process and device; not thread-safe per instance.
class Engine {
public:
    explicit Engine(int device = 0) {
        if (schnorr_b200_create(device, &ctx_) != SCHNORR_B200_OK || !ctx_)
            throw std::runtime_error("schnorr_b200_create failed: a CUDA device is required (no CPU fallback)");
    }
    ~Engine() { schnorr_b200_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    schnorr_b200_ctx* get() const { return ctx_; }
    static Engine& instance() {
        static Engine e(0);
        return e;
    }
    void check(int rc, const char* what) const {
        if (rc != SCHNORR_B200_OK) throw std::runtime_error(std::string(what) + ": " + schnorr_b200_last_error(ctx_));
    }

private:
    schnorr_b200_ctx* ctx_ = nullptr;
};

inline Result result_from_verdict(int v) {
    switch (v) {
        case 0: return std::nullopt;
        case 1: return SignatureError::InvalidPublicKey;
        case 2: return SignatureError::InvalidSignature;
        default: throw std::logic_error("called `Option::unwrap()` on a `None` value (non-canonical encoding)");
    }
}

struct Signature;

// PublicKey(AffinePoint): affine x || y little-endian limbs + identity flag (src/public.rs:24)
struct PublicKey {
    std::array<uint8_t, 96> xy{};
    bool infinity = false;

    std::array<uint8_t, PUBLIC_KEY_LENGTH> to_bytes() const {
        std::array<uint8_t, PUBLIC_KEY_LENGTH> out{};
        uint8_t inf = infinity;
        Engine::instance().check(schnorr_b200_compress(Engine::instance().get(), 1, xy.data(), &inf, out.data()), "compress");
        return out;
    }
    // CtOption<Self>
    static std::optional<PublicKey> from_bytes(const std::array<uint8_t, PUBLIC_KEY_LENGTH>& b) {
        PublicKey k;
        uint8_t inf = 0, ok = 0;
        Engine::instance().check(schnorr_b200_decompress(Engine::instance().get(), 1, b.data(), k.xy.data(), &inf, &ok), "decompress");
        if (!ok) return std::nullopt;
        k.infinity = inf != 0;
        return k;
    }
    Result verify_signature(const Signature& signature, const std::vector<uint8_t>& message) const;
};

// Signature { x: CompressedPoint, e: Scalar } (src/signature.rs:34-40)
struct Signature {
    std::array<uint8_t, 49> x{};
    std::array<uint8_t, 32> e{};

    std::array<uint8_t, SIGNATURE_LENGTH> to_bytes() const {
        std::array<uint8_t, SIGNATURE_LENGTH> out{};
        std::memcpy(out.data(), x.data(), 49);
        std::memcpy(out.data() + 49, e.data(), 32);
        return out;
    }
    static Signature from_raw(const std::array<uint8_t, SIGNATURE_LENGTH>& b) {
        Signature s;
        std::memcpy(s.x.data(), b.data(), 49);
        std::memcpy(s.e.data(), b.data() + 49, 32);
        return s;
    }
    Result verify(const std::vector<uint8_t>& message, const PublicKey& pkey) const {
        auto sig = to_bytes();
        uint64_t off[2] = {0, message.size()};
        uint8_t inf = pkey.infinity, verdict = 255;
        Engine& eng = Engine::instance();
        eng.check(schnorr_b200_verify_many(eng.get(), 1, sig.data(), pkey.xy.data(), &inf,
                                           message.empty() ? nullptr : message.data(), off, &verdict),
                  "verify_many");
        return result_from_verdict(verdict);
    }
};

inline Result PublicKey::verify_signature(const Signature& signature, const std::vector<uint8_t>& message) const {
    return signature.verify(message, *this);
}

struct KeyedSignature {
    PublicKey public_key;
    Signature signature;
    Result verify(const std::vector<uint8_t>& message) const { return signature.verify(message, public_key); }
};

// rng(buf, len) fills len random bytes (CryptoRng + RngCore).  One full-width random scalar is drawn
// per signature (src/batch.rs:75-78); 32 bytes with the top two bits cleared are uniform below 2^254 < q.
using Rng = std::function<void(uint8_t*, size_t)>;

inline Result verify_batch(const std::vector<Signature>& signatures, const std::vector<PublicKey>& public_keys,
                           const std::vector<std::vector<uint8_t>>& messages, const Rng& rng) {
    if (signatures.size() != public_keys.size())
        throw std::logic_error("We should have the same number of signatures than public keys");
    if (messages.size() != public_keys.size())
        throw std::logic_error("We should have the same number of messages than public keys");
    size_t n = signatures.size();
    std::vector<uint8_t> sigs(n * 81), pks(n * 96), inf(n), rand(n * 32), blob;
    std::vector<uint64_t> off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        auto b = signatures[i].to_bytes();
        std::memcpy(&sigs[i * 81], b.data(), 81);
        std::memcpy(&pks[i * 96], public_keys[i].xy.data(), 96);
        inf[i] = public_keys[i].infinity;
        blob.insert(blob.end(), messages[i].begin(), messages[i].end());
        off[i + 1] = blob.size();
    }
    rng(rand.data(), rand.size());
    for (size_t i = 0; i < n; i++) rand[i * 32 + 31] &= 0x3f;
    int verdict = -1;
    Engine& eng = Engine::instance();
    eng.check(schnorr_b200_verify_batch(eng.get(), n, sigs.data(), pks.data(), inf.data(), blob.empty() ? nullptr : blob.data(),
                                        off.data(), rand.data(), &verdict, nullptr, nullptr),
              "verify_batch");
    return result_from_verdict(verdict);
}

}  // namespace schnorr_sig
