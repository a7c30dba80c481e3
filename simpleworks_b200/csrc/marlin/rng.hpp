// Deterministic randomness of the reference stack, host side.
//
//   ChaChaRng        rand_chacha 0.3 ChaCha{12,20}Rng over rand_core::block::BlockRng (64 u32 words
//                    buffered; next_u64 straddles refills the way BlockRng does)
//   test_rng()       ark_std::test_rng() = StdRng(ChaCha12) from the fixed seed; simpleworks'
//                    generate_rand() (reference src/marlin/mod.rs:33-35)
//   Blake2s          blake2 0.9 Blake2s (32-byte digest, no key)
//   FiatShamirRng    ark_marlin::rng::SimpleHashFiatShamirRng<Blake2s, ChaChaRng>
//                    (reference src/marlin/mod.rs:13: `FS`)
//   rand_fr / ...    ark_ff `UniformRand for Fp*`: rejection sampling on raw limbs, kept AS the
//                    Montgomery representation
// Word streams are pinned by tests/golden/rng.json (big-int golden model).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../fp.cuh"

namespace swb {
namespace marlin {

class ChaChaRng {
public:
    ChaChaRng() : rounds_(20) { memset(key_, 0, sizeof key_); }
    ChaChaRng(const uint8_t seed[32], int rounds) : rounds_(rounds) {
        for (int i = 0; i < 8; i++)
            key_[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                      ((uint32_t)seed[4 * i + 3] << 24);
    }
    // BlockRng buffers 64 words (4 ChaCha blocks) and next_u64 takes two consecutive words, also
    // across a refill (index == 63: low word, refill, high word) -- so the output is one linear
    // word stream and the generator state is just a position in it.
    uint32_t next_u32() {
        const uint64_t g = pos_ >> 6;
        if (g != group_) load_group(g);
        return buf_[pos_++ & 63];
    }
    uint64_t next_u64() {
        const uint64_t lo = next_u32();
        const uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
    bool next_bool() { return (int32_t)next_u32() < 0; }            // rand 0.8 Standard for bool
    // rand 0.8 `Standard` for u128: low word first
    void next_u128(uint64_t out[2]) {
        out[0] = next_u64();
        out[1] = next_u64();
    }
    // words [pos, pos + count) of the stream without consuming them (ChaCha is counter based, so
    // blocks are generated in parallel); seek() then moves the position
    void peek_words(uint32_t* out, size_t count) const {
        const uint64_t first_blk = pos_ >> 4, last_blk = (pos_ + count + 15) >> 4;
        const size_t nblk = (size_t)(last_blk - first_blk);
        std::vector<uint32_t> tmp(nblk * 16);
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < nblk; b++) block(first_blk + b, tmp.data() + 16 * b);
        memcpy(out, tmp.data() + (pos_ & 15), count * sizeof(uint32_t));
    }
    uint64_t position() const { return pos_; }
    void seek(uint64_t pos) { pos_ = pos; }
    const uint32_t* key_words() const { return key_; }      // for engines that expand the stream themselves
    int rounds() const { return rounds_; }

private:
    static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
    void block(uint64_t counter, uint32_t* out) const {
        uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        for (int i = 0; i < 8; i++) st[4 + i] = key_[i];
        st[12] = (uint32_t)counter;
        st[13] = (uint32_t)(counter >> 32);
        st[14] = 0;
        st[15] = 0;
        uint32_t x[16];
        memcpy(x, st, sizeof x);
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
            x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
            x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        };
        for (int r = 0; r < rounds_ / 2; r++) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
    }
    void load_group(uint64_t g) {
        for (int i = 0; i < 4; i++) block(4 * g + i, buf_ + 16 * i);
        group_ = g;
    }
    uint32_t key_[8];
    int rounds_;
    uint64_t pos_ = 0;                 // next word of the linear stream
    uint64_t group_ = ~0ull;           // which 64-word group buf_ holds
    uint32_t buf_[64];
};

inline ChaChaRng test_rng() {
    const uint8_t seed[32] = {1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0};
    return ChaChaRng(seed, 12);
}

// ark_ff Fp::rand: raw limbs, top bits shaved, rejected when >= p; the accepted raw value IS the
// Montgomery representation.
template <class P, int SHAVE>
inline Fp<P> rand_fp(ChaChaRng& rng) {
    constexpr int N64 = P::N / 2;
    for (;;) {
        uint64_t w[N64];
        for (int i = 0; i < N64; i++) w[i] = rng.next_u64();
        w[N64 - 1] &= 0xFFFFFFFFFFFFFFFFull >> SHAVE;
        Fp<P> r;
        for (int i = 0; i < N64; i++) {
            r.l[2 * i] = (uint32_t)w[i];
            r.l[2 * i + 1] = (uint32_t)(w[i] >> 32);
        }
        // r < p ?
        bool lt = false;
        for (int i = P::N - 1; i >= 0; i--) {
            if (r.l[i] != P::mod(i)) { lt = r.l[i] < P::mod(i); break; }
        }
        if (lt) return r;
    }
}
inline Fr rand_fr(ChaChaRng& rng) { return rand_fp<FrParams, 3>(rng); }
// n consecutive Fr::rand draws (DensePolynomial::rand): identical stream consumption, but the
// ChaCha blocks are produced in parallel and only the accept/reject walk is sequential
inline void rand_fr_bulk(ChaChaRng& rng, Fr* out, size_t n) {
    size_t done = 0;
    std::vector<uint32_t> words;
    while (done < n) {
        size_t want = n - done;
        if (want > (1u << 17)) want = 1u << 17;             // batches that stay in cache
        const size_t cand = want + want + 64;               // acceptance is ~0.58
        words.resize(cand * 8);
        rng.peek_words(words.data(), words.size());
        size_t used = 0;
        for (size_t k = 0; k < cand && done < n; k++) {
            const uint32_t* w = &words[8 * k];
            used = k + 1;
            Fr r;
            for (int i = 0; i < 8; i++) r.l[i] = w[i];
            r.l[7] &= 0xFFFFFFFFu >> 3;
            bool lt = false;
            for (int i = 7; i >= 0; i--)
                if (r.l[i] != FrParams::mod(i)) { lt = r.l[i] < FrParams::mod(i); break; }
            if (lt) out[done++] = r;
        }
        rng.seek(rng.position() + 8 * used);
    }
}
inline Fq rand_fq(ChaChaRng& rng) { return rand_fp<FqParams, 7>(rng); }

// ---- BLAKE2s-256 (RFC 7693), unkeyed ------------------------------------------------------------
class Blake2s {
public:
    Blake2s() {
        static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        memcpy(h_, iv, sizeof h_);
        h_[0] ^= 0x01010020u;   // digest length 32, fanout 1, depth 1
    }
    void update(const uint8_t* p, size_t n) {
        while (n) {
            if (fill_ == 64) {
                t_ += 64;
                compress(false);
                fill_ = 0;
            }
            size_t k = 64 - fill_ < n ? 64 - fill_ : n;
            memcpy(buf_ + fill_, p, k);
            fill_ += k;
            p += k;
            n -= k;
        }
    }
    void finalize(uint8_t out[32]) {
        t_ += fill_;
        memset(buf_ + fill_, 0, 64 - fill_);
        compress(true);
        for (int i = 0; i < 8; i++)
            for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(h_[i] >> (8 * b));
    }
    static void digest(const std::vector<uint8_t>& in, uint8_t out[32]) {
        Blake2s s;
        s.update(in.data(), in.size());
        s.finalize(out);
    }

private:
    static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(bool last) {
        static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        static const uint8_t sigma[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; i++)
            m[i] = (uint32_t)buf_[4 * i] | ((uint32_t)buf_[4 * i + 1] << 8) | ((uint32_t)buf_[4 * i + 2] << 16) |
                   ((uint32_t)buf_[4 * i + 3] << 24);
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[8 + i] = iv[i]; }
        v[12] ^= (uint32_t)t_;
        v[13] ^= (uint32_t)(t_ >> 32);
        if (last) v[14] = ~v[14];
        auto g = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; r++) {
            const uint8_t* s = sigma[r];
            g(0, 4, 8, 12, m[s[0]], m[s[1]]);   g(1, 5, 9, 13, m[s[2]], m[s[3]]);
            g(2, 6, 10, 14, m[s[4]], m[s[5]]);  g(3, 7, 11, 15, m[s[6]], m[s[7]]);
            g(0, 5, 10, 15, m[s[8]], m[s[9]]);  g(1, 6, 11, 12, m[s[10]], m[s[11]]);
            g(2, 7, 8, 13, m[s[12]], m[s[13]]); g(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[8 + i];
    }
    uint32_t h_[8];
    uint8_t buf_[64];
    size_t fill_ = 0;
    uint64_t t_ = 0;
};

// ark_marlin::rng::SimpleHashFiatShamirRng<Blake2s, ChaChaRng>:
//   initialize(x): seed = H(x);              rng = ChaCha20(seed)
//   absorb(x):     seed = H(x || seed);      rng = ChaCha20(seed)
class FiatShamirRng {
public:
    void initialize(const std::vector<uint8_t>& bytes) {
        Blake2s::digest(bytes, seed_);
        rng_ = ChaChaRng(seed_, 20);
    }
    void absorb(const std::vector<uint8_t>& bytes) {
        std::vector<uint8_t> in(bytes);
        in.insert(in.end(), seed_, seed_ + 32);
        Blake2s::digest(in, seed_);
        rng_ = ChaChaRng(seed_, 20);
    }
    ChaChaRng& rng() { return rng_; }

private:
    uint8_t seed_[32];
    ChaChaRng rng_;
};

}  // namespace marlin
}  // namespace swb
