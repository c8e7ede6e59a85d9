// phj_hash.cuh -- the key hashers as bit-exact __host__ __device__ functions.
//
// Replaces Common::IHasher / Common::XXHasher (reference src/Common/IHasher.hpp:6-11,
// src/Common/XXHasher.hpp:10-28): XXHasher::Hash is XXH3_64bits_withSeed(&key, 8, seed) %
// cardinality, with the arithmetic inside third-party xxHash (not vendored by the reference). The
// 8-byte path of the frozen XXH3 (xxHash >= 0.8.0) is XXH3_len_4to8_64b + XXH3_rrmxmx; it is
// written out here so it can run on the device. Murmur3 and City do not exist in the reference;
// they are the extensions include/phj.h names.
//
// A hasher is split into a seed-only part computed once on the host (HashParams) and the per-key
// part, so that a kernel spends no instructions on the seed.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHJ_HD __host__ __device__ __forceinline__
#else
#define PHJ_HD inline
#endif

namespace phj {

enum HashId : int { kXXH3 = 0, kMurmur3 = 1, kCity = 2 };

struct HashParams {
    uint64_t k0;  // xxh3: bitflip; murmur3: (uint32)seed; city: seed
};

PHJ_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

PHJ_HD uint32_t bswap32(uint32_t x) {
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
}

PHJ_HD HashParams make_hash_params(int hash, uint64_t seed) {
    HashParams p;
    if (hash == kXXH3) {
        // XXH3_len_4to8_64b: seed ^= (u64)swap32((u32)seed) << 32;
        //                    bitflip = (readLE64(secret+8) ^ readLE64(secret+16)) - seed
        const uint64_t secret8 = 0x1cad21f72c81017cULL, secret16 = 0xdb979083e96dd4deULL;
        uint64_t s = seed ^ ((uint64_t)bswap32((uint32_t)seed) << 32);
        p.k0 = (secret8 ^ secret16) - s;
    } else if (hash == kMurmur3) {
        p.k0 = (uint32_t)seed;
    } else {
        p.k0 = seed;
    }
    return p;
}

// XXH3_64bits_withSeed(&key, 8, seed) given the precomputed bitflip.
PHJ_HD uint64_t xxh3_64_len8(uint64_t key, uint64_t bitflip) {
    const uint64_t kPrimeMx2 = 0x9FB21C651E98DF25ULL;
    uint64_t h = ((key >> 32) | (key << 32)) ^ bitflip;  // input2 + (input1 << 32)
    h ^= rotl64(h, 49) ^ rotl64(h, 24);                  // XXH3_rrmxmx
    h *= kPrimeMx2;
    h ^= (h >> 35) + 8;  // len == 8
    h *= kPrimeMx2;
    h ^= h >> 28;
    return h;
}

PHJ_HD uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// h1 of MurmurHash3_x64_128(&key, 8, seed32): no 16-byte body block, an 8-byte tail, finalisation.
PHJ_HD uint64_t murmur3_64_len8(uint64_t key, uint64_t seed32) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed32, h2 = seed32;
    uint64_t k1 = key * c1;
    k1 = rotl64(k1, 31);
    k1 *= c2;
    h1 ^= k1;
    h1 ^= 8;
    h2 ^= 8;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

PHJ_HD uint64_t city_hash_len16(uint64_t u, uint64_t v) {
    const uint64_t mul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (u ^ v) * mul;
    a ^= a >> 47;
    uint64_t b = (v ^ a) * mul;
    b ^= b >> 47;
    return b * mul;
}

// CityHash64WithSeed(&key, 8, seed), v1.0 short-key path: HashLen16(len + (lo32 << 3), hi32),
// then HashLen16(h - k2, seed).
PHJ_HD uint64_t city_64_len8(uint64_t key, uint64_t seed) {
    const uint64_t k2 = 0x9ae16a3b2f90404fULL;
    uint64_t h = city_hash_len16(8 + ((key & 0xffffffffULL) << 3), key >> 32);
    return city_hash_len16(h - k2, seed);
}

template <int HASH>
PHJ_HD uint64_t hash_key(uint64_t key, const HashParams& p) {
    if (HASH == kXXH3) return xxh3_64_len8(key, p.k0);
    if (HASH == kMurmur3) return murmur3_64_len8(key, p.k0);
    return city_64_len8(key, p.k0);
}

PHJ_HD uint64_t hash_key_dyn(int hash, uint64_t key, const HashParams& p) {
    if (hash == kXXH3) return xxh3_64_len8(key, p.k0);
    if (hash == kMurmur3) return murmur3_64_len8(key, p.k0);
    return city_64_len8(key, p.k0);
}

}  // namespace phj
