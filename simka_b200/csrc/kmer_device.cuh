// kmer_device.cuh -- span-32 k-mer model shared by host and device code.
//
// Mirrors the arithmetic of GATB-core's Kmer<32>::ModelCanonical
// (GC/kmer/impl/Model.hpp:637-657,858-884; GC/tools/math/LargeInt1.pri:137-155;
// nucleotide code (c>>1)&3 => A=0 C=1 T=2 G=3, GC/tools/misc/api/Data.hpp:179) on 2-bit packed
// streams instead of ASCII, plus the invertible bucket hash this build uses in place of the
// minimizer repartition table (any deterministic function of the canonical k-mer gives the same
// matrices: SURVEY.md 8a row 4, reference tests/simple_test.py:122-129).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SIMKA_HD __host__ __device__ __forceinline__
#else
#define SIMKA_HD inline
#endif

namespace simka {

// reverse-complement of a k-mer stored in the low 2k bits (first base most significant)
SIMKA_HD uint64_t revcomp(uint64_t x, int k) {
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
#ifdef __CUDA_ARCH__
    x = ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
#else
    x = __builtin_bswap64(x);
#endif
    x ^= 0xAAAAAAAAAAAAAAAAull;  // complement: A<->T (0<->2), C<->G (1<->3)
    return x >> (64 - 2 * k);
}

// Invertible mixing of a 2k-bit value: xorshift, odd multiply modulo 2^(2k), xorshift (one round: the extraction
// kernel is ALU-bound and every input bit still reaches the top bits through the carry chain of the multiply).
// Top bits select the bucket, so canonical-k-mer skew (min(fwd,rev) favours small values) and
// low-complexity repeats spread evenly over buckets.
struct Mix {
    uint64_t mask;   // 2^(2k)-1
    uint64_t m1, m2; // odd multipliers
    uint64_t i1, i2; // their inverses modulo 2^64 (valid modulo 2^(2k) too)
    int shift;       // ceil(2k/2): x ^= x>>shift is an involution on 2k bits
    int bits;        // 2k
};

SIMKA_HD uint64_t mix_fwd(uint64_t x, const Mix &p) {
    x ^= x >> p.shift;
    x = (x * p.m1) & p.mask;
    x ^= x >> p.shift;
    return x;
}
SIMKA_HD uint64_t mix_inv(uint64_t x, const Mix &p) {
    x ^= x >> p.shift;
    x = (x * p.i1) & p.mask;
    x ^= x >> p.shift;
    return x;
}

inline uint64_t inv_odd_u64(uint64_t a) {  // Newton iteration, a odd
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
inline Mix make_mix(int k) {
    Mix p;
    p.bits = 2 * k;
    p.mask = (p.bits == 64) ? ~0ull : ((1ull << p.bits) - 1);
    p.shift = (p.bits + 1) / 2;
    p.m1 = 0xFF51AFD7ED558CCDull;
    p.m2 = 0xC4CEB9FE1A85EC53ull;
    p.i1 = inv_odd_u64(p.m1);
    p.i2 = inv_odd_u64(p.m2);
    return p;
}

// MurmurHash3_x64_128 (Austin Appleby, public domain) of the 8 bytes of a k-mer value, first output word: the hash
// SimkaMin sketches with (src/simkaMin/SimkaMinCount.hpp:247-250, src/simkaMin/MurmurHash3.cpp:246-330)
SIMKA_HD uint64_t murmur3_fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
SIMKA_HD uint64_t murmur3_x64_128_u64(uint64_t key, uint32_t seed) {
    uint64_t h1 = seed, h2 = seed, k1 = key;
    k1 *= 0x87c37b91114253d5ull; k1 = (k1 << 31) | (k1 >> 33); k1 *= 0x4cf5ad432745937full; h1 ^= k1;
    h1 ^= 8; h2 ^= 8;
    h1 += h2; h2 += h1;
    h1 = murmur3_fmix64(h1); h2 = murmur3_fmix64(h2);
    return h1 + h2;
}

// secondary hash used for table slots / sub-pass selection / GPU shard choice
SIMKA_HD uint64_t hash2(uint64_t h) { return h * 0x9E3779B97F4A7C15ull; }

}  // namespace simka
