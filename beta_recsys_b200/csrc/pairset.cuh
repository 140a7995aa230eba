// (user, item) pair set in device memory: open addressing over 64-bit keys, at most half full.
// Shared by the negative sampler (sample_kernels.cu) -- membership = "user has interacted with item".
#pragma once
#include "common.cuh"

#define BRS_PAIR_EMPTY 0xffffffffffffffffull

__host__ __device__ inline unsigned long long brs_mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__host__ __device__ inline long long brs_pairset_capacity(long long n_pairs) {
    long long c = 1024;
    while (c < 2 * n_pairs) c <<= 1;
    return c;
}

__device__ __forceinline__ unsigned long long brs_pair_key(long long u, long long i) {
    return ((unsigned long long)u << 32) | (unsigned long long)i;
}

__device__ __forceinline__ void brs_pairset_insert(unsigned long long* keys, long long cap, unsigned long long key) {
    unsigned long long h = brs_mix64(key) & (unsigned long long)(cap - 1);
    for (;;) {
        const unsigned long long old = atomicCAS(keys + h, BRS_PAIR_EMPTY, key);
        if (old == BRS_PAIR_EMPTY || old == key) return;
        h = (h + 1) & (unsigned long long)(cap - 1);
    }
}

__device__ __forceinline__ bool brs_pairset_contains(const unsigned long long* keys, long long cap, unsigned long long key) {
    unsigned long long h = brs_mix64(key) & (unsigned long long)(cap - 1);
    for (;;) {
        const unsigned long long got = __ldg(keys + h);
        if (got == key) return true;
        if (got == BRS_PAIR_EMPTY) return false;
        h = (h + 1) & (unsigned long long)(cap - 1);
    }
}
