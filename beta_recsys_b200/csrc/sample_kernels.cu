// Negative sampling on the GPU -- sm_100a.  (include/brs_b200.h: brs_pairset_build / brs_sample_negatives)
//
// Replaces the sampling half of BaseData.instance_bpr_loader / instance_bce_loader
// (beta_rec/data/base_data.py:182-253): per user a Python set of ALL items minus the user's positives
// (O(U * I) memory and time -- it cannot build BASELINE configs 2-5), then random.sample of 1 (BPR) or
// num_negative DISTINCT (BCE) items per training row.  Same distribution here -- uniform over the items the
// user has not interacted with, negatives of one row pairwise distinct -- by rejection:
//     candidate(row, t, attempt) = mix64(seed + row * C1 + t * C2 + attempt * C3) mod n_items
// is accepted unless (user, candidate) is in the pair set of the training interactions or equals an earlier
// negative of the same row.  The random stream is a pure function of (seed, row, t, attempt): the output
// does not depend on the launch shape and oracle/sample_oracle.py reproduces it bit for bit.
#include "pairset.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxNeg = 64;
constexpr int kMaxAttempts = 1 << 14;
constexpr unsigned long long C1 = 0x9e3779b97f4a7c15ull, C2 = 0xd1b54a32d192ed03ull, C3 = 0x8cb92ba72f3d8dd7ull;

__global__ void __launch_bounds__(kThreads) pairset_build_kernel(unsigned long long* keys, long long cap, const long long* users,
                                                                 const long long* items, long long n, long long n_users,
                                                                 long long n_items, unsigned int* status) {
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < n; r += (long long)gridDim.x * kThreads) {
        const long long u = users[r], i = items[r];
        if ((unsigned long long)u >= (unsigned long long)n_users || (unsigned long long)i >= (unsigned long long)n_items) {
            atomicOr(status, 1u);
            continue;
        }
        brs_pairset_insert(keys, cap, brs_pair_key(u, i));
    }
}

__global__ void __launch_bounds__(kThreads) sample_negatives_kernel(const unsigned long long* keys, long long cap,
                                                                    const long long* users, long long n, long long n_items,
                                                                    int num_negative, unsigned long long seed,
                                                                    long long* neg, unsigned int* status) {
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < n; r += (long long)gridDim.x * kThreads) {
        const long long u = users[r];
        long long* mine = neg + r * num_negative;
        for (int t = 0; t < num_negative; ++t) {
            long long j = -1;
            for (int attempt = 0; attempt < kMaxAttempts; ++attempt) {
                const unsigned long long x =
                    brs_mix64(seed + (unsigned long long)r * C1 + (unsigned long long)t * C2 + (unsigned long long)attempt * C3);
                const long long c = (long long)(x % (unsigned long long)n_items);
                bool taken = brs_pairset_contains(keys, cap, brs_pair_key(u, c));
                for (int q = 0; q < t && !taken; ++q) taken = mine[q] == c;
                if (!taken) {
                    j = c;
                    break;
                }
            }
            if (j < 0) {  // no item left for this user: random.sample raises ValueError in the reference
                atomicOr(status, 2u);
                j = 0;
            }
            mine[t] = j;
        }
    }
}

struct PairsetHeader {
    long long cap;
    long long n_users, n_items;
    unsigned int status;
    unsigned int pad_[9];
};
static_assert(sizeof(PairsetHeader) == 64, "header");

}  // namespace

extern "C" int64_t brs_pairset_bytes(int64_t n_pairs) {
    if (n_pairs < 0) return 0;
    return (int64_t)(sizeof(PairsetHeader) + sizeof(unsigned long long) * (size_t)brs_pairset_capacity(n_pairs));
}

extern "C" int brs_pairset_build(const int64_t* users, const int64_t* items, int64_t n, int64_t n_users, int64_t n_items,
                                 void* set, int64_t set_bytes, void* stream) {
    if (n < 0 || n_users <= 0 || n_items <= 0 || !set || (n > 0 && (!users || !items))) return BRS_ERR_INVALID_ARG;
    if (n_users >= (1ll << 31) || n_items >= (1ll << 32)) return BRS_ERR_UNSUPPORTED;
    if (set_bytes < brs_pairset_bytes(n)) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PairsetHeader h = {};
    h.cap = brs_pairset_capacity(n);
    h.n_users = n_users;
    h.n_items = n_items;
    BRS_CUDA_CHECK(cudaMemcpyAsync(set, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    unsigned long long* keys = (unsigned long long*)((char*)set + sizeof(PairsetHeader));
    BRS_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * (size_t)h.cap, st));
    if (n > 0) {
        long long blocks = (n + kThreads - 1) / kThreads;
        const long long cap = (long long)brs_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        pairset_build_kernel<<<(int)blocks, kThreads, 0, st>>>(keys, h.cap, (const long long*)users, (const long long*)items, n,
                                                               n_users, n_items, &((PairsetHeader*)set)->status);
    }
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_sample_negatives(const void* set, int64_t n_pairs_in_set, const int64_t* users, int64_t n, int64_t n_items,
                                    int32_t num_negative, uint64_t seed, int64_t* neg_out, void* stream) {
    if (!set || n < 0 || n_items <= 0 || n_pairs_in_set < 0 || (n > 0 && (!users || !neg_out))) return BRS_ERR_INVALID_ARG;
    if (num_negative < 1 || num_negative > kMaxNeg) return BRS_ERR_UNSUPPORTED;
    if (n == 0) return BRS_OK;
    const long long cap = brs_pairset_capacity(n_pairs_in_set);
    const unsigned long long* keys = (const unsigned long long*)((const char*)set + sizeof(PairsetHeader));
    long long blocks = (n + kThreads - 1) / kThreads;
    const long long lim = (long long)brs_sm_count() * 8;
    if (blocks > lim) blocks = lim;
    sample_negatives_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        keys, cap, (const long long*)users, n, n_items, num_negative, (unsigned long long)seed, (long long*)neg_out,
        &((PairsetHeader*)const_cast<void*>(set))->status);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// status word of the set: bit 0 = an interaction was outside [0, n_users) x [0, n_items), bit 1 = a user had no
// item left to sample.  Synchronises the stream.
extern "C" int brs_pairset_status(const void* set, uint32_t* status_out, void* stream) {
    if (!set || !status_out) return BRS_ERR_INVALID_ARG;
    BRS_CUDA_CHECK(cudaMemcpyAsync(status_out, &((const PairsetHeader*)set)->status, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                   (cudaStream_t)stream));
    BRS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return BRS_OK;
}
