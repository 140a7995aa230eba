// Ranking evaluation on the GPU -- sm_100a.  (include/brs_b200.h: brs_rank_metrics)
//
// Replaces beta_rec/core/eval_engine.py:49-87 (evaluate) -> beta_rec/utils/evaluation.py:459-534
// (merge_ranking_true_pred), :537-752 (precision / recall / ndcg / map at k), :755-785 (get_top_k_items):
// pandas groupby().apply(nlargest) + rank + two merges per metric, 6-9 s per epoch on ML-100k in the
// reference's own notebook and broken under pandas 3.
//
// The metrics only need the RANK of every relevant prediction row inside its user's rows
//     rank = 1 + #{rows of the user with a higher score} + #{rows with the same score earlier in the frame}
// (nlargest keeps the first of equal scores, rank(method="first")), so no sort and no top-k list is built:
//   hist      prediction rows per user; relevant (rating >= 1) true rows per user ("actual") and their
//             (user, item) keys into an open-addressing hash set
//   scan      exclusive scan of the rows-per-user histogram (three small kernels)
//   scatter   prediction rows grouped by user: {score, original row index | relevant flag (hash probe)}
//   metrics   one warp per user: for every relevant row of the user, count the rows that outrank it; ranks
//             <= k are distinct integers, kept as a k-bit mask in shared memory, from which hit count,
//             DCG, IDCG and the MAP numerator follow in rank order; double-precision sums, 5 atomics per CTA
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerCta = kThreads / 32;
constexpr int kScanPerCta = 4096;  // elements per CTA of the scan kernels
constexpr int kMaxK = 1024;
constexpr unsigned long long kEmpty = 0xffffffffffffffffull;

struct EvalView {
    int* cnt;            // [U + 1] prediction rows per user -> exclusive offsets after the scan
    int* cursor;         // [U]
    int* actual;         // [U] relevant true rows per user
    int* blk;            // [ceil((U + 1) / kScanPerCta) + 1] CTA sums of the scan
    float* g_score;      // [n_pred] grouped by user
    int* g_meta;         // [n_pred] original row index << 1 | relevant
    unsigned long long* keys;  // [hash_cap] (user << 32 | item) of the relevant true rows
    unsigned int* status;      // [4]
    long long hash_cap;
    size_t bytes;
};

__host__ __device__ inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

inline long long hash_capacity(long long n_true) {
    long long c = 1024;
    while (c < 2 * n_true) c <<= 1;
    return c;
}

inline EvalView eval_view(void* buf, long long n_true, long long n_pred, long long U) {
    EvalView v;
    char* p = (char*)buf;
    size_t o = 0;
#define BRS_CARVE(field, type, count)       \
    v.field = (type*)(p + o);               \
    o = al256(o + sizeof(type) * (size_t)(count));
    BRS_CARVE(cnt, int, U + 1)
    BRS_CARVE(cursor, int, U)
    BRS_CARVE(actual, int, U)
    BRS_CARVE(blk, int, (U + 1 + kScanPerCta - 1) / kScanPerCta + 1)
    BRS_CARVE(status, unsigned int, 4)
    const size_t zeroed = o;  // everything above is cleared at the start of a call
    BRS_CARVE(g_score, float, n_pred)
    BRS_CARVE(g_meta, int, n_pred)
    v.hash_cap = hash_capacity(n_true);
    BRS_CARVE(keys, unsigned long long, v.hash_cap)
#undef BRS_CARVE
    v.bytes = o;
    (void)zeroed;
    return v;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(kThreads) eval_hist_kernel(EvalView v, const long long* t_users, const long long* t_items,
                                                             const float* t_ratings, long long n_true,
                                                             const long long* p_users, const long long* p_items,
                                                             long long n_pred, long long U) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < n_pred; r += stride) {
        const long long u = p_users[r], i = p_items[r];
        if ((unsigned long long)u >= (unsigned long long)U || (unsigned long long)i >= (1ull << 32)) {
            atomicOr(v.status, 1u);
            continue;
        }
        atomicAdd(v.cnt + u, 1);
    }
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < n_true; r += stride) {
        if (!(t_ratings[r] >= 1.0f)) continue;  // evaluation.py:492
        const long long u = t_users[r], i = t_items[r];
        if ((unsigned long long)u >= (unsigned long long)U || (unsigned long long)i >= (1ull << 32)) {
            atomicOr(v.status, 1u);
            continue;
        }
        const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)i;
        unsigned long long h = mix64(key) & (unsigned long long)(v.hash_cap - 1);
        for (;;) {  // the table is at most half full
            const unsigned long long old = atomicCAS(v.keys + h, kEmpty, key);
            if (old == kEmpty) {
                atomicAdd(v.actual + u, 1);  // a (user, item) pair listed twice counts once, as in pandas' merge keys
                break;
            }
            if (old == key) break;
            h = (h + 1) & (unsigned long long)(v.hash_cap - 1);
        }
    }
}

// exclusive scan of cnt[0 .. n): CTA-local scans + CTA sums, scan of the sums, add-back
__global__ void __launch_bounds__(1024) scan_local_kernel(int* data, int* blk, long long n) {
    __shared__ int s_w[32];
    const long long base = (long long)blockIdx.x * kScanPerCta + threadIdx.x * 4;
    int x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = base + q < n ? data[base + q] : 0;
    const int mine = x[0] + x[1] + x[2] + x[3];
    int incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(BRS_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_w[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(BRS_FULL_MASK, wi, o);
            if (lane >= o) wi += t;
        }
        s_w[lane] = wi - w;
        if (lane == 31) blk[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = s_w[warp] + incl - mine;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (base + q < n) data[base + q] = run;
        run += x[q];
    }
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int* blk, int n_blk) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_blk; b0 += 1024) {
        const int k = b0 + threadIdx.x;
        const int x = k < n_blk ? blk[k] : 0;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(BRS_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(BRS_FULL_MASK, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (k < n_blk) blk[k] = carry + s_w[warp] + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) scan_add_kernel(int* data, const int* blk, long long n) {
    const int add = blk[blockIdx.x];
    const long long base = (long long)blockIdx.x * kScanPerCta + threadIdx.x * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (base + q < n) data[base + q] += add;
}

__global__ void __launch_bounds__(kThreads) eval_scatter_kernel(EvalView v, const long long* p_users, const long long* p_items,
                                                                const float* p_scores, long long n_pred, long long U) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < n_pred; r += stride) {
        const long long u = p_users[r], i = p_items[r];
        if ((unsigned long long)u >= (unsigned long long)U || (unsigned long long)i >= (1ull << 32)) continue;
        const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)i;
        unsigned long long h = mix64(key) & (unsigned long long)(v.hash_cap - 1);
        int rel = 0;
        for (;;) {
            const unsigned long long got = v.keys[h];
            if (got == key) {
                rel = 1;
                break;
            }
            if (got == kEmpty) break;
            h = (h + 1) & (unsigned long long)(v.hash_cap - 1);
        }
        const int pos = v.cnt[u] + atomicAdd(v.cursor + u, 1);
        v.g_score[pos] = p_scores[r];
        v.g_meta[pos] = (int)(r << 1) | rel;
    }
}

// one warp per user that has both prediction rows and relevant true rows (the "common users")
__global__ void __launch_bounds__(kThreads) eval_metrics_kernel(EvalView v, long long U, int k, double* out) {
    __shared__ unsigned int s_mask[kWarpsPerCta][kMaxK / 32];
    __shared__ double s_sum[kWarpsPerCta][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_warps = (long long)gridDim.x * kWarpsPerCta;
    double ndcg = 0.0, map = 0.0, prec = 0.0, rec = 0.0, users = 0.0, hits = 0.0;  // lane 0 only
    for (long long u0 = ((long long)blockIdx.x * kWarpsPerCta + warp) * 32; u0 < U; u0 += n_warps * 32) {
        const long long ul = u0 + lane;
        int beg = 0, end = 0, act = 0;
        if (ul < U) {
            beg = v.cnt[ul];
            end = v.cnt[ul + 1];
            act = v.actual[ul];
        }
        unsigned todo = __ballot_sync(BRS_FULL_MASK, end > beg && act > 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int b = __shfl_sync(BRS_FULL_MASK, beg, src), e = __shfl_sync(BRS_FULL_MASK, end, src);
            const int actual = __shfl_sync(BRS_FULL_MASK, act, src);
            for (int w = lane; w < (k + 31) / 32; w += 32) s_mask[warp][w] = 0u;
            __syncwarp();
            // chunks of 32 rows: every relevant row gets its rank from a warp-strided count over the user's rows
            for (int c0 = b; c0 < e; c0 += 32) {
                const int j = c0 + lane;
                const int meta = j < e ? v.g_meta[j] : 0;
                const float sc = j < e ? v.g_score[j] : 0.f;
                unsigned relm = __ballot_sync(BRS_FULL_MASK, j < e && (meta & 1));
                while (relm) {
                    const int rl = __ffs(relm) - 1;
                    relm &= relm - 1;
                    const float s = __shfl_sync(BRS_FULL_MASK, sc, rl);
                    const int idx = __shfl_sync(BRS_FULL_MASK, meta, rl) >> 1;
                    int ahead = 0;
                    for (int q = b + lane; q < e; q += 32) {
                        const float sq = v.g_score[q];
                        ahead += (sq > s) || (sq == s && (v.g_meta[q] >> 1) < idx);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ahead += __shfl_xor_sync(BRS_FULL_MASK, ahead, o);
                    if (lane == 0 && ahead < k) s_mask[warp][ahead >> 5] |= 1u << (ahead & 31);  // rank = ahead + 1
                }
            }
            __syncwarp();
            if (lane == 0) {
                int c = 0;
                double dcg = 0.0, ap = 0.0;
                for (int r = 1; r <= k; ++r) {
                    if (s_mask[warp][(r - 1) >> 5] >> ((r - 1) & 31) & 1u) {
                        c += 1;
                        dcg += 1.0 / log1p((double)r);
                        ap += (double)c / (double)r;
                    }
                }
                users += 1.0;
                if (c > 0) {
                    double idcg = 0.0;
                    const int top = actual < k ? actual : k;
                    for (int r = 1; r <= top; ++r) idcg += 1.0 / log1p((double)r);
                    hits += (double)c;
                    prec += (double)c / (double)k;
                    rec += (double)c / (double)actual;
                    ndcg += dcg / idcg;
                    map += ap / (double)actual;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        s_sum[warp][0] = ndcg;
        s_sum[warp][1] = map;
        s_sum[warp][2] = prec;
        s_sum[warp][3] = rec;
        s_sum[warp][4] = users;
        s_sum[warp][5] = hits;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) t += s_sum[w][threadIdx.x];
        if (t != 0.0) atomicAdd(out + threadIdx.x, t);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[6] = (double)*v.status;
}

}  // namespace

extern "C" int64_t brs_rank_metrics_workspace_bytes(int64_t n_true, int64_t n_pred, int64_t n_user_ids) {
    if (n_true < 0 || n_pred < 0 || n_user_ids <= 0) return 0;
    return (int64_t)eval_view(nullptr, n_true, n_pred, n_user_ids).bytes;
}

extern "C" int brs_rank_metrics(const int64_t* true_users, const int64_t* true_items, const float* true_ratings,
                                int64_t n_true, const int64_t* pred_users, const int64_t* pred_items,
                                const float* pred_scores, int64_t n_pred, int64_t n_user_ids, int32_t k, void* workspace,
                                int64_t workspace_bytes, double* out, void* stream) {
    if (n_true < 0 || n_pred < 0 || n_user_ids <= 0 || !workspace || !out) return BRS_ERR_INVALID_ARG;
    if (k < 1 || k > kMaxK) return BRS_ERR_UNSUPPORTED;
    if (n_pred >= (1ll << 30) || n_true >= (1ll << 30) || n_user_ids >= (1ll << 31) - 1) return BRS_ERR_UNSUPPORTED;
    if ((n_true > 0 && (!true_users || !true_items || !true_ratings)) || (n_pred > 0 && (!pred_users || !pred_items || !pred_scores)))
        return BRS_ERR_INVALID_ARG;
    const EvalView v = eval_view(workspace, n_true, n_pred, n_user_ids);
    if ((size_t)workspace_bytes < v.bytes) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    BRS_CUDA_CHECK(cudaMemsetAsync(workspace, 0, (size_t)((char*)v.g_score - (char*)workspace), st));
    BRS_CUDA_CHECK(cudaMemsetAsync(v.keys, 0xff, sizeof(unsigned long long) * (size_t)v.hash_cap, st));
    BRS_CUDA_CHECK(cudaMemsetAsync(out, 0, 8 * sizeof(double), st));
    const long long cap = (long long)brs_sm_count() * 8;
    const long long most = n_pred > n_true ? n_pred : n_true;
    long long blocks = (most + kThreads - 1) / kThreads;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    eval_hist_kernel<<<(int)blocks, kThreads, 0, st>>>(v, (const long long*)true_users, (const long long*)true_items, true_ratings,
                                                       n_true, (const long long*)pred_users, (const long long*)pred_items, n_pred,
                                                       n_user_ids);
    const long long n_scan = n_user_ids + 1;
    const int n_blk = (int)((n_scan + kScanPerCta - 1) / kScanPerCta);
    scan_local_kernel<<<n_blk, 1024, 0, st>>>(v.cnt, v.blk, n_scan);
    scan_sums_kernel<<<1, 1024, 0, st>>>(v.blk, n_blk);
    scan_add_kernel<<<n_blk, 1024, 0, st>>>(v.cnt, v.blk, n_scan);
    blocks = (n_pred + kThreads - 1) / kThreads;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    eval_scatter_kernel<<<(int)blocks, kThreads, 0, st>>>(v, (const long long*)pred_users, (const long long*)pred_items, pred_scores,
                                                          n_pred, n_user_ids);
    long long wblocks = (n_user_ids + 32 * kWarpsPerCta - 1) / (32 * kWarpsPerCta);
    if (wblocks > cap) wblocks = cap;
    eval_metrics_kernel<<<(int)wblocks, kThreads, 0, st>>>(v, n_user_ids, k, out);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
