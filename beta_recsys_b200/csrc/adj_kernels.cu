// Normalised bipartite adjacency on the GPU -- sm_100a.  (include/brs_b200.h: brs_adj_build)
//
// Replaces BaseData.create_adj_mat (beta_rec/data/base_data.py:337-360: a Python loop over users through
// dok / lil matrices, O(U * nnz)) + normalized_adj_single (beta_rec/utils/common_util.py:24-41) +
// sparse_mx_to_torch_sparse_tensor / coalesce (beta_rec/recommenders/lightgcn.py:15-23,
// beta_rec/models/lightgcn.py:59): interactions (u, i) -> CSR of
//     norm_adj = D^-1 (A + I),   A = [[0, R], [R^T, 0]],  R[u, i] = 1 for every (u, i) (duplicates collapse)
// (or mean_adj = D^-1 A), rows and columns in coalesced (row-major sorted) order, values rounded to fp32
// from the reference's float64 1 / rowsum, plus what the backward SpMM needs: the values of the transpose in
// the same (symmetric) pattern and the map from each transposed non-zero to its forward edge.
//
//   keys     one 64-bit key (row << 32 | col) per directed edge and per self loop
//   sort     cub::DeviceRadixSort (library plumbing, like cuBLAS for a plain GEMM; runs once per training run)
//   unique   cub::DeviceSelect::Unique -> nnz
//   rows     histogram of rows -> exclusive scan = row_ptr
//   fill     col, val = fp32(1.0 / count(row)), val_t = fp32(1.0 / count(col)), edge_id_t by binary search
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct AdjView {
    unsigned long long* keys;   // [M] M = 2E + N
    unsigned long long* sorted; // [M]
    unsigned long long* uniq;   // [M]
    int* cnt;                   // [N + 1]
    long long* n_sel;           // [1]
    unsigned int* status;       // [4]
    void* cub_tmp;
    size_t cub_bytes;
    size_t bytes;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cub_temp_bytes(long long M, long long N) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)M, 0, 64);
    cub::DeviceSelect::Unique(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long*)nullptr,
                              (int)M);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const int*)nullptr, (int*)nullptr, (int)(N + 1));
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

AdjView adj_view(void* buf, long long E, long long N) {
    AdjView v;
    char* p = (char*)buf;
    size_t o = 0;
    const long long M = 2 * E + N;
#define BRS_CARVE(field, type, count)       \
    v.field = (type*)(p + o);               \
    o = al256(o + sizeof(type) * (size_t)(count));
    BRS_CARVE(keys, unsigned long long, M)
    BRS_CARVE(sorted, unsigned long long, M)
    BRS_CARVE(uniq, unsigned long long, M)
    BRS_CARVE(cnt, int, N + 1)
    BRS_CARVE(n_sel, long long, 1)
    BRS_CARVE(status, unsigned int, 4)
#undef BRS_CARVE
    v.cub_bytes = cub_temp_bytes(M, N);
    v.cub_tmp = p + o;
    o = al256(o + v.cub_bytes);
    v.bytes = o;
    return v;
}

__global__ void __launch_bounds__(kThreads) adj_keys_kernel(AdjView v, const long long* users, const long long* items, long long E,
                                                            long long U, long long I, int self_loops) {
    const long long N = U + I;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < E; e += stride) {
        long long u = users[e], i = items[e];
        if ((unsigned long long)u >= (unsigned long long)U || (unsigned long long)i >= (unsigned long long)I) {
            atomicOr(v.status, 1u);
            u = 0;  // keep the key count fixed: the caller discards the result when status != 0
            i = 0;
        }
        const unsigned long long r = (unsigned long long)u, c = (unsigned long long)(U + i);
        v.keys[2 * e] = (r << 32) | c;
        v.keys[2 * e + 1] = (c << 32) | r;
    }
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < N; r += stride)
        v.keys[2 * E + r] = self_loops ? (((unsigned long long)r << 32) | (unsigned long long)r)
                                       : 0xffffffffffffffffull;  // sorts last, dropped below
}

__global__ void __launch_bounds__(kThreads) adj_rows_kernel(AdjView v, long long N) {
    const long long n = *v.n_sel;
    for (long long p = (long long)blockIdx.x * kThreads + threadIdx.x; p < n; p += (long long)gridDim.x * kThreads) {
        const unsigned long long k = v.uniq[p];
        if (k == 0xffffffffffffffffull) continue;
        atomicAdd(v.cnt + (int)(k >> 32), 1);
    }
}

__global__ void __launch_bounds__(kThreads) adj_fill_kernel(AdjView v, long long N, const int* row_ptr, int* col, float* val,
                                                            float* val_t, int* edge_id_t, long long* nnz_out) {
    const long long nnz = row_ptr[N];
    if (blockIdx.x == 0 && threadIdx.x == 0) *nnz_out = nnz;
    for (long long p = (long long)blockIdx.x * kThreads + threadIdx.x; p < nnz; p += (long long)gridDim.x * kThreads) {
        const unsigned long long k = v.uniq[p];
        const int r = (int)(k >> 32), c = (int)(k & 0xffffffffull);
        col[p] = c;
        // the reference divides in float64 (adj + sp.eye promotes to float64) and casts to fp32 afterwards
        val[p] = (float)(1.0 / (double)(row_ptr[r + 1] - row_ptr[r]));
        const int cb = row_ptr[c], ce = row_ptr[c + 1];
        val_t[p] = (float)(1.0 / (double)(ce - cb));  // A^T[r][c] = A[c][r]
        // forward edge (c, r): the pattern is symmetric, so it exists; lower bound on the unique keys of row c
        int lo = cb, hi = ce;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)(v.uniq[mid] & 0xffffffffull) < r) lo = mid + 1;
            else hi = mid;
        }
        edge_id_t[p] = lo;
    }
}

}  // namespace

extern "C" int64_t brs_adj_workspace_bytes(int64_t n_interactions, int64_t n_users, int64_t n_items) {
    if (n_interactions < 0 || n_users <= 0 || n_items <= 0) return 0;
    if (2 * n_interactions + n_users + n_items >= (1ll << 31)) return 0;
    return (int64_t)adj_view(nullptr, n_interactions, n_users + n_items).bytes;
}

extern "C" int brs_adj_build(const int64_t* users, const int64_t* items, int64_t n_interactions, int64_t n_users,
                             int64_t n_items, int32_t self_loops, void* workspace, int64_t workspace_bytes, int32_t* row_ptr,
                             int32_t* col, float* val, float* val_t, int32_t* edge_id_t, int64_t* nnz_out, void* stream) {
    const long long E = n_interactions, U = n_users, I = n_items, N = U + I;
    if (E < 0 || U <= 0 || I <= 0 || !workspace || !row_ptr || !col || !val || !val_t || !edge_id_t || !nnz_out)
        return BRS_ERR_INVALID_ARG;
    if (E > 0 && (!users || !items)) return BRS_ERR_INVALID_ARG;
    const long long M = 2 * E + N;
    if (M >= (1ll << 31)) return BRS_ERR_UNSUPPORTED;  // int32 CSR, like the SpMM kernels
    AdjView v = adj_view(workspace, E, N);
    if ((size_t)workspace_bytes < v.bytes) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    BRS_CUDA_CHECK(cudaMemsetAsync(v.cnt, 0, sizeof(int) * (size_t)(N + 1), st));
    BRS_CUDA_CHECK(cudaMemsetAsync(v.status, 0, 16, st));
    const long long cap = (long long)brs_sm_count() * 8;
    auto grid = [&](long long n) {
        long long b = (n + kThreads - 1) / kThreads;
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    adj_keys_kernel<<<grid(E > N ? E : N), kThreads, 0, st>>>(v, (const long long*)users, (const long long*)items, E, U, I,
                                                              self_loops);
    int bits = 33;
    while ((1ll << (bits - 32)) < N && bits < 64) ++bits;
    size_t tb = v.cub_bytes;
    // without self loops the padding keys are all-ones: sort the full width so that they land at the end
    BRS_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(v.cub_tmp, tb, v.keys, v.sorted, (int)M, 0, self_loops ? bits : 64, st));
    tb = v.cub_bytes;
    BRS_CUDA_CHECK(cub::DeviceSelect::Unique(v.cub_tmp, tb, v.sorted, v.uniq, v.n_sel, (int)M, st));
    adj_rows_kernel<<<grid(M), kThreads, 0, st>>>(v, N);
    tb = v.cub_bytes;
    BRS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(v.cub_tmp, tb, v.cnt, row_ptr, (int)(N + 1), st));
    adj_fill_kernel<<<grid(M), kThreads, 0, st>>>(v, N, row_ptr, col, val, val_t, edge_id_t, (long long*)nnz_out);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// 1 = an interaction was outside [0, n_users) x [0, n_items).  Synchronises the stream.
extern "C" int brs_adj_status(const void* workspace, int64_t n_interactions, int64_t n_users, int64_t n_items,
                              uint32_t* status_out, void* stream) {
    if (!workspace || !status_out) return BRS_ERR_INVALID_ARG;
    AdjView v = adj_view(const_cast<void*>(workspace), n_interactions, n_users + n_items);
    BRS_CUDA_CHECK(cudaMemcpyAsync(status_out, v.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BRS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return BRS_OK;
}
