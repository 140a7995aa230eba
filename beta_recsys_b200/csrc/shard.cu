// Multi-GPU plumbing for the row-sharded tables -- sm_100a, one process per GPU.
//
// New work: the reference has no distributed path.  Tables are
// row-sharded (owner = row mod world); each rank exports its shard through CUDA IPC and
// maps its peers', so the hot-path kernels address remote rows directly over NVLink
// (peer loads / peer REDs / red.or on the touched bitmaps) instead of staging them through collectives.
// This file holds: the exportable allocator + IPC handle helpers, the flag barrier that
// orders the phases of a step across ranks (and exchanges the step's scalar sums), and the
// owner bucketing of triples that feeds the (optional)
// NCCL all-to-all routing of triples to the user-row owner.
#include <string.h>

#include <stdio.h>

#include "common.cuh"

namespace {

// ~20 s at 2 GHz: far beyond any legitimate skew between ranks (a step is ~100 us), short enough that a hung
// job ends inside the caller's own timeout
constexpr long long kBarrierTimeoutCycles = 40ll * 1000 * 1000 * 1000;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------------------
// flag barrier over peer memory (+ exchange of the step sums)
// ---------------------------------------------------------------------------
struct BarrierArgs {
    unsigned long long* flags[BRS_MAX_RANKS];
    double* partials[BRS_MAX_RANKS];
    int world, rank;
    unsigned long long epoch;
    brs_step_ws* ws;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const BarrierArgs a) {
    const int p = threadIdx.x;
    if (a.ws && p < a.world) {  // mail this rank's step sums to every rank (plain stores, one writer per cell)
        double* dst = a.partials[p] + a.rank * 4;
        dst[0] = a.ws->loss_sum;
        dst[1] = a.ws->reg_sum;
        dst[2] = (double)a.ws->g_global_bias;
        dst[3] = (double)a.ws->err_flag;
    }
    __threadfence_system();  // everything this GPU wrote before (incl. earlier kernels' peer REDs) first
    if (p < a.world) st_release_sys(a.flags[p] + a.rank, a.epoch);
    if (p < a.world) {
        // bounded wait: a peer that left the step sequence (error return, exception) never posts its flag;
        // trap instead of spinning forever with a resident kernel -- the launch fails with a sticky error on
        // this rank, which the host reports, and the process group tears down
        const long long t0 = clock64();
        while (ld_acquire_sys(a.flags[a.rank] + p) < a.epoch) {
            if (clock64() - t0 > kBarrierTimeoutCycles) {
                printf("brs: peer barrier timed out on rank %d waiting for rank %d (epoch %lld)\n", a.rank, p, (long long)a.epoch);
                __trap();
            }
        }
    }
    __syncwarp();
    if (a.ws && p == 0) {  // same order on every rank -> bit-identical sums -> replicas stay in sync
        double l = 0.0, r = 0.0, g = 0.0, e = 0.0;
        for (int q = 0; q < a.world; ++q) {
            const double* src = a.partials[a.rank] + q * 4;
            l += src[0];
            r += src[1];
            g += src[2];
            e = fmax(e, src[3]);
        }
        a.ws->loss_sum = l;
        a.ws->reg_sum = r;
        a.ws->g_global_bias = (float)g;
        a.ws->err_flag = (unsigned int)e;
    }
}

// ---------------------------------------------------------------------------
// stable bucketing of triples by owner(user)
// ---------------------------------------------------------------------------
// pass 1: per-block histogram; pass 2 (one block): exclusive scan over (dest, block); pass 3: scatter
__global__ void __launch_bounds__(kThreads) route_hist_kernel(const long long* __restrict__ users, long long n, int world,
                                                             int* __restrict__ block_hist /* [world][grid] */) {
    __shared__ int s_h[BRS_MAX_RANKS];
    if (threadIdx.x < BRS_MAX_RANKS) s_h[threadIdx.x] = 0;
    __syncthreads();
    const long long per_block = (n + gridDim.x - 1) / gridDim.x;
    const long long b0 = blockIdx.x * per_block, b1 = min(n, b0 + per_block);
    for (long long t = b0 + threadIdx.x; t < b1; t += kThreads) atomicAdd(&s_h[(int)(users[t] & (world - 1))], 1);
    __syncthreads();
    if (threadIdx.x < world) block_hist[threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void route_scan_kernel(int* block_hist, int n_blocks, int world, long long* counts) {
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int d = 0; d < world; ++d) {
            long long c = 0;
            for (int b = 0; b < n_blocks; ++b) {
                const int v = block_hist[d * n_blocks + b];
                block_hist[d * n_blocks + b] = (int)(run + c);
                c += v;
            }
            counts[d] = c;
            run += c;
        }
    }
}

__global__ void __launch_bounds__(kThreads) route_scatter_kernel(const long long* __restrict__ users,
                                                                const long long* __restrict__ pos,
                                                                const long long* __restrict__ neg, long long n,
                                                                int world, const int* __restrict__ block_off,
                                                                long long* __restrict__ ou, long long* __restrict__ op,
                                                                long long* __restrict__ on) {
    __shared__ int s_run[BRS_MAX_RANKS];
    __shared__ int s_wcnt[kWarps][BRS_MAX_RANKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < world) s_run[threadIdx.x] = block_off[threadIdx.x * gridDim.x + blockIdx.x];
    __syncthreads();
    const long long per_block = (n + gridDim.x - 1) / gridDim.x;
    const long long b0 = blockIdx.x * per_block, b1 = min(n, b0 + per_block);
    for (long long c0 = b0; c0 < b1; c0 += kThreads) {  // chunks in order => stable
        const long long t = c0 + threadIdx.x;
        const bool on_ = t < b1;
        const long long u = on_ ? users[t] : 0;
        const int d = on_ ? (int)(u & (world - 1)) : -1;
        int lane_off = 0;
        for (int o = 0; o < world; ++o) {
            const unsigned b = __ballot_sync(BRS_FULL_MASK, d == o);
            if (d == o) lane_off = __popc(b & ((1u << lane) - 1u));
            if (lane == 0) s_wcnt[warp][o] = __popc(b);
        }
        __syncthreads();
        if (on_) {
            int off = s_run[d] + lane_off;
            for (int w = 0; w < warp; ++w) off += s_wcnt[w][d];
            ou[off] = u;
            op[off] = pos[t];
            on[off] = neg[t];
        }
        __syncthreads();
        if (threadIdx.x < world) {
            int tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += s_wcnt[w][threadIdx.x];
            s_run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

constexpr int kRouteBlocks = 128;
int* g_route_hist = nullptr;  // [BRS_MAX_RANKS][kRouteBlocks] device scratch, allocated once

}  // namespace

extern "C" int brs_shm_alloc(int64_t bytes, void** ptr) {
    if (!ptr || bytes <= 0) return BRS_ERR_INVALID_ARG;
    BRS_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
    BRS_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
    BRS_CUDA_CHECK(cudaDeviceSynchronize());
    return BRS_OK;
}

extern "C" int brs_shm_free(void* ptr) {
    if (!ptr) return BRS_OK;
    BRS_CUDA_CHECK(cudaFree(ptr));
    return BRS_OK;
}

extern "C" int brs_ipc_get_handle(void* ptr, uint8_t handle[BRS_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == BRS_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ptr || !handle) return BRS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    BRS_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, sizeof(h));
    return BRS_OK;
}

extern "C" int brs_ipc_open_handle(const uint8_t handle[BRS_IPC_HANDLE_BYTES], void** ptr) {
    if (!ptr || !handle) return BRS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    BRS_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BRS_OK;
}

extern "C" int brs_ipc_close_handle(void* ptr) {
    if (!ptr) return BRS_OK;
    BRS_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return BRS_OK;
}

extern "C" int brs_peer_barrier(const brs_peer_sync* sync, uint64_t epoch, void* ws, void* stream) {
    if (!sync || sync->world < 1 || sync->world > BRS_MAX_RANKS || sync->rank < 0 || sync->rank >= sync->world || epoch == 0)
        return BRS_ERR_INVALID_ARG;
    BarrierArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < sync->world; ++r) {
        if (!sync->flags[r] || (ws && !sync->partials[r])) return BRS_ERR_INVALID_ARG;
        a.flags[r] = (unsigned long long*)sync->flags[r];
        a.partials[r] = sync->partials[r];
    }
    a.world = sync->world;
    a.rank = sync->rank;
    a.epoch = epoch;
    a.ws = (brs_step_ws*)ws;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_route_triples(const int64_t* users, const int64_t* pos_items, const int64_t* neg_items, int64_t n,
                                 int32_t world, int64_t* out_users, int64_t* out_pos, int64_t* out_neg,
                                 int64_t* counts, void* stream) {
    if (!users || !pos_items || !neg_items || !out_users || !out_pos || !out_neg || !counts || n < 0)
        return BRS_ERR_INVALID_ARG;
    if (world < 1 || world > BRS_MAX_RANKS || (world & (world - 1)) != 0) return BRS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (!g_route_hist) BRS_CUDA_CHECK(cudaMalloc(&g_route_hist, sizeof(int) * BRS_MAX_RANKS * kRouteBlocks));
    int blocks = (int)((n + kThreads - 1) / kThreads);
    if (blocks > kRouteBlocks) blocks = kRouteBlocks;
    if (blocks < 1) blocks = 1;
    route_hist_kernel<<<blocks, kThreads, 0, st>>>((const long long*)users, n, world, g_route_hist);
    route_scan_kernel<<<1, 32, 0, st>>>(g_route_hist, blocks, world, (long long*)counts);
    route_scatter_kernel<<<blocks, kThreads, 0, st>>>((const long long*)users, (const long long*)pos_items,
                                                      (const long long*)neg_items, n, world, g_route_hist,
                                                      (long long*)out_users, (long long*)out_pos, (long long*)out_neg);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
