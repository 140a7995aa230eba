// Library-level C ABI: status strings, device info, MF step composition, the
// epoch loop over HBM-resident batches, and the gather / scatter-add micro-ops
// of BASELINE.json config 5.  See include/brs_b200.h for the contract.
#include <mutex>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

int brs_mf_fwd_bwd_impl(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                        const void* third, int64_t batch, float reg_weight, void* stream);
int brs_apply_impl(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                   int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                   long long batch, long long max_rows_hint, void* stream);
int brs_apply_impl_next(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                        int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                        long long batch, long long max_rows_hint, void* stream, const brs_rowset* next_rs,
                        const long long* const* next_idx, const long long* next_n, int next_arrays, int parity);
int brs_mf_fwd_bwd_phases(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                          const void* third, int64_t batch, float reg_weight, void* stream, int phases);

namespace {
char g_cuda_err[512] = "";
std::mutex g_err_mu;
int g_sm_count = 0;
}  // namespace

void brs_set_cuda_error(cudaError_t e, const char* what, int line) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), what, line);
    (void)cudaGetLastError();
}

int brs_sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count = n;
        else
            g_sm_count = 148;  // B200
    }
    return g_sm_count;
}

namespace {
// defaults from the round-1 sweep (tools/sweep_mf.py l2, profiles/r01_sweeps.md): streamed weight rows
// are demoted, the compact gradient scratch is pinned -> fused kernel 58 -> 50 us at config 2
brs_l2_policy_cfg g_l2 = {BRS_L2_EVICT_FIRST, BRS_L2_EVICT_LAST, BRS_L2_EVICT_FIRST};
}
const brs_l2_policy_cfg& brs_l2_cfg() { return g_l2; }

extern "C" int brs_debug_set_l2_policy(int gather, int scratch, int weight) {
    if (gather < 0 || gather > 2 || scratch < 0 || scratch > 2 || weight < 0 || weight > 2) return BRS_ERR_INVALID_ARG;
    g_l2.gather = gather;
    g_l2.scratch = scratch;
    g_l2.weight = weight;
    return BRS_OK;
}

extern "C" int brs_abi_version(void) { return BRS_ABI_VERSION; }

extern "C" const char* brs_strerror(int status) {
    switch (status) {
        case BRS_OK: return "ok";
        case BRS_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size or misaligned table)";
        case BRS_ERR_UNSUPPORTED: return "unsupported dimension / optimizer / shape";
        case BRS_ERR_CUDA: return "CUDA runtime error (see brs_last_cuda_error)";
        case BRS_ERR_INDEX_RANGE: return "index out of range";
        case BRS_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

extern "C" const char* brs_last_cuda_error(void) { return g_cuda_err; }

extern "C" int brs_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        (void)cudaGetLastError();
        return BRS_ERR_NO_DEVICE;
    }
    int v = 0;
    if (sm_count) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        *sm_count = v;
    }
    if (cc_major) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
        *cc_major = v;
    }
    if (cc_minor) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
        *cc_minor = v;
    }
    if (l2_bytes) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
        *l2_bytes = v;
    }
    return BRS_OK;
}

// ---------------------------------------------------------------------------
// MF step composition
// ---------------------------------------------------------------------------
extern "C" int brs_mf_apply(const brs_mf_model* model, const brs_opt* opt, int64_t batch, float* out_loss_reg,
                            void* stream) {
    if (!model || !opt || !model->ws || batch < 0) return BRS_ERR_INVALID_ARG;
    brs_entity ents[2] = {model->user, model->item};
    // a batch touches at most `batch` user rows and 2*batch item rows
    const long long hint = 3 * (long long)batch;
    return brs_apply_impl(ents, 2, &model->global_bias, 1, /*dense_grad_from_ws=*/1, opt, model->ws, 0, out_loss_reg,
                          batch, hint, stream);
}

extern "C" int brs_mf_train_batches(const brs_mf_model* model, const brs_opt* opt, int32_t loss_kind,
                                    const int64_t* users, const int64_t* items, const void* third, int64_t n,
                                    int64_t batch, float reg_weight, float* out_loss_reg, void* stream) {
    if (!model || !opt || !users || !items || !third || n < 0 || batch <= 0 || !out_loss_reg) return BRS_ERR_INVALID_ARG;
    const size_t third_sz = loss_kind == 0 ? 8 : 4;
    // With alternate rowsets and a touched-rows optimizer the slot pre-pass of batch b+1 runs inside the
    // apply launch of batch b (2 launches per step instead of 3); otherwise the plain 3-launch sequence.
    const bool overlap = model->user_rows_alt.slot_map && model->item_rows_alt.slot_map &&
                         (opt->kind == BRS_SGD || opt->mode == BRS_TOUCHED_ROWS);
    if (!overlap) {
        int64_t b = 0;
        for (int64_t off = 0; off < n; off += batch, ++b) {
            const int64_t cur = (n - off < batch) ? (n - off) : batch;
            int rc = brs_mf_fwd_bwd_impl(model, loss_kind, users + off, items + off, (const char*)third + off * third_sz,
                                         cur, reg_weight, stream);
            if (rc != BRS_OK) return rc;
            rc = brs_mf_apply(model, opt, cur, out_loss_reg + 4 * b, stream);
            if (rc != BRS_OK) return rc;
        }
        return BRS_OK;
    }
    brs_mf_model m[2] = {*model, *model};  // m[1] works on the alternate rowsets
    m[1].user.rows = model->user_rows_alt;
    m[1].item.rows = model->item_rows_alt;
    m[1].user_rows_alt = model->user.rows;
    m[1].item_rows_alt = model->item.rows;
    if (n > 0) {  // pre-pass of batch 0 on its own
        const int64_t cur = n < batch ? n : batch;
        int rc = brs_mf_fwd_bwd_phases(&m[0], loss_kind, users, items, third, cur, reg_weight, stream, 1);
        if (rc != BRS_OK) return rc;
    }
    int64_t b = 0;
    for (int64_t off = 0; off < n; off += batch, ++b) {
        const int64_t cur = (n - off < batch) ? (n - off) : batch;
        const brs_mf_model& cm = m[b & 1];
        int rc = brs_mf_fwd_bwd_phases(&cm, loss_kind, users + off, items + off, (const char*)third + off * third_sz, cur,
                                       reg_weight, stream, 2);
        if (rc != BRS_OK) return rc;
        brs_entity ents[2] = {cm.user, cm.item};
        const int64_t noff = off + batch;
        if (noff < n) {
            const int64_t ncur = (n - noff < batch) ? (n - noff) : batch;
            const brs_mf_model& nm = m[(b + 1) & 1];
            const brs_rowset rs[3] = {nm.user.rows, nm.item.rows, nm.item.rows};
            const long long* idx[3] = {(const long long*)users + noff, (const long long*)items + noff,
                                       (const long long*)((const char*)third + noff * third_sz)};
            const long long nn[3] = {ncur, ncur, ncur};
            rc = brs_apply_impl_next(ents, 2, &cm.global_bias, 1, 1, opt, cm.ws, 0, out_loss_reg + 4 * b, cur, 3 * cur,
                                     stream, rs, idx, nn, loss_kind == 0 ? 3 : 2, (int)(b & 1));
        } else {
            rc = brs_apply_impl_next(ents, 2, &cm.global_bias, 1, 1, opt, cm.ws, 0, out_loss_reg + 4 * b, cur, 3 * cur,
                                     stream, nullptr, nullptr, nullptr, 0, (int)(b & 1));
        }
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

// the sharded epoch inner loop: every rank calls it with ITS OWN index arrays (same n / batch on all ranks)
extern "C" int brs_mf_sharded_step(const brs_mf_sharded* model, const brs_peer_sync* sync, const brs_opt* opt,
                                  const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                                  int64_t batch, int64_t global_batch, float reg_weight, uint64_t epoch, float* out,
                                  void* stream) {
    if (!model || !sync || !opt || !out || epoch == 0) return BRS_ERR_INVALID_ARG;
    int rc = brs_mf_sharded_bpr_fwd_bwd(model, users, pos_items, neg_items, batch, global_batch, reg_weight, stream);
    if (rc != BRS_OK) return rc;
    if (opt->kind == BRS_SGD) {
        // the push updates the owners' weights in place: every rank must be done gathering first
        rc = brs_peer_barrier(sync, epoch, model->stage.ws, stream);
        if (rc != BRS_OK) return rc;
        rc = brs_mf_sharded_push(model, opt, stream);
        if (rc != BRS_OK) return rc;
        rc = brs_mf_sharded_apply(model, opt, global_batch, out, stream);
        if (rc != BRS_OK) return rc;
        return brs_peer_barrier(sync, epoch + 1, nullptr, stream);
    }
    rc = brs_mf_sharded_push(model, opt, stream);
    if (rc != BRS_OK) return rc;
    rc = brs_peer_barrier(sync, epoch, model->stage.ws, stream);
    if (rc != BRS_OK) return rc;
    rc = brs_mf_sharded_apply(model, opt, global_batch, out, stream);
    if (rc != BRS_OK) return rc;
    return brs_peer_barrier(sync, epoch + 1, nullptr, stream);
}

extern "C" int brs_mf_sharded_train_batches(const brs_mf_sharded* model, const brs_peer_sync* sync, const brs_opt* opt,
                                            const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                                            int64_t n, int64_t batch, int64_t global_batch, float reg_weight,
                                            uint64_t first_epoch, float* out, void* stream) {
    if (!model || !sync || !opt || !users || !pos_items || !neg_items || !out || n < 0 || batch <= 0 || first_epoch == 0)
        return BRS_ERR_INVALID_ARG;
    uint64_t epoch = first_epoch;
    int64_t b = 0;
    for (int64_t off = 0; off < n; off += batch, ++b, epoch += 2) {
        const int64_t cur = (n - off < batch) ? (n - off) : batch;
        const int64_t gb = (cur == batch) ? global_batch : cur * model->world;
        int rc = brs_mf_sharded_step(model, sync, opt, users + off, pos_items + off, neg_items + off, cur, gb, reg_weight,
                                     epoch, out + 4 * b, stream);
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

// ---------------------------------------------------------------------------
// gather / scatter-add micro-ops (config 5): warp-group per index, 128-bit accesses
// ---------------------------------------------------------------------------
namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

enum { OP_GATHER = 0, OP_SCATTER_ADD = 1, OP_GATHER_SGD = 2 };

template <int LPR, int VPL, int OP>
__global__ void __launch_bounds__(kThreads) rows_op_kernel(float* __restrict__ table, long long n_rows, int D,
                                                           const long long* __restrict__ idx, long long n,
                                                           float* __restrict__ buf, float scale) {
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR, grp = lane / LPR;
    const long long warp_global = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * kWarps;
    for (long long base = warp_global * SPW; base < n; base += n_warps * SPW) {
        const long long s = base + grp;
        if (s >= n) continue;
        const long long r = idx[s];
        if ((unsigned long long)r >= (unsigned long long)n_rows) continue;
        float* row = table + r * D;
        float* b = buf + s * D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
            if (col < D) {
                if (OP == OP_GATHER) {
                    *(float4*)(b + col) = ld_row4(row + col);
                } else if (OP == OP_SCATTER_ADD) {
                    float4 x = *(const float4*)(b + col);
                    red_add4(row + col, make_float4(scale * x.x, scale * x.y, scale * x.z, scale * x.w));
                } else {  // gather, hand the row out, apply an SGD-style update in place
                    float4 x = ld_row4(row + col);
                    *(float4*)(b + col) = x;
                    red_add4(row + col, make_float4(-scale * x.x, -scale * x.y, -scale * x.z, -scale * x.w));
                }
            }
        }
    }
}

template <int OP>
int launch_rows_op(float* table, int64_t n_rows, int32_t D, const int64_t* idx, int64_t n, float* buf, float scale,
                   void* stream) {
    if (!table || !idx || !buf || n_rows < 0 || n < 0) return BRS_ERR_INVALID_ARG;
    if (D <= 0 || D % 4 != 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    if ((((uintptr_t)table | (uintptr_t)buf) & 15) != 0) return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define BRS_ROWS(LPR, VPL)                                                                                       \
    do {                                                                                                         \
        auto k = rows_op_kernel<LPR, VPL, OP>;                                                                   \
        int per_sm = 1;                                                                                          \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);                                  \
        long long blocks = (n + kWarps * (32 / LPR) - 1) / (kWarps * (32 / LPR));                                \
        long long grid = (long long)brs_sm_count() * (per_sm < 1 ? 1 : per_sm);                                  \
        if (grid > blocks) grid = blocks;                                                                        \
        k<<<(int)grid, kThreads, 0, st>>>(table, n_rows, D, (const long long*)idx, n, buf, scale);               \
    } while (0)
    if (D <= 4) BRS_ROWS(1, 1);
    else if (D <= 8) BRS_ROWS(2, 1);
    else if (D <= 16) BRS_ROWS(4, 1);
    else if (D <= 32) BRS_ROWS(8, 1);
    else if (D <= 64) BRS_ROWS(16, 1);
    else if (D <= 128) BRS_ROWS(32, 1);
    else if (D <= 256) BRS_ROWS(32, 2);
    else if (D <= 384) BRS_ROWS(32, 3);
    else BRS_ROWS(32, 4);
#undef BRS_ROWS
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
}  // namespace

extern "C" int brs_gather(const float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n, float* out,
                          void* stream) {
    return launch_rows_op<OP_GATHER>(const_cast<float*>(table), n_rows, dim, idx, n, out, 0.f, stream);
}

extern "C" int brs_scatter_add(float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n,
                               const float* src, float scale, void* stream) {
    return launch_rows_op<OP_SCATTER_ADD>(table, n_rows, dim, idx, n, const_cast<float*>(src), scale, stream);
}

extern "C" int brs_gather_sgd_update(float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n,
                                     float lr, float* out, void* stream) {
    return launch_rows_op<OP_GATHER_SGD>(table, n_rows, dim, idx, n, out, lr, stream);
}
