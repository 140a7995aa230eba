// LightGCN training step -- sm_100a.
//
// Replaces, per batch (beta_rec/models/lightgcn.py:27-78,119-152,171-191):
//   dropout(A)                 -> edges consumed through a keep mask (drawn by the host with the
//                                 reference's own CPU generator call so it is bit-identical, :32)
//   L x torch.sparse.mm        -> spmm_csr_kernel: E^(l+1) = A' E^(l), nnz-balanced chunks
//   stack / mean / split       -> never materialised: the tail gathers the L+1 layer rows it needs
//   softplus-BPR + L2 on E^(0) -> lightgcn_tail_kernel (loss, d ebar rows scattered with 128-bit REDs)
//   backward of the propagate  -> the same SpMM kernel on A'^T (CSR of the transpose, edge ids map
//                                 back to the forward keep mask): G_l = d + A'^T G_(l+1)
//   Adam / SGD over all rows   -> dense parameter step (rows_apply.cu); LightGCN gradients are dense
// Row-normalised D^-1(A+I) is asymmetric (beta_rec/utils/common_util.py:24-41), hence the explicit
// transpose.  HBM-bound: per SpMM ~ nnz*(8 B + 4*D B) + N*4*D B.
#include "common.cuh"

int brs_apply_impl(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                   int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                   long long batch, long long max_rows_hint, void* stream);

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 512;  // nonzeros per warp work item

struct SpmmArgs {
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const float* __restrict__ val;
    const int* __restrict__ edge_id;            // NULL: edge e uses keep[e]
    const unsigned char* __restrict__ keep;     // NULL: no dropout
    float inv_keep;
    const float* __restrict__ x;  // [n_cols, D]
    float* y;                     // [n_rows, D], pre-initialised; contributions are RED-added
    long long n_rows, nnz;
    int dim;
};

// first row r with row_ptr[r+1] > e
__device__ __forceinline__ int row_of_edge(const int* __restrict__ row_ptr, long long n_rows, int e) {
    long long lo = 0, hi = n_rows;  // answer in [lo, hi)
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
    }
    return (int)lo;
}

// y[r, :] += sum_{e in row r, kept} val[e]*inv_keep * x[col[e], :]
// LPR lanes cover a row (float4 each, VPL float4 per lane); the warp's G = 32/LPR lane groups take
// alternate edges of the chunk; a group flushes its partial row sum with a RED when its row changes.
template <int LPR, int VPL>
__global__ void __launch_bounds__(kThreads) spmm_csr_kernel(const SpmmArgs a) {
    constexpr int G = 32 / LPR;
    constexpr int U = 4;  // edges in flight per group
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR, grp = lane / LPR;
    const int D = a.dim;
    const long long n_chunks = (a.nnz + kChunk - 1) / kChunk;
    for (long long c = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); c < n_chunks; c += (long long)gridDim.x * kWarps) {
        const int e0 = (int)(c * kChunk);
        const int e1 = (int)min((long long)a.nnz, (long long)e0 + kChunk);
        int row = row_of_edge(a.row_ptr, a.n_rows, e0 + grp < e1 ? e0 + grp : e0);
        int row_end = __ldg(a.row_ptr + row + 1);
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool dirty = false;
        for (int eb = e0 + grp; eb < e1; eb += G * U) {
            int cidx[U];
            float w[U];
            float4 xv[U][VPL];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int e = eb + k * G;
                w[k] = 0.f;
                cidx[k] = 0;
                if (e < e1) {
                    bool on = true;
                    if (a.keep) on = a.keep[a.edge_id ? __ldg(a.edge_id + e) : e] != 0;
                    if (on) {
                        w[k] = __ldg(a.val + e) * a.inv_keep;
                        cidx[k] = __ldg(a.col + e);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float* xr = a.x + (size_t)(unsigned)cidx[k] * (unsigned)D;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int colf = (v * LPR + gl) * 4;
                    xv[k][v] = (w[k] != 0.f && colf < D) ? ld_row4(xr + colf) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int e = eb + k * G;
                if (e >= e1) break;
                if (e >= row_end) {  // this group's edge starts a later row: flush, then advance
                    if (dirty) {
                        float* yr = a.y + (size_t)(unsigned)row * (unsigned)D;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) {
                            const int colf = (v * LPR + gl) * 4;
                            if (colf < D) red_add4(yr + colf, acc[v]);
                            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        dirty = false;
                    }
                    while (e >= row_end) {
                        ++row;
                        row_end = __ldg(a.row_ptr + row + 1);
                    }
                }
                if (w[k] != 0.f) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        acc[v].x = fmaf(w[k], xv[k][v].x, acc[v].x);
                        acc[v].y = fmaf(w[k], xv[k][v].y, acc[v].y);
                        acc[v].z = fmaf(w[k], xv[k][v].z, acc[v].z);
                        acc[v].w = fmaf(w[k], xv[k][v].w, acc[v].w);
                    }
                    dirty = true;
                }
            }
        }
        if (dirty) {
            float* yr = a.y + (size_t)(unsigned)row * (unsigned)D;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int colf = (v * LPR + gl) * 4;
                if (colf < D) red_add4(yr + colf, acc[v]);
            }
        }
    }
}

int launch_spmm(const SpmmArgs& a, cudaStream_t st) {
    const int D = a.dim;
    if (D <= 0 || D % 4 != 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    if (a.nnz == 0) return BRS_OK;
    const long long n_chunks = (a.nnz + kChunk - 1) / kChunk;
#define BRS_SPMM(LPR, VPL)                                                           \
    do {                                                                             \
        auto k = spmm_csr_kernel<LPR, VPL>;                                          \
        int per_sm = 1;                                                              \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);      \
        long long grid = (long long)brs_sm_count() * (per_sm < 1 ? 1 : per_sm);      \
        const long long need = (n_chunks + kWarps - 1) / kWarps;                     \
        if (grid > need) grid = need;                                                \
        k<<<(int)grid, kThreads, 0, st>>>(a);                                        \
    } while (0)
    if (D <= 4) BRS_SPMM(1, 1);
    else if (D <= 8) BRS_SPMM(2, 1);
    else if (D <= 16) BRS_SPMM(4, 1);
    else if (D <= 32) BRS_SPMM(8, 1);
    else if (D <= 64) BRS_SPMM(16, 1);
    else if (D <= 128) BRS_SPMM(32, 1);
    else if (D <= 256) BRS_SPMM(32, 2);
    else BRS_SPMM(32, 4);
#undef BRS_SPMM
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// ---------------------------------------------------------------------------
// tail: softplus-BPR on the layer-mean embeddings + L2 on the layer-0 rows
// ---------------------------------------------------------------------------
struct TailArgs {
    const float* emb[BRS_LGCN_MAX_LAYERS + 1];  // emb[l]: [N, D], users first then items
    int n_layers;
    long long n_users, n_items;
    int dim;
    const long long* users; const long long* pos; const long long* neg;
    long long batch;
    float inv_b, decay, inv_lp1;
    float* d;        // [N, D] zeroed: receives d loss / d E^(l) (same for every l) = d ebar / (L+1)
    float* scores;   // predict
    brs_step_ws* ws;
    int train;
    long long row_lo, row_hi;  // lightgcn_reg_grad_kernel: only node rows in [row_lo, row_hi) (row_hi == 0: all rows)
};

// one warp per sample; lanes own float4 columns (D <= 512)
constexpr int kTailV = 4;
__global__ void __launch_bounds__(kThreads) lightgcn_tail_kernel(const TailArgs a) {
    __shared__ float s_red[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.dim;
    float loss_acc = 0.f;
    for (long long s = (long long)blockIdx.x * kWarps + warp; s < a.batch; s += (long long)gridDim.x * kWarps) {
        long long u = a.users[s], i = a.pos[s], j = a.train ? a.neg[s] : 0;
        bool ok = true;
        if ((unsigned long long)u >= (unsigned long long)a.n_users || (unsigned long long)i >= (unsigned long long)a.n_items ||
            (unsigned long long)j >= (unsigned long long)a.n_items) {
            if (lane == 0) atomicOr(a.train ? &a.ws->err_flag : &a.ws->predict_err, 1u);
            ok = false;
            u = i = j = 0;
        }
        const size_t ru = (size_t)u * D, ri = (size_t)(a.n_users + i) * D, rj = (size_t)(a.n_users + j) * D;
        float4 eu[kTailV], ei[kTailV], ej[kTailV];
        float ps = 0.f, ns = 0.f, r0 = 0.f;
#pragma unroll
        for (int v = 0; v < kTailV; ++v) {
            const int c = v * 128 + lane * 4;
            eu[v] = ei[v] = ej[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < D) {
                for (int l = 0; l <= a.n_layers; ++l) {
                    const float4 x = ld_row4(a.emb[l] + ru + c), y = ld_row4(a.emb[l] + ri + c);
                    eu[v].x += x.x; eu[v].y += x.y; eu[v].z += x.z; eu[v].w += x.w;
                    ei[v].x += y.x; ei[v].y += y.y; ei[v].z += y.z; ei[v].w += y.w;
                    if (l == 0) r0 += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w + y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
                    if (a.train) {
                        const float4 z = ld_row4(a.emb[l] + rj + c);
                        ej[v].x += z.x; ej[v].y += z.y; ej[v].z += z.z; ej[v].w += z.w;
                        if (l == 0) r0 += z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
                    }
                }
                // torch.mean over the stacked layers (lightgcn.py:75-76)
                eu[v].x *= a.inv_lp1; eu[v].y *= a.inv_lp1; eu[v].z *= a.inv_lp1; eu[v].w *= a.inv_lp1;
                ei[v].x *= a.inv_lp1; ei[v].y *= a.inv_lp1; ei[v].z *= a.inv_lp1; ei[v].w *= a.inv_lp1;
                ej[v].x *= a.inv_lp1; ej[v].y *= a.inv_lp1; ej[v].z *= a.inv_lp1; ej[v].w *= a.inv_lp1;
                ps += eu[v].x * ei[v].x + eu[v].y * ei[v].y + eu[v].z * ei[v].z + eu[v].w * ei[v].w;
                ns += eu[v].x * ej[v].x + eu[v].y * ej[v].y + eu[v].z * ej[v].z + eu[v].w * ej[v].w;
            }
        }
        ps = warp_sum(ps);
        if (!a.train) {
            if (lane == 0 && ok) a.scores[s] = sigmoidf_(ps);  // LightGCN.predict (lightgcn.py:100)
            continue;
        }
        ns = warp_sum(ns);
        r0 = warp_sum(r0);
        if (!ok) continue;
        const float x = ns - ps;
        if (lane == 0) loss_acc += softplusf_(x) + a.decay * 0.5f * r0;  // both are divided by B at the end
        // d mean(softplus(x)) / dx = sigmoid(x)/B ; every layer l receives d ebar/(L+1)
        const float cg = sigmoidf_(x) * a.inv_b * a.inv_lp1;
#pragma unroll
        for (int v = 0; v < kTailV; ++v) {
            const int c = v * 128 + lane * 4;
            if (c < D) {
                red_add4(a.d + ru + c, make_float4(cg * (ej[v].x - ei[v].x), cg * (ej[v].y - ei[v].y),
                                                   cg * (ej[v].z - ei[v].z), cg * (ej[v].w - ei[v].w)));
                red_add4(a.d + ri + c, make_float4(-cg * eu[v].x, -cg * eu[v].y, -cg * eu[v].z, -cg * eu[v].w));
                red_add4(a.d + rj + c, make_float4(cg * eu[v].x, cg * eu[v].y, cg * eu[v].z, cg * eu[v].w));
            }
        }
    }
    if (!a.train) return;
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) s_red[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        for (int w = 0; w < kWarps; ++w) l += s_red[w];
        atomicAdd(&a.ws->loss_sum, (double)l);
    }
}

// g[row] += (decay/B) * E0[row] for the batch's layer-0 rows (gradient of the L2 term, lightgcn.py:179-188)
__global__ void __launch_bounds__(kThreads) lightgcn_reg_grad_kernel(const TailArgs a, float* g) {
    const int lane = threadIdx.x & 31;
    const int D = a.dim;
    const float k = a.decay * a.inv_b;
    for (long long t = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); t < 3 * a.batch; t += (long long)gridDim.x * kWarps) {
        const long long s = t / 3;
        const int which = (int)(t - 3 * s);
        const long long id = which == 0 ? a.users[s] : (which == 1 ? a.pos[s] : a.neg[s]);
        const long long lim = which == 0 ? a.n_users : a.n_items;
        if ((unsigned long long)id >= (unsigned long long)lim) continue;
        const long long node = which == 0 ? id : a.n_users + id;
        if (a.row_hi > 0 && (node < a.row_lo || node >= a.row_hi)) continue;  // another rank owns this row
        const size_t r = (size_t)node * D;
        float* gr = g + (size_t)(node - (a.row_hi > 0 ? a.row_lo : 0)) * D;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 x = ld_row4(a.emb[0] + r + c);
            red_add4(gr + c, make_float4(k * x.x, k * x.y, k * x.z, k * x.w));
        }
    }
}

int warp_grid(long long n_items_of_work, const void* k) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);
    long long g = (long long)brs_sm_count() * (per_sm < 1 ? 1 : per_sm);
    const long long need = (n_items_of_work + kWarps - 1) / kWarps;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}

int check(const brs_lightgcn_model* m) {
    if (!m || !m->ws || m->n_layers < 1 || m->n_layers > BRS_LGCN_MAX_LAYERS) return BRS_ERR_INVALID_ARG;
    if (m->dim <= 0 || m->dim % 4 != 0 || m->dim > 512) return BRS_ERR_UNSUPPORTED;
    const long long n = m->n_users + m->n_items;
    if (m->adj.n_rows != n || m->adj_t.n_rows != n || m->adj.nnz != m->adj_t.nnz) return BRS_ERR_INVALID_ARG;
    if (!m->adj.row_ptr || !m->adj.col || !m->adj.val || !m->adj_t.row_ptr || !m->adj_t.col || !m->adj_t.val)
        return BRS_ERR_INVALID_ARG;
    for (int l = 0; l <= m->n_layers; ++l)
        if (!m->emb[l]) return BRS_ERR_INVALID_ARG;
    if (!m->d || !m->g[0] || !m->g[1] || !m->param.weight) return BRS_ERR_INVALID_ARG;
    return BRS_OK;
}

SpmmArgs spmm_args(const brs_csr& csr, const unsigned char* keep, float keep_prob, const float* x, float* y, int dim) {
    SpmmArgs a;
    a.row_ptr = csr.row_ptr;
    a.col = csr.col;
    a.val = csr.val;
    a.edge_id = csr.edge_id;
    a.keep = keep;
    a.inv_keep = keep ? 1.0f / keep_prob : 1.0f;
    a.x = x;
    a.y = y;
    a.n_rows = csr.n_rows;
    a.nnz = csr.nnz;
    a.dim = dim;
    return a;
}

}  // namespace

extern "C" int brs_spmm_csr(const brs_csr* a, const uint8_t* keep_mask, float keep_prob, const float* x, float* y,
                            int32_t dim, void* stream) {
    if (!a || !a->row_ptr || !a->col || !a->val || !x || !y || a->nnz < 0 || a->nnz > 0x7fffffffLL) return BRS_ERR_INVALID_ARG;
    if (keep_mask && !(keep_prob > 0.f)) return BRS_ERR_INVALID_ARG;
    return launch_spmm(spmm_args(*a, keep_mask, keep_prob, x, y, dim), (cudaStream_t)stream);
}

// E^(l+1) = A' E^(l) for l = 0..L-1 (LightGCN.forward, lightgcn.py:71-74)
extern "C" int brs_lightgcn_propagate(const brs_lightgcn_model* m, const uint8_t* keep_mask, float keep_prob,
                                      void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)(m->n_users + m->n_items) * m->dim * sizeof(float);
    for (int l = 0; l < m->n_layers; ++l) {
        BRS_CUDA_CHECK(cudaMemsetAsync(m->emb[l + 1], 0, bytes, st));
        rc = launch_spmm(spmm_args(m->adj, keep_mask, keep_prob, m->emb[l], m->emb[l + 1], m->dim), st);
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

extern "C" int brs_lightgcn_fwd_bwd(const brs_lightgcn_model* m, const uint8_t* keep_mask, float keep_prob,
                                    const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                                    int64_t batch, void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    if (!users || !pos_items || !neg_items || batch <= 0) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)(m->n_users + m->n_items) * m->dim * sizeof(float);
    rc = brs_lightgcn_propagate(m, keep_mask, keep_prob, stream);
    if (rc != BRS_OK) return rc;
    TailArgs t;
    memset(&t, 0, sizeof(t));
    for (int l = 0; l <= m->n_layers; ++l) t.emb[l] = m->emb[l];
    t.n_layers = m->n_layers;
    t.n_users = m->n_users;
    t.n_items = m->n_items;
    t.dim = m->dim;
    t.users = (const long long*)users;
    t.pos = (const long long*)pos_items;
    t.neg = (const long long*)neg_items;
    t.batch = batch;
    t.inv_b = 1.0f / (float)batch;
    t.decay = m->decay;
    t.inv_lp1 = 1.0f / (float)(m->n_layers + 1);
    t.d = m->d;
    t.ws = (brs_step_ws*)m->ws;
    t.train = 1;
    BRS_CUDA_CHECK(cudaMemsetAsync(m->d, 0, bytes, st));
    lightgcn_tail_kernel<<<warp_grid(batch, (const void*)lightgcn_tail_kernel), kThreads, 0, st>>>(t);
    BRS_CUDA_CHECK(cudaGetLastError());
    // backward of the linear propagate: G_L = d ; G_l = d + A'^T G_(l+1) ; result G_0 ends in m->param.grad
    const float* cur = m->d;
    for (int l = m->n_layers - 1; l >= 0; --l) {
        float* out = (l == 0) ? m->param.grad : m->g[l & 1];
        BRS_CUDA_CHECK(cudaMemcpyAsync(out, m->d, bytes, cudaMemcpyDeviceToDevice, st));
        rc = launch_spmm(spmm_args(m->adj_t, keep_mask, keep_prob, cur, out, m->dim), st);
        if (rc != BRS_OK) return rc;
        cur = out;
    }
    lightgcn_reg_grad_kernel<<<warp_grid(3 * batch, (const void*)lightgcn_reg_grad_kernel), kThreads, 0, st>>>(t, m->param.grad);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// ---- pieces of brs_lightgcn_fwd_bwd for the row-partitioned multi-GPU step (sharded_lightgcn.py), which runs the
// propagate and its backward as per-rank block SpMMs (brs_spmm_csr) with collectives in between
static void tail_args(const brs_lightgcn_model* m, const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                      int64_t batch, int64_t global_batch, TailArgs& t) {
    memset(&t, 0, sizeof(t));
    for (int l = 0; l <= m->n_layers; ++l) t.emb[l] = m->emb[l];
    t.n_layers = m->n_layers;
    t.n_users = m->n_users;
    t.n_items = m->n_items;
    t.dim = m->dim;
    t.users = (const long long*)users;
    t.pos = (const long long*)pos_items;
    t.neg = (const long long*)neg_items;
    t.batch = batch;
    t.inv_b = 1.0f / (float)global_batch;
    t.decay = m->decay;
    t.inv_lp1 = 1.0f / (float)(m->n_layers + 1);
    t.d = m->d;
    t.ws = (brs_step_ws*)m->ws;
    t.train = 1;
}

extern "C" int brs_lightgcn_tail(const brs_lightgcn_model* m, const int64_t* users, const int64_t* pos_items,
                                 const int64_t* neg_items, int64_t batch, int64_t global_batch, void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    if (!users || !pos_items || !neg_items || batch <= 0 || global_batch < batch) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TailArgs t;
    tail_args(m, users, pos_items, neg_items, batch, global_batch, t);
    BRS_CUDA_CHECK(cudaMemsetAsync(m->d, 0, (size_t)(m->n_users + m->n_items) * m->dim * sizeof(float), st));
    lightgcn_tail_kernel<<<warp_grid(batch, (const void*)lightgcn_tail_kernel), kThreads, 0, st>>>(t);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_lightgcn_reg_grad(const brs_lightgcn_model* m, const int64_t* users, const int64_t* pos_items,
                                     const int64_t* neg_items, int64_t batch, int64_t global_batch, int64_t row_lo,
                                     int64_t row_hi, float* grad_rows, void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    if (!users || !pos_items || !neg_items || batch <= 0 || !grad_rows || row_lo < 0 || row_hi <= row_lo) return BRS_ERR_INVALID_ARG;
    TailArgs t;
    tail_args(m, users, pos_items, neg_items, batch, global_batch, t);
    t.row_lo = row_lo;
    t.row_hi = row_hi;
    lightgcn_reg_grad_kernel<<<warp_grid(3 * batch, (const void*)lightgcn_reg_grad_kernel), kThreads, 0, (cudaStream_t)stream>>>(t, grad_rows);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_lightgcn_apply(const brs_lightgcn_model* m, const brs_opt* opt, int64_t batch, float* out,
                                  void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    if (!opt || batch <= 0 || !m->param.grad) return BRS_ERR_INVALID_ARG;
    return brs_apply_impl(nullptr, 0, &m->param, 1, 0, opt, m->ws, 0, out, batch, 0, stream);
}

// sigmoid(ebar_u . ebar_i) from the layer buffers left by brs_lightgcn_propagate (LightGCN.predict)
extern "C" int brs_lightgcn_scores(const brs_lightgcn_model* m, const int64_t* users, const int64_t* items, int64_t n,
                                   float* scores, void* stream) {
    int rc = check(m);
    if (rc != BRS_OK) return rc;
    if (!users || !items || !scores || n < 0) return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    TailArgs t;
    memset(&t, 0, sizeof(t));
    for (int l = 0; l <= m->n_layers; ++l) t.emb[l] = m->emb[l];
    t.n_layers = m->n_layers;
    t.n_users = m->n_users;
    t.n_items = m->n_items;
    t.dim = m->dim;
    t.users = (const long long*)users;
    t.pos = (const long long*)items;
    t.batch = n;
    t.inv_lp1 = 1.0f / (float)(m->n_layers + 1);
    t.scores = scores;
    t.ws = (brs_step_ws*)m->ws;
    t.train = 0;
    lightgcn_tail_kernel<<<warp_grid(n, (const void*)lightgcn_tail_kernel), kThreads, 0, (cudaStream_t)stream>>>(t);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
