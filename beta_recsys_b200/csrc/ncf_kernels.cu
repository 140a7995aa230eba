// GMF / MLP / NeuMF (NCF family) training step -- sm_100a.
//
// Replaces, per batch (beta_rec/models/ncf.py:52-71,100-120; gmf.py:29-36,60-80;
// mlp.py:40-51,75-98): 4 embedding gathers, the concat, the fc tower with the
// reference's ReLU placement, the affine head, sigmoid, nn.BCELoss, autograd's
// backward (incl. embedding_dense_backward) -- as a chain of hand-written kernels:
//
//   ncf_gather   : rows -> X0 = [relu](cat(u_mlp, i_mlp)),  MFV = u_mf * i_mf, mark touched rows
//   linear fwd   : H_{l+1} = relu(H_l W_l^T + b_l)                (gemm_simt.cu / gemm_tc.cu)
//   ncf_head     : z = [H_L, MFV].w_o + b_o, sigmoid, BCE, dz, dH_L = dz*w_o*(H_L>0), d w_o, d b_o
//   linear bwd   : dW_l += dH_{l+1}^T H_l ; dH_l = (dH_{l+1} W_l) * (H_l > 0)
//   ncf_scatter  : 128-bit RED of dX0 halves and of the MF-part gradients into the tables' scratch
//
// NeuMF quirk kept (ncf.py:64-66): an extra ReLU follows EVERY sub-module of fc_layers,
// including the leading Dropout -> the concatenated embeddings are ReLU'd before the first
// Linear (so dX0 is masked by X0 > 0).  MLP (mlp.py:47-48) has no such ReLU on its input.
// Dropout must be 0 (the configs' default); p > 0 is rejected by the host engine.
#include "common.cuh"

int brs_linear_fwd_simt(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int M, int N, int K,
                        bool relu, cudaStream_t st);
int brs_linear_dgrad_simt(const float* dY, int ldy, const float* W, float* dX, int ldx, const float* mask_src, int ldm,
                          int M, int N, int K, cudaStream_t st);
int brs_linear_wgrad_simt(const float* dY, int ldy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                          cudaStream_t st);
bool brs_linear_tc_supported(int M, int N, int K);
int brs_linear_tc(const float* A, const float* B, const float* bias, float* Y, const float* mask, int M, int N, int K,
                  bool relu, cudaStream_t st);
int brs_transpose(const float* src, float* dst, int R, int C, cudaStream_t st);
int brs_assign_slots(const brs_rowset* rs, const long long* const* idx, const long long* n, int n_arrays,
                     brs_step_ws* ws, cudaStream_t st);
int brs_apply_impl(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                   int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                   long long batch, long long max_rows_hint, void* stream);

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// 0 = exact-fp32 FFMA kernels (gemm_simt.cu), 1 = tcgen05 3xTF32 (gemm_tc.cu) where the shape allows
int g_gemm_backend = 1;

struct NcfArgs {
    int kind;      // BRS_NCF_GMF / MLP / NEUMF
    int mlp_dim;   // Lm: width of the MLP-side embedding rows (0 for GMF)
    int mf_dim;    // E : width of the MF-side embedding rows (0 for MLP)
    int head_in;   // tower output width (emb_dim) for MLP/NeuMF, 0 for GMF
    const float* u_mlp; const float* i_mlp; const float* u_mf; const float* i_mf;
    float* g_u_mlp; float* g_i_mlp; float* g_u_mf; float* g_i_mf;  // compact scratch, indexed by slot
    const int* user_slot; const int* item_slot;                      // slot maps filled by the pre-pass
    int user_cap, item_cap;                                          // scratch capacities (gs_off layout)
    long long n_users, n_items;
    const long long* users; const long long* items; const float* ratings;
    long long batch;
    float inv_b;
    float* x0;    // [B, 2*Lm]
    float* mfv;   // [B, E]
    float* h_last;   // [B, head_in]
    float* dh_last;  // [B, head_in]
    float* dx0;   // [B, 2*Lm]
    float* dz;    // [B]
    const float* w_out; const float* b_out;  // affine_output [1, head_in + E], [1]
    float* g_w_out; float* g_b_out;
    float* scores;  // predict only
    brs_step_ws* ws;
    int train;
};

// out-of-range ids are flagged by the pre-pass (train) or here (predict); such samples are skipped
__device__ __forceinline__ bool fetch_ids(const NcfArgs& a, long long s, int lane, long long& u, long long& i) {
    u = a.users[s];
    i = a.items[s];
    if ((unsigned long long)u >= (unsigned long long)a.n_users || (unsigned long long)i >= (unsigned long long)a.n_items) {
        if (lane == 0 && !a.train) atomicOr(&a.ws->predict_err, 1u);
        u = i = 0;
        return false;
    }
    return true;
}

// ---- gather: one warp per sample -------------------------------------------
__global__ void __launch_bounds__(kThreads) ncf_gather_kernel(const NcfArgs a) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarps;
    const bool relu_in = a.kind == BRS_NCF_NEUMF;
    for (long long s = w0; s < a.batch; s += nw) {
        long long u, i;
        fetch_ids(a, s, lane, u, i);
        if (a.mlp_dim) {
            float* x = a.x0 + s * 2 * a.mlp_dim;
            for (int c = lane * 4; c < a.mlp_dim; c += 128) {
                float4 um = ld_row4(a.u_mlp + u * a.mlp_dim + c);
                float4 im = ld_row4(a.i_mlp + i * a.mlp_dim + c);
                if (relu_in) {
                    um = make_float4(fmaxf(um.x, 0.f), fmaxf(um.y, 0.f), fmaxf(um.z, 0.f), fmaxf(um.w, 0.f));
                    im = make_float4(fmaxf(im.x, 0.f), fmaxf(im.y, 0.f), fmaxf(im.z, 0.f), fmaxf(im.w, 0.f));
                }
                *(float4*)(x + c) = um;
                *(float4*)(x + a.mlp_dim + c) = im;
            }
        }
        if (a.mf_dim && a.kind == BRS_NCF_NEUMF) {
            float* m = a.mfv + s * a.mf_dim;
            for (int c = lane * 4; c < a.mf_dim; c += 128) {
                const float4 uf = ld_row4(a.u_mf + u * a.mf_dim + c);
                const float4 jf = ld_row4(a.i_mf + i * a.mf_dim + c);
                *(float4*)(m + c) = make_float4(uf.x * jf.x, uf.y * jf.y, uf.z * jf.z, uf.w * jf.w);
            }
        }
    }
}

__device__ __forceinline__ void bce_terms(float s, float r, float inv_b, float& loss_k, float& dz) {
    // nn.BCELoss: logs clamped at -100; backward (s-r)/max((1-s)s, 1e-12)/B, then sigmoid'
    loss_k = -(r * fmaxf(logf(s), -100.f) + (1.0f - r) * fmaxf(log1pf(-s), -100.f));
    const float ds = (s - r) / fmaxf((1.0f - s) * s, 1e-12f) * inv_b;
    dz = ds * s * (1.0f - s);
}

// ---- head (MLP / NeuMF): one warp per sample; lanes own fixed columns of w_out ----
constexpr int kHeadMaxV = 4;  // supports head_in + mf_dim <= 512
__global__ void __launch_bounds__(kThreads) ncf_head_kernel(const NcfArgs a) {
    __shared__ float s_gw[kWarps][kHeadMaxV * 128];
    __shared__ float s_sc[2][kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w0 = (long long)blockIdx.x * kWarps + warp;
    const long long nw = (long long)gridDim.x * kWarps;
    const int H = a.head_in, E = (a.kind == BRS_NCF_NEUMF) ? a.mf_dim : 0, W = H + E;
    const float bo = __ldg(a.b_out);
    float4 wv[kHeadMaxV], gw[kHeadMaxV];
#pragma unroll
    for (int v = 0; v < kHeadMaxV; ++v) {
        const int c = v * 128 + lane * 4;
        wv[v] = (c < W) ? *(const float4*)(a.w_out + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        gw[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float loss_acc = 0.f, gb_acc = 0.f;
    for (long long s = w0; s < a.batch; s += nw) {
        float4 x[kHeadMaxV];
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < kHeadMaxV; ++v) {
            const int c = v * 128 + lane * 4;
            x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < H)
                x[v] = *(const float4*)(a.h_last + s * H + c);
            else if (c < W)
                x[v] = *(const float4*)(a.mfv + s * E + (c - H));
            dot += x[v].x * wv[v].x + x[v].y * wv[v].y + x[v].z * wv[v].z + x[v].w * wv[v].w;
        }
        dot = warp_sum(dot);
        const float sc = sigmoidf_(dot + bo);
        if (!a.train) {
            if (lane == 0) a.scores[s] = sc;
            continue;
        }
        float loss_k, dz;
        bce_terms(sc, a.ratings[s], a.inv_b, loss_k, dz);
        if (lane == 0) {
            a.dz[s] = dz;
            loss_acc += loss_k;
            gb_acc += dz;
        }
#pragma unroll
        for (int v = 0; v < kHeadMaxV; ++v) {
            const int c = v * 128 + lane * 4;
            if (c < W) {
                gw[v].x += dz * x[v].x; gw[v].y += dz * x[v].y; gw[v].z += dz * x[v].z; gw[v].w += dz * x[v].w;
            }
            if (c < H) {  // dH_L = dz * w_o, through the tower's last ReLU
                float4 d;
                d.x = x[v].x > 0.f ? dz * wv[v].x : 0.f;
                d.y = x[v].y > 0.f ? dz * wv[v].y : 0.f;
                d.z = x[v].z > 0.f ? dz * wv[v].z : 0.f;
                d.w = x[v].w > 0.f ? dz * wv[v].w : 0.f;
                *(float4*)(a.dh_last + s * H + c) = d;
            }
        }
    }
    if (!a.train) return;
    // block reduction of d w_out (per-lane column ownership) and the scalars
#pragma unroll
    for (int v = 0; v < kHeadMaxV; ++v) *(float4*)&s_gw[warp][v * 128 + lane * 4] = gw[v];
    loss_acc = warp_sum(loss_acc);
    gb_acc = warp_sum(gb_acc);
    if (lane == 0) {
        s_sc[0][warp] = loss_acc;
        s_sc[1][warp] = gb_acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_gw[w][c];
        red_add1(a.g_w_out + c, t);
    }
    if (threadIdx.x == 0) {
        float l = 0.f, g = 0.f;
        for (int w = 0; w < kWarps; ++w) {
            l += s_sc[0][w];
            g += s_sc[1][w];
        }
        atomicAdd(&a.ws->loss_sum, (double)l);
        red_add1(a.g_b_out, g);
    }
}

// ---- scatter (MLP / NeuMF): one warp per sample ----------------------------
__global__ void __launch_bounds__(kThreads) ncf_scatter_kernel(const NcfArgs a) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarps;
    const int H = a.head_in;
    for (long long s = w0; s < a.batch; s += nw) {
        long long u, i;
        if (!fetch_ids(a, s, lane, u, i)) continue;
        const long long su = a.user_slot[u], si = a.item_slot[i];
        if (su < 0 || si < 0) continue;  // capacity overflow, flagged by the pre-pass
        if (a.mlp_dim) {
            const float* d = a.dx0 + s * 2 * a.mlp_dim;
            for (int c = lane * 4; c < a.mlp_dim; c += 128) {
                red_add4(a.g_u_mlp + gs_off(a.mlp_dim, a.user_cap, (unsigned)su, c), *(const float4*)(d + c));
                red_add4(a.g_i_mlp + gs_off(a.mlp_dim, a.item_cap, (unsigned)si, c), *(const float4*)(d + a.mlp_dim + c));
            }
        }
        if (a.kind == BRS_NCF_NEUMF) {
            const float dz = a.dz[s];
            for (int c = lane * 4; c < a.mf_dim; c += 128) {
                const float4 w = *(const float4*)(a.w_out + H + c);
                const float4 uf = ld_row4(a.u_mf + u * a.mf_dim + c);
                const float4 jf = ld_row4(a.i_mf + i * a.mf_dim + c);
                red_add4(a.g_u_mf + gs_off(a.mf_dim, a.user_cap, (unsigned)su, c), make_float4(dz * w.x * jf.x, dz * w.y * jf.y, dz * w.z * jf.z, dz * w.w * jf.w));
                red_add4(a.g_i_mf + gs_off(a.mf_dim, a.item_cap, (unsigned)si, c), make_float4(dz * w.x * uf.x, dz * w.y * uf.y, dz * w.z * uf.z, dz * w.w * uf.w));
            }
        }
    }
}

// ---- GMF: everything in one kernel (gmf.py:29-36 + BCELoss + backward) ------
__global__ void __launch_bounds__(kThreads) gmf_fused_kernel(const NcfArgs a) {
    __shared__ float s_gw[kWarps][kHeadMaxV * 128];
    __shared__ float s_sc[2][kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w0 = (long long)blockIdx.x * kWarps + warp;
    const long long nw = (long long)gridDim.x * kWarps;
    const int E = a.mf_dim;
    const float bo = __ldg(a.b_out);
    float4 wv[kHeadMaxV], gw[kHeadMaxV];
#pragma unroll
    for (int v = 0; v < kHeadMaxV; ++v) {
        const int c = v * 128 + lane * 4;
        wv[v] = (c < E) ? *(const float4*)(a.w_out + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        gw[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float loss_acc = 0.f, gb_acc = 0.f;
    for (long long s = w0; s < a.batch; s += nw) {
        long long u, i;
        const bool ok = fetch_ids(a, s, lane, u, i);
        float4 uf[kHeadMaxV], jf[kHeadMaxV];
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < kHeadMaxV; ++v) {
            const int c = v * 128 + lane * 4;
            uf[v] = jf[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < E) {
                uf[v] = ld_row4(a.u_mf + u * E + c);
                jf[v] = ld_row4(a.i_mf + i * E + c);
            }
            dot += uf[v].x * jf[v].x * wv[v].x + uf[v].y * jf[v].y * wv[v].y + uf[v].z * jf[v].z * wv[v].z +
                   uf[v].w * jf[v].w * wv[v].w;
        }
        dot = warp_sum(dot);
        const float sc = sigmoidf_(dot + bo);
        if (!a.train) {
            if (lane == 0 && ok) a.scores[s] = sc;
            continue;
        }
        if (!ok) continue;
        const long long su = a.user_slot[u], si = a.item_slot[i];
        if (su < 0 || si < 0) continue;  // capacity overflow, flagged by the pre-pass
        float loss_k, dz;
        bce_terms(sc, a.ratings[s], a.inv_b, loss_k, dz);
        if (lane == 0) {
            loss_acc += loss_k;
            gb_acc += dz;
        }
#pragma unroll
        for (int v = 0; v < kHeadMaxV; ++v) {
            const int c = v * 128 + lane * 4;
            if (c < E) {
                const float4 p = make_float4(uf[v].x * jf[v].x, uf[v].y * jf[v].y, uf[v].z * jf[v].z, uf[v].w * jf[v].w);
                gw[v].x += dz * p.x; gw[v].y += dz * p.y; gw[v].z += dz * p.z; gw[v].w += dz * p.w;
                red_add4(a.g_u_mf + gs_off(E, a.user_cap, (unsigned)su, c), make_float4(dz * wv[v].x * jf[v].x, dz * wv[v].y * jf[v].y,
                                                           dz * wv[v].z * jf[v].z, dz * wv[v].w * jf[v].w));
                red_add4(a.g_i_mf + gs_off(E, a.item_cap, (unsigned)si, c), make_float4(dz * wv[v].x * uf[v].x, dz * wv[v].y * uf[v].y,
                                                           dz * wv[v].z * uf[v].z, dz * wv[v].w * uf[v].w));
            }
        }
    }
    if (!a.train) return;
#pragma unroll
    for (int v = 0; v < kHeadMaxV; ++v) *(float4*)&s_gw[warp][v * 128 + lane * 4] = gw[v];
    loss_acc = warp_sum(loss_acc);
    gb_acc = warp_sum(gb_acc);
    if (lane == 0) {
        s_sc[0][warp] = loss_acc;
        s_sc[1][warp] = gb_acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_gw[w][c];
        red_add1(a.g_w_out + c, t);
    }
    if (threadIdx.x == 0) {
        float l = 0.f, g = 0.f;
        for (int w = 0; w < kWarps; ++w) {
            l += s_sc[0][w];
            g += s_sc[1][w];
        }
        atomicAdd(&a.ws->loss_sum, (double)l);
        red_add1(a.g_b_out, g);
    }
}

int grid_warp_per_sample(long long n, const void* k) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);
    long long g = (long long)brs_sm_count() * (per_sm < 1 ? 1 : per_sm);
    long long need = (n + kWarps - 1) / kWarps;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}

int validate(const brs_ncf_model* m, bool train) {
    if (!m || !m->ws) return BRS_ERR_INVALID_ARG;
    if (m->kind < BRS_NCF_GMF || m->kind > BRS_NCF_NEUMF) return BRS_ERR_INVALID_ARG;
    if (m->emb_dim <= 0 || m->emb_dim % 4 != 0) return BRS_ERR_UNSUPPORTED;
    if (m->kind != BRS_NCF_GMF) {
        if (m->n_layers < 1 || m->n_layers > BRS_NCF_MAX_LAYERS) return BRS_ERR_UNSUPPORTED;
        if (m->mlp_dim != m->emb_dim * (1 << (m->n_layers - 1))) return BRS_ERR_INVALID_ARG;
        for (int l = 0; l <= m->n_layers; ++l)
            if (!m->act[l] || (train && !m->dact[l])) return BRS_ERR_INVALID_ARG;
        for (int l = 0; l < m->n_layers; ++l)
            if (!m->fc_weight[l].weight || !m->fc_bias[l].weight || (train && (!m->fc_weight[l].grad || !m->fc_bias[l].grad)))
                return BRS_ERR_INVALID_ARG;
    }
    const int head_w = (m->kind == BRS_NCF_GMF ? m->emb_dim : (m->kind == BRS_NCF_MLP ? m->emb_dim : 2 * m->emb_dim));
    if (head_w > kHeadMaxV * 128) return BRS_ERR_UNSUPPORTED;
    if (!m->out_weight.weight || !m->out_bias.weight) return BRS_ERR_INVALID_ARG;
    if (train && (!m->out_weight.grad || !m->out_bias.grad || !m->dz)) return BRS_ERR_INVALID_ARG;
    if (m->kind == BRS_NCF_NEUMF && !m->mfv) return BRS_ERR_INVALID_ARG;
    const int need_tables = m->kind == BRS_NCF_NEUMF ? 2 : 1;
    if (m->user.n_tables < need_tables || m->item.n_tables < need_tables) return BRS_ERR_INVALID_ARG;
    return BRS_OK;
}

void fill(const brs_ncf_model* m, NcfArgs& a) {
    memset(&a, 0, sizeof(a));
    a.kind = m->kind;
    a.n_users = m->user.table[0].n_rows;
    a.n_items = m->item.table[0].n_rows;
    a.user_slot = m->user.rows.slot_map;
    a.item_slot = m->item.rows.slot_map;
    a.user_cap = m->user.rows.capacity;
    a.item_cap = m->item.rows.capacity;
    a.ws = (brs_step_ws*)m->ws;
    a.w_out = m->out_weight.weight;
    a.b_out = m->out_bias.weight;
    a.g_w_out = m->out_weight.grad;
    a.g_b_out = m->out_bias.grad;
    a.dz = m->dz;
    if (m->kind == BRS_NCF_GMF) {
        a.mf_dim = m->emb_dim;
        a.u_mf = m->user.table[0].weight;
        a.i_mf = m->item.table[0].weight;
        a.g_u_mf = m->user.table[0].grad;
        a.g_i_mf = m->item.table[0].grad;
    } else {
        a.mlp_dim = m->mlp_dim;
        a.head_in = m->emb_dim;
        a.u_mlp = m->user.table[0].weight;
        a.i_mlp = m->item.table[0].weight;
        a.g_u_mlp = m->user.table[0].grad;
        a.g_i_mlp = m->item.table[0].grad;
        a.x0 = m->act[0];
        a.dx0 = m->dact[0];
        a.h_last = m->act[m->n_layers];
        a.dh_last = m->dact[m->n_layers];
        if (m->kind == BRS_NCF_NEUMF) {
            a.mf_dim = m->emb_dim;
            a.u_mf = m->user.table[1].weight;
            a.i_mf = m->item.table[1].weight;
            a.g_u_mf = m->user.table[1].grad;
            a.g_i_mf = m->item.table[1].grad;
            a.mfv = m->mfv;
        }
    }
}

int layer_in(const brs_ncf_model* m, int l) { return (2 * m->mlp_dim) >> l; }

int forward(const brs_ncf_model* m, const NcfArgs& a, cudaStream_t st) {
    if (m->kind == BRS_NCF_GMF) {
        gmf_fused_kernel<<<grid_warp_per_sample(a.batch, (const void*)gmf_fused_kernel), kThreads, 0, st>>>(a);
        BRS_CUDA_CHECK(cudaGetLastError());
        return BRS_OK;
    }
    ncf_gather_kernel<<<grid_warp_per_sample(a.batch, (const void*)ncf_gather_kernel), kThreads, 0, st>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    for (int l = 0; l < m->n_layers; ++l) {
        const int in = layer_in(m, l), out = in / 2;
        int rc;
        if (g_gemm_backend == 1 && brs_linear_tc_supported((int)a.batch, out, in))
            rc = brs_linear_tc(m->act[l], m->fc_weight[l].weight, m->fc_bias[l].weight, m->act[l + 1], nullptr,
                               (int)a.batch, out, in, /*relu=*/true, st);
        else
            rc = brs_linear_fwd_simt(m->act[l], in, m->fc_weight[l].weight, m->fc_bias[l].weight, m->act[l + 1], out,
                                     (int)a.batch, out, in, /*relu=*/true, st);
        if (rc != BRS_OK) return rc;
    }
    ncf_head_kernel<<<grid_warp_per_sample(a.batch, (const void*)ncf_head_kernel), kThreads, 0, st>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

}  // namespace

extern "C" int brs_ncf_fwd_bwd(const brs_ncf_model* model, const int64_t* users, const int64_t* items,
                               const float* ratings, int64_t batch, void* stream) {
    int rc = validate(model, true);
    if (rc != BRS_OK) return rc;
    if (!users || !items || !ratings || batch < 0 || batch > model->max_batch) return BRS_ERR_INVALID_ARG;
    if (batch == 0) return BRS_OK;
    NcfArgs a;
    fill(model, a);
    a.users = (const long long*)users;
    a.items = (const long long*)items;
    a.ratings = ratings;
    a.batch = batch;
    a.inv_b = 1.0f / (float)batch;
    a.train = 1;
    cudaStream_t st = (cudaStream_t)stream;
    {   // pre-pass: range-check the ids and give every touched row a slot in the compact scratch
        const brs_rowset rs[2] = {model->user.rows, model->item.rows};
        const long long* idx[2] = {a.users, a.items};
        const long long n[2] = {batch, batch};
        rc = brs_assign_slots(rs, idx, n, 2, a.ws, st);
        if (rc != BRS_OK) return rc;
    }
    rc = forward(model, a, st);
    if (rc != BRS_OK || model->kind == BRS_NCF_GMF) return rc;
    for (int l = model->n_layers - 1; l >= 0; --l) {
        const int in = layer_in(model, l), out = in / 2;
        rc = brs_linear_wgrad_simt(model->dact[l + 1], out, model->act[l], in, model->fc_weight[l].grad,
                                   model->fc_bias[l].grad, (int)batch, out, in, st);
        if (rc != BRS_OK) return rc;
        // mask by (layer input > 0): a ReLU precedes every Linear except MLP's first (mlp.py:47-48)
        const float* mask = (l == 0 && model->kind == BRS_NCF_MLP) ? nullptr : model->act[l];
        if (g_gemm_backend == 1 && model->fc_weight_t[l] && brs_linear_tc_supported((int)batch, in, out)) {
            // dX[M,in] = dY[M,out] . W[out,in]  ==  dY . (W^T)^T with W^T [in,out] as the K-major B operand
            rc = brs_transpose(model->fc_weight[l].weight, model->fc_weight_t[l], out, in, st);
            if (rc != BRS_OK) return rc;
            rc = brs_linear_tc(model->dact[l + 1], model->fc_weight_t[l], nullptr, model->dact[l], mask, (int)batch, in,
                               out, /*relu=*/false, st);
        } else {
            rc = brs_linear_dgrad_simt(model->dact[l + 1], out, model->fc_weight[l].weight, model->dact[l], in, mask, in,
                                       (int)batch, out, in, st);
        }
        if (rc != BRS_OK) return rc;
    }
    ncf_scatter_kernel<<<grid_warp_per_sample(batch, (const void*)ncf_scatter_kernel), kThreads, 0, st>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_ncf_apply(const brs_ncf_model* model, const brs_opt* opt, int64_t batch, float* out, void* stream) {
    int rc = validate(model, true);
    if (rc != BRS_OK) return rc;
    if (!opt || batch < 0) return BRS_ERR_INVALID_ARG;
    brs_entity ents[2] = {model->user, model->item};
    brs_dense_param dense[2 * BRS_NCF_MAX_LAYERS + 2];
    int nd = 0;
    if (model->kind != BRS_NCF_GMF)
        for (int l = 0; l < model->n_layers; ++l) {
            dense[nd++] = model->fc_weight[l];
            dense[nd++] = model->fc_bias[l];
        }
    dense[nd++] = model->out_weight;
    dense[nd++] = model->out_bias;
    return brs_apply_impl(ents, 2, dense, nd, 0, opt, model->ws, 0, out, batch, 2 * (long long)batch, stream);
}

extern "C" int brs_ncf_predict(const brs_ncf_model* model, const int64_t* users, const int64_t* items, int64_t n,
                               float* scores, void* stream) {
    int rc = validate(model, false);
    if (rc != BRS_OK) return rc;
    if (!users || !items || !scores || n < 0) return BRS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t off = 0; off < n; off += model->max_batch) {
        NcfArgs a;
        fill(model, a);
        a.users = (const long long*)users + off;
        a.items = (const long long*)items + off;
        a.batch = (n - off < model->max_batch) ? (n - off) : model->max_batch;
        a.scores = scores + off;
        a.train = 0;
        rc = forward(model, a, st);
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

extern "C" int brs_ncf_train_batches(const brs_ncf_model* model, const brs_opt* opt, const int64_t* users,
                                     const int64_t* items, const float* ratings, int64_t n, int64_t batch, float* out,
                                     void* stream) {
    if (!users || !items || !ratings || !out || n < 0 || batch <= 0) return BRS_ERR_INVALID_ARG;
    int64_t b = 0;
    for (int64_t off = 0; off < n; off += batch, ++b) {
        const int64_t cur = (n - off < batch) ? (n - off) : batch;
        int rc = brs_ncf_fwd_bwd(model, users + off, items + off, ratings + off, cur, stream);
        if (rc != BRS_OK) return rc;
        rc = brs_ncf_apply(model, opt, cur, out + 4 * b, stream);
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

// the Linear building blocks, exposed for tests / other callers (brs_mlp_{fwd,bwd})
extern "C" int brs_set_gemm_backend(int backend) {
    if (backend != 0 && backend != 1) return BRS_ERR_INVALID_ARG;
    g_gemm_backend = backend;
    return BRS_OK;
}

// tensor-core Linear building block: y = epilogue(x[m,k] . w[n,k]^T) with 3xTF32 error compensation
extern "C" int brs_mlp_fwd_tc(const float* x, const float* w, const float* b, float* y, const float* keep_where_positive,
                              int64_t m, int32_t n, int32_t k, int32_t relu, void* stream) {
    if (!x || !w || !y || m < 0 || n <= 0 || k <= 0) return BRS_ERR_INVALID_ARG;
    if (m == 0) return BRS_OK;
    return brs_linear_tc(x, w, b, y, keep_where_positive, (int)m, n, k, relu != 0, (cudaStream_t)stream);
}

extern "C" int brs_mlp_fwd(const float* x, const float* w, const float* b, float* y, int64_t m, int32_t n, int32_t k,
                           int32_t relu, void* stream) {
    if (!x || !w || !b || !y || m < 0 || n <= 0 || k <= 0) return BRS_ERR_INVALID_ARG;
    return brs_linear_fwd_simt(x, k, w, b, y, n, (int)m, n, k, relu != 0, (cudaStream_t)stream);
}

extern "C" int brs_mlp_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db,
                           const float* relu_mask_src, int64_t m, int32_t n, int32_t k, void* stream) {
    if (!dy || !x || !w || !dw || m < 0 || n <= 0 || k <= 0) return BRS_ERR_INVALID_ARG;
    int rc = brs_linear_wgrad_simt(dy, n, x, k, dw, db, (int)m, n, k, (cudaStream_t)stream);
    if (rc != BRS_OK || !dx) return rc;
    return brs_linear_dgrad_simt(dy, n, w, dx, k, relu_mask_src, k, (int)m, n, k, (cudaStream_t)stream);
}
