// MF (matrix factorisation) fused forward + backward for one batch -- sm_100a.
//
// Replaces, per batch, the reference's ~45 ATen launches
//   MF.forward x2            beta_rec/models/mf.py:32-55
//   bpr_loss / bce_loss      beta_rec/models/torch_engine.py:92-121
//   loss.backward()          beta_rec/models/mf.py:116-117  (embedding_dense_backward)
// with a slot-assignment pre-pass (rows_apply.cu: one thread per index) and ONE
// fused kernel: the tile's index lists staged into shared memory by 1-D TMA bulk
// copies (mbarrier-tracked), 128-bit coalesced gathers of the three embedding
// rows, group-shuffle dot products, the loss and its closed-form gradient, and
// 128-bit red.global.add scatter of the three gradient rows into the COMPACT,
// L2-resident gradient scratch.  The optimizer update is a separate launch
// (rows_apply.cu) so that every sample reads PRE-step weights (batch-synchronous
// semantics of autograd + torch.optim).
//
// Shape of the launch (from the round-1 ncu captures: v1 was latency-bound, v2
// issue-bound at ~550 instructions per sample because all 32 lanes of a row group
// repeated the scalar loss math): a row of D floats is covered by LPR lanes x VPL
// float4 with VPL = 4 wherever D allows (D = 128 -> 8 lanes x 4), so every warp
// instruction -- address math, shuffles, the sigmoid/log chain -- serves 32/LPR
// samples at once, and each lane still issues 128-bit loads whose LPR-lane groups
// cover whole 128-byte lines.  Persistent grid (resident blocks only), 32-sample
// tiles double-buffered through TMA, warp-uniform loop control, nothing on the
// per-sample path waits for an atomic's return value.
#include <stdlib.h>

#include "common.cuh"
#include "mf_math.cuh"

int brs_assign_slots(const brs_rowset* rs, const long long* const* idx, const long long* n, int n_arrays,
                     brs_step_ws* ws, cudaStream_t st);

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 32;  // samples per staged index tile (double-buffered)

enum { SHARD_NONE = 0, SHARD_DIRECT = 1, SHARD_STAGED = 2 };

struct MfPeerTables {  // one rank's shard, as seen from this process (device-resident array of these)
    const float* user_emb;
    const float* item_emb;
    const float* user_bias;
    const float* item_bias;
    float* g_user_emb;
    float* g_item_emb;
    float* g_user_bias;
    float* g_item_bias;
    unsigned int* user_bits;
    unsigned int* item_bits;
};
static_assert(sizeof(MfPeerTables) == sizeof(brs_mf_peer_tables), "peer table layout is part of the ABI");

template <class T>
__device__ __forceinline__ T* ldg_ptr(T* const* p) {
    return (T*)__ldg((const unsigned long long*)p);
}
struct MfArgs {
    const float* __restrict__ user_emb;
    const float* __restrict__ item_emb;
    const float* __restrict__ user_bias;
    const float* __restrict__ item_bias;
    const float* __restrict__ global_bias;
    float* g_user_emb;   // compact scratch [user capacity, D]
    float* g_item_emb;   // compact scratch [item capacity, D]
    float* g_user_bias;  // compact scratch [user capacity]
    float* g_item_bias;
    const int* __restrict__ user_slot;  // slot_map of the user rowset (filled by the pre-pass)
    const int* __restrict__ item_slot;
    // row-sharded multi-GPU mode (SHARD = true): row r lives on rank r & shard_mask at local row
    // r >> shard_shift; peers[rank] holds that rank's tables as pointers mapped into THIS process
    // (NVLink peer memory).  The fused kernel itself only sees LOCAL memory: user_emb/item_emb/...
    // point at this rank's staging tables (one row per slot), filled by mf_pull_kernel
    const MfPeerTables* __restrict__ peers;
    int shard_shift, shard_mask;
    int user_cap, item_cap;  // capacity of the gradient scratch (sector-blocked layout, common.cuh gs_off)
    int pol_gather, pol_scratch;  // L2 eviction policies (BRS_L2_*)
    brs_step_ws* ws;
    const long long* users;
    const long long* items;  // pos items (bpr) / items (bce)
    const void* third;       // neg items int64 (bpr) / ratings float (bce)
    long long batch;
    long long n_users, n_items;
    int dim;
    float reg_w;   // engine.reg (0.0 in the reference: mf.py:81-83)
    float inv_b;   // 1 / batch
};

struct __align__(16) IdxTile {
    long long a[kTile];
    long long b[kTile];
    long long c[kTile];  // neg ids, or ratings in the first kTile*4 bytes
};

// everything one lane holds for one in-flight sample
template <int VPL, int LOSS>
struct Sample {
    long long u, i, j;
    float rating;
    bool valid;
    float4 ue[VPL], ie[VPL], je[VPL];
    float bu, bi, bj;
    int su, si, sj;  // gradient-scratch slots (SHARD: also the rows of the staging tables)
};

template <int LPR, int VPL, bool FULL, int LOSS, int SHARD>
__device__ __forceinline__ void sample_load(const MfArgs& a, const IdxTile& T, int s, int tile_n, int gl, int D,
                                            unsigned long long pol_g, Sample<VPL, LOSS>& x) {
    x.valid = s < tile_n;
    const int sc = x.valid ? s : 0;
    x.u = T.a[sc];
    x.i = T.b[sc];
    x.j = 0;
    x.rating = 0.f;
    if (LOSS == LOSS_BPR)
        x.j = T.c[sc];
    else
        x.rating = ((const float*)T.c)[sc];
    if ((unsigned long long)x.u >= (unsigned long long)a.n_users ||
        (unsigned long long)x.i >= (unsigned long long)a.n_items ||
        (unsigned long long)x.j >= (unsigned long long)a.n_items) {
        x.valid = false;  // flagged by the pre-pass (the reference raises IndexError)
        x.u = x.i = x.j = 0;
    }
    // local row ids and the tables they live in (this rank's own, or a peer's under SHARD)
    unsigned lu = (unsigned)x.u, li = (unsigned)x.i, lj = (unsigned)x.j;
    const float *t_ue = a.user_emb, *t_ie = a.item_emb, *t_je = a.item_emb;
    const float *t_ub = a.user_bias, *t_ib = a.item_bias, *t_jb = a.item_bias;
    const int *t_us = a.user_slot, *t_is = a.item_slot, *t_js = a.item_slot;
    if (SHARD == SHARD_DIRECT) {
        // rows are gathered from their owners' shards with peer loads (NVLink), once per sample
        const int ou = lu & a.shard_mask, oi = li & a.shard_mask, oj = lj & a.shard_mask;
        lu >>= a.shard_shift;
        li >>= a.shard_shift;
        lj >>= a.shard_shift;
        const MfPeerTables *pu = a.peers + ou, *pi = a.peers + oi, *pj = a.peers + oj;
        t_ue = ldg_ptr(&pu->user_emb);
        t_ie = ldg_ptr(&pi->item_emb);
        t_je = ldg_ptr(&pj->item_emb);
        t_ub = ldg_ptr(&pu->user_bias);
        t_ib = ldg_ptr(&pi->item_bias);
        t_jb = ldg_ptr(&pj->item_bias);
    }
    if (SHARD == SHARD_STAGED) {
        // the batch's unique rows were pulled from their owners into THIS rank's compact staging tables
        // (mf_pull_kernel), one row per slot -- the same slots the gradient scratch uses
        x.su = __ldg(t_us + lu);
        x.si = __ldg(t_is + li);
        x.sj = (LOSS == LOSS_BPR) ? __ldg(t_js + lj) : 0;
        if (x.su < 0 || x.si < 0 || x.sj < 0) {  // capacity overflow, flagged by the pre-pass
            x.valid = false;
            x.su = x.si = x.sj = 0;
        }
        lu = (unsigned)x.su;
        li = (unsigned)x.si;
        lj = (unsigned)x.sj;
    }
    // rows < 2^31 (slot maps are int32), so one 32x32->64 IMAD.WIDE per row address
    const float* ur = t_ue + (unsigned long long)lu * (unsigned)D;
    const float* ir = t_ie + (unsigned long long)li * (unsigned)D;
    const float* jr = t_je + (unsigned long long)lj * (unsigned)D;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int col = (v * LPR + gl) * 4;
        const bool on = FULL || col < D;
        x.ue[v] = on ? ld_row4_pol(ur + col, pol_g) : f4_zero();
        x.ie[v] = on ? ld_row4_pol(ir + col, pol_g) : f4_zero();
        if (LOSS == LOSS_BPR) x.je[v] = on ? ld_row4_pol(jr + col, pol_g) : f4_zero();
    }
    x.bu = __ldg(t_ub + lu);
    x.bi = __ldg(t_ib + li);
    x.bj = (LOSS == LOSS_BPR) ? __ldg(t_jb + lj) : 0.f;
    if (SHARD != SHARD_STAGED) {
        // slots of the compact gradient scratch: indexed by the row id, assigned by the pre-pass
        x.su = __ldg(t_us + (unsigned)x.u);
        x.si = __ldg(t_is + (unsigned)x.i);
        x.sj = (LOSS == LOSS_BPR) ? __ldg(t_js + (unsigned)x.j) : 0;
        if (x.su < 0 || x.si < 0 || x.sj < 0) x.valid = false;  // capacity overflow, flagged by the pre-pass
    }
}

template <int LPR, int VPL, bool FULL, int LOSS, int SHARD>
__device__ __forceinline__ void sample_finish(const MfArgs& a, const Sample<VPL, LOSS>& x, int gl, int D, float bg,
                                              unsigned long long pol_s, float& loss_acc, float& reg_acc,
                                              float& gb_acc) {
    float dp = 0.f, dn = 0.f, uu = 0.f, ii = 0.f, jj = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        dp += f4_dot(x.ue[v], x.ie[v]);
        uu += f4_dot(x.ue[v], x.ue[v]);
        ii += f4_dot(x.ie[v], x.ie[v]);
        if (LOSS == LOSS_BPR) {
            dn += f4_dot(x.ue[v], x.je[v]);
            jj += f4_dot(x.je[v], x.je[v]);
        }
    }
    dp = group_sum<LPR>(dp);
    if (LOSS == LOSS_BPR) dn = group_sum<LPR>(dn);

    float cu_i, cu_j, loss_k;  // d loss / d z for the (u,i) and (u,j) scores
    mf_sample_coef<LOSS>(dp + x.bu + x.bi + bg, dn + x.bu + x.bj + bg, x.rating, a.inv_b, cu_i, cu_j, loss_k);
    if (!x.valid) return;

    // regularizer numerator (mf.py:49-54), one forward call per score
    const float fwd_calls = (LOSS == LOSS_BPR) ? 2.f : 1.f;
    reg_acc += fwd_calls * uu + ii + jj;
    if (gl == 0) {
        reg_acc += fwd_calls * x.bu * x.bu + x.bi * x.bi + x.bj * x.bj;
        loss_acc += loss_k;
        gb_acc += cu_i + cu_j;
    }
    const float rw = 2.0f * a.reg_w * a.inv_b;  // d(reg_w*regularizer)/d row = rw * row per forward call
    // gradients always go to THIS rank's compact scratch (also under SHARD: they are pushed to the owners
    // row by row afterwards, mf_push_kernel)
    float *t_gu = a.g_user_emb, *t_gi = a.g_item_emb, *t_gj = a.g_item_emb;
    float *t_gub = a.g_user_bias, *t_gib = a.g_item_bias, *t_gjb = a.g_item_bias;
    float* gu = t_gu;
    float* gi = t_gi;
    float* gj = t_gj;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int col = (v * LPR + gl) * 4;
        if (FULL || col < D) {
            float4 du = f4_scale(cu_i, x.ie[v]);
            if (LOSS == LOSS_BPR) du = f4_fma(cu_j, x.je[v], du);
            if (a.reg_w != 0.f) du = f4_fma(fwd_calls * rw, x.ue[v], du);
            red_add4_pol(gu + gs_off(D, a.user_cap, (unsigned)x.su, col), du, pol_s);
            float4 di = f4_scale(cu_i, x.ue[v]);
            if (a.reg_w != 0.f) di = f4_fma(rw, x.ie[v], di);
            red_add4_pol(gi + gs_off(D, a.item_cap, (unsigned)x.si, col), di, pol_s);
            if (LOSS == LOSS_BPR) {
                float4 dj = f4_scale(cu_j, x.ue[v]);
                if (a.reg_w != 0.f) dj = f4_fma(rw, x.je[v], dj);
                red_add4_pol(gj + gs_off(D, a.item_cap, (unsigned)x.sj, col), dj, pol_s);
            }
        }
    }
    // bias gradients, spread over the first lanes of the group
    if (gl == 0) red_add1(t_gub + x.su, cu_i + cu_j + fwd_calls * rw * x.bu);
    if (gl == 1 % LPR) red_add1(t_gib + x.si, cu_i + rw * x.bi);
    if (LOSS == LOSS_BPR && gl == 2 % LPR) red_add1(t_gjb + x.sj, cu_j + rw * x.bj);
}

// Stage tile `t` of the batch's index lists into `dst`.  TMA path when the slice is a
// full, 16-byte aligned tile; otherwise (ragged tail / odd views) the block copies it
// with plain loads.  Returns whether the TMA path was taken (block-uniform).
template <int LOSS>
__device__ __forceinline__ bool stage_tile(const MfArgs& a, long long t, IdxTile* dst, uint64_t* bar) {
    const long long base = t * kTile;
    const long long n = min((long long)kTile, a.batch - base);
    const long long* pa = a.users + base;
    const long long* pb = a.items + base;
    const char* pc = (const char*)a.third + base * (LOSS == LOSS_BPR ? 8 : 4);
    constexpr unsigned bytes_c = (LOSS == LOSS_BPR ? 8u : 4u) * kTile;
    const bool used_tma = (n == kTile) && ((((uintptr_t)pa | (uintptr_t)pb | (uintptr_t)pc) & 15) == 0);
    if (used_tma) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, 2u * 8u * kTile + bytes_c);
            tma_load_1d(dst->a, pa, 8u * kTile, bar);
            tma_load_1d(dst->b, pb, 8u * kTile, bar);
            tma_load_1d(dst->c, pc, bytes_c, bar);
        }
    } else {
        for (int k = threadIdx.x; k < n; k += kThreads) {
            dst->a[k] = pa[k];
            dst->b[k] = pb[k];
            if (LOSS == LOSS_BPR)
                dst->c[k] = ((const long long*)pc)[k];
            else
                ((float*)dst->c)[k] = ((const float*)pc)[k];
        }
    }
    return used_tma;
}

template <int LPR, int VPL, bool FULL, int LOSS, int SHARD, int UNROLL, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) mf_fwd_bwd_kernel(const MfArgs a) {
    constexpr int SPW = 32 / LPR;
    __shared__ IdxTile s_tile[2];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float s_red[3][kWarps];

    const int lane = threadIdx.x & 31;
    // broadcast from lane 0 so the compiler knows the value (and every loop bound built
    // from it) is warp-uniform: shuffles below need no divergence guards
    const int warp = __shfl_sync(BRS_FULL_MASK, threadIdx.x >> 5, 0);
    const int gl = lane % LPR;   // lane within the row group
    const int grp = lane / LPR;  // which of the warp's SPW samples
    const int D = a.dim;
    const long long n_tiles = (a.batch + kTile - 1) / kTile;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const float bg = __ldg(a.global_bias);
    const unsigned long long pol_g = l2_policy(a.pol_gather), pol_s = l2_policy(a.pol_scratch);
    float loss_acc = 0.f, reg_acc = 0.f, gb_acc = 0.f;
    unsigned phase_bits = 0u;  // bit b = parity to wait for on s_bar[b]
    unsigned tma_bits = 0u;    // bit b = s_tile[b] is being filled by TMA

    long long t = blockIdx.x;
    int buf = 0;
    if (t < n_tiles && stage_tile<LOSS>(a, t, &s_tile[0], &s_bar[0])) tma_bits |= 1u;

    for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
        const long long tn = t + gridDim.x;
        // prefetch the next tile into the other buffer (its previous readers passed the
        // __syncthreads at the end of the previous iteration)
        if (tn < n_tiles) {
            const bool nt = stage_tile<LOSS>(a, tn, &s_tile[buf ^ 1], &s_bar[buf ^ 1]);
            tma_bits = (tma_bits & ~(1u << (buf ^ 1))) | ((nt ? 1u : 0u) << (buf ^ 1));
        }
        if ((tma_bits >> buf) & 1u) {
            mbar_wait(&s_bar[buf], (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
        } else {
            __syncthreads();  // plain-copy path: make the block's stores visible
        }
        const IdxTile& T = s_tile[buf];
        const int tile_n = (int)min((long long)kTile, a.batch - t * kTile);

        // UNROLL passes (UNROLL * SPW samples) in flight per warp: all their loads are issued
        // before the first one is consumed, and their scalar chains interleave
        for (int b0 = warp * SPW; b0 < tile_n; b0 += UNROLL * kWarps * SPW) {
            Sample<VPL, LOSS> x[UNROLL];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q)
                sample_load<LPR, VPL, FULL, LOSS, SHARD>(a, T, b0 + q * kWarps * SPW + grp, tile_n, gl, D, pol_g, x[q]);
#pragma unroll
            for (int q = 0; q < UNROLL; ++q)
                sample_finish<LPR, VPL, FULL, LOSS, SHARD>(a, x[q], gl, D, bg, pol_s, loss_acc, reg_acc, gb_acc);
        }
        __syncthreads();  // everyone is done with s_tile[buf] before it is refilled
    }

    // block reduction of the scalar outputs -> 3 atomics per block
    loss_acc = warp_sum(loss_acc);
    reg_acc = warp_sum(reg_acc);
    gb_acc = warp_sum(gb_acc);
    if (lane == 0) {
        s_red[0][warp] = loss_acc;
        s_red[1][warp] = reg_acc;
        s_red[2][warp] = gb_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f, r = 0.f, g = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            l += s_red[0][w];
            r += s_red[1][w];
            g += s_red[2][w];
        }
        atomicAdd(&a.ws->loss_sum, (double)l);
        atomicAdd(&a.ws->reg_sum, (double)r);
        atomicAdd(&a.ws->g_global_bias, g);
    }
}

// scores[k] = sigmoid(u.i + b_u + b_i + b_g): MF.predict (mf.py:57-70)
template <int LPR, int VPL, bool FULL>
__global__ void __launch_bounds__(kThreads) mf_predict_kernel(const float* __restrict__ user_emb,
                                                              const float* __restrict__ item_emb,
                                                              const float* __restrict__ user_bias,
                                                              const float* __restrict__ item_bias,
                                                              const float* __restrict__ global_bias, int D,
                                                              long long n_users, long long n_items,
                                                              const long long* __restrict__ users,
                                                              const long long* __restrict__ items, long long n,
                                                              float* __restrict__ scores, unsigned int* err) {
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR, grp = lane / LPR;
    const long long warp_global = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * kWarps;
    const float bg = __ldg(global_bias);
    for (long long base = warp_global * SPW; base < n; base += n_warps * SPW) {
        const long long s = base + grp;
        bool valid = s < n;
        long long u = valid ? users[s] : 0, i = valid ? items[s] : 0;
        if ((unsigned long long)u >= (unsigned long long)n_users || (unsigned long long)i >= (unsigned long long)n_items) {
            // the reference raises IndexError inside nn.Embedding: flag it in the predict-only error word and
            // publish NaN for the sample (nothing is left uninitialised)
            if (valid && gl == 0) {
                atomicOr(err, 1u);
                scores[s] = __int_as_float(0x7fc00000);
            }
            valid = false;
            u = i = 0;
        }
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
            if (FULL || col < D) d += f4_dot(ld_row4(user_emb + u * D + col), ld_row4(item_emb + i * D + col));
        }
        d = group_sum<LPR>(d);
        if (valid && gl == 0) scores[s] = sigmoidf_(d + __ldg(user_bias + u) + __ldg(item_bias + i) + bg);
    }
}

int grid_for(const void* kernel, long long work_blocks) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)brs_sm_count() * per_sm;  // persistent: one wave of resident blocks
    if (g > work_blocks) g = work_blocks;
    return (int)(g < 1 ? 1 : g);
}

// how the row-sharded step reads remote rows: SHARD_DIRECT gathers them per sample with peer loads inside
// the fused kernel; SHARD_STAGED pulls each unique row once into local staging tables first.
// BRS_SHARD_MODE=1|2 or brs_debug_set_shard_mode
// Measured (profiles/r01_multigpu.md): direct wins while at most half of the rows are remote
// (N=2: 773 vs 640 M/s), staged wins once 7/8 of them are (N=8: 2.01 vs 1.96 G/s).
int g_shard_mode = 0;  // 0: by world size
int shard_mode(int world) {
    if (g_shard_mode == 0) {
        const char* e = getenv("BRS_SHARD_MODE");
        const int m = e ? atoi(e) : 0;
        g_shard_mode = (m == SHARD_STAGED || m == SHARD_DIRECT) ? m : -1;
    }
    if (g_shard_mode > 0) return g_shard_mode;
    return world >= 8 ? SHARD_STAGED : SHARD_DIRECT;
}

template <int LOSS, int SHARD = SHARD_NONE>
int launch_fwd_bwd(const MfArgs& a, cudaStream_t st) {
    const int D = a.dim;
    const long long n_tiles = (a.batch + kTile - 1) / kTile;
#define BRS_LAUNCHX(LPR, VPL, FULL, UNROLL, MINB)                                       \
    do {                                                                                \
        auto k = mf_fwd_bwd_kernel<LPR, VPL, FULL, LOSS, SHARD, UNROLL, MINB>;          \
        k<<<grid_for((const void*)k, n_tiles), kThreads, 0, st>>>(a);                   \
    } while (0)
#define BRS_LAUNCH(LPR, VPL, FULL) BRS_LAUNCHX(LPR, VPL, FULL, 1, ((VPL) <= 2 ? 8 : (SHARD ? 4 : 5)))
    if (D % 4 != 0 || D <= 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    switch (D) {  // VPL = 4 float4 per lane wherever D allows: a warp instruction serves 32/LPR samples
        case 4: BRS_LAUNCH(1, 1, true); break;
        case 8: BRS_LAUNCH(1, 2, true); break;
        case 16: BRS_LAUNCH(1, 4, true); break;
        case 32: BRS_LAUNCH(2, 4, true); break;
        case 64: BRS_LAUNCH(4, 4, true); break;
        case 128: BRS_LAUNCH(8, 4, true); break;
        case 256: BRS_LAUNCH(16, 4, true); break;
        case 384: BRS_LAUNCH(32, 3, true); break;
        case 512: BRS_LAUNCH(32, 4, true); break;
        default:  // any other multiple of 4: next power-of-two lane group, tail lanes idle
            if (D < 8) BRS_LAUNCH(2, 1, false);
            else if (D < 16) BRS_LAUNCH(4, 1, false);
            else if (D < 32) BRS_LAUNCH(8, 1, false);
            else if (D < 64) BRS_LAUNCH(16, 1, false);
            else if (D < 128) BRS_LAUNCH(32, 1, false);
            else if (D < 256) BRS_LAUNCH(32, 2, false);
            else if (D < 384) BRS_LAUNCH(32, 3, false);
            else BRS_LAUNCH(32, 4, false);
    }
#undef BRS_LAUNCH
#undef BRS_LAUNCHX
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

int check_model(const brs_mf_model* m) {
    if (!m || !m->ws) return BRS_ERR_INVALID_ARG;
    const brs_table& ue = m->user.table[0];
    const brs_table& ie = m->item.table[0];
    if (m->user.n_tables < 2 || m->item.n_tables < 2) return BRS_ERR_INVALID_ARG;
    if (!ue.weight || !ie.weight || !m->user.table[1].weight || !m->item.table[1].weight || !m->global_bias.weight)
        return BRS_ERR_INVALID_ARG;
    if (ue.dim != ie.dim || m->user.table[1].dim != 1 || m->item.table[1].dim != 1) return BRS_ERR_INVALID_ARG;
    if ((((uintptr_t)ue.weight | (uintptr_t)ie.weight) & 15) != 0) return BRS_ERR_INVALID_ARG;
    return BRS_OK;
}

int fill_args(const brs_mf_model* m, MfArgs& a, bool need_grad) {
    int rc = check_model(m);
    if (rc != BRS_OK) return rc;
    a.user_emb = m->user.table[0].weight;
    a.item_emb = m->item.table[0].weight;
    a.user_bias = m->user.table[1].weight;
    a.item_bias = m->item.table[1].weight;
    a.global_bias = m->global_bias.weight;
    a.g_user_emb = m->user.table[0].grad;
    a.g_item_emb = m->item.table[0].grad;
    a.g_user_bias = m->user.table[1].grad;
    a.g_item_bias = m->item.table[1].grad;
    if (need_grad) {
        if (!a.g_user_emb || !a.g_item_emb || !a.g_user_bias || !a.g_item_bias) return BRS_ERR_INVALID_ARG;
        if ((((uintptr_t)a.g_user_emb | (uintptr_t)a.g_item_emb) & 15) != 0) return BRS_ERR_INVALID_ARG;
        if (!m->user.rows.slot_map || !m->user.rows.list || !m->user.rows.count || !m->item.rows.slot_map ||
            !m->item.rows.list || !m->item.rows.count)
            return BRS_ERR_INVALID_ARG;
    }
    a.user_slot = m->user.rows.slot_map;
    a.item_slot = m->item.rows.slot_map;
    a.user_cap = m->user.rows.capacity;
    a.item_cap = m->item.rows.capacity;
    a.peers = nullptr;
    a.shard_shift = a.shard_mask = 0;
    a.pol_gather = brs_l2_cfg().gather;
    a.pol_scratch = brs_l2_cfg().scratch;
    a.ws = (brs_step_ws*)m->ws;
    a.n_users = m->user.table[0].n_rows;
    a.n_items = m->item.table[0].n_rows;
    a.dim = m->user.table[0].dim;
    return BRS_OK;
}

}  // namespace

// phases: 1 = slot pre-pass, 2 = fused kernel, 3 = both
int brs_mf_fwd_bwd_phases(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                          const void* third, int64_t batch, float reg_weight, void* stream, int phases);

int brs_mf_fwd_bwd_impl(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                        const void* third, int64_t batch, float reg_weight, void* stream) {
    return brs_mf_fwd_bwd_phases(model, loss_kind, users, items, third, batch, reg_weight, stream, 3);
}

int brs_mf_fwd_bwd_phases(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                          const void* third, int64_t batch, float reg_weight, void* stream, int phases) {
    if (!users || !items || !third || batch < 0) return BRS_ERR_INVALID_ARG;
    MfArgs a;
    int rc = fill_args(model, a, true);
    if (rc != BRS_OK) return rc;
    if (batch == 0) return BRS_OK;
    a.users = (const long long*)users;
    a.items = (const long long*)items;
    a.third = third;
    a.batch = batch;
    a.reg_w = reg_weight;
    a.inv_b = 1.0f / (float)batch;
    cudaStream_t st = (cudaStream_t)stream;
    if (phases & 1) {  // pre-pass: range-check the indices and give every touched row a slot in the compact scratch
        const brs_rowset rs[3] = {model->user.rows, model->item.rows, model->item.rows};
        const long long* idx[3] = {(const long long*)users, (const long long*)items, (const long long*)third};
        const long long n[3] = {batch, batch, batch};
        rc = brs_assign_slots(rs, idx, n, loss_kind == LOSS_BPR ? 3 : 2, a.ws, st);
        if (rc != BRS_OK) return rc;
    }
    if (!(phases & 2)) return BRS_OK;
    if (loss_kind == LOSS_BPR) return launch_fwd_bwd<LOSS_BPR>(a, st);
    if (loss_kind == LOSS_BCE) return launch_fwd_bwd<LOSS_BCE>(a, st);
    return BRS_ERR_INVALID_ARG;
}

extern "C" int brs_mf_bpr_fwd_bwd(const brs_mf_model* model, const int64_t* users, const int64_t* pos_items,
                                  const int64_t* neg_items, int64_t batch, float reg_weight, void* stream) {
    return brs_mf_fwd_bwd_impl(model, LOSS_BPR, users, pos_items, neg_items, batch, reg_weight, stream);
}

// the two launches of brs_mf_bpr_fwd_bwd individually (profiling / per-kernel timing)
extern "C" int brs_mf_bpr_prepare(const brs_mf_model* model, const int64_t* users, const int64_t* pos_items,
                                  const int64_t* neg_items, int64_t batch, void* stream) {
    return brs_mf_fwd_bwd_phases(model, LOSS_BPR, users, pos_items, neg_items, batch, 0.f, stream, 1);
}
extern "C" int brs_mf_bpr_fwd_bwd_prepared(const brs_mf_model* model, const int64_t* users, const int64_t* pos_items,
                                           const int64_t* neg_items, int64_t batch, float reg_weight, void* stream) {
    return brs_mf_fwd_bwd_phases(model, LOSS_BPR, users, pos_items, neg_items, batch, reg_weight, stream, 2);
}

extern "C" int brs_mf_bce_fwd_bwd(const brs_mf_model* model, const int64_t* users, const int64_t* items,
                                  const float* ratings, int64_t batch, float reg_weight, void* stream) {
    return brs_mf_fwd_bwd_impl(model, LOSS_BCE, users, items, ratings, batch, reg_weight, stream);
}

extern "C" int brs_mf_predict(const brs_mf_model* model, const int64_t* users, const int64_t* items, int64_t n,
                              float* scores, void* stream) {
    if (!users || !items || !scores || n < 0) return BRS_ERR_INVALID_ARG;
    MfArgs a;
    int rc = fill_args(model, a, false);
    if (rc != BRS_OK) return rc;
    if (n == 0) return BRS_OK;
    const int D = a.dim;
    if (D % 4 != 0 || D <= 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* err = &a.ws->predict_err;
#define BRS_PRED(LPR, VPL, FULL)                                                                                   \
    do {                                                                                                           \
        auto k = mf_predict_kernel<LPR, VPL, FULL>;                                                                \
        long long blocks = (n + kWarps * (32 / LPR) - 1) / (kWarps * (32 / LPR));                                  \
        k<<<grid_for((const void*)k, blocks), kThreads, 0, st>>>(a.user_emb, a.item_emb, a.user_bias, a.item_bias, \
                                                                 a.global_bias, D, a.n_users, a.n_items,           \
                                                                 (const long long*)users, (const long long*)items, \
                                                                 n, scores, err);                                  \
    } while (0)
    if (D == 4) BRS_PRED(1, 1, true);
    else if (D == 8) BRS_PRED(2, 1, true);
    else if (D == 16) BRS_PRED(4, 1, true);
    else if (D == 32) BRS_PRED(8, 1, true);
    else if (D == 64) BRS_PRED(16, 1, true);
    else if (D == 128) BRS_PRED(32, 1, true);
    else if (D == 256) BRS_PRED(32, 2, true);
    else if (D < 8) BRS_PRED(2, 1, false);
    else if (D < 16) BRS_PRED(4, 1, false);
    else if (D < 32) BRS_PRED(8, 1, false);
    else if (D < 64) BRS_PRED(16, 1, false);
    else if (D < 128) BRS_PRED(32, 1, false);
    else if (D < 256) BRS_PRED(32, 2, false);
    else if (D <= 384) BRS_PRED(32, 3, false);
    else BRS_PRED(32, 4, false);
#undef BRS_PRED
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}


// ---------------------------------------------------------------------------
// multi-GPU: push this rank's aggregated gradient rows to the owners' dense shard gradients
// ---------------------------------------------------------------------------
namespace {

struct PushArgs {
    const MfPeerTables* __restrict__ peers;
    brs_rowset rows[2];       // user / item rowsets (GLOBAL ids)
    float* scratch_emb[2];    // local compact scratch (sector-blocked)
    float* scratch_bias[2];
    int dim, shift, mask;
    float scale;              // DIRECT: -lr
};

__device__ __forceinline__ void red_or_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.relaxed.sys.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one warp per unique row of the batch (slot s of the local rowsets): copy the row and its bias from the
// owner's shard (peer loads over NVLink, or local) into this rank's staging tables at row s.  Every unique
// row crosses NVLink ONCE per step however many samples use it; two rows per warp in flight.
struct PullArgs {
    const MfPeerTables* __restrict__ peers;
    brs_rowset rows[2];
    float* stage_emb[2];   // [capacity, D] row-major
    float* stage_bias[2];  // [capacity]
    int dim, shift, mask;
};

__global__ void __launch_bounds__(256) mf_pull_kernel(const PullArgs a) {
    constexpr int R = 4;  // rows in flight per warp
    const int lane = threadIdx.x & 31;
    const int D = a.dim;
    const int cu = min(*a.rows[0].count, a.rows[0].capacity), ci = min(*a.rows[1].count, a.rows[1].capacity);
    const int total = cu + ci;
    const int n_warps = gridDim.x * 8;
    for (int w0 = blockIdx.x * 8 + (threadIdx.x >> 5); w0 < total; w0 += R * n_warps) {
        const float* src[R];
        float* dst[R];
        const float* bsrc[R];
        float* bdst[R];
        bool on[R];
        unsigned g[R];
        int e[R], sl[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int w = w0 + k * n_warps;
            on[k] = w < total;
            e[k] = (on[k] && w >= cu) ? 1 : 0;
            sl[k] = on[k] ? (e[k] == 0 ? w : w - cu) : 0;
            g[k] = (unsigned)a.rows[e[k]].list[sl[k]];
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const MfPeerTables* pt = a.peers + (int)(g[k] & (unsigned)a.mask);
            const unsigned lrow = g[k] >> a.shift;
            src[k] = (e[k] == 0 ? ldg_ptr(&pt->user_emb) : ldg_ptr(&pt->item_emb)) + (size_t)lrow * (unsigned)D;
            bsrc[k] = (e[k] == 0 ? ldg_ptr(&pt->user_bias) : ldg_ptr(&pt->item_bias)) + lrow;
            dst[k] = a.stage_emb[e[k]] + (size_t)sl[k] * (unsigned)D;
            bdst[k] = a.stage_bias[e[k]] + sl[k];
        }
        for (int c = lane * 4; c < D; c += 128) {
            float4 v[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
                if (on[k]) v[k] = ld_row4(src[k] + c);
#pragma unroll
            for (int k = 0; k < R; ++k)
                if (on[k]) *(float4*)(dst[k] + c) = v[k];
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (lane == k && on[k]) *bdst[k] = __ldg(bsrc[k]);
    }
}

// one warp per touched slot: 128-bit REDs of the whole row to its owner (coalesced 16*lanes-byte bursts
// over NVLink); the local scratch row is zeroed and the local slot released on the way.
//   DIRECT (SGD, linear in the gradient): adds -lr * g straight into the owner's WEIGHT row -- no dense
//           gradient table, no bitmap, no owner-side pass; needs a barrier between the last gather and here.
//   else  : adds g into the owner's dense per-shard gradient table and sets the row's touched bit.
template <bool DIRECT>
__global__ void __launch_bounds__(256) mf_push_kernel(const PushArgs a) {
    const int lane = threadIdx.x & 31;
    const int D = a.dim;
    const int cu = min(*a.rows[0].count, a.rows[0].capacity), ci = min(*a.rows[1].count, a.rows[1].capacity);
    const int total = cu + ci;
    const float sc = a.scale;
    for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < total; w += gridDim.x * 8) {
        const int e = w < cu ? 0 : 1;
        const int s = e == 0 ? w : w - cu;
        const brs_rowset& rs = a.rows[e];
        const unsigned g = (unsigned)rs.list[s];
        const int owner = (int)(g & (unsigned)a.mask);
        const unsigned lrow = g >> a.shift;
        const MfPeerTables* pt = a.peers + owner;
        float* base;
        if (DIRECT)
            base = (float*)(e == 0 ? ldg_ptr(&pt->user_emb) : ldg_ptr(&pt->item_emb));
        else
            base = e == 0 ? ldg_ptr(&pt->g_user_emb) : ldg_ptr(&pt->g_item_emb);
        float* dst = base + (size_t)lrow * (unsigned)D;
        for (int c = lane * 4; c < D; c += 128) {
            float* src = a.scratch_emb[e] + gs_off(D, rs.capacity, (unsigned)s, c);
            float4 v = *(const float4*)src;
            if (DIRECT) v = make_float4(sc * v.x, sc * v.y, sc * v.z, sc * v.w);
            red_add4_sys(dst + c, v);
            *(float4*)src = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) {
            float* sb = a.scratch_bias[e] + s;
            if (DIRECT) {
                red_add1_sys((float*)(e == 0 ? ldg_ptr(&pt->user_bias) : ldg_ptr(&pt->item_bias)) + lrow, sc * *sb);
            } else {
                red_add1_sys((e == 0 ? ldg_ptr(&pt->g_user_bias) : ldg_ptr(&pt->g_item_bias)) + lrow, *sb);
                red_or_u32((e == 0 ? ldg_ptr(&pt->user_bits) : ldg_ptr(&pt->item_bits)) + (lrow >> 5), 1u << (lrow & 31));
            }
            *sb = 0.f;
            rs.slot_map[g] = BRS_SLOT_NONE;
        }
    }
}

__global__ void mf_push_reset_kernel(int* c0, int* c1) {
    if (threadIdx.x == 0) {
        *c0 = 0;
        *c1 = 0;
    }
}

}  // namespace

extern "C" int brs_mf_sharded_bpr_fwd_bwd(const brs_mf_sharded* model, const int64_t* users, const int64_t* pos_items,
                                          const int64_t* neg_items, int64_t batch, int64_t global_batch,
                                          float reg_weight, void* stream) {
    if (!model || !model->peers || !users || !pos_items || !neg_items || batch < 0 || global_batch < batch)
        return BRS_ERR_INVALID_ARG;
    const int w = model->world;
    if (w < 1 || w > BRS_MAX_RANKS || (w & (w - 1)) != 0 || model->rank < 0 || model->rank >= w) return BRS_ERR_UNSUPPORTED;
    MfArgs a;
    int rc = fill_args(&model->stage, a, true);
    if (rc != BRS_OK) return rc;
    if (model->stage.user.rows.n_rows != model->n_users || model->stage.item.rows.n_rows != model->n_items)
        return BRS_ERR_INVALID_ARG;  // the staging slot maps are indexed by GLOBAL ids
    if (batch == 0) return BRS_OK;
    int shift = 0;
    while ((1 << shift) < w) ++shift;
    a.n_users = model->n_users;
    a.n_items = model->n_items;
    a.peers = (const MfPeerTables*)model->peers;
    a.shard_shift = shift;
    a.shard_mask = w - 1;
    a.users = (const long long*)users;
    a.items = (const long long*)pos_items;
    a.third = neg_items;
    a.batch = batch;
    a.reg_w = reg_weight;
    a.inv_b = 1.0f / (float)global_batch;
    cudaStream_t st = (cudaStream_t)stream;
    {   // local pre-pass over GLOBAL ids
        const brs_rowset rs[3] = {model->stage.user.rows, model->stage.item.rows, model->stage.item.rows};
        const long long* idx[3] = {a.users, a.items, (const long long*)neg_items};
        const long long n[3] = {batch, batch, batch};
        rc = brs_assign_slots(rs, idx, n, 3, a.ws, st);
        if (rc != BRS_OK) return rc;
    }
    if (shard_mode(w) == SHARD_DIRECT) return launch_fwd_bwd<LOSS_BPR, SHARD_DIRECT>(a, st);
    if (!model->pull_user_emb || !model->pull_item_emb || !model->pull_user_bias || !model->pull_item_bias)
        return BRS_ERR_INVALID_ARG;
    PullArgs p;
    p.peers = a.peers;
    p.rows[0] = model->stage.user.rows;
    p.rows[1] = model->stage.item.rows;
    p.stage_emb[0] = model->pull_user_emb;
    p.stage_emb[1] = model->pull_item_emb;
    p.stage_bias[0] = model->pull_user_bias;
    p.stage_bias[1] = model->pull_item_bias;
    p.dim = a.dim;
    p.shift = shift;
    p.mask = w - 1;
    mf_pull_kernel<<<brs_sm_count() * 8, 256, 0, st>>>(p);
    // the fused kernel now runs on local memory only: staging rows, indexed by slot
    a.user_emb = model->pull_user_emb;
    a.item_emb = model->pull_item_emb;
    a.user_bias = model->pull_user_bias;
    a.item_bias = model->pull_item_bias;
    return launch_fwd_bwd<LOSS_BPR, SHARD_STAGED>(a, st);
}

extern "C" int brs_debug_set_shard_mode(int mode) {
    if (mode != 0 && mode != SHARD_DIRECT && mode != SHARD_STAGED) return BRS_ERR_INVALID_ARG;
    g_shard_mode = mode == 0 ? -1 : mode;  // 0: pick by world size
    return BRS_OK;
}

extern "C" int brs_mf_sharded_push(const brs_mf_sharded* model, const brs_opt* opt, void* stream) {
    if (!model || !model->peers || !opt) return BRS_ERR_INVALID_ARG;
    const int w = model->world;
    if (w < 1 || w > BRS_MAX_RANKS || (w & (w - 1)) != 0) return BRS_ERR_UNSUPPORTED;
    MfArgs a;
    int rc = fill_args(&model->stage, a, true);
    if (rc != BRS_OK) return rc;
    int shift = 0;
    while ((1 << shift) < w) ++shift;
    cudaStream_t st = (cudaStream_t)stream;
    PushArgs p;
    p.peers = (const MfPeerTables*)model->peers;
    p.rows[0] = model->stage.user.rows;
    p.rows[1] = model->stage.item.rows;
    p.scratch_emb[0] = a.g_user_emb;
    p.scratch_emb[1] = a.g_item_emb;
    p.scratch_bias[0] = a.g_user_bias;
    p.scratch_bias[1] = a.g_item_bias;
    p.dim = a.dim;
    p.shift = shift;
    p.mask = w - 1;
    p.scale = -(float)opt->lr;
    if (opt->kind == BRS_SGD)
        mf_push_kernel<true><<<brs_sm_count() * 4, 256, 0, st>>>(p);
    else
        mf_push_kernel<false><<<brs_sm_count() * 4, 256, 0, st>>>(p);
    mf_push_reset_kernel<<<1, 32, 0, st>>>(model->stage.user.rows.count, model->stage.item.rows.count);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
