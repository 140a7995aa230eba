// Touched-row bookkeeping and optimizer.step() for embedding tables -- sm_100a.
//
// Replaces torch.optim.{SGD,Adam,RMSprop}.step over dense [N,D] gradients
// (beta_rec/models/torch_engine.py:23-39, called at beta_rec/models/mf.py:118).
//
//  * assign_slots: one thread per batch index; the first toucher of a row claims the
//    next slot of the entity's compact gradient scratch (slot_map / list / count).
//  * rows kernels: a warp per group of TOUCHED rows (4 rows in flight per warp): read
//    the scratch row + weight row (+ m, v), update, write back, zero the scratch row,
//    release the slot.  Exact for SGD (g = 0 elsewhere => no change).
//  * dense sweep: every element of every table, g = 0 for untouched rows -- what the
//    reference's dense Adam/RMSprop actually does each step.
// The last block to finish also applies the ws-sourced scalar (MF global bias),
// publishes {loss, regularizer, status} and resets the step scratch.
#include "common.cuh"
#include "opt_math.cuh"

int brs_assign_slots(const brs_rowset* rs, const long long* const* idx, const long long* n, int n_arrays,
                     brs_step_ws* ws, cudaStream_t st);

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxEntities = 4;
constexpr int kMaxDense = 16;
constexpr int kMaxAssign = 4;

// ---------------------------------------------------------------------------
// slot assignment pre-pass
// ---------------------------------------------------------------------------
struct AssignArgs {
    brs_rowset rs[kMaxAssign];
    const long long* idx[kMaxAssign];
    long long n[kMaxAssign];
    int n_arrays;
    unsigned int* err_flag;
};

// grid = (blocks, n_arrays): every block works on ONE index array, so all of its claims go to
// one rowset and are aggregated into a single atomicAdd on that rowset's counter (the v1
// per-thread atomicAdd on one address was 55% of this kernel's stall samples).
// body shared by the stand-alone pre-pass and by the apply kernel's "prepare the next batch" blocks:
// block `bx` of `nbx` works on index array k
__device__ __forceinline__ void assign_slots_block(const AssignArgs& a, int k, int bx, int nbx) {
    __shared__ int s_warp_cnt[kWarps];
    __shared__ int s_base;
    const brs_rowset rs = a.rs[k];
    const long long* __restrict__ idx = a.idx[k];
    const long long n = a.n[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)nbx * kThreads;
    const long long n_iter = (n + stride - 1) / stride;
    for (long long it = 0; it < n_iter; ++it) {
        const long long t = it * stride + (long long)bx * kThreads + threadIdx.x;
        long long row = -1;
        bool won = false;
        if (t < n) {
            row = idx[t];
            if ((unsigned long long)row >= (unsigned long long)rs.n_rows) {
                atomicOr(a.err_flag, 1u);  // the reference raises IndexError (nn.Embedding)
                row = -1;
            } else {
                int* m = rs.slot_map + row;
                // cheap read first: hot (Zipf) rows are claimed by the time most samples arrive
                if (*((volatile int*)m) == BRS_SLOT_NONE) won = atomicCAS(m, BRS_SLOT_NONE, BRS_SLOT_PENDING) == BRS_SLOT_NONE;
            }
        }
        const unsigned ballot = __ballot_sync(BRS_FULL_MASK, won);
        const int lane_off = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int c = s_warp_cnt[w];
                s_warp_cnt[w] = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(rs.count, tot) : 0;
        }
        __syncthreads();
        if (won) {
            const int slot = s_base + s_warp_cnt[warp] + lane_off;
            if (slot < rs.capacity) {
                rs.list[slot] = (int)row;
                rs.slot_map[row] = slot;  // consumers run in later kernels of the same stream
            } else {
                rs.slot_map[row] = BRS_SLOT_NONE;
                atomicOr(a.err_flag, 2u);
            }
        }
        __syncthreads();  // s_warp_cnt / s_base are reused by the next iteration
    }
}

__global__ void __launch_bounds__(kThreads) assign_slots_kernel(const AssignArgs a) {
    assign_slots_block(a, blockIdx.y, blockIdx.x, gridDim.x);
}

struct ApplyArgs {
    brs_entity ent[kMaxEntities];
    int n_ent;
    brs_dense_param dense[kMaxDense];
    int n_dense;
    int dense_grad_from_ws;  // dense[0].grad is ws->g_global_bias (MF)
    OptParams opt;
    brs_step_ws* ws;    // may be NULL for the stand-alone generic entry points
    long long t_explicit;  // step number when ws == NULL
    float* out;         // device brs_step_out (float[4]) or NULL
    double inv_batch;
    int advance_step;   // last block: ws->step += 1, reset sums
    int pol_scratch, pol_weight;  // L2 eviction policies (BRS_L2_*)
    // optional: blocks [n_apply_blocks, gridDim.x) run the slot pre-pass of the NEXT batch (on the
    // alternate rowsets) inside this launch -- both halves are latency-bound and overlap well
    int n_apply_blocks;   // 0: every block applies
    int parity;           // which ws->err_pending slot belongs to the batch being applied
    AssignArgs next;
    // row-owner MF step (mf_rowwise.cu): the touched rows (slot >= 0) were already updated by their owners,
    // the sweep only moves the others (g = 0); an index error voids the whole step
    int skip_touched;
};

// one touched row of one table: scratch row `slot` -> weight row `row`
template <int KIND>
__device__ __forceinline__ void update_row(const brs_table& tb, int cap, long long row, long long slot, int lane,
                                           const OptScalars& s) {
    const int d = tb.dim;
    float* w = tb.weight + row * d;
    float* g = tb.grad;  // addressed through gs_off (sector-blocked when d % 8 == 0)
    float* m = (KIND == BRS_ADAM) ? tb.m + row * d : nullptr;
    float* v = (KIND != BRS_SGD) ? tb.v + row * d : nullptr;
    if ((d & 3) == 0) {
        for (int c = lane * 4; c < d; c += 128) {
            float4* gp = (float4*)(g + gs_off(d, cap, (unsigned)slot, c));
            float4 gv = *gp;
            float4 wv = *(const float4*)(w + c);
            float4 mv = make_float4(0.f, 0.f, 0.f, 0.f), vv = mv;
            if (KIND == BRS_ADAM) mv = *(const float4*)(m + c);
            if (KIND != BRS_SGD) vv = *(const float4*)(v + c);
            opt_elem4<KIND>(wv, gv, mv, vv, s);
            *(float4*)(w + c) = wv;
            if (KIND == BRS_ADAM) *(float4*)(m + c) = mv;
            if (KIND != BRS_SGD) *(float4*)(v + c) = vv;
            *gp = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        g += (size_t)slot * d;  // row-major for dims that are not multiples of 8
        for (int c = lane; c < d; c += 32) {
            float gv = g[c], wv = w[c];
            float mv = (KIND == BRS_ADAM) ? m[c] : 0.f;
            float vv = (KIND != BRS_SGD) ? v[c] : 0.f;
            opt_elem<KIND>(wv, gv, mv, vv, s);
            w[c] = wv;
            if (KIND == BRS_ADAM) m[c] = mv;
            if (KIND != BRS_SGD) v[c] = vv;
            g[c] = 0.f;
        }
    }
}

// fast path: ROWS touched rows of a (dim % 4 == 0, dim <= 128) table in flight per warp.
// SGD needs no weight load at all: w += -lr*g leaves as a fire-and-forget 128-bit RED.
template <int KIND, int ROWS>
__device__ __forceinline__ void update_rows_small(const brs_table& tb, int cap, const int (&row)[ROWS], int s0, int n,
                                                  int lane, const OptScalars& s, unsigned long long pol_s,
                                                  unsigned long long pol_w) {
    const int d = tb.dim, c = lane * 4;
    if (c >= d) return;
    float4 gv[ROWS], wv[ROWS], mv[ROWS], vv[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        wv[r] = mv[r] = vv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) {
            const size_t ro = (size_t)(unsigned)row[r] * (unsigned)d + c;
            gv[r] = ld4_pol(tb.grad + gs_off(d, cap, (unsigned)(s0 + r), c), pol_s);
            if (KIND != BRS_SGD) wv[r] = ld4_pol(tb.weight + ro, pol_w);
            if (KIND == BRS_ADAM) mv[r] = *(const float4*)(tb.m + ro);
            if (KIND != BRS_SGD) vv[r] = *(const float4*)(tb.v + ro);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (r < n) {
            const size_t ro = (size_t)(unsigned)row[r] * (unsigned)d + c;
            if (KIND == BRS_SGD) {
                red_add4_pol(tb.weight + ro, make_float4(-s.lr * gv[r].x, -s.lr * gv[r].y, -s.lr * gv[r].z, -s.lr * gv[r].w), pol_w);
            } else {
                opt_elem4<KIND>(wv[r], gv[r], mv[r], vv[r], s);
                st4_pol(tb.weight + ro, wv[r], pol_w);
                if (KIND == BRS_ADAM) *(float4*)(tb.m + ro) = mv[r];
                *(float4*)(tb.v + ro) = vv[r];
            }
            st4_pol(tb.grad + gs_off(d, cap, (unsigned)(s0 + r), c), make_float4(0.f, 0.f, 0.f, 0.f), pol_s);
        }
    }
}

// dense (Linear / bias) parameters whose gradients were completed by earlier kernels:
// element-wise over the whole grid; their grads are cleared for the next step
template <int KIND>
__device__ __forceinline__ void dense_params_update(const ApplyArgs& a, const OptScalars& s) {
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long nthreads = (long long)(a.n_apply_blocks > 0 ? a.n_apply_blocks : gridDim.x) * kThreads;
    for (int k = (a.dense_grad_from_ws ? 1 : 0); k < a.n_dense; ++k) {
        const brs_dense_param& dp = a.dense[k];
        for (long long e = tid; e < dp.numel; e += nthreads) {
            const float g = dp.grad[e];
            float w = dp.weight[e];
            float m = (KIND == BRS_ADAM) ? dp.m[e] : 0.f;
            float v = (KIND != BRS_SGD) ? dp.v[e] : 0.f;
            opt_elem<KIND>(w, g, m, v, s);
            dp.weight[e] = w;
            if (KIND == BRS_ADAM) dp.m[e] = m;
            if (KIND != BRS_SGD) dp.v[e] = v;
            dp.grad[e] = 0.f;
        }
    }
}

// the ws-sourced scalar parameter (MF global bias) + publication of the step's scalars;
// run by the LAST block only (every other block has finished reading ws by then)
template <int KIND>
__device__ void finalize(const ApplyArgs& a, const OptScalars& s) {
    const unsigned int status = a.ws ? (a.ws->err_flag | a.ws->err_pending[a.parity & 1]) : 0u;
    const bool void_step = a.skip_touched && status != 0u;  // row-owner step: nothing was updated
    if (a.dense_grad_from_ws && a.n_dense > 0 && threadIdx.x == 0 && !void_step) {
        const brs_dense_param& dp = a.dense[0];
        const float g = a.ws->g_global_bias;
        float w = dp.weight[0];
        float m = (KIND == BRS_ADAM) ? dp.m[0] : 0.f;
        float v = (KIND != BRS_SGD) ? dp.v[0] : 0.f;
        opt_elem<KIND>(w, g, m, v, s);
        dp.weight[0] = w;
        if (KIND == BRS_ADAM) dp.m[0] = m;
        if (KIND != BRS_SGD) dp.v[0] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && a.ws) {
        if (a.out) {  // brs_step_out
            a.out[0] = (float)(a.ws->loss_sum * a.inv_batch);
            a.out[1] = (float)(a.ws->reg_sum * a.inv_batch);
            // 0 ok | 1 index out of range | 2 touched-row capacity overflow
            ((int*)a.out)[2] = (int)status;
            a.out[3] = 0.f;
        }
        if (a.advance_step) {
            a.ws->err_flag = 0u;
            a.ws->err_pending[a.parity & 1] = 0u;
            a.ws->loss_sum = 0.0;
            a.ws->reg_sum = 0.0;
            a.ws->g_global_bias = 0.f;
            if (!void_step) a.ws->step += 1;
        }
        a.ws->ticket = 0u;
    }
}

__device__ __forceinline__ bool last_block(const ApplyArgs& a) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(&a.ws->ticket, 1u);
        s_last = (prev == (unsigned)(a.n_apply_blocks > 0 ? a.n_apply_blocks : gridDim.x) - 1);
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

__device__ __forceinline__ int clamped_count(const brs_rowset& rs) {
    const int c = *rs.count;
    return c < rs.capacity ? c : rs.capacity;
}

// ---- touched rows only ------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kThreads) rows_apply_kernel(const ApplyArgs a) {
    constexpr int ROWS = 4;
    const int n_apply = a.n_apply_blocks > 0 ? a.n_apply_blocks : gridDim.x;
    if ((int)blockIdx.x >= n_apply) {  // "prepare" role: slot pre-pass of the next batch
        const int nb = gridDim.x - n_apply;          // blocks shared by the index arrays
        const int per = nb / a.next.n_arrays;        // launch guarantees per >= 1
        const int q = blockIdx.x - n_apply;
        const int k = q / per;
        if (k < a.next.n_arrays) assign_slots_block(a.next, k, q - k * per, per);
        return;
    }
    __shared__ OptScalars s_opt;
    __shared__ int s_cnt[kMaxEntities + 1];  // prefix of work items (groups of ROWS slots)
    __shared__ int s_rows[kMaxEntities];     // touched rows per entity
    if (threadIdx.x == 0) {
        const long long t = a.ws ? a.ws->step + 1 : a.t_explicit;
        s_opt = make_scalars(a.opt, t);
        int acc = 0;
        for (int e = 0; e < a.n_ent; ++e) {
            s_cnt[e] = acc;
            s_rows[e] = clamped_count(a.ent[e].rows);
            acc += (s_rows[e] + ROWS - 1) / ROWS;
        }
        s_cnt[a.n_ent] = acc;
    }
    __syncthreads();
    const OptScalars s = s_opt;
    const unsigned long long pol_s = l2_policy(a.pol_scratch), pol_w = l2_policy(a.pol_weight);
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(BRS_FULL_MASK, threadIdx.x >> 5, 0);
    const int total = s_cnt[a.n_ent];
    for (int r = blockIdx.x * kWarps + warp; r < total; r += n_apply * kWarps) {
        int e = 0;
        while (e + 1 < a.n_ent && r >= s_cnt[e + 1]) ++e;
        const brs_entity& en = a.ent[e];
        const int s0 = (r - s_cnt[e]) * ROWS;
        const int n = min(ROWS, s_rows[e] - s0);
        // lanes 0..ROWS-1 fetch the row ids, everyone gets them by shuffle (registers, no local memory)
        const int my_row = (lane < n) ? en.rows.list[s0 + lane] : 0;
        int row[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) row[k] = __shfl_sync(BRS_FULL_MASK, my_row, k);
        for (int k = 0; k < en.n_tables; ++k) {
            const brs_table& tb = en.table[k];
            if ((tb.dim & 3) == 0 && tb.dim <= 128) {
                update_rows_small<KIND, ROWS>(tb, en.rows.capacity, row, s0, n, lane, s, pol_s, pol_w);
            } else if (tb.dim == 1) {  // bias tables: lane q handles row q
                if (lane < n) update_row<KIND>(tb, en.rows.capacity, my_row, s0 + lane, 0, s);
            } else {
#pragma unroll
                for (int q = 0; q < ROWS; ++q)
                    if (q < n) update_row<KIND>(tb, en.rows.capacity, row[q], s0 + q, lane, s);
            }
        }
        if (lane < n) en.rows.slot_map[my_row] = BRS_SLOT_NONE;  // release the slot
    }
    dense_params_update<KIND>(a, s);
    if (a.ws) {
        if (last_block(a)) {
            if (threadIdx.x < a.n_ent) *a.ent[threadIdx.x].rows.count = 0;
            finalize<KIND>(a, s);
        }
    }
}

// ---- every row (reference-exact Adam / RMSprop) -----------------------------
// grid-stride over float4 vectors (or scalars when dim % 4 != 0) of one table
template <int KIND>
__device__ __forceinline__ void sweep_table(const brs_table& tb, const int* __restrict__ slot_map, int cap,
                                            const OptScalars& s, long long tid, long long nthreads, bool skip_touched) {
    const int d = tb.dim;
    if ((d & 3) == 0) {
        const int vpr = d >> 2;
        const long long nvec = tb.n_rows * vpr;
        const int shift = (vpr & (vpr - 1)) == 0 ? __ffs(vpr) - 1 : -1;
        for (long long i = tid; i < nvec; i += nthreads) {
            const long long row = shift >= 0 ? (i >> shift) : (nvec < (1ll << 31) ? (long long)((unsigned)i / (unsigned)vpr) : i / vpr);
            const int slot = slot_map[row];
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (skip_touched) {
                if (slot >= 0) continue;
            } else if (slot >= 0) {
                float4* gp = (float4*)(tb.grad + gs_off(d, cap, (unsigned)slot, (int)(i - row * vpr) * 4));
                gv = *gp;
                *gp = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 wv = ((const float4*)tb.weight)[i];
            float4 mv = make_float4(0.f, 0.f, 0.f, 0.f), vv = mv;
            if (KIND == BRS_ADAM) mv = ((const float4*)tb.m)[i];
            if (KIND != BRS_SGD) vv = ((const float4*)tb.v)[i];
            opt_elem4<KIND>(wv, gv, mv, vv, s);
            ((float4*)tb.weight)[i] = wv;
            if (KIND == BRS_ADAM) ((float4*)tb.m)[i] = mv;
            if (KIND != BRS_SGD) ((float4*)tb.v)[i] = vv;
        }
    } else {
        const long long n = tb.n_rows * d;
        for (long long i = tid; i < n; i += nthreads) {
            const long long row = i / d;
            const int slot = slot_map[row];
            float g = 0.f;
            if (skip_touched) {
                if (slot >= 0) continue;
            } else if (slot >= 0) {
                float* gp = tb.grad + (long long)slot * d + (i - row * d);
                g = *gp;
                *gp = 0.f;
            }
            float w = tb.weight[i];
            float m = (KIND == BRS_ADAM) ? tb.m[i] : 0.f;
            float v = (KIND != BRS_SGD) ? tb.v[i] : 0.f;
            opt_elem<KIND>(w, g, m, v, s);
            tb.weight[i] = w;
            if (KIND == BRS_ADAM) tb.m[i] = m;
            if (KIND != BRS_SGD) tb.v[i] = v;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) dense_sweep_kernel(const ApplyArgs a) {
    __shared__ OptScalars s_opt;
    if (threadIdx.x == 0) s_opt = make_scalars(a.opt, a.ws ? a.ws->step + 1 : a.t_explicit);
    __syncthreads();
    const OptScalars s = s_opt;
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * kThreads;
    const bool void_step = a.skip_touched && a.ws && a.ws->err_pending[a.parity & 1] != 0u;
    if (!void_step) {
        for (int e = 0; e < a.n_ent; ++e)
            for (int k = 0; k < a.ent[e].n_tables; ++k)
                sweep_table<KIND>(a.ent[e].table[k], a.ent[e].rows.slot_map, a.ent[e].rows.capacity, s, tid, nthreads,
                                  a.skip_touched != 0);
        dense_params_update<KIND>(a, s);
    }
    if (a.ws) {
        if (last_block(a)) finalize<KIND>(a, s);
    }
}

// releases the slots after a dense sweep (the sweep itself reads the slot maps)
__global__ void __launch_bounds__(kThreads) rowset_clear_kernel(const ApplyArgs a) {
    for (int e = 0; e < a.n_ent; ++e) {
        const brs_rowset& rs = a.ent[e].rows;
        const int c = clamped_count(rs);
        for (int r = blockIdx.x * kThreads + threadIdx.x; r < c; r += gridDim.x * kThreads)
            rs.slot_map[rs.list[r]] = BRS_SLOT_NONE;
    }
}

__global__ void rowset_reset_counts_kernel(const ApplyArgs a) {
    if (threadIdx.x < a.n_ent) *a.ent[threadIdx.x].rows.count = 0;
}

// ---------------------------------------------------------------------------
// multi-GPU owner side: apply the optimizer to the rows of this shard whose touched bit is set
// (the gradients were pushed into the shard's DENSE gradient tables by all ranks, mf_push_kernel)
// ---------------------------------------------------------------------------
struct ShardEnt {
    float* w; float* wb;      // shard weights [rows, dim], [rows]
    float* g; float* gb;      // dense gradients, same shapes (row-major)
    float* m; float* v; float* mb; float* vb;
    unsigned int* bits;
    long long rows;
};
struct ShardApplyArgs {
    ShardEnt ent[2];
    int dim;
    int dense_all;   // 1: every row is updated (g = 0 where the bit is clear): reference-exact Adam / RMSprop
    ApplyArgs base;  // opt, ws, out, inv_batch, global-bias dense param, finalize bookkeeping
};

// Owner-side optimizer pass over the touched bitmap.  A warp takes one BYTE of the bitmap (8 rows) at a
// time and keeps up to four rows in flight (all loads issued before the first dependent store), so a
// dense word -- every item row at large global batches -- costs two load round trips, not 32.
template <int KIND>
__global__ void __launch_bounds__(kThreads) shard_apply_kernel(const ShardApplyArgs a) {
    __shared__ OptScalars s_opt;
    if (threadIdx.x == 0) s_opt = make_scalars(a.base.opt, a.base.ws->step + 1);
    __syncthreads();
    const OptScalars s = s_opt;
    const int lane = threadIdx.x & 31;
    const int D = a.dim;
    constexpr int R = 4;
    const long long tasks0 = (a.ent[0].rows + 7) >> 3, tasks1 = (a.ent[1].rows + 7) >> 3;
    for (long long ti = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); ti < tasks0 + tasks1;
         ti += (long long)gridDim.x * kWarps) {
        const ShardEnt& e = a.ent[ti < tasks0 ? 0 : 1];
        const long long task = ti < tasks0 ? ti : ti - tasks0;
        unsigned char* bytes = (unsigned char*)e.bits;  // little-endian: byte k of the bitmap = rows 8k..8k+7
        const unsigned int bits = bytes[task];
        unsigned int todo = bits;
        if (a.dense_all) {
            const long long left = e.rows - task * 8;
            todo = left >= 8 ? 0xffu : ((1u << left) - 1u);
        }
        const long long row0 = task * 8;
        while (todo) {
            int rb[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                rb[k] = todo ? __ffs(todo) - 1 : -1;
                todo &= todo - 1;  // 0 stays 0
            }
            for (int c = lane * 4; c < D; c += 128) {
                float4 gv[R], wv[R], mv[R], vv[R];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    gv[k] = mv[k] = vv[k] = wv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rb[k] >= 0) {
                        const size_t o = (size_t)(row0 + rb[k]) * D + c;
                        if ((bits >> rb[k]) & 1u) gv[k] = *(const float4*)(e.g + o);
                        wv[k] = *(const float4*)(e.w + o);
                        if (KIND == BRS_ADAM) mv[k] = *(const float4*)(e.m + o);
                        if (KIND != BRS_SGD) vv[k] = *(const float4*)(e.v + o);
                    }
                }
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (rb[k] >= 0) {
                        const size_t o = (size_t)(row0 + rb[k]) * D + c;
                        opt_elem4<KIND>(wv[k], gv[k], mv[k], vv[k], s);
                        *(float4*)(e.w + o) = wv[k];
                        if (KIND == BRS_ADAM) *(float4*)(e.m + o) = mv[k];
                        if (KIND != BRS_SGD) *(float4*)(e.v + o) = vv[k];
                        if ((bits >> rb[k]) & 1u) *(float4*)(e.g + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            // biases: lane k owns row rb[k]
            int my = -1;
#pragma unroll
            for (int k = 0; k < R; ++k)
                if (lane == k) my = rb[k];
            if (my >= 0) {
                const long long row = row0 + my;
                float g = 0.f;
                if ((bits >> my) & 1u) {
                    g = e.gb[row];
                    e.gb[row] = 0.f;
                }
                float w = e.wb[row];
                float m = (KIND == BRS_ADAM) ? e.mb[row] : 0.f;
                float v = (KIND != BRS_SGD) ? e.vb[row] : 0.f;
                opt_elem<KIND>(w, g, m, v, s);
                e.wb[row] = w;
                if (KIND == BRS_ADAM) e.mb[row] = m;
                if (KIND != BRS_SGD) e.vb[row] = v;
            }
        }
        if (lane == 0 && bits) bytes[task] = 0;
    }
    if (last_block(a.base)) finalize<KIND>(a.base, s);
}

int validate_entities(const brs_entity* ents, int n, int kind) {
    if (n < 0 || n > kMaxEntities || (n > 0 && !ents)) return BRS_ERR_INVALID_ARG;
    for (int e = 0; e < n; ++e) {
        const brs_entity& en = ents[e];
        if (en.n_tables < 0 || en.n_tables > BRS_MAX_ENTITY_TABLES) return BRS_ERR_INVALID_ARG;
        if (!en.rows.slot_map || !en.rows.list || !en.rows.count || en.rows.capacity <= 0) return BRS_ERR_INVALID_ARG;
        for (int k = 0; k < en.n_tables; ++k) {
            const brs_table& t = en.table[k];
            if (!t.weight || !t.grad || t.dim <= 0 || t.n_rows < 0) return BRS_ERR_INVALID_ARG;
            if (kind == BRS_ADAM && (!t.m || !t.v)) return BRS_ERR_INVALID_ARG;
            if (kind == BRS_RMSPROP && !t.v) return BRS_ERR_INVALID_ARG;
            if ((t.dim & 3) == 0 && ((((uintptr_t)t.weight | (uintptr_t)t.grad | (uintptr_t)t.m | (uintptr_t)t.v) & 15) != 0))
                return BRS_ERR_INVALID_ARG;
        }
    }
    return BRS_OK;
}

int validate_dense(const brs_dense_param* p, int n, int kind, bool grad_from_ws) {
    if (n < 0 || n > kMaxDense || (n > 0 && !p)) return BRS_ERR_INVALID_ARG;
    for (int k = 0; k < n; ++k) {
        if (!p[k].weight || p[k].numel < 0) return BRS_ERR_INVALID_ARG;
        if (!(k == 0 && grad_from_ws) && !p[k].grad) return BRS_ERR_INVALID_ARG;
        if (kind == BRS_ADAM && (!p[k].m || !p[k].v)) return BRS_ERR_INVALID_ARG;
        if (kind == BRS_RMSPROP && !p[k].v) return BRS_ERR_INVALID_ARG;
    }
    return BRS_OK;
}

void fill_opt(ApplyArgs& a, const brs_opt* opt) {
    a.opt.kind = opt->kind;
    a.opt.lr = opt->lr;
    a.opt.beta1 = opt->beta1;
    a.opt.beta2 = opt->beta2;
    a.opt.eps = opt->eps;
    a.opt.alpha = opt->alpha;
}

int persistent_grid(const void* kernel) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    return brs_sm_count() * per_sm;
}

template <int KIND>
int launch_apply(const ApplyArgs& a, int mode, long long max_rows_hint, cudaStream_t st) {
    if (KIND != BRS_SGD && mode == BRS_DENSE) {
        auto k = dense_sweep_kernel<KIND>;
        k<<<persistent_grid((const void*)k), kThreads, 0, st>>>(a);
        rowset_clear_kernel<<<brs_sm_count(), kThreads, 0, st>>>(a);
        rowset_reset_counts_kernel<<<1, 32, 0, st>>>(a);
    } else {
        auto k = rows_apply_kernel<KIND>;
        int grid = persistent_grid((const void*)k);
        long long need = (max_rows_hint / 4 + kWarps) / kWarps;  // 4 rows per warp work item
        for (int d = 0; d < a.n_dense; ++d) need = max(need, (long long)((a.dense[d].numel + kThreads - 1) / kThreads));
        if (need < 1) need = 1;
        if (grid > need) grid = (int)need;
        if (a.next.n_arrays > 0) {  // fused "apply batch t + prepare batch t+1" launch
            ApplyArgs b = a;
            long long nmax = 0;
            for (int q = 0; q < a.next.n_arrays; ++q) nmax = max(nmax, a.next.n[q]);
            long long per = (nmax + kThreads - 1) / kThreads;
            const long long cap = (long long)brs_sm_count() * 2;
            if (per > cap) per = cap;
            if (per < 1) per = 1;
            b.n_apply_blocks = grid;
            k<<<grid + (int)per * a.next.n_arrays, kThreads, 0, st>>>(b);
        } else {
            k<<<grid, kThreads, 0, st>>>(a);
        }
        if (!a.ws) rowset_reset_counts_kernel<<<1, 32, 0, st>>>(a);  // stand-alone use: no last-block finalize
    }
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// grad_scratch[slot_map[idx[k]]] += scale * src[k]: the scatter half of an embedding backward
__global__ void __launch_bounds__(kThreads) rows_scatter_grad_kernel(brs_table tb, const int* __restrict__ slot_map,
                                                                     int cap, const long long* __restrict__ idx,
                                                                     long long n, const float* __restrict__ src,
                                                                     float scale) {
    const int lane = threadIdx.x & 31;
    const int d = tb.dim;
    for (long long k = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); k < n; k += (long long)gridDim.x * kWarps) {
        const long long row = idx[k];
        if ((unsigned long long)row >= (unsigned long long)tb.n_rows) continue;
        const long long slot = slot_map[row];
        if (slot < 0) continue;
        const float* sp = src + k * d;
        if ((d & 3) == 0) {
            for (int c = lane * 4; c < d; c += 128) {
                const float4 x = *(const float4*)(sp + c);
                red_add4(tb.grad + gs_off(d, cap, (unsigned)slot, c), make_float4(scale * x.x, scale * x.y, scale * x.z, scale * x.w));
            }
        } else {
            for (int c = lane; c < d; c += 32) red_add1(tb.grad + slot * d + c, scale * sp[c]);
        }
    }
}

// out[k, :] = grad_scratch[slot(idx[k])]: the inverse of rows_scatter_grad_kernel (rows without a slot read as 0)
__global__ void __launch_bounds__(kThreads) rows_read_grad_kernel(brs_table tb, const int* __restrict__ slot_map, int cap,
                                                                  const long long* __restrict__ idx, long long n,
                                                                  float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int d = tb.dim;
    for (long long k = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); k < n; k += (long long)gridDim.x * kWarps) {
        const long long row = idx[k];
        long long slot = -1;
        if ((unsigned long long)row < (unsigned long long)tb.n_rows) slot = slot_map[row];
        float* op = out + k * d;
        if ((d & 3) == 0) {
            for (int c = lane * 4; c < d; c += 128)
                *(float4*)(op + c) = slot < 0 ? make_float4(0.f, 0.f, 0.f, 0.f)
                                              : *(const float4*)(tb.grad + gs_off(d, cap, (unsigned)slot, c));
        } else {
            for (int c = lane; c < d; c += 32) op[c] = slot < 0 ? 0.f : tb.grad[slot * d + c];
        }
    }
}

}  // namespace

extern "C" int brs_rows_read_grad(const brs_entity* entity, int32_t table, const int64_t* idx, int64_t n, float* out,
                                  void* stream) {
    if (!entity || table < 0 || table >= entity->n_tables || !idx || !out || n < 0) return BRS_ERR_INVALID_ARG;
    const brs_table& tb = entity->table[table];
    if (!tb.grad || !entity->rows.slot_map) return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    long long blocks = (n + kWarps - 1) / kWarps;
    const long long cap = (long long)brs_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    rows_read_grad_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(tb, entity->rows.slot_map, entity->rows.capacity,
                                                                             (const long long*)idx, n, out);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_rows_assign(const brs_rowset* rows, const int64_t* idx, int64_t n, void* ws, void* stream) {
    if (!rows || !ws) return BRS_ERR_INVALID_ARG;
    const long long* ip = (const long long*)idx;
    const long long nn = n;
    return brs_assign_slots(rows, &ip, &nn, 1, (brs_step_ws*)ws, (cudaStream_t)stream);
}

extern "C" int brs_rows_scatter_grad(const brs_entity* entity, int32_t table, const int64_t* idx, int64_t n,
                                     const float* src, float scale, void* stream) {
    if (!entity || table < 0 || table >= entity->n_tables || !idx || !src || n < 0) return BRS_ERR_INVALID_ARG;
    const brs_table& tb = entity->table[table];
    if (!tb.grad || !entity->rows.slot_map) return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    long long blocks = (n + kWarps - 1) / kWarps;
    const long long cap = (long long)brs_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    rows_scatter_grad_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        tb, entity->rows.slot_map, entity->rows.capacity, (const long long*)idx, n, src, scale);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// range-check the index arrays and give every touched row a slot (shared with the model kernels)
int brs_assign_slots(const brs_rowset* rs, const long long* const* idx, const long long* n, int n_arrays,
                     brs_step_ws* ws, cudaStream_t st) {
    if (n_arrays < 1 || n_arrays > kMaxAssign || !ws) return BRS_ERR_INVALID_ARG;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    long long nmax = 0;
    for (int k = 0; k < n_arrays; ++k) {
        if (!rs[k].slot_map || !rs[k].list || !rs[k].count || !idx[k] || n[k] < 0) return BRS_ERR_INVALID_ARG;
        a.rs[k] = rs[k];
        a.idx[k] = idx[k];
        a.n[k] = n[k];
        if (n[k] > nmax) nmax = n[k];
    }
    a.n_arrays = n_arrays;
    a.err_flag = &ws->err_flag;
    if (nmax == 0) return BRS_OK;
    long long blocks = (nmax + kThreads - 1) / kThreads;
    const long long cap = (long long)brs_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    assign_slots_kernel<<<dim3((unsigned)blocks, (unsigned)n_arrays), kThreads, 0, st>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// shared with abi.cu: apply `opt` to entities + dense params (+ finalize through ws)
int brs_apply_impl_next(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                        int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                        long long batch, long long max_rows_hint, void* stream, const brs_rowset* next_rs,
                        const long long* const* next_idx, const long long* next_n, int next_arrays, int parity);

int brs_apply_impl(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                   int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                   long long batch, long long max_rows_hint, void* stream) {
    return brs_apply_impl_next(ents, n_ent, dense, n_dense, dense_grad_from_ws, opt, ws, t_explicit, out, batch,
                               max_rows_hint, stream, nullptr, nullptr, nullptr, 0, 0);
}

// as brs_apply_impl; with next_arrays > 0 (touched-rows optimizers only) the same launch also runs the slot
// pre-pass of the NEXT batch on the rowsets next_rs[k] (which must differ from the ones being applied)
int brs_apply_impl_next(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                        int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                        long long batch, long long max_rows_hint, void* stream, const brs_rowset* next_rs,
                        const long long* const* next_idx, const long long* next_n, int next_arrays, int parity) {
    if (!opt) return BRS_ERR_INVALID_ARG;
    int rc = validate_entities(ents, n_ent, opt->kind);
    if (rc != BRS_OK) return rc;
    rc = validate_dense(dense, n_dense, opt->kind, dense_grad_from_ws != 0);
    if (rc != BRS_OK) return rc;
    ApplyArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < n_ent; ++e) a.ent[e] = ents[e];
    a.n_ent = n_ent;
    for (int k = 0; k < n_dense; ++k) a.dense[k] = dense[k];
    a.n_dense = n_dense;
    a.dense_grad_from_ws = dense_grad_from_ws;
    fill_opt(a, opt);
    a.ws = (brs_step_ws*)ws;
    a.t_explicit = t_explicit;
    a.out = out;
    a.inv_batch = batch > 0 ? 1.0 / (double)batch : 0.0;
    a.advance_step = ws ? 1 : 0;
    a.pol_scratch = brs_l2_cfg().scratch;
    a.pol_weight = brs_l2_cfg().weight;
    if (next_arrays > 0) {
        if (next_arrays > kMaxAssign || !ws || !next_rs || !next_idx || !next_n) return BRS_ERR_INVALID_ARG;
        if (opt->kind != BRS_SGD && opt->mode == BRS_DENSE) return BRS_ERR_INVALID_ARG;  // the sweep reads slot maps
        for (int k = 0; k < next_arrays; ++k) {
            if (!next_rs[k].slot_map || !next_rs[k].list || !next_rs[k].count || !next_idx[k] || next_n[k] < 0)
                return BRS_ERR_INVALID_ARG;
            a.next.rs[k] = next_rs[k];
            a.next.idx[k] = next_idx[k];
            a.next.n[k] = next_n[k];
        }
        a.next.n_arrays = next_arrays;
        a.next.err_flag = &((brs_step_ws*)ws)->err_pending[(parity + 1) & 1];
    }
    a.parity = parity & 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (opt->kind) {
        case BRS_SGD: return launch_apply<BRS_SGD>(a, BRS_TOUCHED_ROWS, max_rows_hint, st);
        case BRS_ADAM: return launch_apply<BRS_ADAM>(a, opt->mode, max_rows_hint, st);
        case BRS_RMSPROP: return launch_apply<BRS_RMSPROP>(a, opt->mode, max_rows_hint, st);
        default: return BRS_ERR_UNSUPPORTED;
    }
}

// row-owner MF step, BRS_DENSE Adam / RMSprop: move every row that is NOT in the batch with g = 0 (the
// reference's dense optimizers do), finalise the step through ws, then release the rowsets
int brs_dense_sweep_untouched(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                              int dense_grad_from_ws, const brs_opt* opt, void* ws, float* out, long long batch,
                              int parity, void* stream) {
    if (!opt || !ws || opt->kind == BRS_SGD) return BRS_ERR_INVALID_ARG;
    int rc = validate_entities(ents, n_ent, opt->kind);
    if (rc != BRS_OK) return rc;
    rc = validate_dense(dense, n_dense, opt->kind, dense_grad_from_ws != 0);
    if (rc != BRS_OK) return rc;
    ApplyArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < n_ent; ++e) a.ent[e] = ents[e];
    a.n_ent = n_ent;
    for (int k = 0; k < n_dense; ++k) a.dense[k] = dense[k];
    a.n_dense = n_dense;
    a.dense_grad_from_ws = dense_grad_from_ws;
    fill_opt(a, opt);
    a.ws = (brs_step_ws*)ws;
    a.out = out;
    a.inv_batch = batch > 0 ? 1.0 / (double)batch : 0.0;
    a.advance_step = 1;
    a.parity = parity & 1;
    a.skip_touched = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (opt->kind == BRS_ADAM) return launch_apply<BRS_ADAM>(a, BRS_DENSE, 0, st);
    return launch_apply<BRS_RMSPROP>(a, BRS_DENSE, 0, st);
}

extern "C" int brs_rows_sgd(const brs_entity* entities, int32_t n_entities, double lr, void* stream) {
    brs_opt o;
    memset(&o, 0, sizeof(o));
    o.kind = BRS_SGD;
    o.lr = lr;
    long long hint = 0;
    for (int e = 0; entities && e < n_entities && e < kMaxEntities; ++e) hint += entities[e].rows.capacity;
    return brs_apply_impl(entities, n_entities, nullptr, 0, 0, &o, nullptr, 1, nullptr, 0, hint, stream);
}

extern "C" int brs_rows_adam(const brs_entity* entities, int32_t n_entities, const brs_opt* opt, int64_t t,
                             void* stream) {
    if (!opt || t < 1) return BRS_ERR_INVALID_ARG;
    brs_opt o = *opt;
    o.mode = BRS_TOUCHED_ROWS;
    long long hint = 0;
    for (int e = 0; entities && e < n_entities && e < kMaxEntities; ++e) hint += entities[e].rows.capacity;
    return brs_apply_impl(entities, n_entities, nullptr, 0, 0, &o, nullptr, t, nullptr, 0, hint, stream);
}

extern "C" int brs_dense_adam_sweep(const brs_entity* entities, int32_t n_entities, const brs_opt* opt, int64_t t,
                                    void* stream) {
    if (!opt || t < 1 || opt->kind == BRS_SGD) return BRS_ERR_INVALID_ARG;
    brs_opt o = *opt;
    o.mode = BRS_DENSE;
    return brs_apply_impl(entities, n_entities, nullptr, 0, 0, &o, nullptr, t, nullptr, 0, 0, stream);
}

extern "C" int brs_dense_params_step(const brs_dense_param* params, int32_t n_params, const brs_opt* opt, int64_t t,
                                     void* stream) {
    if (!opt || t < 1) return BRS_ERR_INVALID_ARG;
    brs_opt o = *opt;
    o.mode = BRS_TOUCHED_ROWS;
    return brs_apply_impl(nullptr, 0, params, n_params, 0, &o, nullptr, t, nullptr, 0, 0, stream);
}

extern "C" int brs_mf_sharded_apply(const brs_mf_sharded* model, const brs_opt* opt, int64_t global_batch, float* out,
                                    void* stream) {
    if (!model || !opt || !model->stage.ws || global_batch <= 0) return BRS_ERR_INVALID_ARG;
    const brs_mf_model& sm = model->stage;
    const int D = sm.user.table[0].dim;
    if (D <= 0 || (D & 3) != 0) return BRS_ERR_UNSUPPORTED;
    ShardApplyArgs a;
    memset(&a, 0, sizeof(a));
    const brs_entity* ents[2] = {&sm.user, &sm.item};
    const brs_mf_peer_tables& own = model->own;
    float* g[2] = {own.g_user_emb, own.g_item_emb};
    float* gb[2] = {own.g_user_bias, own.g_item_bias};
    unsigned int* bits[2] = {own.user_bits, own.item_bits};
    const long long rows[2] = {model->local_users, model->local_items};
    for (int e = 0; e < 2; ++e) {
        const brs_table& te = ents[e]->table[0];
        const brs_table& tb = ents[e]->table[1];
        if (!te.weight || !tb.weight || !g[e] || !gb[e] || !bits[e]) return BRS_ERR_INVALID_ARG;
        if (opt->kind == BRS_ADAM && (!te.m || !te.v || !tb.m || !tb.v)) return BRS_ERR_INVALID_ARG;
        if (opt->kind == BRS_RMSPROP && (!te.v || !tb.v)) return BRS_ERR_INVALID_ARG;
        a.ent[e] = ShardEnt{te.weight, tb.weight, g[e], gb[e], te.m, te.v, tb.m, tb.v, bits[e], rows[e]};
    }
    a.dim = D;
    a.dense_all = (opt->kind != BRS_SGD && opt->mode == BRS_DENSE) ? 1 : 0;
    int rc = validate_dense(&sm.global_bias, 1, opt->kind, true);
    if (rc != BRS_OK) return rc;
    a.base.dense[0] = sm.global_bias;
    a.base.n_dense = 1;
    a.base.dense_grad_from_ws = 1;
    fill_opt(a.base, opt);
    a.base.ws = (brs_step_ws*)sm.ws;
    a.base.out = out;
    a.base.inv_batch = 1.0 / (double)global_batch;
    a.base.advance_step = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (opt->kind == BRS_SGD) {  // brs_mf_sharded_push already applied -lr * g at the owners: only the
        a.ent[0].rows = 0;       // replicated global bias and the step record are left
        a.ent[1].rows = 0;
    }
    const long long tasks = ((a.ent[0].rows + 7) >> 3) + ((a.ent[1].rows + 7) >> 3);
#define BRS_SHARD_APPLY(KIND)                                                         \
    do {                                                                              \
        auto k = shard_apply_kernel<KIND>;                                            \
        long long grid = persistent_grid((const void*)k);                             \
        const long long need = (tasks + kWarps - 1) / kWarps;                         \
        if (grid > need) grid = need;                                                 \
        k<<<(int)(grid < 1 ? 1 : grid), kThreads, 0, st>>>(a);                        \
    } while (0)
    switch (opt->kind) {
        case BRS_SGD: BRS_SHARD_APPLY(BRS_SGD); break;
        case BRS_ADAM: BRS_SHARD_APPLY(BRS_ADAM); break;
        case BRS_RMSPROP: BRS_SHARD_APPLY(BRS_RMSPROP); break;
        default: return BRS_ERR_UNSUPPORTED;
    }
#undef BRS_SHARD_APPLY
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
