// MF row-owner training step -- sm_100a.  (include/brs_b200.h: brs_mf_plan_build / brs_mf_step_planned)
//
// Replaces, per batch, MF.forward x2 + bpr_loss/bce_loss + loss.backward() + optimizer.step()
// (beta_rec/models/mf.py:92-119, torch_engine.py:23-39,92-121) with a schedule in which every
// touched table row is READ ONCE AND WRITTEN ONCE by the lane group that owns it:
//
//   plan   index-only (no table access, runs ahead on a side stream):
//            claim    one thread per sample: range check, one slot per unique row (atomicCAS on the
//                     rowset's slot map), rank of the sample among the samples of its user / item rows
//            segment  one thread per slot: a contiguous segment of the entity's stream per unique row
//                     (block sum + one atomicAdd on the stream cursor; the order of segments is free)
//            fill     one thread per sample: its record in the USER stream (samples grouped by user row)
//                     and its one or two entries in the ITEM stream (entries grouped by item row)
//   users  one warp per work unit of the user stream (about 16 samples; a unit never cuts a short row),
//          rows in registers: per block of 4 samples the rows are gathered with 128-bit read-only loads
//          (one warp instruction per 512-byte row at dim 128, the whole block in flight at once), the
//          block's 8 dots are reduced by ONE transposing butterfly so that the sigmoid / loss chain runs
//          once per block, the user-row gradient is accumulated in registers while the user stays the
//          same; at the end of a row: PRE-step row -> staging table, updated row -> table, in place; per
//          sample (coefficient, user slot) -> the item stream
//   items  same decomposition of the item stream -- a row-per-warp SpMM: sum of coefficient * staged user
//          row in registers, updated item row in place; last block: global-bias step + brs_step_out
//
// Batch-synchronous semantics hold because item rows are only written by `items` (after every gather
// of `users` has completed: kernel boundary), `items` reads user rows only from the staging copy, and a
// user row is written only after all of its samples were read.  Rows longer than a unit (the Zipf head)
// are cut at unit boundaries: the parts add their partial sums into the row-major gradient scratch with
// 128-bit REDs and the last part to arrive (ticket) applies the update.  Round 1 did one 512-byte RED per
// sample-row (196 608 per batch at config 2; RED issue rate was the limiter) plus a second pass over the
// touched rows and 2.9x the compulsory DRAM traffic.
#include <string.h>

#include "common.cuh"
#include "mf_math.cuh"
#include "opt_math.cuh"

int brs_dense_sweep_untouched(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                              int dense_grad_from_ws, const brs_opt* opt, void* ws, float* out, long long batch,
                              int parity, void* stream);

namespace {

constexpr int kPlanThreads = 256;
constexpr int kPlanWarps = kPlanThreads / 32;
#define BRS_SLOT_OVERFLOW (-3)


// ---------------------------------------------------------------------------
// plan buffer carve-up (host and device agree through this one function)
// ---------------------------------------------------------------------------
struct PlanView {
    int* hdr;        // [64]: 0 = samples in the user stream, 1 = entries in the item stream (segment cursors),
    int* u_slot;     // [B]   user slot of sample s (-1: sample dropped)
    int* u_rank;     // [B]   rank of s inside its user segment
    int* i_slot;     // [2B]  c*B + s
    int* i_rank;     // [2B]
    int* u_cnt;      // [Cu]  segment sizes (zero between plans)
    int* i_cnt;      // [Ci]
    int2* u_seg;     // [Cu]  {begin, end} of the slot's segment in the user stream
    int2* i_seg;     // [Ci]
    int* u_ticket;   // [Cu]  parts of a multi-part row that have finished (zero between steps)
    int* i_ticket;   // [Ci]
    int4* s_a;       // [B]   user stream position p -> {user row, user slot, pos item, neg item | rating bits}
    int4* s_b;       // [B]   p -> {item-stream position of the pos entry, of the neg entry, segment begin, end}
    int4* i_a;       // [2B]  item stream position q -> {item row, item slot, segment begin, end}
    float2* ipair;   // [2B]  q -> {coefficient, user slot bits}   (written by the users kernel)
    size_t bytes;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline PlanView plan_view(void* buf, long long B, int Cu, int Ci) {
    PlanView v;
    char* p = (char*)buf;
    size_t o = 0;
#define BRS_CARVE(field, type, count)            \
    v.field = (type*)(p + o);                    \
    o = align256(o + sizeof(type) * (size_t)(count));
    BRS_CARVE(hdr, int, 64)
    BRS_CARVE(u_slot, int, B)
    BRS_CARVE(u_rank, int, B)
    BRS_CARVE(i_slot, int, 2 * B)
    BRS_CARVE(i_rank, int, 2 * B)
    BRS_CARVE(u_cnt, int, Cu)
    BRS_CARVE(i_cnt, int, Ci)
    BRS_CARVE(u_seg, int2, Cu)
    BRS_CARVE(i_seg, int2, Ci)
    BRS_CARVE(u_ticket, int, Cu)
    BRS_CARVE(i_ticket, int, Ci)
    BRS_CARVE(s_a, int4, B)
    BRS_CARVE(s_b, int4, B)
    BRS_CARVE(i_a, int4, 2 * B)
    BRS_CARVE(ipair, float2, 2 * B)
#undef BRS_CARVE
    v.bytes = o;
    return v;
}

// ---------------------------------------------------------------------------
// plan kernels
// ---------------------------------------------------------------------------
struct PlanArgs {
    PlanView pv;
    brs_rowset urs, irs;
    const long long* users;
    const long long* items;
    const void* third;  // neg ids (int64) or ratings (float)
    long long batch;
    int n_cols;         // 2 = bpr (pos, neg), 1 = bce
    unsigned int* err;  // ws->err_pending[which]
};

__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *((const volatile int*)p); }

__device__ __forceinline__ bool claim_row(int* m) {
    // cheap read first: hot (Zipf) rows are claimed by the time most samples arrive
    if (ld_volatile_i32(m) != BRS_SLOT_NONE) return false;
    return atomicCAS(m, BRS_SLOT_NONE, BRS_SLOT_PENDING) == BRS_SLOT_NONE;
}

// the winner of the claim is past its CAS (resident and running) and publishes the slot before it
// waits for anything itself, so this spin cannot deadlock
__device__ __forceinline__ int wait_slot(const int* m) {
    int v;
    do {
        v = ld_volatile_i32(m);
    } while (v == BRS_SLOT_PENDING);
    return v;
}

// claim: one thread per sample -- range check, one slot per unique row, rank of the sample in its segments
__global__ void __launch_bounds__(kPlanThreads) mf_plan_claim_kernel(const PlanArgs a) {
    __shared__ int s_wu[kPlanWarps], s_wi[kPlanWarps];
    __shared__ int s_bu, s_bi;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long B = a.batch;
    const long long stride = (long long)gridDim.x * kPlanThreads;
    const long long n_iter = (B + stride - 1) / stride;
    const bool two = a.n_cols == 2;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // segment cursors of this plan (consumed by the next kernel)
        a.pv.hdr[0] = 0;
        a.pv.hdr[1] = 0;
    }
    for (long long it = 0; it < n_iter; ++it) {
        const long long s = it * stride + (long long)blockIdx.x * kPlanThreads + threadIdx.x;
        long long u = 0, i = 0, j = 0;
        bool valid = false;
        if (s < B) {
            u = a.users[s];
            i = a.items[s];
            j = two ? ((const long long*)a.third)[s] : 0;
            valid = (unsigned long long)u < (unsigned long long)a.urs.n_rows &&
                    (unsigned long long)i < (unsigned long long)a.irs.n_rows &&
                    (unsigned long long)j < (unsigned long long)a.irs.n_rows;
            if (!valid) atomicOr(a.err, 1u);  // the reference raises IndexError (nn.Embedding)
        }
        bool wu = false, wi = false, wj = false;
        if (valid) {
            wu = claim_row(a.urs.slot_map + u);
            wi = claim_row(a.irs.slot_map + i);
            if (two) wj = claim_row(a.irs.slot_map + j);  // j == i: already pending, not claimed twice
        }
        // the block's claims leave as ONE atomicAdd per rowset counter
        const unsigned bu = __ballot_sync(BRS_FULL_MASK, wu), bi = __ballot_sync(BRS_FULL_MASK, wi),
                       bj = __ballot_sync(BRS_FULL_MASK, wj);
        const unsigned lt = (1u << lane) - 1u;
        const int ou = __popc(bu & lt), oi = __popc(bi & lt), oj = __popc(bi) + __popc(bj & lt);
        if (lane == 0) {
            s_wu[warp] = __popc(bu);
            s_wi[warp] = __popc(bi) + __popc(bj);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tu = 0, ti = 0;
#pragma unroll
            for (int w = 0; w < kPlanWarps; ++w) {
                const int cu = s_wu[w], ci = s_wi[w];
                s_wu[w] = tu;
                s_wi[w] = ti;
                tu += cu;
                ti += ci;
            }
            s_bu = tu ? atomicAdd(a.urs.count, tu) : 0;
            s_bi = ti ? atomicAdd(a.irs.count, ti) : 0;
        }
        __syncthreads();
        int su = -1, si = -1, sj = -1;
        if (wu) {
            su = s_bu + s_wu[warp] + ou;
            if (su < a.urs.capacity) a.urs.list[su] = (int)u; else { su = BRS_SLOT_OVERFLOW; atomicOr(a.err, 2u); }
            atomicExch(a.urs.slot_map + u, su);
        }
        if (wi) {
            si = s_bi + s_wi[warp] + oi;
            if (si < a.irs.capacity) a.irs.list[si] = (int)i; else { si = BRS_SLOT_OVERFLOW; atomicOr(a.err, 2u); }
            atomicExch(a.irs.slot_map + i, si);
        }
        if (wj) {
            sj = s_bi + s_wi[warp] + oj;
            if (sj < a.irs.capacity) a.irs.list[sj] = (int)j; else { sj = BRS_SLOT_OVERFLOW; atomicOr(a.err, 2u); }
            atomicExch(a.irs.slot_map + j, sj);
        }
        __syncthreads();  // s_w* / s_b* are reused by the next iteration
        if (valid) {
            if (!wu) su = wait_slot(a.urs.slot_map + u);
            if (!wi) si = wait_slot(a.irs.slot_map + i);
            if (two && !wj) sj = wait_slot(a.irs.slot_map + j);
            if (su < 0 || si < 0 || (two && sj < 0)) valid = false;
        }
        // ranks inside the row's group: the lanes of a warp that share a slot take consecutive ranks from ONE
        // atomicAdd (under Zipf ids the hottest user / item is ~10% of the batch: its counter was a serial chain
        // of ~6-7 k same-address atomics, now ~3x shorter)
        auto ranks = [&](int* cnt, int slot, bool active) -> int {
            const int key = active ? slot : (-1 - lane);  // inactive lanes match nobody
            const unsigned peers = __match_any_sync(BRS_FULL_MASK, key);
            const int leader = __ffs(peers) - 1;
            int first = 0;
            if (active && lane == leader) first = atomicAdd(cnt + slot, __popc(peers));
            first = __shfl_sync(BRS_FULL_MASK, first, leader);
            return first + __popc(peers & ((1u << lane) - 1u));
        };
        const int ru_ = ranks(a.pv.u_cnt, su, valid);
        const int ri_ = ranks(a.pv.i_cnt, si, valid);
        // the neg column after the pos column of the whole warp: a lane with pos == neg must not share one add
        const int rj_ = ranks(a.pv.i_cnt, sj, valid && two);
        if (s < B) {
            if (valid) {
                a.pv.u_slot[s] = su;
                a.pv.u_rank[s] = ru_;
                a.pv.i_slot[s] = si;
                a.pv.i_rank[s] = ri_;
                if (two) {
                    a.pv.i_slot[B + s] = sj;
                    a.pv.i_rank[B + s] = rj_;
                }
            } else {
                a.pv.u_slot[s] = -1;  // dropped (the step is void anyway: status 1 / 2)
            }
        }
    }
}

// segments: every slot gets a contiguous range of its entity's stream.  The ORDER of the segments is
// irrelevant (only contiguity matters), so no scan: a block sums its 256 sizes and takes its range with
// one atomicAdd on the stream cursor.  Clears the sizes for the next plan that uses this buffer.
__global__ void __launch_bounds__(kPlanThreads) mf_plan_segment_kernel(const PlanArgs a) {
    __shared__ int s_w[kPlanWarps];
    __shared__ int s_base;
    const int ent = blockIdx.y;
    const brs_rowset& rs = ent == 0 ? a.urs : a.irs;
    int* cnt = ent == 0 ? a.pv.u_cnt : a.pv.i_cnt;
    int2* seg = ent == 0 ? a.pv.u_seg : a.pv.i_seg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n = *rs.count;
    if (n > rs.capacity) n = rs.capacity;
    for (int k0 = blockIdx.x * kPlanThreads; k0 < n; k0 += gridDim.x * kPlanThreads) {  // block-uniform
        const int k = k0 + threadIdx.x;
        const int c = k < n ? cnt[k] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(BRS_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
#pragma unroll
            for (int w = 0; w < kPlanWarps; ++w) {
                const int t = s_w[w];
                s_w[w] = run;
                run += t;
            }
            s_base = run ? atomicAdd(a.pv.hdr + ent, run) : 0;
        }
        __syncthreads();
        if (k < n) {
            const int begin = s_base + s_w[warp] + incl - c;
            seg[k] = make_int2(begin, begin + c);
            cnt[k] = 0;
        }
        __syncthreads();
    }
}

// fill: one thread per sample -- its record in the user stream and its entries in the item stream
__global__ void __launch_bounds__(kPlanThreads) mf_plan_fill_kernel(const PlanArgs a) {
    const long long B = a.batch;
    const bool two = a.n_cols == 2;
    for (long long s = (long long)blockIdx.x * kPlanThreads + threadIdx.x; s < B; s += (long long)gridDim.x * kPlanThreads) {
        const int su = a.pv.u_slot[s];
        if (su < 0) continue;
        const int2 us = a.pv.u_seg[su];
        const int p = us.x + a.pv.u_rank[s];
        const int u = (int)a.users[s];
        const int i = (int)a.items[s];
        int second;
        if (two)
            second = (int)((const long long*)a.third)[s];
        else
            second = __float_as_int(((const float*)a.third)[s]);
        const int si = a.pv.i_slot[s];
        const int2 is = a.pv.i_seg[si];
        const int qi = is.x + a.pv.i_rank[s];
        a.pv.i_a[qi] = make_int4(i, si, is.x, is.y);
        int qj = -1;
        if (two) {
            const int sj = a.pv.i_slot[B + s];
            const int2 js = a.pv.i_seg[sj];
            qj = js.x + a.pv.i_rank[B + s];
            a.pv.i_a[qj] = make_int4(second, sj, js.x, js.y);
        }
        a.pv.s_a[p] = make_int4(u, su, i, second);
        a.pv.s_b[p] = make_int4(qi, qj, us.x, us.y);
    }
}

// Work units.  A stream [0, n) is nominally cut every kUnit positions.  A nominal boundary that falls inside
// a SHORT segment (<= kUnit positions: almost every row) moves to that segment's end, so the row is owned by
// exactly one unit and never takes the RED + ticket path; inside a LONG segment (the Zipf head) it stays, so
// a long row is cut at every multiple of kUnit.  Unit k is therefore a sub-range of the positions
// [k*kUnit, k*kUnit + kTile): its records sit at a STATIC address (no cut table, no dependent load) and its
// two ends follow from the records at tile offsets 0 and kUnit.
#ifndef BRS_ROWS_UNIT_SHIFT
#define BRS_ROWS_UNIT_SHIFT 4
#endif
constexpr int kUnitShift = BRS_ROWS_UNIT_SHIFT;
constexpr int kUnit = 1 << kUnitShift;
constexpr int kTile = 2 * kUnit;  // one record per lane
static_assert(kTile <= 32, "a tile is copied one record per lane");

__device__ __forceinline__ int snap_cut(int pos, int sb, int se) {
    if (pos <= sb) return pos;          // already a row boundary
    if (se - sb <= kUnit) return se;    // short row: all of it goes to the earlier unit
    return pos;                         // long row: cut here
}
// parts of the row whose segment is [sb, se)
__device__ __forceinline__ int row_parts(int sb, int se) {
    if (se - sb <= kUnit) return 1;
    return ((se - 1) >> kUnitShift) - (sb >> kUnitShift) + 1;
}

// ---------------------------------------------------------------------------
// row-owner kernels
// ---------------------------------------------------------------------------
struct RowTable {
    float* w;   // [N, D] weights, updated in place
    float* m;   // Adam exp_avg
    float* v;   // Adam exp_avg_sq / RMSprop square_avg
    float* g;   // row-major [capacity][D] partial sums of multi-part rows (zero between steps)
};

struct RowArgs {
    RowTable ue, ub, ie, ib;  // user emb / user bias / item emb / item bias
    const float* global_bias;
    float* user_stage;        // [user capacity, D] PRE-step rows of the batch's users
    int* u_slot_map;
    int* i_slot_map;
    int* u_count;
    int* i_count;
    PlanView pv;
    const unsigned int* err;  // ws->err_pending[which]: non-zero => leave the parameters untouched
    brs_step_ws* ws;
    OptParams opt;
    int dim;
    float reg_w, inv_b;
    int release;   // release the rows' slots (0 when a dense sweep still needs the slot maps)
    // items kernel, last block
    int finalize;
    int parity;
    brs_dense_param gb;
    float* out;
    double inv_batch;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *(const float4*)p; }
__device__ __forceinline__ void st4(float* p, float4 v) { *(float4*)p = v; }
__device__ __forceinline__ float4 ld4_cg(const float* p) { return __ldcg((const float4*)p); }

// The last sample of a row part was accumulated by this warp (lane l holds columns 4*(v*32 + l) .. +3):
// rows with several parts combine their partial sums through the scratch and the last part to arrive
// continues; then the optimizer is applied to the row and its bias, the PRE-step row is staged (users) and
// the slot is released.  Warp-uniform control flow.
template <int VPL, bool FULL, int KIND>
__device__ __forceinline__ void flush_row(const RowTable& te, const RowTable& tb, int* ticket, int* slot_map,
                                          float* stage, int lane, int D, bool skip, int release,
                                          const OptScalars& os, int slot, int row, int nparts,
                                          const float4 (&w)[VPL], float4 (&acc)[VPL], float bias, float gbias) {
    const size_t so = (size_t)(unsigned)slot * (unsigned)D;
    if (nparts > 1) {
        if (!skip) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int col = (v * 32 + lane) * 4;
                if (FULL || col < D) red_add4(te.g + so + col, acc[v]);
            }
            if (lane == 0) red_add1(tb.g + slot, gbias);
            __threadfence();
        }
        __syncwarp();
        int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket + slot, 1);
        tk = __shfl_sync(BRS_FULL_MASK, tk, 0);
        if (tk != nparts - 1) return;  // somebody else finishes this row
        __threadfence();
        if (!skip) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int col = (v * 32 + lane) * 4;
                if (FULL || col < D) {
                    acc[v] = ld4_cg(te.g + so + col);
                    st4(te.g + so + col, f4_zero());
                }
            }
            if (lane == 0) {
                gbias = __ldcg(tb.g + slot);
                tb.g[slot] = 0.f;
            }
        }
        if (lane == 0) ticket[slot] = 0;
    }
    if (!skip) {
        const size_t ro = (size_t)(unsigned)row * (unsigned)D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * 32 + lane) * 4;
            if (FULL || col < D) {
                if (stage) st4(stage + so + col, w[v]);  // PRE-step copy for the items kernel
                float4 mv = f4_zero(), vv = f4_zero();
                if (KIND == BRS_ADAM) mv = ld4(te.m + ro + col);
                if (KIND != BRS_SGD) vv = ld4(te.v + ro + col);
                float4 nw = w[v];
                opt_elem4<KIND>(nw, acc[v], mv, vv, os);
                st4(te.w + ro + col, nw);
                if (KIND == BRS_ADAM) st4(te.m + ro + col, mv);
                if (KIND != BRS_SGD) st4(te.v + ro + col, vv);
            }
        }
        if (lane == 0) {
            float mb = (KIND == BRS_ADAM) ? tb.m[row] : 0.f;
            float vb = (KIND != BRS_SGD) ? tb.v[row] : 0.f;
            opt_elem<KIND>(bias, gbias, mb, vb, os);
            tb.w[row] = bias;
            if (KIND == BRS_ADAM) tb.m[row] = mb;
            if (KIND != BRS_SGD) tb.v[row] = vb;
        }
    }
    if (release && lane == 0) slot_map[row] = BRS_SLOT_NONE;
}

// Sum over the 32 lanes of NV per-lane values with NV - 1 + log2(32 / NV) shuffles instead of 5 * NV: each
// halving step exchanges half of the values a lane still carries.  Afterwards lane l holds the complete sum
// of value (l >> (5 - log2 NV)) & (NV - 1)  (NV = 8: index (l >> 2) & 7; NV = 4: index l >> 3).
template <int NV>
__device__ __forceinline__ float transpose_reduce(float (&v)[NV], int lane) {
    int o = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
            const float keep = hi ? v[k + n / 2] : v[k];
            const float send = hi ? v[k] : v[k + n / 2];
            v[k] = keep + __shfl_xor_sync(BRS_FULL_MASK, send, o);
        }
        o >>= 1;
    }
    float r = v[0];
    for (; o > 0; o >>= 1) r += __shfl_xor_sync(BRS_FULL_MASK, r, o);
    return r;
}


#ifndef BRS_ROW_WARPS
#define BRS_ROW_WARPS 4
#endif
constexpr int kRowWarps = BRS_ROW_WARPS;  // warps (= work units) per CTA of the row kernels
constexpr int kRowThreads = kRowWarps * 32;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg((const float4*)p); }

// users kernel: ONE WARP PER WORK UNIT of the user stream, everything in registers (the only shared memory
// is the unit's record tile); rows are gathered with plain 128-bit read-only loads -- one warp instruction
// per 512-byte row at dim 128, 12 in flight per warp, ~20 warps per SM -- so the Zipf-hot rows hit L1
// and nothing has to be staged.  Per block of 4 samples:
//   load     user, pos-item and neg-item rows of every sample (consecutive samples of a user hit L1); the three
//            biases of sample k by the 8 lanes that later run sample k's loss chain
//   phase 1  per-lane partial dots u.i, u.j
//   phase 2  ONE transposing reduction for the block's 8 dots, then the sigmoid / loss / d loss chain once
//            per block (lanes 8k .. 8k+7 work on sample k); (coefficient, user slot) -> item stream
//   phase 3  gradient of the user row accumulated in registers; at the end of a row part: regularizer term
//            of the row, PRE-step row -> staging table, updated row -> table
template <int VPL, bool FULL, int LOSS, int KIND>
__global__ void __launch_bounds__(kRowThreads) mf_user_rows_kernel(const RowArgs a) {
    constexpr int C = (LOSS == LOSS_BPR) ? 2 : 1;
    constexpr int K = 4, KSH = 3;
    __shared__ int4 s_tile[kRowWarps][kTile][2];  // [t][0] = s_a, [t][1] = s_b of stream position base + t
    __shared__ OptScalars s_opt;
    __shared__ float s_red[3][kRowWarps];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int D = a.dim;
    if (threadIdx.x == 0) s_opt = make_scalars(a.opt, a.ws->step + 1);
    __syncthreads();
    const OptScalars os = s_opt;
    const bool skip = __ldg(a.err) != 0u;
    const float bg = __ldg(a.global_bias);
    const int n = __ldg(a.pv.hdr + 0);
    const float rw = 2.0f * a.reg_w * a.inv_b;  // d(reg_w*regularizer)/d row = rw * row per forward call
    const float fwd_calls = (float)C;
    float loss_acc = 0.f, reg_acc = 0.f, gb_acc = 0.f;
    const int base = (blockIdx.x * kRowWarps + warp) << kUnitShift;
    if (base < n) {
        int4(*tl)[2] = s_tile[warp];
        {
            const int pos = base + lane;
            const bool ok = lane < kTile && pos < n;
            if (lane < kTile) {
                tl[lane][0] = ok ? __ldg(a.pv.s_a + pos) : make_int4(0, 0, 0, 0);
                tl[lane][1] = ok ? __ldg(a.pv.s_b + pos) : make_int4(0, 0, 0, 0);
            }
        }
        __syncwarp();
        const int lo = snap_cut(base, tl[0][1].z, tl[0][1].w);
        const int hi = base + kUnit >= n ? n : snap_cut(base + kUnit, tl[kUnit][1].z, tl[kUnit][1].w);
        // per-lane constants of the gathers
        const float* ue_l = a.ue.w + lane * 4;
        const float* ie_l = a.ie.w + lane * 4;
        bool colv[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) colv[v] = FULL || (v * 32 + lane) * 4 < D;
        const int kl = lane >> KSH;  // the sample of a block whose scalar work (biases, loss chain) this lane does
        // state of the row part being accumulated
        float4 ru[VPL], acc[VPL];
        float uu = 0.f, bu = 0.f, gbias = 0.f;
        int n_row = 0;
        bool fresh = true;  // warp-uniform: the next sample is the first of a row part (unit start / after a flush)
#pragma unroll
        for (int v = 0; v < VPL; ++v) ru[v] = acc[v] = f4_zero();

        for (int p0 = lo; p0 < hi; p0 += K) {  // warp-uniform
            const int cnt = min(K, hi - p0);
            const int t0 = p0 - base;
            // ---- load: every row of the block is in flight before the first one is used.  The user row is
            // loaded with EVERY sample (consecutive samples of a user hit L1): no per-sample "row changed?"
            // register shuffling in the dot phase
            float4 ir[K][VPL], jr[C == 2 ? K : 1][VPL], ur[K][VPL];
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    ir[k][v] = ur[k][v] = f4_zero();
                    if (C == 2) jr[k][v] = f4_zero();
                }
                if (k < cnt) {  // warp-uniform
                    const int4 ra = tl[t0 + k][0];
                    const float* ip = ie_l + (size_t)(unsigned)ra.z * (unsigned)D;
                    const float* jp = ie_l + (size_t)(unsigned)(C == 2 ? ra.w : 0) * (unsigned)D;
                    const float* up = ue_l + (size_t)(unsigned)ra.x * (unsigned)D;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (colv[v]) {
                            ur[k][v] = ldg4(up + v * 128);
                            ir[k][v] = ldg4(ip + v * 128);
                            if (C == 2) jr[k][v] = ldg4(jp + v * 128);
                        }
                    }
                }
            }
            // the lane group of sample kl loads that sample's three biases (one address per group)
            const bool onl = kl < cnt;
            const int4 la = tl[t0 + (onl ? kl : 0)][0], lb = tl[t0 + (onl ? kl : 0)][1];
            const float b_u = __ldg(a.ub.w + (unsigned)la.x);
            const float b_i = __ldg(a.ib.w + (unsigned)la.z);
            const float b_j = (C == 2) ? __ldg(a.ib.w + (unsigned)la.w) : 0.f;
            // ---- phase 1: partial dots
            float dots[K * C];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float dp = 0.f, dn = 0.f;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    dp += f4_dot(ur[k][v], ir[k][v]);
                    if (C == 2) dn += f4_dot(ur[k][v], jr[k][v]);
                }
                dots[C * k] = dp;
                if (C == 2) dots[C * k + 1] = dn;
            }
            // ---- phase 2: one reduction, one loss chain per block; lanes [k << KSH, (k+1) << KSH) work on sample k
            float z = transpose_reduce<K * C>(dots, lane);
            float zp = z, zn = 0.f;
            if (C == 2) {
                const float other = __shfl_xor_sync(BRS_FULL_MASK, z, 1 << (KSH - 1));
                const bool odd = (lane & (1 << (KSH - 1))) != 0;
                zp = odd ? other : z;
                zn = odd ? z : other;
            }
            zp += b_u + b_i + bg;
            zn += b_u + b_j + bg;
            const float rating = (C == 1) ? __int_as_float(la.w) : 0.f;
            float cu_i, cu_j, loss_k;
            mf_sample_coef_fast<LOSS>(zp, zn, rating, a.inv_b, cu_i, cu_j, loss_k);
            if (onl && (lane & ((1 << KSH) - 1)) == 0) {
                loss_acc += loss_k;
                gb_acc += cu_i + cu_j;
                const float sf = __int_as_float(la.y);  // hand the coefficients (+ user slot) to the item stream
                a.pv.ipair[lb.x] = make_float2(cu_i, sf);
                if (C == 2) a.pv.ipair[lb.y] = make_float2(cu_j, sf);
            }
            // ---- phase 3: gradient of the user row
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < cnt) {  // warp-uniform
                    const float ci = __shfl_sync(BRS_FULL_MASK, cu_i, k << KSH);
                    const float cj = (C == 2) ? __shfl_sync(BRS_FULL_MASK, cu_j, k << KSH) : 0.f;
                    const int p = p0 + k;
                    const int4 rb = tl[t0 + k][1];
                    if (fresh) {  // a new row part starts here: its PRE-step weights stay in registers
                        fresh = false;
                        uu = 0.f;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) {
                            ru[v] = ur[k][v];
                            uu += f4_dot(ru[v], ru[v]);
                            acc[v] = f4_zero();
                        }
                        bu = __shfl_sync(BRS_FULL_MASK, b_u, k << KSH);
                        gbias = 0.f;
                        n_row = 0;
                    }
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        acc[v] = f4_fma(ci, ir[k][v], acc[v]);
                        if (C == 2) acc[v] = f4_fma(cj, jr[k][v], acc[v]);
                    }
                    gbias += ci + cj;
                    n_row += 1;
                    if (p + 1 == rb.w || p + 1 == hi) {  // last sample of this row inside my unit
                        const int4 ra = tl[t0 + k][0];
                        // regularizer numerator (mf.py:49-54): every forward call of every sample adds |u|^2 + b_u^2
                        reg_acc += fwd_calls * (float)n_row * (uu + (lane == 0 ? bu * bu : 0.f));
                        if (a.reg_w != 0.f) {
                            const float nl = fwd_calls * (float)n_row * rw;
#pragma unroll
                            for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(nl, ru[v], acc[v]);
                            gbias += nl * bu;
                        }
                        flush_row<VPL, FULL, KIND>(a.ue, a.ub, a.pv.u_ticket, a.u_slot_map, a.user_stage, lane, D, skip,
                                                   a.release, os, ra.y, ra.x, row_parts(rb.z, rb.w), ru, acc, bu, gbias);
                        fresh = true;  // the next sample starts a new row part
                    }
                }
            }
        }
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
    if (threadIdx.x == 0 && !skip) {
        float l = 0.f, r = 0.f, g = 0.f;
#pragma unroll
        for (int q = 0; q < kRowWarps; ++q) {
            l += s_red[0][q];
            r += s_red[1][q];
            g += s_red[2][q];
        }
        if (l != 0.f || r != 0.f || g != 0.f) {
            atomicAdd(&a.ws->loss_sum, (double)l);
            atomicAdd(&a.ws->reg_sum, (double)r);
            atomicAdd(&a.ws->g_global_bias, g);
        }
    }
}

// items kernel: one warp per work unit of the item stream -- a row-per-warp SpMM over the batch's
// (coefficient, user slot) entries: 8 staged PRE-step user rows in flight per warp, the item's gradient in
// registers; at the end of a row part the item's own row and bias are read, the regularizer term taken and
// the update applied in place.  The last block to finish applies the global-bias step and publishes
// brs_step_out.
template <int VPL, bool FULL, int KIND>
__global__ void __launch_bounds__(kRowThreads, 24 / kRowWarps) mf_item_rows_kernel(const RowArgs a) {
#ifndef BRS_ITEMS_K
#define BRS_ITEMS_K 4
#endif
    constexpr int K = BRS_ITEMS_K;  // entries per block
    __shared__ int4 s_ia[kRowWarps][kTile];    // i_a of stream position base + t
    __shared__ float2 s_pr[kRowWarps][kTile];  // {coefficient, user slot}
    __shared__ OptScalars s_opt;
    __shared__ float s_red[kRowWarps];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int D = a.dim;
    if (threadIdx.x == 0) s_opt = make_scalars(a.opt, a.ws->step + 1);
    __syncthreads();
    const OptScalars os = s_opt;
    const bool skip = __ldg(a.err) != 0u;
    const int n = __ldg(a.pv.hdr + 1);
    const float rw = 2.0f * a.reg_w * a.inv_b;
    float reg_acc = 0.f;
    const int base = (blockIdx.x * kRowWarps + warp) << kUnitShift;
    if (base < n) {
        int4* ta = s_ia[warp];
        float2* tp = s_pr[warp];
        {
            const int pos = base + lane;
            const bool ok = lane < kTile && pos < n;
            if (lane < kTile) {
                ta[lane] = ok ? __ldg(a.pv.i_a + pos) : make_int4(0, 0, 0, 0);
                tp[lane] = ok ? __ldcg(a.pv.ipair + pos) : make_float2(0.f, 0.f);
            }
        }
        __syncwarp();
        const int lo = snap_cut(base, ta[0].z, ta[0].w);
        const int hi = base + kUnit >= n ? n : snap_cut(base + kUnit, ta[kUnit].z, ta[kUnit].w);
        const float* us_l = a.user_stage + lane * 4;
        const float* ie_l = a.ie.w + lane * 4;
        bool colv[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) colv[v] = FULL || (v * 32 + lane) * 4 < D;
        float4 acc[VPL];
        float gbias = 0.f;
        int n_row = 0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = f4_zero();

        for (int q0 = lo; q0 < hi; q0 += K) {  // warp-uniform
            const int cnt = min(K, hi - q0);
            const int t0 = q0 - base;
            // every row of the block is in flight before the first one is used: the staged user row of each
            // entry and, where a row part ends, the item's own row and bias (no dependent load at the flush)
            float4 rr[K][VPL], wr[K][VPL];
            float wb[K], coef[K];
            int4 ia[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < cnt) {  // warp-uniform
                    const float2 pr = tp[t0 + k];
                    coef[k] = pr.x;
                    ia[k] = ta[t0 + k];
                    const float* up = us_l + (size_t)(unsigned)__float_as_int(pr.y) * (unsigned)D;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) rr[k][v] = colv[v] ? ldg4(up + v * 128) : f4_zero();
                    if (q0 + k + 1 == ia[k].w || q0 + k + 1 == hi) {  // warp-uniform: a row part ends with this entry
                        const float* ip = ie_l + (size_t)(unsigned)ia[k].x * (unsigned)D;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) wr[k][v] = colv[v] ? ld4(ip + v * 128) : f4_zero();
                        wb[k] = a.ib.w[(unsigned)ia[k].x];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < cnt) {  // warp-uniform
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(coef[k], rr[k][v], acc[v]);
                    gbias += coef[k];
                    n_row += 1;
                    if (q0 + k + 1 == ia[k].w || q0 + k + 1 == hi) {
                        float ww = 0.f;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) ww += f4_dot(wr[k][v], wr[k][v]);
                        const float bi = wb[k];
                        reg_acc += (float)n_row * (ww + (lane == 0 ? bi * bi : 0.f));  // mf.py:49-54, item side
                        if (a.reg_w != 0.f) {
                            const float nl = (float)n_row * rw;
#pragma unroll
                            for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(nl, wr[k][v], acc[v]);
                            gbias += nl * bi;
                        }
                        flush_row<VPL, FULL, KIND>(a.ie, a.ib, a.pv.i_ticket, a.i_slot_map, nullptr, lane, D, skip, a.release,
                                                   os, ia[k].y, ia[k].x, row_parts(ia[k].z, ia[k].w), wr[k], acc, bi, gbias);
                        // the next entry starts a new row part
#pragma unroll
                        for (int v = 0; v < VPL; ++v) acc[v] = f4_zero();
                        gbias = 0.f;
                        n_row = 0;
                    }
                }
            }
        }
    }
    reg_acc = warp_sum(reg_acc);
    if (lane == 0) s_red[warp] = reg_acc;
    __syncthreads();
    if (threadIdx.x == 0 && !skip) {
        float r = 0.f;
#pragma unroll
        for (int q = 0; q < kRowWarps; ++q) r += s_red[q];
        if (r != 0.f) atomicAdd(&a.ws->reg_sum, (double)r);
    }
    if (!a.finalize) return;
    // last block: global-bias step (its gradient was summed by the users kernel), publish, reset
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    brs_step_ws* ws = a.ws;
    const unsigned int status = *(volatile unsigned int*)a.err;
    if (status == 0u) {
        const float g = *(volatile float*)&ws->g_global_bias;
        float wv = a.gb.weight[0];
        float m = (KIND == BRS_ADAM) ? a.gb.m[0] : 0.f;
        float v = (KIND != BRS_SGD) ? a.gb.v[0] : 0.f;
        opt_elem<KIND>(wv, g, m, v, os);
        a.gb.weight[0] = wv;
        if (KIND == BRS_ADAM) a.gb.m[0] = m;
        if (KIND != BRS_SGD) a.gb.v[0] = v;
    }
    if (a.out) {
        a.out[0] = (float)(*(volatile double*)&ws->loss_sum * a.inv_batch);
        a.out[1] = (float)(*(volatile double*)&ws->reg_sum * a.inv_batch);
        ((int*)a.out)[2] = (int)status;  // brs_step_out.status
        a.out[3] = 0.f;
    }
    ws->err_pending[a.parity & 1] = 0u;
    ws->loss_sum = 0.0;
    ws->reg_sum = 0.0;
    ws->g_global_bias = 0.f;
    if (status == 0u) ws->step += 1;  // a void step does not advance the optimizer's step count
    ws->ticket = 0u;
    *a.u_count = 0;
    *a.i_count = 0;
}

int check_model(const brs_mf_model* m, int which) {
    if (!m || !m->ws || which < 0 || which > 1) return BRS_ERR_INVALID_ARG;
    if (m->user.n_tables < 2 || m->item.n_tables < 2) return BRS_ERR_INVALID_ARG;
    const brs_table& ue = m->user.table[0];
    const brs_table& ie = m->item.table[0];
    if (!ue.weight || !ie.weight || !m->user.table[1].weight || !m->item.table[1].weight || !m->global_bias.weight)
        return BRS_ERR_INVALID_ARG;
    if (ue.dim != ie.dim || m->user.table[1].dim != 1 || m->item.table[1].dim != 1) return BRS_ERR_INVALID_ARG;
    if ((((uintptr_t)ue.weight | (uintptr_t)ie.weight | (uintptr_t)m->user_stage) & 15) != 0) return BRS_ERR_INVALID_ARG;
    if (!m->plan[which].buf || !m->user_stage) return BRS_ERR_INVALID_ARG;
    const brs_rowset& ur = which ? m->user_rows_alt : m->user.rows;
    const brs_rowset& ir = which ? m->item_rows_alt : m->item.rows;
    if (!ur.slot_map || !ur.list || !ur.count || !ir.slot_map || !ir.list || !ir.count) return BRS_ERR_INVALID_ARG;
    const brs_mf_plan& pl = m->plan[which];
    if (pl.user_capacity != ur.capacity || pl.item_capacity != ir.capacity || pl.batch_capacity <= 0) return BRS_ERR_INVALID_ARG;
    if ((size_t)pl.bytes < plan_view(nullptr, pl.batch_capacity, pl.user_capacity, pl.item_capacity).bytes)
        return BRS_ERR_INVALID_ARG;
    if (ue.n_rows >= (1ll << 31) || ie.n_rows >= (1ll << 31)) return BRS_ERR_UNSUPPORTED;  // int32 row ids in the plan
    if (pl.batch_capacity >= (1ll << 29)) return BRS_ERR_UNSUPPORTED;
    return BRS_OK;
}


int g_rows_only = 0;  // diagnostics: 1 = users kernel only, 2 = items kernel only

template <class K>
int launch_units(K kernel, long long max_positions, const RowArgs& a, cudaStream_t st) {
    // the kernels keep their rows in registers and lean on L1 for the Zipf-hot rows: a carve-out that just holds
    // the record tiles of the CTAs the register file admits (<= 4 x 8.3 KB), the rest of the array stays L1
    static bool configured[16] = {};  // per kernel instantiation and device: keep driver calls off the per-step path
    int dev = 0;
    BRS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16 || !configured[dev]) {
        BRS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 20));
        if (dev >= 0 && dev < 16) configured[dev] = true;
    }
    const long long units = (max_positions + kUnit - 1) >> kUnitShift;
    long long blocks = (units + kRowWarps - 1) / kRowWarps;
    if (blocks < 1) blocks = 1;  // the items kernel's last block finalises the step even for an empty batch
    kernel<<<(unsigned)blocks, kRowThreads, 0, st>>>(a);
    return BRS_OK;
}

template <int LOSS, int KIND>
int launch_rows(const RowArgs& a, long long batch, cudaStream_t st) {
    const int D = a.dim;
    const long long n_u = batch, n_i = (LOSS == LOSS_BPR ? 2 : 1) * batch;  // upper bounds of the stream lengths
#define BRS_ROWS(VPL, FULL)                                                                      \
    do {                                                                                         \
        int rc_ = BRS_OK;                                                                        \
        if (g_rows_only != 2) rc_ = launch_units(mf_user_rows_kernel<VPL, FULL, LOSS, KIND>, n_u, a, st); \
        if (rc_ != BRS_OK) return rc_;                                                           \
        if (g_rows_only != 1) rc_ = launch_units(mf_item_rows_kernel<VPL, FULL, KIND>, n_i, a, st);       \
        if (rc_ != BRS_OK) return rc_;                                                           \
    } while (0)
    // a row of D floats = 32 lanes x VPL float4 (lanes past dim idle for dim < 128)
    if (D == 128) BRS_ROWS(1, true);
    else if (D < 128) BRS_ROWS(1, false);
    else if (D == 256) BRS_ROWS(2, true);
    else if (D < 256) BRS_ROWS(2, false);
    else if (D <= 384) BRS_ROWS(3, false);
    else BRS_ROWS(4, false);
#undef BRS_ROWS
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

RowTable row_table(const brs_table& t) { return RowTable{t.weight, t.m, t.v, t.grad}; }

}  // namespace

extern "C" int64_t brs_mf_plan_bytes(int64_t batch_capacity, int32_t user_capacity, int32_t item_capacity) {
    if (batch_capacity <= 0 || user_capacity <= 0 || item_capacity <= 0) return 0;
    return (int64_t)plan_view(nullptr, batch_capacity, user_capacity, item_capacity).bytes;
}

extern "C" int brs_debug_set_mf_rows_only(int which) {
    g_rows_only = which;
    return BRS_OK;
}

extern "C" int brs_mf_plan_build(const brs_mf_model* model, int32_t which, int32_t loss_kind, const int64_t* users,
                                 const int64_t* items, const void* third, int64_t batch, void* stream) {
    if (!users || !items || !third || batch < 0) return BRS_ERR_INVALID_ARG;
    if (loss_kind != LOSS_BPR && loss_kind != LOSS_BCE) return BRS_ERR_INVALID_ARG;
    int rc = check_model(model, which);
    if (rc != BRS_OK) return rc;
    const brs_mf_plan& pl = model->plan[which];
    if (batch > pl.batch_capacity) return BRS_ERR_INVALID_ARG;
    PlanArgs a;
    a.pv = plan_view(pl.buf, pl.batch_capacity, pl.user_capacity, pl.item_capacity);
    a.urs = which ? model->user_rows_alt : model->user.rows;
    a.irs = which ? model->item_rows_alt : model->item.rows;
    a.users = (const long long*)users;
    a.items = (const long long*)items;
    a.third = third;
    a.batch = batch;
    a.n_cols = loss_kind == LOSS_BPR ? 2 : 1;
    a.err = &((brs_step_ws*)model->ws)->err_pending[which];
    cudaStream_t st = (cudaStream_t)stream;
    if (batch > 0) {
        long long blocks = (batch + kPlanThreads - 1) / kPlanThreads;
        const long long cap = (long long)brs_sm_count() * 4;
        if (blocks > cap) blocks = cap;
        mf_plan_claim_kernel<<<(int)blocks, kPlanThreads, 0, st>>>(a);
    } else {
        // the claim kernel is what resets the stream cursors: an empty batch must not inherit the previous plan's
        BRS_CUDA_CHECK(cudaMemsetAsync(a.pv.hdr, 0, 2 * sizeof(int), st));
    }
    {
        const int cmax = pl.user_capacity > pl.item_capacity ? pl.user_capacity : pl.item_capacity;
        int blocks = (cmax + kPlanThreads - 1) / kPlanThreads;
        const int cap = brs_sm_count() * 2;
        if (blocks > cap) blocks = cap;
        mf_plan_segment_kernel<<<dim3((unsigned)blocks, 2u), kPlanThreads, 0, st>>>(a);
    }
    if (batch > 0) {
        long long blocks = (batch + kPlanThreads - 1) / kPlanThreads;
        const long long cap = (long long)brs_sm_count() * 4;
        if (blocks > cap) blocks = cap;
        mf_plan_fill_kernel<<<(int)blocks, kPlanThreads, 0, st>>>(a);
    }
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

extern "C" int brs_mf_step_planned(const brs_mf_model* model, int32_t which, const brs_opt* opt, int32_t loss_kind,
                                   int64_t batch, float reg_weight, float* out, void* stream) {
    if (!opt || batch < 0) return BRS_ERR_INVALID_ARG;
    if (loss_kind != LOSS_BPR && loss_kind != LOSS_BCE) return BRS_ERR_INVALID_ARG;
    int rc = check_model(model, which);
    if (rc != BRS_OK) return rc;
    const brs_table& ue = model->user.table[0];
    const int D = ue.dim;
    if (D % 4 != 0 || D <= 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    const brs_table* tabs[4] = {&model->user.table[0], &model->user.table[1], &model->item.table[0], &model->item.table[1]};
    for (const brs_table* t : tabs) {
        if (!t->grad) return BRS_ERR_INVALID_ARG;
        if (opt->kind == BRS_ADAM && (!t->m || !t->v)) return BRS_ERR_INVALID_ARG;
        if (opt->kind == BRS_RMSPROP && !t->v) return BRS_ERR_INVALID_ARG;
    }
    if (opt->kind == BRS_ADAM && (!model->global_bias.m || !model->global_bias.v)) return BRS_ERR_INVALID_ARG;
    if (opt->kind == BRS_RMSPROP && !model->global_bias.v) return BRS_ERR_INVALID_ARG;
    const brs_mf_plan& pl = model->plan[which];
    const brs_rowset& urs = which ? model->user_rows_alt : model->user.rows;
    const brs_rowset& irs = which ? model->item_rows_alt : model->item.rows;
    const bool dense = opt->kind != BRS_SGD && opt->mode == BRS_DENSE;
    RowArgs a;
    memset(&a, 0, sizeof(a));
    a.ue = row_table(model->user.table[0]);
    a.ub = row_table(model->user.table[1]);
    a.ie = row_table(model->item.table[0]);
    a.ib = row_table(model->item.table[1]);
    a.global_bias = model->global_bias.weight;
    a.user_stage = model->user_stage;
    a.u_slot_map = urs.slot_map;
    a.i_slot_map = irs.slot_map;
    a.u_count = urs.count;
    a.i_count = irs.count;
    a.pv = plan_view(pl.buf, pl.batch_capacity, pl.user_capacity, pl.item_capacity);
    a.ws = (brs_step_ws*)model->ws;
    a.err = &a.ws->err_pending[which];
    a.opt.kind = opt->kind;
    a.opt.lr = opt->lr;
    a.opt.beta1 = opt->beta1;
    a.opt.beta2 = opt->beta2;
    a.opt.eps = opt->eps;
    a.opt.alpha = opt->alpha;
    a.dim = D;
    a.reg_w = reg_weight;
    a.inv_b = batch > 0 ? 1.0f / (float)batch : 0.f;
    a.release = dense ? 0 : 1;
    a.finalize = dense ? 0 : 1;
    a.parity = which;
    a.gb = model->global_bias;
    a.out = out;
    a.inv_batch = batch > 0 ? 1.0 / (double)batch : 0.0;
    cudaStream_t st = (cudaStream_t)stream;
#define BRS_STEP(KIND)                                            \
    (loss_kind == LOSS_BPR ? launch_rows<LOSS_BPR, KIND>(a, batch, st) : launch_rows<LOSS_BCE, KIND>(a, batch, st))
    switch (opt->kind) {
        case BRS_SGD: rc = BRS_STEP(BRS_SGD); break;
        case BRS_ADAM: rc = BRS_STEP(BRS_ADAM); break;
        case BRS_RMSPROP: rc = BRS_STEP(BRS_RMSPROP); break;
        default: return BRS_ERR_UNSUPPORTED;
    }
#undef BRS_STEP
    if (rc != BRS_OK) return rc;
    if (dense) {
        // reference-exact Adam / RMSprop: every row NOT in the batch moves too (g = 0); the touched rows
        // (slot >= 0) were just updated by their owners.  The sweep's last block finalises the step.
        brs_entity ents[2] = {model->user, model->item};
        ents[0].rows = urs;
        ents[1].rows = irs;
        rc = brs_dense_sweep_untouched(ents, 2, &model->global_bias, 1, 1, opt, model->ws, out, batch, which, stream);
    }
    return rc;
}

extern "C" int brs_mf_step(const brs_mf_model* model, const brs_opt* opt, int32_t loss_kind, const int64_t* users,
                           const int64_t* items, const void* third, int64_t batch, float reg_weight, float* out,
                           void* stream) {
    int rc = brs_mf_plan_build(model, 0, loss_kind, users, items, third, batch, stream);
    if (rc != BRS_OK) return rc;
    return brs_mf_step_planned(model, 0, opt, loss_kind, batch, reg_weight, out, stream);
}
