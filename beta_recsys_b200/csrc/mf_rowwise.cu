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
//   users  the user stream is cut into equal ranges, one per lane group (a whole warp at dim 128): per
//          block of 4 samples the three rows are gathered with 128-bit loads (next block prefetched in
//          registers), the dots are reduced by a halving butterfly so that the sigmoid / loss chain runs
//          once per block, the user-row gradient is accumulated in registers while the user stays the
//          same; at the end of a row: PRE-step row -> staging table, updated row -> table, in place;
//          per sample (coefficient, user slot) -> the item stream
//   items  same walk over the item stream: sum of coefficient * staged user row in registers, updated
//          item row in place; last block: global-bias step + brs_step_out
//
// Batch-synchronous semantics hold because item rows are only written by `items` (after every gather
// of `users` has completed: kernel boundary), `items` reads user rows only from the staging copy, and a
// user row is written only after all of its samples were read.  Rows whose segment crosses a range
// boundary (always the Zipf head) add their partial sums into the row-major gradient scratch with 128-bit
// REDs; the last part to arrive (ticket) applies the update.  Round 1 did one 512-byte RED per sample-row
// (196 608 per batch at config 2; RED issue rate was the limiter) plus a second pass over the touched
// rows and 2.9x the compulsory DRAM traffic; here at most two rows per lane group RED.
#include <string.h>

#include "common.cuh"
#include "mf_math.cuh"
#include "opt_math.cuh"

int brs_dense_sweep_untouched(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                              int dense_grad_from_ws, const brs_opt* opt, void* ws, float* out, long long batch,
                              int parity, void* stream);

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kPlanThreads = 256;
constexpr int kPlanWarps = kPlanThreads / 32;
#define BRS_SLOT_OVERFLOW (-3)

// ---------------------------------------------------------------------------
// plan buffer carve-up (host and device agree through this one function)
// ---------------------------------------------------------------------------
struct PlanView {
    int* hdr;        // [64]: 0 = samples in the user stream, 1 = entries in the item stream (segment cursors),
                     //       2 / 3 = work-unit counters of the users / items kernels
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
    int* u_cuts;     // [B+2]  work-unit boundaries of the user stream (see mf_plan_cuts_kernel)
    int* i_cuts;     // [2B+2] ... of the item stream
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
    BRS_CARVE(u_cuts, int, B + 2)
    BRS_CARVE(i_cuts, int, 2 * B + 2)
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
    int unit_shift;     // nominal work-unit length = 1 << unit_shift stream positions
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
        a.pv.hdr[2] = 0;
        a.pv.hdr[3] = 0;
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
        if (s < B) {
            if (valid) {
                if (!wu) su = wait_slot(a.urs.slot_map + u);
                if (!wi) si = wait_slot(a.irs.slot_map + i);
                if (two && !wj) sj = wait_slot(a.irs.slot_map + j);
                if (su < 0 || si < 0 || (two && sj < 0)) valid = false;
            }
            if (valid) {
                a.pv.u_slot[s] = su;
                a.pv.u_rank[s] = atomicAdd(a.pv.u_cnt + su, 1);
                a.pv.i_slot[s] = si;
                a.pv.i_rank[s] = atomicAdd(a.pv.i_cnt + si, 1);
                if (two) {
                    a.pv.i_slot[B + s] = sj;
                    a.pv.i_rank[B + s] = atomicAdd(a.pv.i_cnt + sj, 1);
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

// Work units.  A stream [0, n) is nominally cut every L = 2^unit_shift positions; unit k of a stream is
// [cut[k], cut[k+1]).  A nominal boundary that falls inside a SHORT segment (<= kSnap positions: almost every
// row) moves to that segment's end, so the row is owned by exactly one unit and never takes the RED + ticket
// path; inside a LONG segment (the Zipf head) it moves up to the next multiple of kSnap, so a long row is
// cut every kSnap positions whatever L is.  A unit is therefore at most L + kSnap - 1 positions long.
constexpr int kSnap = 8;
constexpr int kSnapShift = 3;
__global__ void __launch_bounds__(kPlanThreads) mf_plan_cuts_kernel(const PlanArgs a) {
    const int ent = blockIdx.y;
    const int n = a.pv.hdr[ent];
    const int4* rec = ent == 0 ? a.pv.s_b : a.pv.i_a;  // .z / .w = segment begin / end in both streams
    int* cuts = ent == 0 ? a.pv.u_cuts : a.pv.i_cuts;
    const int L = 1 << a.unit_shift;
    const int units = (n + L - 1) >> a.unit_shift;
    for (int k = blockIdx.x * kPlanThreads + threadIdx.x; k <= units; k += gridDim.x * kPlanThreads) {
        const long long pos = (long long)k << a.unit_shift;
        int c;
        if (pos >= n) {
            c = n;
        } else {
            const int4 r = rec[pos];
            const int sb = r.z, se = r.w;
            if ((int)pos <= sb) c = (int)pos;                 // already on a row boundary
            else if (se - sb <= kSnap) c = se;                // short row: whole row goes to the earlier unit
            else c = min(se, ((int)pos + kSnap - 1) & ~(kSnap - 1));
        }
        cuts[k] = c;
    }
}
// parts of the row whose segment is [sb, se): cuts inside a long segment are the multiples of kSnap
__device__ __forceinline__ int row_parts(int sb, int se, int lo, int hi) {
    if (sb >= lo && se <= hi) return 1;  // the whole row is in my unit (the common case)
    if (se - sb <= kSnap) return 1;      // short rows are never cut
    return ((se - 1) >> kSnapShift) - (sb >> kSnapShift) + 1;
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

#ifdef BRS_ROWS_PROFILE
#define BRS_PROF_T(x) const long long x = clock64()
#define BRS_PROF_ADD(k, t0, t1) prof[k] += (t1) - (t0)
#else
#define BRS_PROF_T(x)
#define BRS_PROF_ADD(k, t0, t1)
#endif

struct RowArgs {
    long long* prof;  // BRS_ROWS_PROFILE builds: [warps][8] cycle counters
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
    int unit_shift;  // must equal the one the plan's cuts were computed with
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

template <int LPR>
__device__ __forceinline__ float gsum(unsigned gmask, float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

// A row's last sample (inside this group's range) was accumulated: rows with several parts combine their
// partial sums through the scratch and the last part to arrive continues; then the optimizer is applied
// to the row and its bias, the PRE-step row is staged (users) and the slot is released.
// Group-uniform control flow; only lanes of `gmask` take part.
template <int LPR, int VPL, bool FULL, int KIND>
__device__ __forceinline__ void flush_row(const RowTable& te, const RowTable& tb, int* ticket, int* slot_map,
                                          float* stage, unsigned gmask, int gl, int lane_base, int D, bool skip,
                                          int release, const OptScalars& os, int slot, int row, int nparts,
                                          const float4 (&w)[VPL], float4 (&acc)[VPL], float bias, float gbias) {
    const size_t so = (size_t)(unsigned)slot * (unsigned)D;
    if (nparts > 1) {
        if (!skip) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int col = (v * LPR + gl) * 4;
                if (FULL || col < D) red_add4(te.g + so + col, acc[v]);
            }
            if (gl == 0) red_add1(tb.g + slot, gbias);
            __threadfence();
        }
        __syncwarp(gmask);
        int tk = 0;
        if (gl == 0) tk = atomicAdd(ticket + slot, 1);
        tk = __shfl_sync(gmask, tk, lane_base);
        if (tk != nparts - 1) return;  // somebody else finishes this row
        __threadfence();
        if (!skip) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int col = (v * LPR + gl) * 4;
                if (FULL || col < D) {
                    acc[v] = ld4_cg(te.g + so + col);
                    st4(te.g + so + col, f4_zero());
                }
            }
            if (gl == 0) {
                gbias = __ldcg(tb.g + slot);
                tb.g[slot] = 0.f;
            }
        }
        if (gl == 0) ticket[slot] = 0;
    }
    if (!skip) {
        const size_t ro = (size_t)(unsigned)row * (unsigned)D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
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
        if (gl == 0) {
            float mb = (KIND == BRS_ADAM) ? tb.m[row] : 0.f;
            float vb = (KIND != BRS_SGD) ? tb.v[row] : 0.f;
            opt_elem<KIND>(bias, gbias, mb, vb, os);
            tb.w[row] = bias;
            if (KIND == BRS_ADAM) tb.m[row] = mb;
            if (KIND != BRS_SGD) tb.v[row] = vb;
        }
    }
    if (release && gl == 0) slot_map[row] = BRS_SLOT_NONE;
}

// Ampere-style asynchronous copies (SASS: LDGSTS) global -> shared, tracked by commit groups.  They are
// what keeps several samples per lane group in flight without holding them in registers; every lane
// later reads back exactly the bytes it copied, so no barrier is needed, only cp.async.wait_group.
// src_bytes == 0 zero-fills the destination (predicated-off lanes / columns past dim).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int N>
struct WaitGroups {
    static __device__ __forceinline__ void upto(int pending) {  // wait until at most `pending` (< N) groups are in flight
        if (pending >= N - 1) cp_async_wait<N - 1>();
        else WaitGroups<N - 1>::upto(pending);
    }
};
template <>
struct WaitGroups<1> {
    static __device__ __forceinline__ void upto(int) { cp_async_wait<0>(); }
};

// shared-memory footprint of one warp: two record tiles per lane group (current / next work unit) + a ring of
// S stages; a stage holds, for ONE position of every group, NR rows (VPL x 16 bytes per lane) + 4 scalars
template <int LPR, int VPL, int S, int NR>
struct RingGeom {
    static constexpr int SPW = 32 / LPR;
    static constexpr int TR = 16;  // records (32 bytes each) per group tile >= unit length (L + kSnap - 1, L <= 8)
    static constexpr int ROW_B = VPL * 512;
    static constexpr int STAGE_B = NR * ROW_B + SPW * 16;
    static constexpr int REC_BYTES = 2 * SPW * TR * 32;
    static constexpr int WARP_B = REC_BYTES + S * STAGE_B;
};

// Walks a stream for one warp.  Work units are handed out by an atomic counter (one warp-unit = SPW
// consecutive units, one per lane group), the next unit's records are loaded while the current one is
// processed, and the rows of up to S positions (the one being consumed included) are in flight in the ring
// across unit boundaries.  All loop control is warp-uniform; the per-group range lives in lo/nt.
//   load_rec(buf, t, pos)   record of stream position pos -> slot t of record tile `buf`
//   issue(buf, t, nt, lo, stage)    start the asynchronous copies of slot t (no-op for t >= nt); NO commit
//   consume(buf, t, nt, lo, stage)  process slot t (masked for t >= nt)
template <int LPR, int S, int TR, class LoadRec, class Issue, class Consume>
__device__ __forceinline__ void walk_stream(int n, int unit_shift, const int* __restrict__ cuts, int* counter, int lane,
                                            LoadRec load_rec, Issue issue, Consume consume) {
    constexpr int SPW = 32 / LPR;
    const int gl = lane % LPR, grp = lane / LPR;
    const int L = 1 << unit_shift;
    const int units = (n + L - 1) >> unit_shift;
    const int warp_units = (units + SPW - 1) / SPW;
    auto fetch = [&]() {
        int u = 0;
        if (lane == 0) u = atomicAdd(counter, 1);
        return __shfl_sync(BRS_FULL_MASK, u, 0);
    };
    auto warp_max = [&](int v) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) v = max(v, __shfl_xor_sync(BRS_FULL_MASK, v, o));
        return v;
    };
    // range of this group in warp-unit wu + its records into tile `buf`
    auto setup = [&](int wu, int buf, int& lo, int& nt) {
        lo = 0;
        nt = 0;
        const int k = wu * SPW + grp;
        if (wu < warp_units && k < units) {
            lo = __ldg(cuts + k);
            nt = min(TR, max(0, __ldg(cuts + k + 1) - lo));
        }
        for (int t = gl; t < nt; t += LPR) load_rec(buf, t, lo + t);
    };
    int wu_c = fetch(), lo_c, nt_c, lo_n, nt_n;
    if (wu_c >= warp_units) return;  // warp-uniform
    setup(wu_c, 0, lo_c, nt_c);
    int wu_n = fetch();
    setup(wu_n, 1, lo_n, nt_n);
    int ntw_c = warp_max(nt_c), ntw_n = warp_max(nt_n);
    int cb = 0;               // record tile of the current unit
    int iu = 0, it = 0;       // issue pointer: unit (0 = current, 1 = next), slot
    int gi = 0, gc = 0;       // positions issued / consumed so far (ring stage = count % S)
    __syncwarp();
    auto advance = [&]() -> bool {  // issue the next position of the concatenated units, if any is known yet
        if (iu == 0 && it >= ntw_c) {
            iu = 1;
            it = 0;
        }
        if (iu == 0) issue(cb, it, nt_c, lo_c, gi % S);
        else if (it < ntw_n) issue(cb ^ 1, it, nt_n, lo_n, gi % S);
        else return false;
        cp_async_commit();
        ++it;
        ++gi;
        return true;
    };
    while (wu_c < warp_units) {  // warp-uniform
        for (int t = 0; t < ntw_c; ++t) {
            __syncwarp();  // every lane has finished reading the stage that may be refilled now
            while (gi - gc < S && advance()) {
            }
            WaitGroups<S>::upto(gi - gc - 1);  // position gc has landed (this lane's own copies)
            __syncwarp();                      // ... and the scalar cells written by other lanes of the group
            consume(cb, t, nt_c, lo_c, gc % S);
            ++gc;
        }
        __syncwarp();  // the current unit's records are dead: its tile is refilled with the unit after next
        cb ^= 1;
        wu_c = wu_n;
        lo_c = lo_n;
        nt_c = nt_n;
        ntw_c = ntw_n;
        if (iu == 1) iu = 0;  // `it` positions of the new current unit are already in flight
        else it = 0;
        wu_n = wu_c < warp_units ? fetch() : warp_units;
        setup(wu_n, cb ^ 1, lo_n, nt_n);
        ntw_n = warp_max(nt_n);
        __syncwarp();
    }
    cp_async_wait<0>();
}

template <int LPR, int VPL, bool FULL, int LOSS, int KIND, int S, int NW>
__global__ void __launch_bounds__(NW * 32) mf_user_rows_kernel(const RowArgs a) {
    constexpr int SPW = 32 / LPR;
    constexpr int C = (LOSS == LOSS_BPR) ? 2 : 1;
    using G = RingGeom<LPR, VPL, S, 3>;
    constexpr int TR = G::TR;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ OptScalars s_opt;
    __shared__ float s_red[3][NW];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane % LPR, grp = lane / LPR;
    const int lane_base = grp * LPR;
    const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << (LPR & 31)) - 1u) << lane_base);
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

    unsigned char* wbase = smem + (size_t)warp * G::WARP_B;
    // record slot t of the group's tile `buf`: rec[2t] = s_a, rec[2t+1] = s_b
    auto rec_of = [&](int buf) { return (int4*)wbase + (buf * SPW + grp) * TR * 2; };
    unsigned char* ring = wbase + G::REC_BYTES;
    // this lane's 16-byte cells inside a stage: row r, vector v at ((r * VPL + v) * 32 + lane) * 16
    const int cell = lane * 16;
    const int bcell = 3 * G::ROW_B + grp * 16;  // the group's {b_u, b_i, b_j} of the position

    float4 acc[VPL], ru[VPL];  // the user row being accumulated: gradient, PRE-step weights
    float gbias = 0.f, bu = 0.f, uu = 0.f;
    int n_row = 0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = ru[v] = f4_zero();

    auto load_rec = [&](int buf, int t, int pos) {
        int4* rec = rec_of(buf);
        rec[2 * t] = __ldg(a.pv.s_a + pos);
        rec[2 * t + 1] = __ldg(a.pv.s_b + pos);
    };
    // asynchronous gather of a position's rows into a ring stage.  The stream is sorted by user, so the
    // user row (and bias) only travels with the FIRST sample of a row inside the unit
    auto issue = [&](int buf, int t, int nt, int lo, int stage) {
        if (t >= nt) return;
        const int4* rec = rec_of(buf);
        unsigned char* st = ring + stage * G::STAGE_B;
        const int4 ra = rec[2 * t];
        const bool first = lo + t == rec[2 * t + 1].z || t == 0;
        const float* up = a.ue.w + (size_t)(unsigned)ra.x * (unsigned)D;
        const float* ip = a.ie.w + (size_t)(unsigned)ra.z * (unsigned)D;
        const float* jp = a.ie.w + (size_t)(unsigned)(C == 2 ? ra.w : 0) * (unsigned)D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
            const bool ld = FULL || col < D;
            const int cc = ld ? col : 0;
            unsigned char* d0 = st + v * 512 + cell;
            if (first) cp_async16(d0, up + cc, ld ? 16 : 0);
            cp_async16(d0 + G::ROW_B, ip + cc, ld ? 16 : 0);
            if (C == 2) cp_async16(d0 + 2 * G::ROW_B, jp + cc, ld ? 16 : 0);
        }
        // lanes 0..2 of the group fetch the three biases (they wrap around for tiny dims)
#pragma unroll
        for (int k = gl; k < 1 + C; k += LPR) {
            const float* bp = k == 0 ? a.ub.w + (unsigned)ra.x : (k == 1 ? a.ib.w + (unsigned)ra.z : a.ib.w + (unsigned)ra.w);
            if (k != 0 || first) cp_async4(st + bcell + k * 4, bp, 4);
        }
    };
    auto consume = [&](int buf, int t, int nt, int lo, int stage) {
        const bool on = t < nt;  // groups past their unit compute on stale data, side effects are masked
        const int tc = on ? t : 0;
        const int4* rec = rec_of(buf);
        const unsigned char* st = ring + stage * G::STAGE_B;
        const int4 ra = rec[2 * tc], rb = rec[2 * tc + 1];
        const int p = lo + tc, hi = lo + nt;
        const int slot = ra.y, sb = rb.z, se = rb.w;
        const bool first = on && (p == sb || tc == 0);
        const float4 bb = *(const float4*)(st + bcell);  // {b_u, b_i, b_j, -}
        if (first) {  // a new user row starts here: its PRE-step weights stay in registers
            uu = 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                ru[v] = *(const float4*)(st + v * 512 + cell);
                uu += f4_dot(ru[v], ru[v]);
                acc[v] = f4_zero();
            }
            bu = bb.x;
            gbias = 0.f;
            n_row = 0;
        }
        float4 ri[VPL], rj[C == 2 ? VPL : 1];
        float dp = 0.f, dn = 0.f, sq = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const unsigned char* d0 = st + v * 512 + cell;
            ri[v] = *(const float4*)(d0 + G::ROW_B);
            dp += f4_dot(ru[v], ri[v]);
            sq += f4_dot(ri[v], ri[v]);
            if (C == 2) {
                rj[v] = *(const float4*)(d0 + 2 * G::ROW_B);
                dn += f4_dot(ru[v], rj[v]);
                sq += f4_dot(rj[v], rj[v]);
            }
        }
        const float bi = bb.y, bj = (C == 2) ? bb.z : 0.f;
        dp = group_sum<LPR>(dp);
        if (C == 2) dn = group_sum<LPR>(dn);
        const float rating = (C == 1) ? __int_as_float(ra.w) : 0.f;
        float cu_i, cu_j, loss_k;
        mf_sample_coef<LOSS>(dp + bu + bi + bg, dn + bu + bj + bg, rating, a.inv_b, cu_i, cu_j, loss_k);
        if (!on) return;
        reg_acc += fwd_calls * uu + sq;  // regularizer numerator (mf.py:49-54)
        if (gl == 0) {
            reg_acc += fwd_calls * bu * bu + bi * bi + bj * bj;
            loss_acc += loss_k;
            gb_acc += cu_i + cu_j;
            const float sf = __int_as_float(slot);  // hand the coefficients to the item stream
            a.pv.ipair[rb.x] = make_float2(cu_i, sf);
            if (C == 2) a.pv.ipair[rb.y] = make_float2(cu_j, sf);
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            acc[v] = f4_fma(cu_i, ri[v], acc[v]);
            if (C == 2) acc[v] = f4_fma(cu_j, rj[v], acc[v]);
        }
        gbias += cu_i + cu_j;
        n_row += 1;
        if (p + 1 == se || p + 1 == hi) {  // last sample of this row inside my unit
            if (a.reg_w != 0.f) {
                const float nl = fwd_calls * (float)n_row * rw;
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(nl, ru[v], acc[v]);
                gbias += nl * bu;
            }
            flush_row<LPR, VPL, FULL, KIND>(a.ue, a.ub, a.pv.u_ticket, a.u_slot_map, a.user_stage, gmask, gl, lane_base, D, skip,
                                            a.release, os, slot, ra.x, row_parts(sb, se, lo, hi), ru, acc, bu, gbias);
        }
    };
    walk_stream<LPR, S, TR>(n, a.unit_shift, a.pv.u_cuts, a.pv.hdr + 2, lane, load_rec, issue, consume);

    // block reduction of the scalar outputs -> 3 atomics per block
    __syncwarp();
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
        for (int q = 0; q < NW; ++q) {
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

template <int LPR, int VPL, bool FULL, int KIND, int S, int NW>
__global__ void __launch_bounds__(NW * 32) mf_item_rows_kernel(const RowArgs a) {
    constexpr int SPW = 32 / LPR;
    using G = RingGeom<LPR, VPL, S, 2>;
    constexpr int TR = G::TR;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ OptScalars s_opt;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane % LPR, grp = lane / LPR;
    const int lane_base = grp * LPR;
    const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << (LPR & 31)) - 1u) << lane_base);
    const int D = a.dim;
    if (threadIdx.x == 0) s_opt = make_scalars(a.opt, a.ws->step + 1);
    __syncthreads();
    const OptScalars os = s_opt;
    const bool skip = __ldg(a.err) != 0u;
    const int n = __ldg(a.pv.hdr + 1);
    const float rw = 2.0f * a.reg_w * a.inv_b;

    unsigned char* wbase = smem + (size_t)warp * G::WARP_B;
    // entry slot t of the group's tile `buf`: rec[2t] = i_a, rec[2t+1].xy = ipair {coefficient, user slot}
    auto rec_of = [&](int buf) { return (int4*)wbase + (buf * SPW + grp) * TR * 2; };
    unsigned char* ring = wbase + G::REC_BYTES;
    const int cell = lane * 16;
    const int bcell = 2 * G::ROW_B + grp * 16;

    float4 acc[VPL];
    float gbias = 0.f;
    int n_row = 0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = f4_zero();

    auto load_rec = [&](int buf, int t, int pos) {
        int4* rec = rec_of(buf);
        rec[2 * t] = __ldg(a.pv.i_a + pos);
        const float2 pr = __ldcg(a.pv.ipair + pos);
        rec[2 * t + 1] = make_int4(__float_as_int(pr.x), __float_as_int(pr.y), 0, 0);
    };
    // the item's own row (and bias) only travels with the LAST entry of the row inside the unit, where the
    // update is applied; every entry brings the staged PRE-step row of its user
    auto issue = [&](int buf, int t, int nt, int lo, int stage) {
        if (t >= nt) return;
        const int4* rec = rec_of(buf);
        unsigned char* st = ring + stage * G::STAGE_B;
        const int4 ia = rec[2 * t];
        const int us = rec[2 * t + 1].y;
        const bool last = lo + t + 1 == ia.w || t + 1 == nt;
        const float* up = a.user_stage + (size_t)(unsigned)us * (unsigned)D;
        const float* ip = a.ie.w + (size_t)(unsigned)ia.x * (unsigned)D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
            const bool ld = FULL || col < D;
            const int cc = ld ? col : 0;
            unsigned char* d0 = st + v * 512 + cell;
            cp_async16(d0, up + cc, ld ? 16 : 0);
            if (last) cp_async16(d0 + G::ROW_B, ip + cc, ld ? 16 : 0);
        }
        if (last && gl == 0) cp_async4(st + bcell, a.ib.w + (unsigned)ia.x, 4);
    };
    auto consume = [&](int buf, int t, int nt, int lo, int stage) {
        if (t >= nt) return;
        const int4* rec = rec_of(buf);
        const unsigned char* st = ring + stage * G::STAGE_B;
        const int q = lo + t, hi = lo + nt;
        const int4 ia = rec[2 * t];
        const float coef = __int_as_float(rec[2 * t + 1].x);
        const int slot = ia.y, sb = ia.z, se = ia.w;
        if (q == sb || t == 0) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) acc[v] = f4_zero();
            gbias = 0.f;
            n_row = 0;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(coef, *(const float4*)(st + v * 512 + cell), acc[v]);
        gbias += coef;
        n_row += 1;
        if (q + 1 == se || q + 1 == hi) {
            float4 w[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) w[v] = *(const float4*)(st + G::ROW_B + v * 512 + cell);
            const float bi = *(const float*)(st + bcell);
            if (a.reg_w != 0.f) {
                const float nl = (float)n_row * rw;
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(nl, w[v], acc[v]);
                gbias += nl * bi;
            }
            flush_row<LPR, VPL, FULL, KIND>(a.ie, a.ib, a.pv.i_ticket, a.i_slot_map, nullptr, gmask, gl, lane_base, D, skip,
                                            a.release, os, slot, ia.x, row_parts(sb, se, lo, hi), w, acc, bi, gbias);
        }
    };
    walk_stream<LPR, S, TR>(n, a.unit_shift, a.pv.i_cuts, a.pv.hdr + 3, lane, load_rec, issue, consume);

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

// The ring lives in shared memory, which is carved out of the same 228 KB as the L1 cache: the blocks per
// SM are capped so that the rings take about half of it and the Zipf-hot item rows still hit L1 (with all
// of it given to rings the L1 hit rate fell to 5% and every hot-row read went to the same two L2 slices).
int g_rows_blocks_per_sm = 0;  // 0 = default below; diagnostics: brs_debug_set_mf_rows_shape
template <class K>
int launch_persistent(K kernel, int threads, int smem_bytes, const RowArgs& a, cudaStream_t st) {
    BRS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    BRS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 1;
    BRS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, (size_t)smem_bytes));
    if (per_sm < 1) per_sm = 1;
    int want = g_rows_blocks_per_sm > 0 ? g_rows_blocks_per_sm : (120 * 1024) / (smem_bytes + 1024);
    if (want < 1) want = 1;
    if (per_sm > want) per_sm = want;
    // ask for the smallest carve-out that holds per_sm blocks; the rest of the 228 KB is L1
    int pct = (int)(((long long)per_sm * (smem_bytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    BRS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    kernel<<<brs_sm_count() * per_sm, threads, smem_bytes, st>>>(a);
    return BRS_OK;
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

int g_rows_stages = 2, g_rows_warps = 2;  // diagnostics: brs_debug_set_mf_rows_shape
int g_unit_shift = 1;                     // nominal work unit = 2 stream positions per lane group
int g_rows_only = 0;                      // diagnostics: 1 = users kernel only, 2 = items kernel only
long long* g_rows_prof = nullptr;         // BRS_ROWS_PROFILE builds: device [65536][8] cycle counters

template <int LOSS, int KIND>
int launch_rows(const RowArgs& a, cudaStream_t st) {
    const int D = a.dim;
    // a row of D floats = LPR lanes x VPL float4 with VPL = 4 wherever D allows (D = 128 -> 8 lanes x 4): every
    // warp instruction -- record reads, address math, the sigmoid / loss chain, range tests -- serves 32/LPR
    // stream positions at once
#define BRS_ROWS_SN(LPR, VPL, FULL, S, NW)                                                                     \
    do {                                                                                                       \
        int rc_ = BRS_OK;                                                                                      \
        if (g_rows_only != 2)                                                                                  \
            rc_ = launch_persistent(mf_user_rows_kernel<LPR, VPL, FULL, LOSS, KIND, S, NW>, NW * 32,           \
                                    NW * RingGeom<LPR, VPL, S, 3>::WARP_B, a, st);                             \
        if (rc_ != BRS_OK) return rc_;                                                                         \
        if (g_rows_only != 1)                                                                                  \
            rc_ = launch_persistent(mf_item_rows_kernel<LPR, VPL, FULL, KIND, S, NW>, NW * 32,                 \
                                    NW * RingGeom<LPR, VPL, S, 2>::WARP_B, a, st);                             \
        if (rc_ != BRS_OK) return rc_;                                                                         \
    } while (0)
#define BRS_ROWS(LPR, VPL, FULL) BRS_ROWS_SN(LPR, VPL, FULL, 2, 2)
    if (D == 128 && (g_rows_stages != 2 || g_rows_warps != 2)) {  // tuning sweep of the benchmark shape
        if (g_rows_stages == 3 && g_rows_warps == 2) BRS_ROWS_SN(8, 4, true, 3, 2);
        else if (g_rows_stages == 2 && g_rows_warps == 4) BRS_ROWS_SN(8, 4, true, 2, 4);
        else if (g_rows_stages == 3 && g_rows_warps == 4) BRS_ROWS_SN(8, 4, true, 3, 4);
        else if (g_rows_stages == 4 && g_rows_warps == 2) BRS_ROWS_SN(8, 4, true, 4, 2);
        else return BRS_ERR_INVALID_ARG;
        BRS_CUDA_CHECK(cudaGetLastError());
        return BRS_OK;
    }
    switch (D) {
        case 4: BRS_ROWS(1, 1, true); break;
        case 8: BRS_ROWS(1, 2, true); break;
        case 16: BRS_ROWS(1, 4, true); break;
        case 32: BRS_ROWS(2, 4, true); break;
        case 64: BRS_ROWS(4, 4, true); break;
        case 128: BRS_ROWS(8, 4, true); break;
        case 256: BRS_ROWS(16, 4, true); break;
        case 384: BRS_ROWS(32, 3, true); break;
        case 512: BRS_ROWS(32, 4, true); break;
        default:  // any other multiple of 4: next power-of-two lane group, tail lanes idle
            if (D < 8) BRS_ROWS(2, 1, false);
            else if (D < 16) BRS_ROWS(4, 1, false);
            else if (D < 32) BRS_ROWS(8, 1, false);
            else if (D < 64) BRS_ROWS(16, 1, false);
            else if (D < 128) BRS_ROWS(32, 1, false);
            else if (D < 256) BRS_ROWS(32, 2, false);
            else if (D < 384) BRS_ROWS(32, 3, false);
            else BRS_ROWS(32, 4, false);
    }
#undef BRS_ROWS
#undef BRS_ROWS_SN
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

RowTable row_table(const brs_table& t) { return RowTable{t.weight, t.m, t.v, t.grad}; }

}  // namespace

extern "C" int64_t brs_mf_plan_bytes(int64_t batch_capacity, int32_t user_capacity, int32_t item_capacity) {
    if (batch_capacity <= 0 || user_capacity <= 0 || item_capacity <= 0) return 0;
    return (int64_t)plan_view(nullptr, batch_capacity, user_capacity, item_capacity).bytes;
}

extern "C" int brs_debug_mf_rows_profile(long long* host_out, int n_warps) {
#ifdef BRS_ROWS_PROFILE
    if (!g_rows_prof) {
        BRS_CUDA_CHECK(cudaMalloc(&g_rows_prof, 65536 * 8 * sizeof(long long)));
        BRS_CUDA_CHECK(cudaMemset(g_rows_prof, 0, 65536 * 8 * sizeof(long long)));
    }
    if (host_out && n_warps > 0)
        BRS_CUDA_CHECK(cudaMemcpy(host_out, g_rows_prof, (size_t)(n_warps < 65536 ? n_warps : 65536) * 8 * sizeof(long long),
                                  cudaMemcpyDeviceToHost));
    return BRS_OK;
#else
    (void)host_out;
    (void)n_warps;
    return BRS_ERR_UNSUPPORTED;
#endif
}

extern "C" int brs_debug_set_mf_rows_only(int which) {
    g_rows_only = which;
    return BRS_OK;
}

extern "C" int brs_debug_set_mf_rows_shape(int stages, int warps_per_block, int blocks_per_sm, int unit_shift) {
    if (unit_shift < 0 || unit_shift > 3) return BRS_ERR_INVALID_ARG;  // unit <= 8: tiles hold 16 records
    g_rows_stages = stages;
    g_rows_warps = warps_per_block;
    g_rows_blocks_per_sm = blocks_per_sm;
    g_unit_shift = unit_shift;
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
    a.unit_shift = g_unit_shift;
    a.err = &((brs_step_ws*)model->ws)->err_pending[which];
    cudaStream_t st = (cudaStream_t)stream;
    if (batch > 0) {
        long long blocks = (batch + kPlanThreads - 1) / kPlanThreads;
        const long long cap = (long long)brs_sm_count() * 4;
        if (blocks > cap) blocks = cap;
        mf_plan_claim_kernel<<<(int)blocks, kPlanThreads, 0, st>>>(a);
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
    {
        const long long units = ((2 * batch) >> a.unit_shift) + 2;
        long long blocks = (units + kPlanThreads - 1) / kPlanThreads;
        const long long cap = (long long)brs_sm_count() * 2;
        if (blocks > cap) blocks = cap;
        mf_plan_cuts_kernel<<<dim3((unsigned)blocks, 2u), kPlanThreads, 0, st>>>(a);
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
    a.prof = g_rows_prof;
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
    a.unit_shift = g_unit_shift;
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
    (loss_kind == LOSS_BPR ? launch_rows<LOSS_BPR, KIND>(a, st) : launch_rows<LOSS_BCE, KIND>(a, st))
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
