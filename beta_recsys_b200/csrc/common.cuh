// Device helpers shared by the sm_100a kernels of libbrs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/brs_b200.h"

#define BRS_WARP 32
#define BRS_FULL_MASK 0xffffffffu

// Device scratch behind brs_*_model.ws (BRS_STEP_WS_BYTES, zero-initialised once by the host).
struct __align__(16) brs_step_ws {
    double loss_sum;          // sum over samples of the per-sample loss
    double reg_sum;           // sum over samples of the regularizer numerator
    float g_global_bias;      // d loss / d global_bias (MF)
    unsigned int ticket;      // blocks finished in the apply kernel (last one finalises)
    long long step;           // optimizer step counter t (incremented by apply)
    unsigned int err_flag;    // set when an index is out of range
    unsigned int err_pending[2];  // same, raised by a pre-pass that ran inside the PREVIOUS step's apply launch
    unsigned int predict_err; // raised by the no_grad scoring kernels only (never by a training step): the
                              // host reads and clears it after a predict call (engines: _check_predict)
    unsigned int pad_[6];
};
#define BRS_WS_PREDICT_ERR_OFFSET 44
static_assert(offsetof(brs_step_ws, predict_err) == BRS_WS_PREDICT_ERR_OFFSET, "engines read this word through the ws tensor");
static_assert(sizeof(brs_step_ws) <= BRS_STEP_WS_BYTES, "ws layout");

// ---------------------------------------------------------------------------
// memory ops
// ---------------------------------------------------------------------------
// 128-bit gather load of an embedding row on the read-only (L1-allocating) path: tables are
// never written by the kernels that gather them, and under Zipf indices the hot rows are
// re-read dozens of times per SM -- L1 hits keep those reads off the two L2 slices a 512-byte
// row hashes to (round-1 profile: the hottest row is ~10% of a batch).
__device__ __forceinline__ float4 ld_row4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// L2 eviction-priority policies (createpolicy + L2::cache_hint).  Round-1 sweep: the fused kernel's
// time did not react to occupancy / ILP / lane mapping, while ~60% of its RED sectors missed L2 and
// became random 32-byte DRAM read-modify-writes.  The compact gradient scratch is therefore pinned
// with evict_last, and streamed weight traffic can be demoted with evict_first.
enum { BRS_L2_NORMAL = 0, BRS_L2_EVICT_FIRST = 1, BRS_L2_EVICT_LAST = 2 };
__device__ __forceinline__ unsigned long long l2_policy(int kind) {
    unsigned long long p;
    if (kind == BRS_L2_EVICT_FIRST)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == BRS_L2_EVICT_LAST)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_row4_pol(const float* p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ld4_pol(const float* p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return r;
}
__device__ __forceinline__ void st4_pol(float* p, float4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void red_add4_pol(float* p, float4 v, unsigned long long pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
// process-wide policy selection (diagnostics: brs_debug_set_l2_policy)
struct brs_l2_policy_cfg {
    int gather;   // embedding-row gathers in the fused kernels
    int scratch;  // compact gradient scratch: REDs, reads, zeroing
    int weight;   // weight-row updates in the apply kernels
};
const brs_l2_policy_cfg& brs_l2_cfg();

// Gradient-scratch layout.  For dim % BRS_GS_BLOCK == 0 the compact scratch of a table is stored
// blocked: [dim/BLK][capacity][BLK floats], i.e. element (slot, col) lives at
//     ((col / BLK) * capacity + slot) * BLK + (col % BLK).
// Default BLK = 8 ("sector-blocked"): the 32-byte sectors of one gradient row are capacity*32 bytes
// apart and hash to different L2 slices.  A B200 L2 slice retires ~1 RED sector per clock, and a row-major
// 512-byte row maps to only TWO slices, so under Zipf indices the hottest row serialised ~36 us of REDs on
// one slice (round-1 ncu: lts__d_atomic_input_cycles_active max 62%, 31% with BLK = 8).  The price is four
// times as many L2 requests per warp-level RED (profiles/r01f_fused_kernel_source_stalls.md);
// -DBRS_GS_BLOCK=32 (one 128-byte line per block, a row over 4 slices) is the round-2 experiment
// (BRS_NVCC_DEFINES, build.py).  Other dims keep the row-major [capacity][dim] layout.
#ifndef BRS_GS_BLOCK
#define BRS_GS_BLOCK 8
#endif
static_assert(BRS_GS_BLOCK >= 4 && (BRS_GS_BLOCK & (BRS_GS_BLOCK - 1)) == 0, "block = power of two >= one float4");
__device__ __forceinline__ size_t gs_off(int dim, int capacity, unsigned slot, int col) {
    return (dim & (BRS_GS_BLOCK - 1))
               ? (size_t)slot * (unsigned)dim + col
               : ((size_t)(col / BRS_GS_BLOCK) * (unsigned)capacity + slot) * BRS_GS_BLOCK + (col & (BRS_GS_BLOCK - 1));
}

// 128-bit fire-and-forget scatter-add (sm_90+): one L2 reduction per 16 bytes.
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// The same reductions at SYSTEM scope, for destinations that may be another GPU's memory (row-sharded push:
// up to 8 GPUs reduce into the same hot rows over NVLink).  At .gpu scope those cross-device reductions are
// not morally strong under the PTX memory model.
__device__ __forceinline__ void red_add4_sys(float* p, float4 v) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add1_sys(float* p, float v) {
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BRS_FULL_MASK, v, o);
    return v;
}
// sum across the LANES-wide aligned group the lane belongs to
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(BRS_FULL_MASK, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (global -> shared), used to stage index tiles
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// cp.async.bulk (SASS: UBLKCP): bytes % 16 == 0, src/dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------
// touched-row bookkeeping: slot_map values (assign_slots_kernel in rows_apply.cu: the
// first toucher of a row claims the next slot of the entity's compact gradient scratch)
// ---------------------------------------------------------------------------
#define BRS_SLOT_NONE (-1)
#define BRS_SLOT_PENDING (-2)

// ---------------------------------------------------------------------------
// scalar math exactly as ATen evaluates it in fp32
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// F.logsigmoid(x) = min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float logsigmoidf_(float x) { return fminf(x, 0.0f) - log1pf(expf(-fabsf(x))); }
// F.softplus(x), beta 1, threshold 20
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

#define BRS_CUDA_CHECK(expr)                          \
    do {                                              \
        cudaError_t e__ = (expr);                     \
        if (e__ != cudaSuccess) {                     \
            brs_set_cuda_error(e__, #expr, __LINE__); \
            return BRS_ERR_CUDA;                      \
        }                                             \
    } while (0)

void brs_set_cuda_error(cudaError_t e, const char* what, int line);
int brs_sm_count();
