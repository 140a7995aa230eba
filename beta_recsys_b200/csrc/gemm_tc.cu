// tcgen05 (5th-gen tensor core) Linear layer for the NCF tower -- sm_100a only.
//
//   Y[M,N] = epilogue( A[M,K] . B[N,K]^T )      A, B row-major fp32 ("K-major" operands)
//   epilogue: (+ bias[N]) (ReLU) (* (mask_src[M,N] > 0))
//
// used for nn.Linear forward (A = activations, B = weight [out,in]; beta_rec/models/ncf.py:64-69)
// and for dgrad (A = dY, B = W^T, mask = ReLU mask of the layer input).
//
// fp32 parity on tensor cores (north-star budget 1e-5; plain TF32 is ~1e-3): 3xTF32 error
// compensation.  While staging a k-block from global to shared memory every value x is split in
// registers into hi = tf32(x) and lo = tf32(x - hi); the MMA issuer then accumulates
// hi.hi + hi.lo + lo.hi into the SAME TMEM accumulator (three tcgen05.mma per k-step).
//
// Accuracy note (round-1 measurement, tools/gemm_accuracy.py): with ONE accumulator the error grew
// linearly in K (rms 1.0e-6 / 1.9e-6 / 3.6e-6 at K = 128 / 256 / 512) -- the tensor core truncates on
// every accumulation, a bias that adds up along the chain.  The kernel therefore keeps THREE TMEM
// accumulators per tile -- hi.hi of the even k-blocks, hi.hi of the odd k-blocks, and the (2^-11
// smaller) cross terms -- and adds them with round-to-nearest in the epilogue.
//
// Structure (one CTA = one 128 x BN output tile, 256 threads):
//   * operands are written to shared memory in the K-major SWIZZLE_128B layout (128-byte rows, 16-byte
//     chunks XOR-ed with row % 8) with plain st.shared -- the split has to pass through registers anyway,
//     so no tensor maps are needed; the no-swizzle layout of the first version left the tensor pipe
//     waiting for operands (ncu: L1TEX 77 % busy, tensor math 13 %);
//   * the k-block after the one being split is already in flight from global memory (register prefetch);
//   * thread 0 issues tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x BN x 8), accumulator in TMEM;
//     tcgen05.commit -> mbarrier tells the loaders when a stage may be overwritten (3-4 stages), so
//     splitting k-block kb+1 overlaps the MMAs of k-block kb;
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld (32x32b.x8), applies
//     bias / ReLU / mask and stores rows to global memory.
#include "common.cuh"

namespace {

constexpr int TC_M = 128;      // rows per CTA tile (UMMA_M)
constexpr int TC_BK = 32;      // fp32 elements of K per stage (4 MMA k-steps of 8)
constexpr int TC_THREADS = 512;  // 16 warps: two k-blocks of global loads in flight at 32 registers per thread
// shared-memory stages: 3 x 64 KB (BN = 128) or 4 x 48 KB (BN = 64), one CTA per SM
__host__ __device__ constexpr int tc_stages(int bn) { return bn == 128 ? 3 : 4; }

// round-to-nearest (ties away from zero) to the 10-bit tf32 mantissa, on the integer ALU: the same result as
// cvt.rna.tf32.f32 for finite inputs, but that instruction issues on the XU pipe (16 lanes/clk/SM), which the
// first version of this kernel saturated (ncu: sm__inst_executed_pipe_xu 128 % of peak, tensor math 13 %)
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// mbarrier wait that traps instead of hanging the GPU if the tensor-core pipeline never signals
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, unsigned parity) {
    for (unsigned spins = 0; spins < (1u << 26); ++spins) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (a row of the tile = 128 bytes = 32 tf32 of K, the
// eight 16-byte chunks of a row XOR-ed with row % 8; 8 rows = one 1024-byte swizzle atom):
//   start address >> 4 | LBO >> 4 at bit 16 (unused: K fits one atom) | SBO >> 4 at bit 32 (next 8-row
//   group = 1024 B) | version 1 at bit 46 | layout type 2 (SWIZZLE_128B) at bit 61
// A k-step of 8 tf32 advances the start address by 32 bytes inside the row.
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);
    d |= (unsigned long long)(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr unsigned make_idesc_tf32(int m, int n) {
    return (1u << 4)                      // c_format = F32
           | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
           | ((unsigned)(n >> 3) << 17)   // n_dim
           | ((unsigned)(m >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// float offset of the 16-byte chunk k4 (= k/4) of row `row` in a [ROWS x 32] fp32 tile, K-major SWIZZLE_128B:
// [row/8][row%8][(k/4) ^ (row%8)][k%4]
__device__ __forceinline__ int tile_off(int row, int k4) { return (row >> 3) * 256 + (row & 7) * 32 + ((k4 ^ (row & 7)) << 2); }

// float4 values per thread of a [ROWS x 32] fp32 tile: ROWS/4 warp tasks (4 rows x 128 bytes each) over the CTA's warps
__host__ __device__ constexpr int tc_per_thread(int rows) { return (rows / 8) * 2 / (TC_THREADS / 32); }

// A [ROWS x 32] block of a row-major matrix (leading dim ld) travels global -> registers -> shared in two
// steps so that the global loads of k-block kb+1 are in flight while k-block kb is split and multiplied.
// Thread mapping: a quarter-warp covers the eight 16-byte chunks of ONE row (128 contiguous bytes in global
// memory: a warp instruction is four full 128-byte lines; in shared memory the same 128-byte row, chunks
// permuted by the swizzle: conflict-free), the four quarter-warps take four consecutive rows.
template <int ROWS>
__device__ __forceinline__ void tile_load(const float* __restrict__ src, int ld, int row0, int n_rows_total, int k0,
                                          float4 (&r)[tc_per_thread(ROWS)]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = TC_THREADS / 32;
#pragma unroll
    for (int i = 0; i < tc_per_thread(ROWS); ++i) {
        const int gr = row0 + (w + i * NW) * 4 + (lane >> 3);
        const int k4 = lane & 7;
        r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < n_rows_total) r[i] = __ldg((const float4*)(src + (size_t)gr * ld + k0 + k4 * 4));
    }
}

// split every value x into hi = tf32(x), lo = tf32(x - hi) and store both tiles
template <int ROWS>
__device__ __forceinline__ void tile_split_store(const float4 (&r)[tc_per_thread(ROWS)], float* __restrict__ hi,
                                                 float* __restrict__ lo) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = TC_THREADS / 32;
#pragma unroll
    for (int i = 0; i < tc_per_thread(ROWS); ++i) {
        const int row = (w + i * NW) * 4 + (lane >> 3);
        const int k4 = lane & 7;
        const float4 v = r[i];
        float4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
        const int o = tile_off(row, k4);
        *(float4*)(hi + o) = h;
        *(float4*)(lo + o) = l;
    }
}

struct TcArgs {
    const float* A; int lda;   // [M, K]
    const float* B; int ldb;   // [N_total, K]
    float* Y; int ldy;         // [M, N_total]
    const float* bias;         // [N_total] or NULL
    const float* mask; int ldm;  // [M, N_total] or NULL: output kept where mask > 0
    int M, N_total, K;
    int relu;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) linear_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int A_FLOATS = TC_M * TC_BK, B_FLOATS = BN * TC_BK;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    constexpr int STAGES = tc_stages(BN);
    float* smem = (float*)smem_raw;
    __shared__ __align__(8) uint64_t s_bar[STAGES];  // MMAs reading stage s have completed
    __shared__ __align__(8) uint64_t s_done;         // all MMAs of the tile have completed
    __shared__ unsigned s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // column tiles vary fastest: the CTAs that share a block of A rows run together and re-read it from L2
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_bar[s], 1);
        mbar_init(&s_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) {  // TMEM: three BN-column fp32 accumulators (hi.hi even k-blocks, hi.hi odd, cross terms)
        constexpr unsigned cols = 512;
        static_assert(3 * BN <= 512, "TMEM columns");
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_acc = s_tmem;

    constexpr unsigned idesc = make_idesc_tf32(TC_M, BN);
    const int n_kb = a.K / TC_BK;
    // two k-blocks in flight from global memory (register sets 0 / 1 alternate)
    float4 ra0[tc_per_thread(TC_M)], rb0[tc_per_thread(BN)], ra1[tc_per_thread(TC_M)], rb1[tc_per_thread(BN)];
    tile_load<TC_M>(a.A, a.lda, m0, a.M, 0, ra0);
    tile_load<BN>(a.B, a.ldb, n0, a.N_total, 0, rb0);
    if (n_kb > 1) {
        tile_load<TC_M>(a.A, a.lda, m0, a.M, TC_BK, ra1);
        tile_load<BN>(a.B, a.ldb, n0, a.N_total, TC_BK, rb1);
    }
    // one k-block: wait for the stage, split the register set into it, refill the set with k-block kb+2,
    // hand the stage to the tensor core
#define BRS_TC_KBLOCK(KB, RA, RB)                                                                                     \
    do {                                                                                                              \
        const int kb = (KB);                                                                                          \
        const int s = kb % STAGES;                                                                                    \
        float* a_hi = smem + s * STAGE_FLOATS;                                                                        \
        float* a_lo = a_hi + A_FLOATS;                                                                                \
        float* b_hi = a_lo + A_FLOATS;                                                                                \
        float* b_lo = b_hi + B_FLOATS;                                                                                \
        if (kb >= STAGES) mbar_wait_or_trap(&s_bar[s], (unsigned)((kb / STAGES - 1) & 1));                            \
        tile_split_store<TC_M>(RA, a_hi, a_lo);                                                                       \
        tile_split_store<BN>(RB, b_hi, b_lo);                                                                         \
        /* generic-proxy stores -> async-proxy (MMA) reads; issued before the prefetch so that the fence never   */  \
        /* has loads in flight to order (measured neutral: 151 vs 154 us at 65536x256x512)                       */  \
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                  \
        if (kb + 2 < n_kb) {                                                                                          \
            tile_load<TC_M>(a.A, a.lda, m0, a.M, (kb + 2) * TC_BK, RA);                                               \
            tile_load<BN>(a.B, a.ldb, n0, a.N_total, (kb + 2) * TC_BK, RB);                                           \
        }                                                                                                             \
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                              \
        __syncthreads();                                                                                              \
        if (threadIdx.x == 0) {                                                                                       \
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                           \
            const unsigned a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo);                                          \
            const unsigned b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);                                          \
            _Pragma("unroll") for (int j = 0; j < TC_BK / 8; ++j) { /* one UMMA k-step = 8 tf32 = two 16-B chunks */ \
                const unsigned long long ah = make_smem_desc(a_hi_s + 32u * j);                                       \
                const unsigned long long al = make_smem_desc(a_lo_s + 32u * j);                                       \
                const unsigned long long bh = make_smem_desc(b_hi_s + 32u * j);                                       \
                const unsigned long long bl = make_smem_desc(b_lo_s + 32u * j);                                       \
                tc_mma_tf32(tmem_acc + (unsigned)((kb & 1) * BN), ah, bh, idesc, ((kb >> 1) | j) ? 1u : 0u);          \
                tc_mma_tf32(tmem_acc + (unsigned)(2 * BN), ah, bl, idesc, (kb | j) ? 1u : 0u);                        \
                tc_mma_tf32(tmem_acc + (unsigned)(2 * BN), al, bh, idesc, 1u);                                        \
            }                                                                                                         \
            tc_commit(&s_bar[s]);                   /* stage s reusable once these MMAs finish */                     \
            if (kb == n_kb - 1) tc_commit(&s_done); /* accumulator complete */                                        \
        }                                                                                                             \
    } while (0)
    for (int kb2 = 0; kb2 < n_kb; kb2 += 2) {
        BRS_TC_KBLOCK(kb2, ra0, rb0);
        if (kb2 + 1 < n_kb) BRS_TC_KBLOCK(kb2 + 1, ra1, rb1);
    }
#undef BRS_TC_KBLOCK
    mbar_wait_or_trap(&s_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: a warp may only touch the TMEM lane quarter (warp % 4); the four warps of a quarter
    // split the columns.  y = acc_even + acc_odd + acc_cross, then bias / ReLU / mask.
    constexpr int PARTS = TC_THREADS / 128;  // warps per TMEM lane quarter
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const unsigned taddr_row = tmem_acc + ((unsigned)(q * 32) << 16);
    const bool have_odd = n_kb > 1;
#pragma unroll 1
    for (int c = half * (BN / PARTS); c < (half + 1) * (BN / PARTS); c += 8) {
        unsigned r0[8], r1[8], r2[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7])
                     : "r"(taddr_row + (unsigned)c));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7])
                     : "r"(taddr_row + (unsigned)(2 * BN + c)));
        if (have_odd) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7])
                         : "r"(taddr_row + (unsigned)(BN + c)));
        } else {
#pragma unroll
            for (int z = 0; z < 8; ++z) r1[z] = 0u;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < a.M) {
            float v[8], mk[8];
            if (a.mask) {
                const float4* mp = (const float4*)(a.mask + (size_t)row * a.ldm + n0 + c);
                const float4 m0v = __ldg(mp), m1v = __ldg(mp + 1);
                mk[0] = m0v.x; mk[1] = m0v.y; mk[2] = m0v.z; mk[3] = m0v.w;
                mk[4] = m1v.x; mk[5] = m1v.y; mk[6] = m1v.z; mk[7] = m1v.w;
            }
#pragma unroll
            for (int z = 0; z < 8; ++z) {
                const int col = n0 + c + z;
                float x = (__uint_as_float(r0[z]) + __uint_as_float(r1[z])) + __uint_as_float(r2[z]);
                if (a.bias) x += __ldg(a.bias + col);
                if (a.relu) x = fmaxf(x, 0.f);
                if (a.mask) x = (mk[z] > 0.f) ? x : 0.f;
                v[z] = x;
            }
            float* y = a.Y + (size_t)row * a.ldy + n0 + c;
            *(float4*)y = make_float4(v[0], v[1], v[2], v[3]);
            *(float4*)(y + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        constexpr unsigned cols = 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(cols) : "memory");
    }
}

template <int BN>
int launch_tc(const TcArgs& a, cudaStream_t st) {
    constexpr size_t smem = (size_t)tc_stages(BN) * (2 * TC_M * TC_BK + 2 * BN * TC_BK) * sizeof(float);
    auto k = linear_tc_kernel<BN>;
    static bool configured[16] = {};  // per device (function attributes are per context) and per BN instantiation
    int dev = 0;
    BRS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16 || !configured[dev]) {
        BRS_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 16) configured[dev] = true;
    }
    if ((a.M + TC_M - 1) / TC_M > 65535) return BRS_ERR_UNSUPPORTED;
    dim3 grid(a.N_total / BN, (a.M + TC_M - 1) / TC_M);
    k<<<grid, TC_THREADS, smem, st>>>(a);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

}  // namespace

// shapes the tensor-core path covers; everything else stays on the fp32 FFMA kernels
bool brs_linear_tc_supported(int M, int N, int K) {
    return M > 0 && K % TC_BK == 0 && K >= TC_BK && (N % 64 == 0) && N >= 64;
}

// Y = epilogue(A[M,K] . B[N,K]^T); lda = ldb = K, ldy = ldm = N
int brs_linear_tc(const float* A, const float* B, const float* bias, float* Y, const float* mask, int M, int N, int K,
                  bool relu, cudaStream_t st) {
    if (!brs_linear_tc_supported(M, N, K)) return BRS_ERR_UNSUPPORTED;
    if ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)Y) & 15) != 0) return BRS_ERR_INVALID_ARG;
    TcArgs a;
    a.A = A; a.lda = K;
    a.B = B; a.ldb = K;
    a.Y = Y; a.ldy = N;
    a.bias = bias;
    a.mask = mask; a.ldm = N;
    a.M = M; a.N_total = N; a.K = K;
    a.relu = relu ? 1 : 0;
    if (N % 128 == 0) return launch_tc<128>(a, st);  // 3 accumulators x 128 columns of the 512 TMEM columns
    return launch_tc<64>(a, st);
}

// [R, C] row-major -> [C, R] row-major (weights only: <= a few hundred KB)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

int brs_transpose(const float* src, float* dst, int R, int C, cudaStream_t st) {
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(src, dst, R, C);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
