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
//   * operands are written to shared memory in the canonical no-swizzle K-major core-matrix layout
//     (8 rows x 16 bytes per core matrix) -- plain st.shared, no tensor maps needed;
//   * thread 0 issues tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x BN x 8), accumulator in TMEM;
//     tcgen05.commit -> mbarrier tells the loaders when a stage may be overwritten (2 stages);
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld (32x32b.x8), applies
//     bias / ReLU / mask and stores rows to global memory.
#include "common.cuh"

namespace {

constexpr int TC_M = 128;      // rows per CTA tile (UMMA_M)
constexpr int TC_BK = 32;      // fp32 elements of K per stage (4 MMA k-steps of 8)
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 2;

__device__ __forceinline__ float to_tf32(float x) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
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

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major:
//   start address >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 | version 1 at bit 46
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
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

// element (row, k) of a [ROWS x 32] fp32 tile in the canonical K-major no-swizzle layout:
// [k/4][row/8][row%8][k%4]  ->  LBO (next 16-byte k chunk) = ROWS*16 B, SBO (next 8-row group) = 128 B
__device__ __forceinline__ int tile_off(int rows, int row, int k4) { return k4 * (rows * 4) + (row >> 3) * 32 + (row & 7) * 4; }

// stage a [ROWS x 32] block of a row-major matrix (leading dim ld) as hi / lo tf32 tiles
template <int ROWS>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int ld, int row0, int n_rows_total, int k0,
                                           float* __restrict__ hi, float* __restrict__ lo) {
    // a quarter-warp covers the 8 rows of one core-matrix column (128 contiguous bytes of shared memory:
    // conflict-free stores); the four quarter-warps take four adjacent 16-byte k chunks of the same rows
    // (two full 32-byte sectors per row on the global side)
    const int t = threadIdx.x;
    const int lane = t & 31, w = t >> 5;
    const int rsub = lane & 7, ksub = lane >> 3;
    constexpr int NW = TC_THREADS / 32;
#pragma unroll 2
    for (int it = w; it < (ROWS / 8) * 2; it += NW) {
        const int r = (it >> 1) * 8 + rsub;
        const int k4 = (it & 1) * 4 + ksub;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gr = row0 + r;
        if (gr < n_rows_total) v = *(const float4*)(src + (size_t)gr * ld + k0 + k4 * 4);
        float4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
        const int o = tile_off(ROWS, r, k4);
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
    float* smem = (float*)smem_raw;
    __shared__ __align__(8) uint64_t s_bar[TC_STAGES];  // MMAs reading stage s have completed
    __shared__ __align__(8) uint64_t s_done;            // all MMAs of the tile have completed
    __shared__ unsigned s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_M, n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&s_bar[s], 1);
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
    unsigned stage_phase = 0;  // bit s: parity to wait for before refilling stage s
    for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb & 1;
        float* a_hi = smem + s * STAGE_FLOATS;
        float* a_lo = a_hi + A_FLOATS;
        float* b_hi = a_lo + A_FLOATS;
        float* b_lo = b_hi + B_FLOATS;
        if (kb >= TC_STAGES) {  // the MMAs that read this stage two k-blocks ago must be done
            mbar_wait_or_trap(&s_bar[s], (stage_phase >> s) & 1u);
            stage_phase ^= 1u << s;
        }
        stage_tile<TC_M>(a.A, a.lda, m0, a.M, kb * TC_BK, a_hi, a_lo);
        stage_tile<BN>(a.B, a.ldb, n0, a.N_total, kb * TC_BK, b_hi, b_lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy (MMA) reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {  // one UMMA k-step = 8 tf32 = two 16-byte chunks
                const unsigned a_off = 2 * j * (TC_M * 16), b_off = 2 * j * (BN * 16);
                const unsigned long long ah = make_smem_desc(a_hi_s + a_off, TC_M * 16, 128);
                const unsigned long long al = make_smem_desc(a_lo_s + a_off, TC_M * 16, 128);
                const unsigned long long bh = make_smem_desc(b_hi_s + b_off, BN * 16, 128);
                const unsigned long long bl = make_smem_desc(b_lo_s + b_off, BN * 16, 128);
                tc_mma_tf32(tmem_acc + (unsigned)((kb & 1) * BN), ah, bh, idesc, ((kb >> 1) | j) ? 1u : 0u);
                tc_mma_tf32(tmem_acc + (unsigned)(2 * BN), ah, bl, idesc, (kb | j) ? 1u : 0u);
                tc_mma_tf32(tmem_acc + (unsigned)(2 * BN), al, bh, idesc, 1u);
            }
            tc_commit(&s_bar[s]);                       // stage s reusable once these MMAs finish
            if (kb == n_kb - 1) tc_commit(&s_done);     // accumulator complete
        }
    }
    mbar_wait_or_trap(&s_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: a warp may only touch the TMEM lane quarter (warp % 4); the two warps of a quarter
    // split the columns.  y = acc_even + acc_odd + acc_cross, then bias / ReLU / mask.
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const unsigned taddr_row = tmem_acc + ((unsigned)(q * 32) << 16);
    const bool have_odd = n_kb > 1;
#pragma unroll 1
    for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 8) {
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
            float v[8];
#pragma unroll
            for (int z = 0; z < 8; ++z) {
                const int col = n0 + c + z;
                float x = (__uint_as_float(r0[z]) + __uint_as_float(r1[z])) + __uint_as_float(r2[z]);
                if (a.bias) x += __ldg(a.bias + col);
                if (a.relu) x = fmaxf(x, 0.f);
                if (a.mask) x = (a.mask[(size_t)row * a.ldm + col] > 0.f) ? x : 0.f;
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
    constexpr size_t smem = (size_t)TC_STAGES * (2 * TC_M * TC_BK + 2 * BN * TC_BK) * sizeof(float);
    auto k = linear_tc_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        BRS_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((a.M + TC_M - 1) / TC_M, a.N_total / BN);
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
