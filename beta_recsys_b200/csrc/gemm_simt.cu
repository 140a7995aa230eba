// fp32 (FFMA) Linear-layer kernels for the NCF tower -- sm_100a.
//
// Replaces nn.Linear forward/backward (cuBLAS SGEMM + separate bias/ReLU
// launches in the reference: beta_rec/models/ncf.py:64-69, mlp.py:47-49):
//   fwd : Y[M,N]  = act(X[M,K] . W[N,K]^T + b[N])              (bias + ReLU fused)
//   dgrad: dX[M,K] = (dY[M,N] . W[N,K]) * (Xact[M,K] > 0)       (ReLU mask of the layer input fused)
//   wgrad: dW[N,K] += dY[M,N]^T . X[M,K],  db[N] += sum_m dY    (split over the batch, RED-accumulated)
// Exact fp32 products and accumulation: this is the path parity is pinned on; the
// tcgen05 3xTF32 kernels (gemm_tc.cu) are checked against it.
// 128x64 block tile, 16-deep k slices, 8x4 register tile per thread, register
// prefetch of the next slice.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int TM = 8, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256
static_assert(kThreads == 256, "tile config");

// C[M,N] = A[M,Kd] * B  with B given either as [N,Kd] (B_T = true: "NT", Linear forward)
// or as [Kd,N] (B_T = false: "NN", dgrad).  lda/ldb/ldc are row strides in floats.
template <bool B_T, bool RELU, bool BIAS, bool MASK>
__global__ void __launch_bounds__(kThreads) gemm_rowmajor_kernel(const float* __restrict__ A, int lda,
                                                                 const float* __restrict__ B, int ldb,
                                                                 float* __restrict__ C, int ldc,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ mask_src, int ldm, int M,
                                                                 int N, int Kd) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);  // 16 x 16

    // global -> smem assignments
    // A tile: BM x BK = 128 x 16 floats: each thread loads 2 float4 along k
    const int a_row = tid / 4, a_k4 = (tid % 4) * 4;  // rows 0..63 (+64 for second)
    // B tile (NT): BN x BK = 64 x 16: one float4 along k per thread
    const int bt_row = tid / 4, bt_k4 = (tid % 4) * 4;
    // B tile (NN): BK x BN = 16 x 64: one float4 along n per thread
    const int bn_k = tid / 16, bn_n4 = (tid % 16) * 4;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb;
    auto load_global = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = m0 + a_row + h * 64, k = k0 + a_k4;
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < M) {
                if (k + 3 < Kd && (lda & 3) == 0)
                    ra[h] = *(const float4*)(A + (long long)r * lda + k);
                else {
                    const float* p = A + (long long)r * lda;
                    if (k < Kd) ra[h].x = p[k];
                    if (k + 1 < Kd) ra[h].y = p[k + 1];
                    if (k + 2 < Kd) ra[h].z = p[k + 2];
                    if (k + 3 < Kd) ra[h].w = p[k + 3];
                }
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (B_T) {
            const int r = n0 + bt_row, k = k0 + bt_k4;
            if (bt_row < BN && r < N) {
                if (k + 3 < Kd && (ldb & 3) == 0)
                    rb = *(const float4*)(B + (long long)r * ldb + k);
                else {
                    const float* p = B + (long long)r * ldb;
                    if (k < Kd) rb.x = p[k];
                    if (k + 1 < Kd) rb.y = p[k + 1];
                    if (k + 2 < Kd) rb.z = p[k + 2];
                    if (k + 3 < Kd) rb.w = p[k + 3];
                }
            }
        } else {
            const int k = k0 + bn_k, n = n0 + bn_n4;
            if (k < Kd) {
                if (n + 3 < N && (ldb & 3) == 0)
                    rb = *(const float4*)(B + (long long)k * ldb + n);
                else {
                    const float* p = B + (long long)k * ldb;
                    if (n < N) rb.x = p[n];
                    if (n + 1 < N) rb.y = p[n + 1];
                    if (n + 2 < N) rb.z = p[n + 2];
                    if (n + 3 < N) rb.w = p[n + 3];
                }
            }
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = a_row + h * 64;
            As[buf][a_k4 + 0][r] = ra[h].x;
            As[buf][a_k4 + 1][r] = ra[h].y;
            As[buf][a_k4 + 2][r] = ra[h].z;
            As[buf][a_k4 + 3][r] = ra[h].w;
        }
        if (B_T) {
            if (bt_row < BN) {
                Bs[buf][bt_k4 + 0][bt_row] = rb.x;
                Bs[buf][bt_k4 + 1][bt_row] = rb.y;
                Bs[buf][bt_k4 + 2][bt_row] = rb.z;
                Bs[buf][bt_k4 + 3][bt_row] = rb.w;
            }
        } else {
            *(float4*)&Bs[buf][bn_k][bn_n4] = rb;
        }
    };

    const int n_k = (Kd + BK - 1) / BK;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int kt = 0; kt < n_k; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_k) load_global((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *(const float4*)&As[buf][k][ty * TM];
            const float4 a1 = *(const float4*)&As[buf][k][ty * TM + 4];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            const float4 b0 = *(const float4*)&Bs[buf][k][tx * TN];
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < n_k) store_smem(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = m0 + ty * TM + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = n0 + tx * TN + j;
            if (c >= N) continue;
            float v = acc[i][j];
            if (BIAS) v += bias[c];
            if (RELU) v = fmaxf(v, 0.f);
            if (MASK) v = (mask_src[(long long)r * ldm + c] > 0.f) ? v : 0.f;
            C[(long long)r * ldc + c] = v;
        }
    }
}

// wgrad: dW[N,K] += dY[M,N]^T X[M,K] over the M rows [m_begin, m_end) of this block; db[N] += colsum(dY).
// 64x64 output tile, 4x4 per thread, reduction slices of 16 batch rows.
constexpr int WT = 64, WK = 16;
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X,
                                                    int ldx, float* __restrict__ dW, int ldw, float* __restrict__ db,
                                                    int M, int N, int K, int m_chunk) {
    __shared__ float Ys[WK][WT + 4];
    __shared__ float Xs[WK][WT + 4];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * WT, k0 = blockIdx.y * WT;
    const int m_begin = blockIdx.z * m_chunk;
    const int m_end = min(M, m_begin + m_chunk);
    const int tx = tid % 16, ty = tid / 16;  // ty -> n, tx -> k
    const int l_m = tid / 16, l_c4 = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;  // column sum of dY for column n0 + tid (threads 0..63 of k-block 0)

    for (int m0 = m_begin; m0 < m_end; m0 += WK) {
        const int m = m0 + l_m;
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f), x = y;
        if (m < m_end) {
            const float* py = dY + (long long)m * ldy;
            const float* px = X + (long long)m * ldx;
            const int n = n0 + l_c4, k = k0 + l_c4;
            if (n + 3 < N && (ldy & 3) == 0) y = *(const float4*)(py + n);
            else {
                if (n < N) y.x = py[n];
                if (n + 1 < N) y.y = py[n + 1];
                if (n + 2 < N) y.z = py[n + 2];
                if (n + 3 < N) y.w = py[n + 3];
            }
            if (k + 3 < K && (ldx & 3) == 0) x = *(const float4*)(px + k);
            else {
                if (k < K) x.x = px[k];
                if (k + 1 < K) x.y = px[k + 1];
                if (k + 2 < K) x.z = px[k + 2];
                if (k + 3 < K) x.w = px[k + 3];
            }
        }
        __syncthreads();
        *(float4*)&Ys[l_m][l_c4] = y;
        *(float4*)&Xs[l_m][l_c4] = x;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WK; ++r) {
            const float4 a = *(const float4*)&Ys[r][ty * 4];
            const float4 b = *(const float4*)&Xs[r][tx * 4];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (db && blockIdx.y == 0 && tid < WT) {
#pragma unroll
            for (int r = 0; r < WK; ++r) bsum += Ys[r][tid];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) red_add1(dW + (long long)n * ldw + k, acc[i][j]);
        }
    }
    if (db && blockIdx.y == 0 && tid < WT && n0 + tid < N) red_add1(db + n0 + tid, bsum);
}

}  // namespace

// ---- host-side launchers shared with ncf_kernels.cu / abi ----
int brs_linear_fwd_simt(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int M, int N, int K,
                        bool relu, cudaStream_t st) {
    if (M <= 0) return BRS_OK;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    if (relu)
        gemm_rowmajor_kernel<true, true, true, false><<<grid, kThreads, 0, st>>>(X, ldx, W, K, Y, ldy, b, nullptr, 0, M, N, K);
    else
        gemm_rowmajor_kernel<true, false, true, false><<<grid, kThreads, 0, st>>>(X, ldx, W, K, Y, ldy, b, nullptr, 0, M, N, K);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// dX[M,K] = (dY[M,N] . W[N,K]) (* (mask_src[M,K] > 0) when mask_src != NULL)
int brs_linear_dgrad_simt(const float* dY, int ldy, const float* W, float* dX, int ldx, const float* mask_src, int ldm,
                          int M, int N, int K, cudaStream_t st) {
    if (M <= 0) return BRS_OK;
    dim3 grid((M + BM - 1) / BM, (K + BN - 1) / BN);
    if (mask_src)
        gemm_rowmajor_kernel<false, false, false, true><<<grid, kThreads, 0, st>>>(dY, ldy, W, K, dX, ldx, nullptr, mask_src, ldm, M, K, N);
    else
        gemm_rowmajor_kernel<false, false, false, false><<<grid, kThreads, 0, st>>>(dY, ldy, W, K, dX, ldx, nullptr, nullptr, 0, M, K, N);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}

// dW[N,K] += dY^T X ; db[N] += colsum(dY)   (dW/db must be zero before the step)
int brs_linear_wgrad_simt(const float* dY, int ldy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                          cudaStream_t st) {
    if (M <= 0) return BRS_OK;
    const int tiles = ((N + WT - 1) / WT) * ((K + WT - 1) / WT);
    int splits = (brs_sm_count() * 4 + tiles - 1) / tiles;  // ~4 blocks per SM in total
    int m_chunk = (M + splits - 1) / splits;
    m_chunk = ((m_chunk + WK - 1) / WK) * WK;
    if (m_chunk < WK) m_chunk = WK;
    splits = (M + m_chunk - 1) / m_chunk;
    dim3 grid((N + WT - 1) / WT, (K + WT - 1) / WT, splits);
    wgrad_kernel<<<grid, 256, 0, st>>>(dY, ldy, X, ldx, dW, K, db, M, N, K, m_chunk);
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
