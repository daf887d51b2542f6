// Materialised kernel expectations: Psi1 = eKxz [N,M] and Psi2 = eKzxKxz [N,M1,M2].
//
// Psi2 replaces `_E`, gpflow_pilco/utils/kernel_expectation.py:72-187.  The reference builds >= 4 [N,M1,M2]
// temporaries (tile of Z, two triangular solves, a K=D matmul, two exp tensors); here each entry is produced
// once from  log Q_ij = r_i + s_j + z1'_i^T R z2'_j  (common.cuh) and written with 128-bit coalesced stores.
// The kernel is HBM-write bound (8 B/entry); the FP64 pipe (KS/32 DMMA + 8 FP64 ops per entry) is a close second.
#include "mma_exp.cuh"

namespace gpp {

template <int D>
__global__ void __launch_bounds__(128) k_ekxz(const double* __restrict__ mu, const double* __restrict__ cov, int N, int M,
                                              const double* __restrict__ Z, const double* __restrict__ ell, double variance,
                                              double* __restrict__ out, int* info) {
  const int n = blockIdx.y;
  Mat<D> A, Li;
  double m[D], half_log_v = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double e = ell[d];
    half_log_v += log(e);
    m[d] = mu[(size_t)n * D + d];
#pragma unroll
    for (int e2 = 0; e2 < D; ++e2) A(d, e2) = cov[(size_t)n * D * D + d * D + e2] + (d == e2 ? e * e : 0.0);
  }
  bool ok = cholesky<D>(A);
  if (!ok && threadIdx.x == 0 && blockIdx.x == 0) flag_not_pd(info, n);
  double log_det = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) log_det += log(A(d, d));
  tri_inverse<D>(A, Li);
  const double c0 = log(variance) + half_log_v - log_det;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
    double dz[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dz[d] = Z[(size_t)j * D + d] - m[d];
    double maha = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double y = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) y = fma(Li(i, k), dz[k], y);
      maha = fma(y, y, maha);
    }
    out[(size_t)n * M + j] = fast_exp(c0 - 0.5 * maha);
  }
}

template <int D>
__global__ void __launch_bounds__(64) k_pack_single(const double* __restrict__ mu, const double* __restrict__ cov, int N,
                                                    const double* __restrict__ ell1, const double* __restrict__ ell2,
                                                    double log_amp, double* __restrict__ packs, int* info) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double V1[D], V2[D], m[D], Sg[D * D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    V1[d] = ell1[d] * ell1[d];
    V2[d] = ell2[d] * ell2[d];
    m[d] = mu[(size_t)n * D + d];
  }
#pragma unroll
  for (int d = 0; d < D * D; ++d) Sg[d] = cov[(size_t)n * D * D + d];
  double out[PairPack<D>::SIZE];
  if (!make_pair_pack<D>(m, Sg, V1, V2, log_amp, out)) flag_not_pd(info, n);
#pragma unroll
  for (int t = 0; t < PairPack<D>::SIZE; ++t) packs[(size_t)n * PairPack<D>::SIZE + t] = out[t];
}

// Psi2, materialised.  CTA = (input n, block of 128 rows), 16 warps; warp w owns the 8-row strip w and walks all columns in
// blocks of 128.  With the extended vectors A_i, B_j (mma_exp.cuh) the exponent of an 8 x 8 block of entries is KS
// mma.sync.m8n8k4.f64 (DMMA) instructions; each lane then holds two horizontally adjacent entries: table exp (8 FP64 ops each)
// and one 128-bit streaming store.  Rows are prepared once per CTA, columns once per column block by all 512 threads
// (4 threads per column share the quadratic form).  Measured: DMMA runs at the DFMA rate on B200 but needs 4 register operands
// per 512 flop instead of 3 per 2 (scripts/microbench/dmma.cu, fp64_pipe.cu).
constexpr int kPsi2Rows = 128, kPsi2Cols = 128, kPsi2Threads = 512;

template <int D>
struct Psi2Cfg {
  // when carrying (r_i, 1).(1, s_j) inside the inner product would cost a whole extra k-step (D = 3, 4, 7, 8), the two scalars are
  // added into the DMMA accumulator instead (1 DADD per entry)
  static constexpr bool EXCL = (D + 3) / 4 < ExtLayout<D>::KS;
  static constexpr int KS = EXCL ? (D + 3) / 4 : ExtLayout<D>::KS;
  static constexpr int REP = 16;
  static constexpr int PK = 0;                                   // coefficient pack
  static constexpr int ROWA = (PairPack<D>::SIZE + 1) & ~1;      // [KS][128][4]
  static constexpr int COLB = ROWA + KS * kPsi2Rows * 4;         // [2][KS][128][4]
  static constexpr int ETAB = COLB + 2 * KS * kPsi2Cols * 4;     // [256][REP]
  static constexpr int ROWR = ETAB + 256 * REP;                  // [128] r_i        (EXCL only)
  static constexpr int COLS = ROWR + kPsi2Rows;                  // [2][128] s_j     (EXCL only)
  static constexpr int TOTAL = COLS + 2 * kPsi2Cols;             // doubles
};

// B vectors (and, in the EXCL layout, s_j) of every column block of every input, written once as images of the main kernel's
// shared-memory column buffer [KS][128][4] (+ [128]); the 16 row-block CTAs of an input then copy instead of recomputing them.
// CTA = (column block, input): thread = (column, quarter); the quarters share the quadratic form z2'^T P2 z2' by shuffle.
template <int D>
__global__ void __launch_bounds__(kPsi2Threads) k_psi2_cols(const double* __restrict__ packs, const double* __restrict__ Z2, int M2,
                                                            double* __restrict__ colext) {
  using PP = PairPack<D>;
  using CF = Psi2Cfg<D>;
  constexpr int KS = CF::KS, FB = KS * kPsi2Cols * 4, IMG = FB + kPsi2Cols;
  constexpr bool EXCL = CF::EXCL;
  __shared__ double pk[PP::SIZE];
  const int cbk = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  for (int t = tid; t < PP::SIZE; t += kPsi2Threads) pk[t] = packs[(size_t)n * PP::SIZE + t];
  __syncthreads();
  double* img = colext + ((size_t)n * gridDim.x + cbk) * IMG;
  const int jl = tid >> 2, q = tid & 3, j = cbk * kPsi2Cols + jl;
  double zc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) zc[d] = (j < M2 ? Z2[(size_t)j * D + d] : 0.0) - pk[PP::MU + d];
  double part = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if ((d & 3) == q) {
      double rowsum = 0.0;
#pragma unroll
      for (int e = d; e < D; ++e) rowsum = fma(pk[PP::P2 + d * D - d * (d - 1) / 2 + (e - d)], zc[e], rowsum);
      part = fma(rowsum, zc[d], part);
    }
  }
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  if (q < KS) {
    double* dst = img + (q * kPsi2Cols + jl) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int e = 4 * q + c;                 // B_j = [z2' (D), 1, s_j, 0..]
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) v = (e == d) ? zc[d] : v;
      if (!EXCL) v = (e == D) ? 1.0 : ((e == D + 1) ? part : v);
      dst[c] = v;
    }
  }
  if (q == 3) img[FB + jl] = part;
}

template <int D>
__global__ void __launch_bounds__(kPsi2Threads, 2) k_ekzxkxz(const double* __restrict__ packs, const double* __restrict__ Z1, int M1,
                                                             const double* __restrict__ colext, int M2, double* __restrict__ out) {
  using PP = PairPack<D>;
  using CF = Psi2Cfg<D>;
  constexpr int KS = CF::KS, FB = KS * kPsi2Cols * 4;
  extern __shared__ __align__(16) double smem[];
  double* pk = smem + CF::PK;
  double* rowA = smem + CF::ROWA;
  double* colB = smem + CF::COLB;
  double* etab = smem + CF::ETAB;
  double* rowR = smem + CF::ROWR;
  double* colS = smem + CF::COLS;
  constexpr bool EXCL = CF::EXCL;
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kPsi2Rows;
  for (int t = tid; t < PP::SIZE; t += kPsi2Threads) pk[t] = packs[(size_t)n * PP::SIZE + t];
  for (int t = tid; t < 256 * CF::REP; t += kPsi2Threads) etab[t] = kExp2Tab256[t / CF::REP];
  __syncthreads();
  if (tid < kPsi2Rows) {   // A_i of row i0 + tid
    const int i = i0 + tid;
    double zc[D], ext[4 * KS];
#pragma unroll
    for (int d = 0; d < D; ++d) zc[d] = (i < M1 ? Z1[(size_t)i * D + d] : 0.0) - pk[PP::MU + d];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) t = fma(zc[d], pk[PP::R + d * D + e], t);
      ext[e] = t;
    }
    const double ri = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zc);
    if (EXCL) {
      rowR[tid] = ri;
#pragma unroll
      for (int e = D; e < 4 * KS; ++e) ext[e] = 0.0;
    } else {
      ext[D] = ri;
      ext[D + 1] = 1.0;
#pragma unroll
      for (int e = D + 2; e < 4 * KS; ++e) ext[e] = 0.0;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) rowA[(ks * kPsi2Rows + tid) * 4 + q] = ext[ks * 4 + q];
  }
  // column block `cbk` -> colB[buf] / colS[buf]: copy the image k_psi2_cols prepared (L2-resident: shared by the input's row blocks)
  const int ncb = (M2 + kPsi2Cols - 1) / kPsi2Cols;
  auto prepare_columns = [&](int cbk, int buf) {
    constexpr int IMG = FB + kPsi2Cols;
    const double* img = colext + ((size_t)n * ncb + cbk) * IMG;
    for (int t = tid; t < FB; t += kPsi2Threads) colB[buf * FB + t] = img[t];
    if (EXCL && tid < kPsi2Cols) colS[buf * kPsi2Cols + tid] = img[FB + tid];
  };
  prepare_columns(0, 0);
  __syncthreads();
  double a[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) a[ks] = rowA[(ks * kPsi2Rows + warp * 8) * 4 + lane];
  const int row = i0 + warp * 8 + (lane >> 2);
  const int cpair = 2 * (lane & 3);
  const double ri = EXCL ? rowR[warp * 8 + (lane >> 2)] : 0.0;
  const unsigned etab_lane = (unsigned)__cvta_generic_to_shared(etab + (lane & (CF::REP - 1)));
  const bool pair_ok = ((M2 & 1) == 0);          // 16-byte aligned pair stores need an even row length
  double* outrow = out + ((size_t)n * M1 + row) * M2;
  for (int cbk = 0; cbk < ncb; ++cbk) {
    const int buf = cbk & 1;
    if (cbk + 1 < ncb) prepare_columns(cbk + 1, buf ^ 1);      // next block's columns into the other buffer
    const double* cb = colB + buf * FB + lane;
#pragma unroll 2
    for (int cg = 0; cg < kPsi2Cols / 8; cg += 2) {
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      if (EXCL) {
        const double2 s0 = *reinterpret_cast<const double2*>(colS + buf * kPsi2Cols + cg * 8 + cpair);
        const double2 s1 = *reinterpret_cast<const double2*>(colS + buf * kPsi2Cols + cg * 8 + 8 + cpair);
        t[0] = ri + s0.x; t[1] = ri + s0.y; t[2] = ri + s1.x; t[3] = ri + s1.y;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        dmma_m8n8k4(t[0], t[1], a[ks], cb[(ks * kPsi2Cols + cg * 8) * 4]);
        dmma_m8n8k4(t[2], t[3], a[ks], cb[(ks * kPsi2Cols + cg * 8 + 8) * 4]);
      }
      bool tiny[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) tiny[k] = (unsigned)__double2hiint(t[k]) > 0xC0861800u;   // log Q < -707: exactly 0
      exp_tab_contract<4, CF::REP>(t, etab_lane);
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = tiny[k] ? 0.0 : t[k];
      if (row < M1) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = cbk * kPsi2Cols + (cg + u) * 8 + cpair;
          if (pair_ok && j + 1 < M2) {
            __stcs(reinterpret_cast<double2*>(outrow + j), make_double2(t[2 * u], t[2 * u + 1]));   // streaming: never re-read
          } else {
            if (j < M2) __stcs(outrow + j, t[2 * u]);
            if (j + 1 < M2) __stcs(outrow + j + 1, t[2 * u + 1]);
          }
        }
      }
    }
    __syncthreads();   // next block's columns complete; this block's buffer free for the block after next
  }
}

template <int D>
static int run_ekxz(const double* mu, const double* cov, int N, const double* Z, int M, const double* ell, double variance,
                    double* out, int* info, cudaStream_t stream) {
  dim3 grid(std::min((M + 127) / 128, 64), N);
  k_ekxz<D><<<grid, 128, 0, stream>>>(mu, cov, N, M, Z, ell, variance, out, info);
  count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

template <int D>
static int run_ekzxkxz(const double* mu, const double* cov, int N, const double* Z1, int M1, const double* ell1, double var1,
                       const double* Z2, int M2, const double* ell2, double var2, double* out, int* info,
                       cudaStream_t stream) {
  // the per-call scratch (coefficient packs, column vectors: tens of MB) comes from the stream-ordered pool; with the default release
  // threshold of 0 the pool hands everything back to the driver at every synchronisation and each call pays a real cudaMalloc
  // (measured: 3.2 ms per call against a 1.7 ms kernel) — keep the pool's memory
  static bool pool_configured = false;
  if (!pool_configured) {
    int dev = 0;
    cudaMemPool_t pool;
    GPP_CUDA_OK(cudaGetDevice(&dev));
    GPP_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    GPP_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_configured = true;
  }
  double* packs = nullptr;
  GPP_CUDA_OK(cudaMallocAsync(&packs, sizeof(double) * PairPack<D>::SIZE * (size_t)N, stream));
  k_pack_single<D><<<(N + 63) / 64, 64, 0, stream>>>(mu, cov, N, ell1, ell2, log(var1 * var2), packs, info);
  dim3 grid((M1 + kPsi2Rows - 1) / kPsi2Rows, N);
  const int ncb = (M2 + kPsi2Cols - 1) / kPsi2Cols;
  const size_t img = (size_t)Psi2Cfg<D>::KS * kPsi2Cols * 4 + kPsi2Cols;
  double* colext = nullptr;
  GPP_CUDA_OK(cudaMallocAsync(&colext, sizeof(double) * img * ncb * (size_t)N, stream));
  k_psi2_cols<D><<<dim3(ncb, N), kPsi2Threads, 0, stream>>>(packs, Z2, M2, colext);
  count_launch();
  const size_t smem = sizeof(double) * Psi2Cfg<D>::TOTAL;
  static PerDeviceSmemOptIn configured;
  if (configured.raise(smem)) {
    GPP_CUDA_OK(cudaFuncSetAttribute(k_ekzxkxz<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  profile_begin(stream);
  k_ekzxkxz<D><<<grid, kPsi2Threads, smem, stream>>>(packs, Z1, M1, colext, M2, out);
  profile_end(stream);
  count_launch(2);
  GPP_CUDA_OK(cudaGetLastError());
  GPP_CUDA_OK(cudaFreeAsync(colext, stream));
  GPP_CUDA_OK(cudaFreeAsync(packs, stream));
  return GPP_OK;
}

}  // namespace gpp

extern "C" {

int gpp_ekxz(const double* mu, const double* cov, int N, int D, const double* Z, int M, const double* lengthscales,
             double variance, double* out, int* info, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(N >= 0 && M >= 0 && D >= 1, GPP_ERR_BAD_SHAPE, "gpp_ekxz: bad sizes N=%d M=%d D=%d", N, M, D);
  GPP_REQUIRE(N <= 65535, GPP_ERR_UNSUPPORTED, "gpp_ekxz: N=%d > 65535 inputs per call", N);
  if (N == 0 || M == 0) return GPP_OK;      // empty batch: nothing to do (the array pointers may be null)
  GPP_REQUIRE(mu && cov && Z && lengthscales && out, GPP_ERR_NULL, "gpp_ekxz: null argument");
  switch (D) {
#define GPP_CASE(d) case d: return gpp::run_ekxz<d>(mu, cov, N, Z, M, lengthscales, variance, out, info, (cudaStream_t)stream);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: gpp::set_error("gpp_ekxz: unsupported D=%d (max %d)", D, GPP_MAX_D); return GPP_ERR_UNSUPPORTED;
  }
}

int gpp_ekzxkxz(const double* mu, const double* cov, int N, int D, const double* Z1, int M1, const double* lengthscales1,
                double variance1, const double* Z2, int M2, const double* lengthscales2, double variance2, double* out,
                int* info, void* stream) {
  GPP_NVTX_RANGE();
  if (!Z2) { Z2 = Z1; M2 = M1; }
  if (!lengthscales2) { lengthscales2 = lengthscales1; variance2 = variance1; }
  GPP_REQUIRE(N >= 0 && M1 >= 0 && M2 >= 0 && D >= 1, GPP_ERR_BAD_SHAPE, "gpp_ekzxkxz: bad sizes N=%d M1=%d M2=%d D=%d", N, M1, M2, D);
  GPP_REQUIRE(N <= 65535, GPP_ERR_UNSUPPORTED, "gpp_ekzxkxz: N=%d > 65535 inputs per call", N);
  if (N == 0 || M1 == 0 || M2 == 0) return GPP_OK;      // empty batch: nothing to do (the array pointers may be null)
  GPP_REQUIRE(mu && cov && Z1 && lengthscales1 && out, GPP_ERR_NULL, "gpp_ekzxkxz: null argument");
  switch (D) {
#define GPP_CASE(d) \
  case d: return gpp::run_ekzxkxz<d>(mu, cov, N, Z1, M1, lengthscales1, variance1, Z2, M2, lengthscales2, variance2, out, info, (cudaStream_t)stream);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: gpp::set_error("gpp_ekzxkxz: unsupported D=%d (max %d)", D, GPP_MAX_D); return GPP_ERR_UNSUPPORTED;
  }
}

}  // extern "C"
