// Device code of the backward of the fused moment-matched GP predict: the gradient contraction
// (contract_grad_item: one (input, unordered kernel pair, 128-row block) work item on 16 warps), the per-input prologue pieces
// (un-mixing of the output adjoints, Psi1 adjoints) and the per-input D x D finalize.  Shared by the stand-alone kernels of
// mm_predict_bwd.cu and the persistent reverse sweep of rollout_persist.cu.
#pragma once
#include <type_traits>

#include "mma_exp.cuh"
#include "model.cuh"
#include "persist_common.cuh"
#include "predict_kernels.cuh"

namespace gpp {

constexpr int kGradRows = 128, kGradCols = 128, kGradThreads = 512;

template <int D>
struct GradStats {
  static constexpr int TRI = D * (D + 1) / 2;
  static constexpr int S0 = 0, R1 = 1, R2 = 1 + D, X = 1 + D + TRI;
  static constexpr int SIZE = 1 + D + TRI + D * D;
};

template <int D>
struct GradCfg {
  static constexpr int KS = ExtLayout<D>::KS;
  static constexpr int REP = KS <= 2 ? 16 : 8;
  static constexpr int NW = kGradThreads / 32;
  static constexpr int PK = 0;                                   // coefficient pack
  static constexpr int ROWA = (PairPack<D>::SIZE + 1) & ~1;      // [KS][128][4]
  static constexpr int COLB = ROWA + KS * kGradRows * 4;         // [2][KS][128][4]
  static constexpr int COLW = COLB + 2 * KS * kGradCols * 4;     // [2][128] beta_b of the columns (off-diagonal pairs)
  static constexpr int ETAB = COLW + 2 * kGradCols;              // [256][REP]
  static constexpr int RED = ETAB + 256 * REP;                   // [8][GradStats::SIZE]
  static constexpr int CSUM = RED + (kGradThreads / 64) * GradStats<D>::SIZE;   // [2][16][128] per-warp column sums (off-diagonal pairs)
  static constexpr int COLA = CSUM + 2 * NW * kGradCols;         // [ncb * 128] column sums a'_j of this row block
  static constexpr int FIXED = COLA;                             // + ncb * 128 doubles at launch
};

// Statistics of one UNORDERED pair {a <= b}: CTA = (input n, pair, block of 128 rows of latent a), 16 warps, warp w = 8-row strip w,
// all columns (rows of latent b) in blocks of 128.  Same DMMA scheme as the forward kernels (mma_exp.cuh): the exponents of an
// 8 x 8 block are KS mma.sync.m8n8k4.f64, each lane then holds entries of ONE row, and the row sums a_i = sum_j A_ij,
// u_i = sum_j A_ij z2'_j (A = C o Q, or beta_a beta_b^T o Q) accumulate on the tensor path as well.
// For a < b the statistics of the ordered pair (b, a) follow from the same entries (Q_ba = Q_ab^T, same centre mu):
//   S0_ba = S0_ab,  X_ba = X_ab^T,  r1_ba = sum_i u_i,  R2_ba = sum_j a'_j z2'_j z2'_j^T  with the COLUMN sums a'_j = sum_i A_ij,
// so only a'_j is extra: a 3-level transpose-reduce over the 8 row lanes (8 shuffles per 16 columns), one store per warp and
// column into shared memory, summed over the 16 strips in a fixed order.  The exponentials of (b, a) are never evaluated.
template <bool PERSIST>
__device__ __forceinline__ void grad_sync() {
  if (PERSIST) {                               // the 16 contraction warps of a persistent CTA (rollout_persist.cu)
    __syncwarp();
    asm volatile("bar.sync 7, 512;" ::: "memory");
  } else {
    __syncthreads();
  }
}
template <bool PERSIST>
__device__ __forceinline__ double grad_ld(const double* p) { return PERSIST ? __ldcg(p) : *p; }   // written by another CTA of the launch?

// `tid` = index among the 512 contraction threads (16 warps); `smem` = GradCfg<D> layout (+ ncb * 128 doubles).  With `fill_table`
// the exp table is (re)written first (once per CTA in the persistent kernel, once per item in k_contract_grad).  With ll_tag != 0
// (persistent sweep) `stats` is an array of tagged words, two per statistic, read by bwd_finalize_body without any fence.
// With `pack_given` the caller has already fetched the pair's coefficient pack (thread tid holds element tid in `pack_value`), checked
// that the pair's weight is not zero and synchronised the 512 threads once since the previous item.
template <int D, bool PERSIST>
__device__ void contract_grad_item(double* __restrict__ smem, const int tid, const int n, const int pr, const int rb,
                                   const double* __restrict__ Z, const double* __restrict__ beta, const double* __restrict__ C,
                                   const double* __restrict__ packs, const double* __restrict__ omega, double* __restrict__ stats,
                                   const int M, const int L, const int nrb, const bool fill_table, const unsigned ll_tag = 0u,
                                   const bool pack_given = false, const double pack_value = 0.0) {
  using PP = PairPack<D>;
  using GS = GradStats<D>;
  using CF = GradCfg<D>;
  constexpr int KS = CF::KS, FB = KS * kGradCols * 4;
  double* pk = smem + CF::PK;
  double* rowA = smem + CF::ROWA;
  double* colB = smem + CF::COLB;
  double* colW = smem + CF::COLW;
  double* etab = smem + CF::ETAB;
  double* red = smem + CF::RED;
  double* csum = smem + CF::CSUM;
  double* colA = smem + CF::COLA;
  const int lane = tid & 31, warp = tid >> 5;
  int a = 0, b = pr;                           // unordered pair index -> (a <= b), rows of the upper triangle in order
  while (b >= L - a) { b -= L - a; ++a; }
  b += a;
  const bool diag = a == b;
  const int p = a * L + b;
  double* out = stats + (((size_t)n * L * L + p) * nrb + rb) * GS::SIZE * (ll_tag ? 2 : 1);
  auto put = [&](double* base, int k, double v) {
    if (ll_tag) ll_store(reinterpret_cast<unsigned long long*>(base) + 2 * k, v, ll_tag);
    else base[k] = v;
  };
  // pairs whose output adjoint is zero (e.g. diagonal-only covariance) are skipped; k_bwd_finalize skips them too
  if (pack_given) {
    static_assert(PP::SIZE <= kGradThreads, "one pack element per thread");
    if (tid < PP::SIZE) pk[tid] = pack_value;
  } else {
    const double wgt = grad_ld<PERSIST>(omega + ((size_t)n * L + a) * L + b) + grad_ld<PERSIST>(omega + ((size_t)n * L + b) * L + a);
    if (wgt == 0.0) return;
    // (persistent caller: every thread passed the ticket barrier of persist_bwd_contract after it finished the previous item, so the
    //  previous item's readers of this scratch are done)
    for (int t = tid; t < PP::SIZE; t += kGradThreads) pk[t] = grad_ld<PERSIST>(packs + ((size_t)n * L * L + p) * PP::SIZE + t);
  }
  if (fill_table)
    for (int t = tid; t < 256 * CF::REP; t += kGradThreads) etab[t] = kExp2Tab256[t / CF::REP];
  grad_sync<PERSIST>();
  const int i0 = rb * kGradRows;
  if (tid < kGradRows) {   // A_i of row i0 + tid
    const int i = i0 + tid;
    double zc[D], ext[4 * KS];
#pragma unroll
    for (int d = 0; d < D; ++d) zc[d] = (i < M ? Z[((size_t)a * M + i) * D + d] : 0.0) - pk[PP::MU + d];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) t = fma(zc[d], pk[PP::R + d * D + e], t);
      ext[e] = t;
    }
    ext[D] = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zc);
    ext[D + 1] = 1.0;
#pragma unroll
    for (int e = D + 2; e < 4 * KS; ++e) ext[e] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) rowA[(ks * kGradRows + tid) * 4 + q] = ext[ks * 4 + q];
  }
  // column block -> colB[buf] (thread = (column, quarter), as in k_ekzxkxz) and the columns' beta_b -> colW[buf]
  auto prepare_columns = [&](int cbk, int buf) {
    const int jl = tid >> 2, q = tid & 3, j = cbk * kGradCols + jl;
    double zc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) zc[d] = (j < M ? Z[((size_t)b * M + j) * D + d] : 0.0) - pk[PP::MU + d];
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
      double* dst = colB + buf * FB + (q * kGradCols + jl) * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int e = 4 * q + c;                 // B_j = [z2' (D), 1, s_j, 0..]
        double v = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) v = (e == d) ? zc[d] : v;
        v = (e == D) ? 1.0 : ((e == D + 1) ? part : v);
        dst[c] = v;
      }
    }
    if (q == 3) colW[buf * kGradCols + jl] = (j < M) ? beta[(size_t)b * M + j] : 0.0;
  };
  // off-diagonal pairs: a'_j of column block cbk = sum over the 16 strips, in strip order
  auto fold_column_sums = [&](int cbk) {
    if (tid < kGradCols) {
      const double* src = csum + (cbk & 1) * CF::NW * kGradCols + tid;
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < CF::NW; ++w) s += src[w * kGradCols];
      colA[cbk * kGradCols + tid] = s;
    }
  };
  prepare_columns(0, 0);
  grad_sync<PERSIST>();
  double af[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) af[ks] = rowA[(ks * kGradRows + warp * 8) * 4 + lane];
  const int rloc = warp * 8 + (lane >> 2);
  const int row = i0 + rloc;
  const unsigned etab_lane = (unsigned)__cvta_generic_to_shared(etab + (lane & (CF::REP - 1)));
  const double* Crow = C + (size_t)a * M * M + (size_t)min(row, M - 1) * M;   // C symmetric: row `row`, contiguous over columns
  // row weight: off-diagonal pairs apply beta_a[row] inside the loop (the column sums need it); diagonal pairs only mask rows >= M
  const double rs = row < M ? (diag ? 1.0 : beta[(size_t)a * M + row]) : 0.0;
  // U = A Z2' accumulates on the tensor path too.  The exponent DMMA's output column n is mapped to the PHYSICAL column
  // pi(n) = (n >> 1) + 4 (n & 1) of the 8-column group, so that the two entries a lane (r, c) receives are A[r][c] and A[r][4 + c]:
  // exactly the A-operand fragments of the two k-slices of  U += A Z2'  — no layout conversion.  The B operand is Z2' as stored in
  // the columns' B vectors (lane (n, k) reads dimension n of column k).  Since B_j[D] = 1, the row sum a_i = sum_j A_ij is column
  // D of U (D <= 7).
  constexpr bool ROWSUM_IN_U = D <= 7;
  double u0 = 0.0, u1 = 0.0, ai = 0.0;       // U[row][2c], U[row][2c+1] (c = lane & 3), scalar row sum when D = 8
  const int c4 = lane & 3;
  const int nn = lane >> 2;
  const int boff = (((nn >> 1) + 4 * (nn & 1)) * 4) + c4;                // exponent B fragment: ext index k = c4 of column pi(n)
  const int zoff = ((lane >> 4) * kGradCols) * 4 + ((lane >> 2) & 3);    // U B fragment: dimension n -> k-step n >> 2, element n & 3
  const int ncb = (M + kGradCols - 1) / kGradCols;
  const bool b2 = lane & 4, b3 = lane & 8;
  const int csoff = warp * kGradCols + (b2 ? 8 : 0) + (b3 ? 4 : 0) + c4;    // this lane's column after the transpose-reduce
  auto main_loop = [&](auto diag_tag) {
    constexpr bool DIAG = decltype(diag_tag)::value;
    // diagonal pairs read their weights C_a[row][j] from global memory (L2): the 4 values of an iteration are fetched one iteration
    // ahead so that the load latency hides behind the DMMA + exp chain
    auto load_weights = [&](int cbk, int cg, double (&w)[4]) {
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const int j = cbk * kGradCols + (cg + uu) * 8 + c4;
        w[2 * uu] = j < M ? Crow[j] : 0.0;
        w[2 * uu + 1] = j + 4 < M ? Crow[j + 4] : 0.0;
      }
    };
    double wn[4] = {0.0, 0.0, 0.0, 0.0};
    if (DIAG) load_weights(0, 0, wn);
#pragma unroll 2
    for (int cbk = 0; cbk < ncb; ++cbk) {
      const int buf = cbk & 1;                // (unrolled by two: a compile-time constant in each copy)
      if (cbk + 1 < ncb) prepare_columns(cbk + 1, buf ^ 1);
      if (!DIAG && cbk > 0) fold_column_sums(cbk - 1);
      const double* cb = colB + buf * FB;
      double* cs = csum + buf * CF::NW * kGradCols + csoff;
#pragma unroll 1
      for (int cg = 0; cg < kGradCols / 8; cg += 2) {
        double w[4];
        if (DIAG) {
#pragma unroll
          for (int k = 0; k < 4; ++k) w[k] = wn[k];
          // next iteration's weights (the index past the last block is masked by j < M)
          const int cgn = cg + 2 < kGradCols / 8 ? cg + 2 : 0, cbn = cg + 2 < kGradCols / 8 ? cbk : cbk + 1;
          load_weights(cbn, cgn, wn);
        }
        double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          dmma_m8n8k4(t[0], t[1], af[ks], cb[(ks * kGradCols + cg * 8) * 4 + boff]);
          dmma_m8n8k4(t[2], t[3], af[ks], cb[(ks * kGradCols + cg * 8 + 8) * 4 + boff]);
        }
        exp_tab_contract<4, CF::REP>(t, etab_lane);
        double A[4];
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const int jl = (cg + uu) * 8 + c4;                               // this lane's columns: jl and jl + 4 of the block
          if (!DIAG) {
            w[2 * uu] = colW[buf * kGradCols + jl] * rs;
            w[2 * uu + 1] = colW[buf * kGradCols + jl + 4] * rs;
          }
          A[2 * uu] = t[2 * uu] * w[2 * uu];                               // A[row][j], A[row][j + 4]
          A[2 * uu + 1] = t[2 * uu + 1] * w[2 * uu + 1];
          if (!ROWSUM_IN_U) ai += A[2 * uu] + A[2 * uu + 1];
          const double* zb = cb + jl * 4 + zoff;                           // Z2'[column 8 (cg+uu) + k][dimension n], k = c4
          dmma_m8n8k4(u0, u1, A[2 * uu], zb[0]);
          dmma_m8n8k4(u0, u1, A[2 * uu + 1], zb[16]);
        }
        if (!DIAG) {
          // column sums over the strip's 8 rows (lane bits 2..4): halve the live values at each level
          const double k0 = b2 ? A[2] : A[0], k1 = b2 ? A[3] : A[1];
          const double s0 = b2 ? A[0] : A[2], s1 = b2 ? A[1] : A[3];
          const double h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);     // block uu = b2, columns c4 and c4 + 4
          const double h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
          double v = (b3 ? h1 : h0) + __shfl_xor_sync(0xffffffffu, b3 ? h0 : h1, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (lane < 16) cs[cg * 8] = v;                                   // column 8 (cg + b2) + 4 b3 + c4
        }
      }
      grad_sync<PERSIST>();
    }
  };
  if (diag) main_loop(std::true_type{});
  else main_loop(std::false_type{});
  // lane (r, c) holds U[r][2c], U[r][2c+1]: park each row's (a_i, u_i, z1'_i) in shared memory (the column buffers are free now),
  // then one thread per (statistic, 16-row chunk) forms the row products and sums them in a fixed order
  constexpr int RW = 2 * D + 2;              // per row: a, u[D], z1'[D], pad
  double* rows = colB;                       // 128 x RW doubles <= 2 KS 128 4
  static_assert(kGradRows * RW <= 2 * KS * kGradCols * 4, "row scratch fits in the column buffers");
  {
    const double re = diag ? rs : 1.0;       // off-diagonal rows are already weighted
    double* rw = rows + rloc * RW;
    if (2 * c4 < D) rw[1 + 2 * c4] = u0 * re;
    if (2 * c4 + 1 < D) rw[1 + 2 * c4 + 1] = u1 * re;
    if (ROWSUM_IN_U) {
      if (2 * c4 == D) rw[0] = u0 * re;
      if (2 * c4 + 1 == D) rw[0] = u1 * re;
    } else {
      ai += __shfl_xor_sync(0xffffffffu, ai, 1);
      ai += __shfl_xor_sync(0xffffffffu, ai, 2);
      if (c4 == 0) rw[0] = ai * re;
    }
    for (int d = c4; d < D; d += 4) rw[1 + D + d] = row < M ? Z[((size_t)a * M + row) * D + d] - pk[PP::MU + d] : 0.0;
  }
  if (!diag) fold_column_sums(ncb - 1);
  grad_sync<PERSIST>();
  double* out2 = stats + (((size_t)n * L * L + b * L + a) * nrb + rb) * GS::SIZE * (ll_tag ? 2 : 1);   // ordered pair (b, a), a < b
  if constexpr (D <= 7) {
    // Every statistic is an entry of the Gram product  G = sum_rows F^T H  with  F = [z1' (D), 1],  H = [a z1' (D), u (D), a]:
    //   G[m][n]     (m, n < D)   = R2[m][n]        G[m][D + e] = X[m][e]        G[m][2D] = r1[m]
    //   G[D][D + e]              = sum_i u_i[e] = r1_ba[e]                       G[D][2D] = S0
    // and, for a < b, R2_ba = sum_cols z2'^T (a' z2').  Each warp forms the product of its own 8 rows (2 k-steps x 2 column tiles)
    // and of 8 columns of every column block (2 k-steps) on the tensor path; the 16 partial 8 x 24 matrices are summed in warp order.
    constexpr int GW = 24;
    const int q = lane >> 2;                                  // A operand: feature q of row k; B operand: feature q (+8) of row k
    double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double* rw = rows + (warp * 8 + c4 + 4 * s) * RW;
      const double av = rw[0];
      const double fa = q < D ? rw[1 + D + q] : (q == D ? 1.0 : 0.0);
      auto hval = [&](int nh) -> double {                     // H[row][nh] = a z[nh] | u[nh - D] | a | 0
        if (nh < D) return av * rw[1 + D + nh];
        if (nh < 2 * D) return rw[1 + (nh - D)];
        return nh == 2 * D ? av : 0.0;
      };
      dmma_m8n8k4(g00, g01, fa, hval(q));
      dmma_m8n8k4(g10, g11, fa, hval(q + 8));
    }
    if (!diag) {
      for (int cbk = 0; cbk < ncb; ++cbk)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int j = cbk * kGradCols + warp * 8 + c4 + 4 * s;
          const double zq = (q < D && j < M) ? Z[((size_t)b * M + j) * D + q] - pk[PP::MU + q] : 0.0;
          dmma_m8n8k4(h0, h1, zq, colA[j] * zq);
        }
    }
    static_assert(CF::NW * 8 * GW <= 2 * CF::NW * kGradCols, "partial Gram matrices fit in the column-sum buffers");
    double* gw = csum + warp * (8 * GW) + q * GW + 2 * c4;      // the per-warp column sums are folded: their buffers are free
    gw[0] = g00; gw[1] = g01; gw[8] = g10; gw[9] = g11; gw[16] = h0; gw[17] = h1;
    grad_sync<PERSIST>();
    if (tid < 8 * GW) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < CF::NW; ++w) sum += csum[w * (8 * GW) + tid];
      const int m = tid / GW, nn2 = tid % GW;
      if (nn2 < 16) {
        if (m < D) {
          if (nn2 < D) { if (nn2 >= m) put(out, GS::R2 + m * D - m * (m - 1) / 2 + (nn2 - m), sum); }
          else if (nn2 < 2 * D) put(out, GS::X + m * D + (nn2 - D), sum);
          else if (nn2 == 2 * D) put(out, GS::R1 + m, sum);
        } else if (m == D) {
          if (nn2 == 2 * D) put(out, GS::S0, sum);
          else if (!diag && nn2 >= D && nn2 < 2 * D) put(out2, GS::R1 + (nn2 - D), sum);
        }
      } else if (!diag) {
        const int e = nn2 - 16;
        if (m < D && e < D && e >= m) put(out2, GS::R2 + m * D - m * (m - 1) / 2 + (e - m), sum);
      }
    }
  } else {
    // D = 8: [z, 1] does not fit one 8-row tile; one thread per (statistic, 16-row chunk) forms the products, fixed-order sums
    constexpr int CH = kGradThreads / 64;      // 8 chunks of 16 entries; statistics beyond 64 loop
    // statistic k of a scratch block: S0 = sum a;  R1[d] = sum a z[d];  R2[d,e] = sum a z[d] z[e];  X[d,e] = sum z[d] u[e]
    auto decode = [](int k, int& kind, int& d1, int& d2) {
      d1 = 0; d2 = 0;
      if (k == GS::S0) kind = 0;
      else if (k < GS::R2) { kind = 1; d1 = k - GS::R1; }
      else if (k < GS::X) {
        kind = 2;
        int t = k - GS::R2;
        while (t >= D - d1) { t -= D - d1; ++d1; }
        d2 = d1 + t;
      } else { kind = 3; d1 = (k - GS::X) / D; d2 = (k - GS::X) % D; }
    };
    auto chunk_sum = [&](const double* scratch, int q, int kind, int d1, int d2) {
      double acc = 0.0;
      for (int r = q * (kGradRows / CH); r < (q + 1) * (kGradRows / CH); ++r) {
        const double* rw = scratch + r * RW;
        const double av = rw[0];
        if (kind == 0) acc += av;
        else if (kind == 1) acc = fma(av, rw[1 + D + d1], acc);
        else if (kind == 2) acc = fma(av * rw[1 + D + d1], rw[1 + D + d2], acc);
        else if (kind == 3) acc = fma(rw[1 + D + d1], rw[1 + d2], acc);
        else acc += rw[1 + d1];                // kind 4: sum of u[d1]
      }
      return acc;
    };
    for (int k = tid & 63; k < GS::SIZE; k += 64) {
      int kind, d1, d2;
      decode(k, kind, d1, d2);
      red[(tid >> 6) * GS::SIZE + k] = chunk_sum(rows, tid >> 6, kind, d1, d2);
    }
    grad_sync<PERSIST>();
    if (tid < GS::SIZE) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < CH; ++q) s += red[q * GS::SIZE + tid];
      put(out, tid, s);
    }
    if (diag) return;
    // ordered pair (b, a): r1_ba = sum_i u_i (row scratch), R2_ba = sum_j a'_j z2'_j z2'_j^T (column scratch, 128 columns at a time
    // in the per-warp column-sum buffers, free after the last fold); S0_ba and X_ba are not stored (symmetry, see k_bwd_finalize)
    double* cscr = csum;
    static_assert(kGradCols * RW <= 2 * CF::NW * kGradCols, "column scratch fits in the column-sum buffers");
    double acc2[(GS::SIZE + 63) / 64];
#pragma unroll
    for (int s = 0; s < (GS::SIZE + 63) / 64; ++s) acc2[s] = 0.0;
    {
      int s = 0;
      for (int k = tid & 63; k < GS::SIZE; k += 64, ++s)
        if (k >= GS::R1 && k < GS::R2) acc2[s] = chunk_sum(rows, tid >> 6, 4, k - GS::R1, 0);
    }
    for (int cbk = 0; cbk < ncb; ++cbk) {
      grad_sync<PERSIST>();                         // previous block's readers are done
      if (tid < kGradCols) {
        const int j = cbk * kGradCols + tid;
        double* cw = cscr + tid * RW;
        cw[0] = colA[j];
#pragma unroll
        for (int d = 0; d < D; ++d) cw[1 + D + d] = j < M ? Z[((size_t)b * M + j) * D + d] - pk[PP::MU + d] : 0.0;
      }
      grad_sync<PERSIST>();
      int s = 0;
      for (int k = tid & 63; k < GS::SIZE; k += 64, ++s) {
        if (k < GS::R2 || k >= GS::X) continue;
        int kind, d1, d2;
        decode(k, kind, d1, d2);
        acc2[s] += chunk_sum(cscr, tid >> 6, 2, d1, d2);
      }
    }
    {
      int s = 0;
      for (int k = tid & 63; k < GS::SIZE; k += 64, ++s) red[(tid >> 6) * GS::SIZE + k] = acc2[s];
    }
    grad_sync<PERSIST>();
    if (tid >= GS::R1 && tid < GS::X) {        // the only statistics of (b, a) k_bwd_finalize reads
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < CH; ++q) s += red[q * GS::SIZE + tid];
      put(out2, tid, s);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kGradThreads, 2) k_contract_grad(const double* __restrict__ Z, const double* __restrict__ beta,
                                                                   const double* __restrict__ C, const double* __restrict__ packs,
                                                                   const double* __restrict__ omega, double* __restrict__ stats,
                                                                   int M, int L, int nrb) {
  extern __shared__ __align__(16) double smem[];
  const int npairs = L * (L + 1) / 2;
  const int rb = blockIdx.x % nrb;
  const int pr = (blockIdx.x / nrb) % npairs;
  const int n = blockIdx.x / (nrb * npairs);
  contract_grad_item<D, false>(smem, threadIdx.x, n, pr, rb, Z, beta, C, packs, omega, stats, M, L, nrb, true);
}

// un-mix the output adjoints to latent space and apply the chain rule of Sff = f2 - f1 f1^T (+ const)
struct BwdPrepareParams {
  const double *f1_bar, *Sff_bar, *cross_bar;   // [N,P], [N,P,P], [N,D,P]  (any may be null = zero)
  const double* f1lat;                           // [N,L]
  const double* W;                               // [P,L] or null
  double *f1lat_bar, *crosslat_bar, *omega;      // [N,L], [N,D,L], [N,L,L]
  int N, L, P, D, full_cov;
};

// all 128 threads of the group (one output entry each); ends with a group barrier
__device__ inline void bwd_prepare_input(const BwdPrepareParams& p, int n) {
  const int L = p.L, P = p.P, D = p.D, tid = threadIdx.x, nt = kGroupThreads;
  __shared__ double SL[GPP_MAX_L * GPP_MAX_L];
  for (int t = tid; t < L * L; t += nt) {
    const int l = t / L, k = t % L;
    double v = 0.0;
    if (p.Sff_bar) {
      const double* Sb = p.Sff_bar + (size_t)n * P * P;
      if (p.W) {
        for (int a = 0; a < P; ++a)
          for (int b = 0; b < P; ++b)
            if (p.full_cov || a == b) v = fma(p.W[a * L + l] * p.W[b * L + k], Sb[a * P + b], v);
      } else if (p.full_cov || l == k) {
        v = Sb[l * P + k];
      }
    }
    SL[t] = v;
    p.omega[((size_t)n * L + l) * L + k] = v;
  }
  group_sync();
  const double* f1l = p.f1lat + (size_t)n * L;
  for (int t = tid; t < L + D * L; t += nt) {
    if (t < L) {
      const int l = t;
      double v = 0.0;
      if (p.f1_bar) {
        if (p.W) {
          for (int o = 0; o < P; ++o) v = fma(p.W[o * L + l], p.f1_bar[(size_t)n * P + o], v);
        } else {
          v = p.f1_bar[(size_t)n * P + l];
        }
      }
      for (int k = 0; k < L; ++k) v -= (SL[l * L + k] + SL[k * L + l]) * f1l[k];
      p.f1lat_bar[(size_t)n * L + l] = v;
    } else {
      const int d = (t - L) / L, l = (t - L) % L;
      double c = 0.0;
      if (p.cross_bar) {
        const double* cb = p.cross_bar + ((size_t)n * D + d) * P;
        if (p.W) {
          for (int o = 0; o < P; ++o) c = fma(p.W[o * L + l], cb[o], c);
        } else {
          c = cb[l];
        }
      }
      p.crosslat_bar[((size_t)n * D + d) * L + l] = c;
    }
  }
  group_sync();
}

// G x for symmetric G = Li^T Li given the lower-triangular Li
template <int D>
__device__ __forceinline__ void gram_apply(const Mat<D>& Li, const double* x, double* out) {
  double y[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k <= i; ++k) t = fma(Li(i, k), x[k], t);
    y[i] = t;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double t = 0.0;
#pragma unroll
    for (int i = d; i < D; ++i) t = fma(Li(i, d), y[i], t);
    out[d] = t;
  }
}

template <int D>
__device__ __forceinline__ void psi1_bwd_body(int n, const double* __restrict__ m, const double* __restrict__ S, int L, int M,
                                              const double* __restrict__ Z, const double* __restrict__ ell,
                                              const double* __restrict__ var, const double* __restrict__ beta,
                                              const double* __restrict__ f1lat_bar, const double* __restrict__ crosslat_bar,
                                              double* __restrict__ gm /*[N,L,D]*/, double* __restrict__ gS /*[N,L,D,D]*/,
                                              const double* li_in /* shared [L][D*D + 1] from psi1_body */) {
  constexpr int TRI = D * (D + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kGroupThreads >> 5;
  double mu[D];
#pragma unroll
  for (int d = 0; d < D; ++d) mu[d] = m[(size_t)n * D + d];
  for (int l = warp; l < L; l += nwarps) {
    // the forward half of this block factorised S + Lambda_l a moment ago: reuse its inverse Cholesky factor and log normaliser
    Mat<D> Li, G;
#pragma unroll
    for (int t = 0; t < D * D; ++t) Li.a[t] = li_in[l * (D * D + 1) + t];
    const double c0 = li_in[l * (D * D + 1) + D * D];
    gram_inverse<D>(Li, G);
    const double fb = f1lat_bar[(size_t)n * L + l];
    double cb[D], y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) cb[d] = crosslat_bar[((size_t)n * D + d) * L + l];
    gram_apply<D>(Li, cb, y);
    double A0 = 0.0, B0 = 0.0, A1[D], B1[D], A2[TRI];
#pragma unroll
    for (int d = 0; d < D; ++d) { A1[d] = 0.0; B1[d] = 0.0; }
#pragma unroll
    for (int t = 0; t < TRI; ++t) A2[t] = 0.0;
    const double* Zl = Z + (size_t)l * M * D;
    for (int j = lane; j < M; j += 32) {
      double dz[D];
#pragma unroll
      for (int d = 0; d < D; ++d) dz[d] = Zl[(size_t)j * D + d] - mu[d];
      double maha = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) t = fma(Li(i, k), dz[k], t);
        maha = fma(t, t, maha);
      }
      const double w = beta[(size_t)l * M + j] * fast_exp(c0 - 0.5 * maha);
      double e = fb;
#pragma unroll
      for (int d = 0; d < D; ++d) e = fma(y[d], dz[d], e);
      const double we = w * e;
      A0 += we;
      B0 += w;
      int t = 0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        A1[d] = fma(we, dz[d], A1[d]);
        B1[d] = fma(w, dz[d], B1[d]);
#pragma unroll
        for (int e2 = d; e2 < D; ++e2, ++t) A2[t] = fma(we * dz[d], dz[e2], A2[t]);
      }
    }
    A0 = warp_sum(A0);
    B0 = warp_sum(B0);
#pragma unroll
    for (int d = 0; d < D; ++d) { A1[d] = warp_sum(A1[d]); B1[d] = warp_sum(B1[d]); }
#pragma unroll
    for (int t = 0; t < TRI; ++t) A2[t] = warp_sum(A2[t]);
    if (lane == 0) {
      double GA1[D], c[D];
      gram_apply<D>(Li, A1, GA1);
      gram_apply<D>(Li, B1, c);
      double* om = gm + ((size_t)n * L + l) * D;
      double* oS = gS + ((size_t)n * L + l) * D * D;
#pragma unroll
      for (int d = 0; d < D; ++d) om[d] = GA1[d] - B0 * y[d];
      // T = A2 (full symmetric) ; GTG = G T G
      double T[D * D], GT[D * D];
      {
        int t = 0;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int e2 = d; e2 < D; ++e2, ++t) { T[d * D + e2] = A2[t]; T[e2 * D + d] = A2[t]; }
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) t = fma(G(i, k), T[k * D + j], t);
          GT[i * D + j] = t;
        }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) t = fma(GT[i * D + k], G(k, j), t);
          oS[i * D + j] = 0.5 * (t - A0 * G(i, j)) - 0.5 * (y[i] * c[j] + c[i] * y[j]);
        }
    }
  }
}

// Runs inside the prologue launch, in the Psi1 block of input n, once the block has written f1lat[n, :]: the output adjoints are
// un-mixed to latent space (thread 0), then the Psi1 adjoints follow with a warp per latent.
template <int D>
struct BwdEpilogue {
  BwdPrepareParams bp;
  const double *m, *S, *Z, *ell, *var, *beta;
  double *gm, *gS;
  int M;
  static constexpr bool kNeedsFactors = true;
  __device__ __forceinline__ void operator()(int n, const double* li) const {
    group_sync();
    bwd_prepare_input(bp, n);
    psi1_bwd_body<D>(n, m, S, bp.L, M, Z, ell, var, beta, bp.f1lat_bar, bp.crosslat_bar, gm, gS, li);
  }
};

struct BwdFinalizeParams {
  const double *m, *S;          // [N,D], [N,D,D]
  const double* ell;            // [L,D]
  const double* stats;          // [N,L*L,nrb,GS::SIZE]  (ll_tag != 0: tagged words, two per statistic)
  unsigned ll_tag;
  const double* omega;          // [N,L,L]
  const double* Gs;             // [N,L*L,D,D]  (Sigma_n + V_ab)^-1 from the prologue
  const double *gm, *gS;        // psi1 contributions [N,L,D], [N,L,D,D]
  double *m_bar, *S_bar;        // [N,D], [N,D,D]
  int N, L, nrb;
};

template <int D>
struct FinalizeSmem {
  using GS = GradStats<D>;
  static constexpr int CS = D + D * D, DD = D * D;
  static constexpr int PER_PAIR = 2 * GS::SIZE + 2 * CS + 2 * DD + 4;   // statistics (ab, ba), E1|E2, contribution, G, G E2, (a, b, weight, pad)
};

// CTA per input.  d f2_ab / d mu = G E1,  d f2_ab / d Sigma = 1/2 (G E2 G - S0 G)  with E1 = A1 r1_ab + A2 r1_ba and
// E2 = A1 R2_ab A1 + A2 R2_ba A2 + A1 X A2 + (A1 X A2)^T per unordered pair; every phase is spread over the CTA's threads
// (one matrix entry each) with the operands in shared memory, pairs and Psi1 terms are summed in a fixed order at the end.
// (`fsm`: npairs * FinalizeSmem<D>::PER_PAIR doubles of shared memory; the statistics may come from other CTAs of the launch)
template <int D>
__device__ void bwd_finalize_body(const BwdFinalizeParams& p, const int n, double* __restrict__ fsm) {
  using GS = GradStats<D>;
  using FS = FinalizeSmem<D>;
  constexpr int CS = FS::CS, DD = FS::DD;
  const int tid = threadIdx.x, L = p.L, nt = kGroupThreads;
  const int npairs = L * (L + 1) / 2;
  double* sst = fsm;                                   // [npairs][2][SIZE]
  double* e12 = sst + (size_t)npairs * 2 * GS::SIZE;   // [npairs][CS]
  double* contrib = e12 + (size_t)npairs * CS;         // [npairs][CS]
  double* gmat = contrib + (size_t)npairs * CS;        // [npairs][DD]
  double* gt = gmat + (size_t)npairs * DD;             // [npairs][DD]
  double* meta = gt + (size_t)npairs * DD;             // [npairs][4]: a, b, weight
  for (int pr = tid; pr < npairs; pr += nt) {          // unordered pair index -> (a <= b)
    int a = 0, b = pr;
    while (b >= L - a) { b -= L - a; ++a; }
    b += a;
    meta[pr * 4] = a;
    meta[pr * 4 + 1] = b;
    meta[pr * 4 + 2] = (a == b) ? p.omega[((size_t)n * L + a) * L + a]
                                : p.omega[((size_t)n * L + a) * L + b] + p.omega[((size_t)n * L + b) * L + a];
  }
  group_sync();
  // slot (a, b), a <= b: all statistics; slot (b, a), a < b: only r1 and R2 (S0_ba = S0_ab, X_ba = X_ab^T are not stored)
  for (int idx = tid; idx < npairs * 2 * GS::SIZE; idx += nt) {
    const int k = idx % GS::SIZE, which = (idx / GS::SIZE) & 1, pr = idx / (2 * GS::SIZE);
    const int a = (int)meta[pr * 4], b = (int)meta[pr * 4 + 1];
    const bool mirrored = k >= GS::R1 && k < GS::X;
    const int slot = (which == 0 || !mirrored) ? a * L + b : b * L + a;
    double x = 0.0;
    if (meta[pr * 4 + 2] != 0.0)             // skipped pairs were not written by k_contract_grad
      for (int rb = 0; rb < p.nrb; ++rb) {
        const size_t at = (((size_t)n * L * L + slot) * p.nrb + rb) * GS::SIZE + k;
        x += p.ll_tag ? ll_load(reinterpret_cast<const unsigned long long*>(p.stats) + 2 * at, p.ll_tag) : p.stats[at];
      }
    sst[idx] = x;
  }
  for (int idx = tid; idx < npairs * DD; idx += nt) {
    const int pr = idx / DD, k = idx % DD;
    const int a = (int)meta[pr * 4], b = (int)meta[pr * 4 + 1];
    gmat[idx] = p.Gs[((size_t)n * L * L + a * L + b) * DD + k];
  }
  group_sync();
  for (int idx = tid; idx < npairs * CS; idx += nt) {  // E1 [D] | E2 [D][D]
    const int pr = idx / CS, k = idx % CS;
    const int a = (int)meta[pr * 4], b = (int)meta[pr * 4 + 1];
    const double* sab = sst + (size_t)pr * 2 * GS::SIZE;
    const double* sba = sab + GS::SIZE;
    auto mix = [&](int d, double& A1, double& A2) {
      const double v1 = p.ell[a * D + d] * p.ell[a * D + d], v2 = p.ell[b * D + d] * p.ell[b * D + d];
      A1 = v2 / (v1 + v2);
      A2 = v1 / (v1 + v2);
    };
    double v;
    if (k < D) {
      double A1, A2;
      mix(k, A1, A2);
      v = A1 * sab[GS::R1 + k] + A2 * sba[GS::R1 + k];
    } else {
      const int d = (k - D) / D, e = (k - D) % D;
      double A1d, A2d, A1e, A2e;
      mix(d, A1d, A2d);
      mix(e, A1e, A2e);
      const int lo = d < e ? d : e, hi = d < e ? e : d;
      const int t = lo * D - lo * (lo - 1) / 2 + (hi - lo);
      v = A1d * sab[GS::R2 + t] * A1e + A2d * sba[GS::R2 + t] * A2e + A1d * sab[GS::X + d * D + e] * A2e + A2d * sab[GS::X + e * D + d] * A1e;
    }
    e12[idx] = v;
  }
  group_sync();
  for (int idx = tid; idx < npairs * CS; idx += nt) {  // G E1 -> mean contribution; G E2 -> gt
    const int pr = idx / CS, k = idx % CS;
    const double* G = gmat + (size_t)pr * DD;
    const double* E = e12 + (size_t)pr * CS;
    if (k < D) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) t = fma(G[k * D + c], E[c], t);
      contrib[idx] = meta[pr * 4 + 2] != 0.0 ? meta[pr * 4 + 2] * t : 0.0;
    } else {
      const int i = (k - D) / D, j = (k - D) % D;
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) t = fma(G[i * D + c], E[D + c * D + j], t);
      gt[(size_t)pr * DD + (k - D)] = t;
    }
  }
  group_sync();
  for (int idx = tid; idx < npairs * DD; idx += nt) {  // (G E2) G - S0 G -> covariance contribution
    const int pr = idx / DD, k = idx % DD, i = k / D, j = k % D;
    const double* G = gmat + (size_t)pr * DD;
    const double* T = gt + (size_t)pr * DD;
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) t = fma(T[i * D + c], G[c * D + j], t);
    const double wgt = meta[pr * 4 + 2];
    contrib[(size_t)pr * CS + D + k] = wgt != 0.0 ? wgt * 0.5 * (t - sst[(size_t)pr * 2 * GS::SIZE + GS::S0] * G[k]) : 0.0;
  }
  group_sync();
  for (int k = tid; k < CS; k += nt) {
    double s = 0.0;
    for (int pr = 0; pr < npairs; ++pr) s += contrib[(size_t)pr * CS + k];
    for (int l = 0; l < L; ++l) s += (k < D) ? p.gm[((size_t)n * L + l) * D + k] : p.gS[((size_t)n * L + l) * D * D + (k - D)];
    if (k < D) p.m_bar[(size_t)n * D + k] = s;
    else p.S_bar[(size_t)n * D * D + (k - D)] = s;
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_bwd_finalize(BwdFinalizeParams p) {
  extern __shared__ __align__(16) double fsm[];
  bwd_finalize_body<D>(p, (int)blockIdx.x, fsm);
}

}  // namespace gpp
