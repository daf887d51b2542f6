// k_contract: the Psi2 contraction kernel of the fused moment-matched GP predict (included by mm_predict.cu).
//
// Work item = (kernel pair (a,b), T x T tile of the M x M index space, chunk of inputs).  Persistent CTAs pull
// items from an atomic counter.  A CTA keeps the tile of C_a (diagonal pairs) in shared memory for the whole
// chunk and, per input n:
//   stage   (3 thread groups in parallel, reading the (n,pair) coefficient pack that cp.async prefetched)
//           group 0: g_i = R^T z1'_i           group 2: r_i = c0 + z1'^T P1 z1'
//           group 1: z2'_j, s_j = z2'^T P2 z2'  group 3 / warp 0: reduces the lane partials of input n-1
//   main    lane l of every warp owns rows {l, l+32, ..} (RPT = T/32 register tile), warp w owns CW = T/NW
//           columns.  Per entry: 1 DADD + D DFMA + 16 FP64 (exp) + 1 DFMA.  The RPT row chains of a thread share
//           their column operands, which makes ptxas interleave them (a dependent DFMA issues 8 cycles after its
//           producer, the pipe accepts one warp instruction every 2 cycles: >= 4 independent chains per scheduler).
//   reduce  each lane stores one partial; no shuffles on the compute warps.
// Shared-memory layouts are chosen so that every warp access is either a broadcast or unit-stride (no bank conflicts):
//   Ct[j][i] (i fastest), rowbuf[field][i] (structure of arrays), colbuf[j][field] (broadcast reads).
#pragma once

namespace gpp {

struct ContractParams {
  const double* Z;
  const double* beta;
  const double* C;
  const double* packs;
  double* part;            // [N, nslots]
  const gpp_slot* slots;
  unsigned* counter;
  int N, M, L, npairs, nslots, nchunks, chunk;
};

template <int D>
struct ColLayout {
  static constexpr int STRIDE = (D + 2 + 1) & ~1;   // z2'[D], s_j, w_j  (even => 16-byte records)
};

template <int D, int T>
struct ContractSmem {
  using PP = PairPack<D>;
  static constexpr int NT = 4 * T;
  static constexpr int NW = NT / 32;
  static constexpr int CT = 0;                                  // [T][T]
  static constexpr int COL = CT + T * T;                        // [T][STRIDE]
  static constexpr int ROW = COL + T * ColLayout<D>::STRIDE;    // [D+2][T]
  static constexpr int PACK = ROW + (D + 2) * T;                // [2][PP::SIZE]
  static constexpr int RED = PACK + 2 * PP::SIZE;               // [2][NW][32]
  static constexpr int TOTAL = RED + 2 * NW * 32;               // doubles
};

template <int D, int T>
__global__ void __launch_bounds__(4 * T) k_contract(ContractParams p) {
  using PP = PairPack<D>;
  using SM = ContractSmem<D, T>;
  constexpr int NT = SM::NT, NW = SM::NW;
  constexpr int RPT = T / 32;                // rows per thread
  constexpr int CW = T / NW;                 // columns per warp
  constexpr int CS = ColLayout<D>::STRIDE;
  static_assert(T % 32 == 0 && T % NW == 0, "tile must split into whole warps of rows and whole columns per warp");

  extern __shared__ __align__(16) double smem[];
  double* Ct = smem + SM::CT;
  double* colbuf = smem + SM::COL;
  double* rowbuf = smem + SM::ROW;
  double* packbuf = smem + SM::PACK;
  double* red = smem + SM::RED;
  __shared__ int s_item;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = tid / T, st = tid % T;   // staging role and index within the tile
  const int nitems = p.nslots * p.nchunks;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = (int)atomicAdd(p.counter, 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= nitems) break;
    const int slot_id = item / p.nchunks, chunk_id = item % p.nchunks;
    const gpp_slot sl = p.slots[slot_id];
    const int n0 = chunk_id * p.chunk, n1 = min(p.N, n0 + p.chunk);
    const bool diag = (sl.a == sl.b);

    // static operand of this thread's staging role: one inducing point (row of Z_a for groups 0/2, column of Z_b for 1)
    const int latent = (group == 1) ? sl.b : sl.a;
    const int glob = ((group == 1) ? sl.tj : sl.ti) * T + st;
    double zs[D], wgt = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) zs[d] = (group < 3 && glob < p.M) ? p.Z[((size_t)latent * p.M + glob) * D + d] : 0.0;
    if (!diag && group < 3 && glob < p.M) wgt = p.beta[(size_t)latent * p.M + glob];
    if (diag) {   // C is symmetric: read C[j][i] so that global reads and the later lane-wise smem reads are unit-stride
      const double* Ca = p.C + (size_t)sl.a * p.M * p.M;
      for (int idx = tid; idx < T * T; idx += NT) {
        int jj = idx / T, ii = idx % T;
        int jg = sl.tj * T + jj, ig = sl.ti * T + ii;
        Ct[idx] = (jg < p.M && ig < p.M) ? Ca[(size_t)jg * p.M + ig] : 0.0;
      }
    }
    {
      const double* src = p.packs + ((size_t)n0 * p.npairs + sl.pair) * PP::SIZE;
      for (int t = tid; t < PP::SIZE / 2; t += NT) cp_async16(packbuf + 2 * t, src + 2 * t);
      cp_async_commit();
    }

    for (int n = n0; n <= n1; ++n) {         // one extra trip drains the reduction of the last input
      const int buf = (n - n0) & 1;
      cp_async_wait<0>();
      __syncthreads();                       // pack(n) landed; colbuf/rowbuf free; red[buf^1] complete
      if (group == 3) {
        if (warp == 3 * T / 32 && n > n0) {  // reducer warp: fixed-order sum of the lane partials of input n-1
          const double* rp = red + (buf ^ 1) * NW * 32 + lane;
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) s += rp[w * 32];
          s = warp_sum(s);
          if (lane == 0) p.part[(size_t)(n - 1) * p.nslots + slot_id] = s;
        }
        if (n + 1 < n1) {                    // prefetch the next coefficient pack
          const double* src = p.packs + ((size_t)(n + 1) * p.npairs + sl.pair) * PP::SIZE;
          for (int t = tid - 3 * T; t < PP::SIZE / 2; t += T) cp_async16(packbuf + (buf ^ 1) * PP::SIZE + 2 * t, src + 2 * t);
        }
      }
      cp_async_commit();
      if (n == n1) break;
      const double* pk = packbuf + buf * PP::SIZE;
      if (group < 3) {
        double zc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zc[d] = zs[d] - pk[PP::MU + d];
        if (group == 0) {                    // g_i = R^T z1'_i
#pragma unroll
          for (int e = 0; e < D; ++e) {
            double t = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) t = fma(zc[d], pk[PP::R + d * D + e], t);
            rowbuf[e * T + st] = t;
          }
        } else if (group == 2) {             // r_i = c0 + z1'^T P1 z1', beta_i
          rowbuf[D * T + st] = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zc);
          rowbuf[(D + 1) * T + st] = wgt;
        } else {                             // z2'_j, s_j = z2'^T P2 z2', beta_j
          double* dst = colbuf + st * CS;
#pragma unroll
          for (int d = 0; d < D; ++d) dst[d] = zc[d];
          dst[D] = packed_quad<D>(pk + PP::P2, zc);
          dst[D + 1] = wgt;
        }
      }
      __syncthreads();

      double g[RPT][D], r[RPT], acc[RPT];
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
#pragma unroll
        for (int d = 0; d < D; ++d) g[k][d] = rowbuf[d * T + lane + 32 * k];
        r[k] = rowbuf[D * T + lane + 32 * k];
        acc[k] = 0.0;
      }
      const double* cb = colbuf + warp * CW * CS;
      const double* ct = Ct + (warp * CW) * T + lane;
#pragma unroll 2
      for (int jj = 0; jj < CW; ++jj) {
        const double* c = cb + jj * CS;
        double zc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zc[d] = c[d];
        const double sj = c[D];
        double t[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) t[k] = r[k] + sj;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int k = 0; k < RPT; ++k) t[k] = fma(g[k][d], zc[d], t[k]);
        fast_exp_n<RPT>(t);
        if (diag) {
#pragma unroll
          for (int k = 0; k < RPT; ++k) acc[k] = fma(t[k], ct[jj * T + 32 * k], acc[k]);
        } else {
          const double wj = c[D + 1];
#pragma unroll
          for (int k = 0; k < RPT; ++k) acc[k] = fma(t[k], wj, acc[k]);
        }
      }
      double total = 0.0;
#pragma unroll
      for (int k = 0; k < RPT; ++k) total += diag ? acc[k] : acc[k] * rowbuf[(D + 1) * T + lane + 32 * k];
      red[buf * NW * 32 + warp * 32 + lane] = total;
    }
  }
}

}  // namespace gpp
