// k_contract: the Psi2 contraction kernel of the fused moment-matched GP predict (included by mm_predict.cu).
//
// Work item = (kernel pair (a,b), T x T tile of the M x M index space, chunk of inputs).  Persistent CTAs pull
// items from an atomic counter.  A CTA keeps the tile of C_a (diagonal pairs) in shared memory for the whole
// chunk.  Warps are specialised and hand data over through double-buffered shared memory with named barriers
// (bar.arrive / bar.sync), so there is NO CTA-wide barrier per input and the warps drift apart: shared-memory
// bursts of one warp hide behind the FP64 work of the others.
//
//   producer warps (NP)  one input ahead: read the (n, pair) coefficient pack from L2, compute
//                        g_i = R^T z1'_i, r_i = c0 + z1'^T P1 z1'  (rows)  and  z2'_j, s_j = z2'^T P2 z2'  (columns)
//                        into buffer b = k & 1; they also sum the lane partials of finished inputs (fixed order).
//   consumer warps (NC)  lane l owns rows {l, l+32, ..} (RPT = T/32 register tile), warp w owns a slice of columns.
//                        Per entry: 1 DADD + D DFMA + 10 FP64 (table exp, gpp_math.h) + 1 DFMA.  The RPT row chains of a thread share
//                        their column operands, which makes ptxas interleave them (a dependent DFMA issues 8 cycles
//                        after its producer, the pipe accepts one warp instruction every 2 cycles).
// Shared-memory layouts make every warp access a broadcast or unit-stride (no bank conflicts):
//   Ct[j][i] (i fastest), rowbuf[b][field][i] (structure of arrays), colbuf[b][j][field] (broadcast reads).
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

template <int D, int T, int NP, int NC>
struct ContractCfg {
  static constexpr int NT = 32 * (NP + NC);
  static constexpr int PT = 32 * NP;                            // producer threads
  static constexpr int CT = 0;                                  // [T][T]
  static constexpr int COL = CT + T * T;                        // [2][T][STRIDE]
  static constexpr int ROW = COL + 2 * T * ColLayout<D>::STRIDE;   // [2][D+2][T]
  static constexpr int RED = ROW + 2 * (D + 2) * T;             // [2][NC][32]
  static constexpr int ETAB = RED + 2 * NC * 32;                // [64][GPP_EXP_TAB_REP] replicated 2^(j/64) table (fast_exp_tab_n)
  static constexpr int PKBUF = ETAB + 64 * GPP_EXP_TAB_REP;     // [2][PairPack<D>::SIZE] coefficient packs of the inputs in flight
  static constexpr int TOTAL = PKBUF + 2 * PairPack<D>::SIZE;   // doubles
};

// barrier ids are immediates (a register id would make ptxas reserve all 16 hardware barriers for the CTA)
template <int ID>
__device__ __forceinline__ void named_bar_sync_imm(int count) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_bar_arrive_imm(int count) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int BASE>
__device__ __forceinline__ void named_bar_sync(int b, int count) {
  if (b) named_bar_sync_imm<BASE + 1>(count); else named_bar_sync_imm<BASE>(count);
}
template <int BASE>
__device__ __forceinline__ void named_bar_arrive(int b, int count) {
  if (b) named_bar_arrive_imm<BASE + 1>(count); else named_bar_arrive_imm<BASE>(count);
}

// One work item = (slot, input chunk).  Both roles walk the same item sequence: thread 0 draws the next item from the global
// counter and every thread of the CTA reads it between two CTA-wide barriers; the C tile load is shared too.
struct ContractItem {
  gpp_slot sl;
  int slot_id, n0, K;
  bool diag;
};

template <int D, int T, int NP, int NC>
__device__ __forceinline__ bool contract_next_item(const ContractParams& p, double* Ct, int* s_item, ContractItem& it) {
  constexpr int NT = ContractCfg<D, T, NP, NC>::NT;
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) *s_item = (int)atomicAdd(p.counter, 1u);
  __syncthreads();
  const int item = *s_item;
  if (item >= p.nslots * p.nchunks) return false;
  it.slot_id = item / p.nchunks;
  const int chunk_id = item % p.nchunks;
  it.sl = p.slots[it.slot_id];
  it.n0 = chunk_id * p.chunk;
  it.K = min(p.N, it.n0 + p.chunk) - it.n0;
  it.diag = (it.sl.a == it.sl.b);
  if (it.diag) {   // C is symmetric: read C[j][i] so that global reads and the later lane-wise smem reads are unit-stride
    const double* Ca = p.C + (size_t)it.sl.a * p.M * p.M;
    for (int idx = tid; idx < T * T; idx += NT) {
      int jj = idx / T, ii = idx % T;
      int jg = it.sl.tj * T + jj, ig = it.sl.ti * T + ii;
      Ct[idx] = (jg < p.M && ig < p.M) ? Ca[(size_t)jg * p.M + ig] : 0.0;
    }
  }
  __syncthreads();
  return true;
}

template <int D, int T, int NP, int NC>
__global__ void __launch_bounds__(32 * (NP + NC)) k_contract(ContractParams p) {
  using PP = PairPack<D>;
  using CF = ContractCfg<D, T, NP, NC>;
  constexpr int NT = CF::NT, PT = CF::PT;
  constexpr int RPT = T / 32;                // rows per consumer thread
  constexpr int CS = ColLayout<D>::STRIDE;
  constexpr int RBUF = (D + 2) * T, CBUF = T * CS, DBUF = NC * 32;
  constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_PROD = 5;  // named barriers 1,2 (full), 3,4 (empty), 5 (producers); 0 is __syncthreads
  static_assert(T % 32 == 0, "rows of a tile are covered by 32 lanes x RPT");

  extern __shared__ __align__(16) double smem[];
  double* Ct = smem + CF::CT;
  double* colbuf = smem + CF::COL;
  double* rowbuf = smem + CF::ROW;
  double* red = smem + CF::RED;
  double* etab = smem + CF::ETAB;
  double* pkbuf = smem + CF::PKBUF;
  __shared__ int s_item;
  for (int i = threadIdx.x; i < 64 * GPP_EXP_TAB_REP; i += NT) etab[i] = kExp2Tab[i / GPP_EXP_TAB_REP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ContractItem it;

  if (warp < NP) {
    // ------------------------------------------------------------------ producers (own copy of the item loop, then exit:
    // the consumer loop below is then straight-line code for ptxas, which keeps its constants in uniform registers)
    static_assert(PT == T, "one producer thread per tile row / column");
    constexpr int NPV = (PP::SIZE + PT - 1) / PT;   // pack elements staged per producer thread
    while (contract_next_item<D, T, NP, NC>(p, Ct, &s_item, it)) {
      const gpp_slot sl = it.sl;
      const bool diag = it.diag;
      // the tile's centres and weights do not change over the chunk: keep this thread's row and column in registers
      const int ig = sl.ti * T + tid, jg = sl.tj * T + tid;
      double zrow[D], zcol[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        zrow[d] = ig < p.M ? p.Z[((size_t)sl.a * p.M + ig) * D + d] : 0.0;
        zcol[d] = jg < p.M ? p.Z[((size_t)sl.b * p.M + jg) * D + d] : 0.0;
      }
      const double brow = (!diag && ig < p.M) ? p.beta[(size_t)sl.a * p.M + ig] : 0.0;
      const double bcol = (!diag && jg < p.M) ? p.beta[(size_t)sl.b * p.M + jg] : 0.0;
      // coefficient pack of input k: one element per thread, loaded one input ahead, staged in shared memory (double buffered)
      const double* pk0 = p.packs + ((size_t)it.n0 * p.npairs + sl.pair) * PP::SIZE;
      const size_t pk_stride = (size_t)p.npairs * PP::SIZE;
      double pv[NPV];
#pragma unroll
      for (int q = 0; q < NPV; ++q) pv[q] = tid + q * PT < PP::SIZE ? pk0[tid + q * PT] : 0.0;
      for (int k = 0; k < it.K + 2; ++k) {
        const int b = k & 1;
        if (k >= 2) {                        // consumers are done with input k-2 (buffer b): reduce its lane partials
          named_bar_sync<BAR_EMPTY>(b, NT);
          if (warp == 0) {
            const double* rp = red + b * DBUF + lane;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NC; ++w) s += rp[w * 32];
            s = warp_sum(s);
            if (lane == 0) p.part[(size_t)(it.n0 + k - 2) * p.nslots + it.slot_id] = s;
          }
        }
        if (k >= it.K) continue;
        double* pk = pkbuf + b * PP::SIZE;
#pragma unroll
        for (int q = 0; q < NPV; ++q)
          if (tid + q * PT < PP::SIZE) pk[tid + q * PT] = pv[q];
        named_bar_sync_imm<BAR_PROD>(PT);   // producers only: pack k visible; pack k-2 (same buffer) no longer read
        if (k + 1 < it.K) {
#pragma unroll
          for (int q = 0; q < NPV; ++q)
            if (tid + q * PT < PP::SIZE) pv[q] = pk0[(size_t)(k + 1) * pk_stride + tid + q * PT];
        }
        double* rb = rowbuf + b * RBUF;
        double* cb = colbuf + b * CBUF;
        {
          // row `tid` of the tile
          double zc[D];
#pragma unroll
          for (int d = 0; d < D; ++d) zc[d] = zrow[d] - pk[PP::MU + d];
#pragma unroll
          for (int e = 0; e < D; ++e) {
            double t = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) t = fma(zc[d], pk[PP::R + d * D + e], t);
            rb[e * T + tid] = t;
          }
          rb[D * T + tid] = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zc);
          rb[(D + 1) * T + tid] = brow;
        }
        {
          // column `tid` of the tile
          double zc[D];
#pragma unroll
          for (int d = 0; d < D; ++d) zc[d] = zcol[d] - pk[PP::MU + d];
          double* dst = cb + tid * CS;
#pragma unroll
          for (int d = 0; d < D; ++d) dst[d] = zc[d];
          dst[D] = packed_quad<D>(pk + PP::P2, zc);
          dst[D + 1] = bcol;
        }
        __threadfence_block();
        named_bar_arrive<BAR_FULL>(b, NT);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int cwarp = warp - NP;               // consumer warp index
  // consumer column slice: T columns over NC warps, first (T % NC) warps take one more
  const int cw = T / NC + (cwarp < T % NC ? 1 : 0);
  const int c0 = cwarp * (T / NC) + min(cwarp, T % NC);
  const double* ct = Ct + c0 * T + lane;
  const double* etab_lane = etab + (lane & (GPP_EXP_TAB_REP - 1));
  while (contract_next_item<D, T, NP, NC>(p, Ct, &s_item, it)) {
    const bool diag = it.diag;
    for (int k = 0; k < it.K; ++k) {
      const int b = k & 1;
      named_bar_sync<BAR_FULL>(b, NT);
      const double* rb = rowbuf + b * RBUF + lane;
      double g[RPT][D], r[RPT], acc[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
#pragma unroll
        for (int d = 0; d < D; ++d) g[q][d] = rb[d * T + 32 * q];
        r[q] = rb[D * T + 32 * q];
        acc[q] = 0.0;
      }
      const double* cb = colbuf + b * CBUF + c0 * CS;
#pragma unroll 2
      for (int jj = 0; jj < cw; ++jj) {
        const double* c = cb + jj * CS;
        double zc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zc[d] = c[d];
        const double sj = c[D];
        double t[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) t[q] = r[q] + sj;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int q = 0; q < RPT; ++q) t[q] = fma(g[q][d], zc[d], t[q]);
        fast_exp_tab_n<RPT>(t, etab_lane);
        if (diag) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) acc[q] = fma(t[q], ct[jj * T + 32 * q], acc[q]);
        } else {
          const double wj = c[D + 1];
#pragma unroll
          for (int q = 0; q < RPT; ++q) acc[q] = fma(t[q], wj, acc[q]);
        }
      }
      double total = 0.0;
#pragma unroll
      for (int q = 0; q < RPT; ++q) total += diag ? acc[q] : acc[q] * rb[(D + 1) * T + 32 * q];
      red[b * DBUF + cwarp * 32 + lane] = total;
      __threadfence_block();
      named_bar_arrive<BAR_EMPTY>(b, NT);
    }
  }
}

}  // namespace gpp
