// k_contract: the Psi2 contraction kernel of the fused moment-matched GP predict (included by mm_predict.cu).
//
// Work item = (kernel pair (a,b), T x T tile of the M x M index space, chunk of inputs).  Persistent CTAs pull
// items from an atomic counter.  A CTA keeps the tile of C_a (diagonal pairs) in shared memory for the whole
// chunk.  Warps are specialised and hand data over through double-buffered shared memory with named barriers
// (bar.arrive / bar.sync), so there is NO CTA-wide barrier per input and the warps drift apart.
//
//   producer warps (NP)  one input ahead: stage the (n, pair) coefficient pack in shared memory and write, for their tile row
//                        and column, the EXTENDED vectors  A_i = [g_i (D), r_i, 1, 0..],  B_j = [z2'_j (D), 1, s_j, 0..]
//                        (g_i = R^T z1'_i, r_i = c0 + z1'^T P1 z1', s_j = z2'^T P2 z2'), so that the whole exponent is one
//                        inner product  log Q_ij = A_i . B_j  of length K = 4 KS.  They also sum the lane partials of
//                        finished inputs (fixed order).
//   consumer warps (NC)  warp w owns the 8-row strip w of the tile.  The exponent of an 8 x 8 block of entries is KS
//                        `mma.sync.m8n8k4.f64` instructions (SASS DMMA): on B200 DMMA runs on the FP64 pipe at the DFMA rate
//                        (37.1 vs 36.1 TFLOP/s, scripts/microbench/dmma.cu) but takes its operands from 4 registers per 512
//                        flop, whereas the scalar form needs D three-register DFMAs per entry, each of which costs 3 issue
//                        cycles instead of 2 (scripts/microbench/fp64_pipe.cu).  Each lane then holds 2 entries of one row
//                        (adjacent columns): table exp (10 FP64 ops, gpp_math.h) and one DFMA each into the contraction with
//                        C_a[i][j..j+1] (one 128-bit shared load) or beta_b[j..j+1].
// Shared-memory layouts make every warp access unit-stride or bank-conflict free:
//   rowA[b][ks][row][4], colB[b][ks][col][4]   a warp's 32 fragment elements of one k-step are 256 contiguous bytes
//   Ct[i][j] with row stride T + 8 doubles      8 rows x 4 column pairs per 128-bit load: quarter-warps hit disjoint banks
#pragma once
#include "mma_exp.cuh"
#include "model.cuh"

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

template <int D, int T, int NP, int NC>
struct ContractCfg {
  static_assert(NC * 8 == T, "one consumer warp per 8-row strip of the tile");
  static constexpr int KS = ExtLayout<D>::KS;
  static constexpr int LDC = T + 8;                             // row stride of the C tile (bank spreading for 128-bit loads)
  static constexpr int REP = KS <= 2 ? 16 : 8;                  // replication of the exp table
  static constexpr int NT = 32 * (NP + NC);
  static constexpr int PT = 32 * NP;                            // producer threads
  static constexpr int CT = 0;                                  // [T][LDC]
  static constexpr int COL = CT + T * LDC;                      // [2][KS][T][4]
  static constexpr int ROW = COL + 2 * KS * T * 4;              // [2][KS][T][4]
  static constexpr int WGT = ROW + 2 * KS * T * 4;              // [2][2][T]  beta of the rows / columns (off-diagonal pairs)
  static constexpr int RED = WGT + 2 * 2 * T;                   // [2][NC][32]
  static constexpr int ETAB = RED + 2 * NC * 32;                // [256][REP] replicated 2^(j/256) table (exp_tab_contract)
  static constexpr int PKBUF = ETAB + 256 * REP;                // [2][PairPack<D>::SIZE] coefficient packs of the inputs in flight
  static constexpr int TOTAL = PKBUF + 2 * PairPack<D>::SIZE;   // doubles
};

// Hand-over fence before bar.arrive: the producer/consumer pattern of the PTX ISA (st.shared; bar.arrive / bar.sync; ld.shared) relies on
// the barrier's own ordering at CTA scope, so an acquire-release CTA fence is already more than required; __threadfence_block() is a
// sequentially consistent fence (SASS MEMBAR.SC.CTA) and cost every consumer warp a round trip per input.
__device__ __forceinline__ void handover_fence() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// barrier ids are immediates (a register id would make ptxas reserve all 16 hardware barriers for the CTA)
template <int ID>
__device__ __forceinline__ void named_bar_sync_imm(int count) {
  __syncwarp();                                // inline-asm barriers do not make the compiler reconverge the warp (common.cuh)
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_bar_arrive_imm(int count) {
  __syncwarp();
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
  if (it.diag) {
    constexpr int LDC = ContractCfg<D, T, NP, NC>::LDC;
    const double* Ca = p.C + (size_t)it.sl.a * p.M * p.M;
    for (int idx = tid; idx < T * T; idx += NT) {
      int ii = idx / T, jj = idx % T;
      int ig = it.sl.ti * T + ii, jg = it.sl.tj * T + jj;
      Ct[ii * LDC + jj] = (jg < p.M && ig < p.M) ? Ca[(size_t)ig * p.M + jg] : 0.0;
    }
  }
  __syncthreads();
  return true;
}

template <int D, int T, int NP, int NC>
__global__ void __launch_bounds__(32 * (NP + NC)) k_contract(ContractParams p) {
  using PP = PairPack<D>;
  using CF = ContractCfg<D, T, NP, NC>;
  constexpr int NT = CF::NT, PT = CF::PT, KS = CF::KS, LDC = CF::LDC;
  constexpr int FBUF = KS * T * 4, WBUF = 2 * T, DBUF = NC * 32;
  constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_PROD = 5;  // named barriers 1,2 (full), 3,4 (empty), 5 (producers); 0 is __syncthreads
  static_assert(PT == T, "one producer thread per tile row / column");
  static_assert((T / 8) % 2 == 0, "column groups are processed in pairs");

  extern __shared__ __align__(16) double smem[];
  double* Ct = smem + CF::CT;
  double* colB = smem + CF::COL;
  double* rowA = smem + CF::ROW;
  double* wgt = smem + CF::WGT;
  double* red = smem + CF::RED;
  double* etab = smem + CF::ETAB;
  double* pkbuf = smem + CF::PKBUF;
  __shared__ int s_item;
  for (int i = threadIdx.x; i < kContractTab * CF::REP; i += NT) etab[i] = kExp2Tab256[i / CF::REP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ContractItem it;

  if (warp < NP) {
    // ------------------------------------------------------------------ producers (own copy of the item loop, then exit)
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
      // coefficient pack of input k: loaded one input ahead, staged in shared memory (double buffered)
      const double* pk0 = p.packs + ((size_t)it.n0 * p.npairs + sl.pair) * PP::SIZE;
      const size_t pk_stride = (size_t)p.npairs * PP::SIZE;
      double pv[NPV];
#pragma unroll
      for (int q = 0; q < NPV; ++q) pv[q] = tid + q * PT < PP::SIZE ? pk0[tid + q * PT] : 0.0;
#pragma unroll 2
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
        double ext[4 * KS];
        {
          // row `tid` of the tile: A = [g (D), r, 1, 0..]
          double zc[D];
#pragma unroll
          for (int d = 0; d < D; ++d) zc[d] = zrow[d] - pk[PP::MU + d];
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
          double* ra = rowA + b * FBUF + tid * 4;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int q = 0; q < 4; q += 2)
              *reinterpret_cast<double2*>(ra + ks * T * 4 + q) = make_double2(ext[ks * 4 + q], ext[ks * 4 + q + 1]);
          wgt[b * WBUF + tid] = brow;
        }
        {
          // column `tid` of the tile: B = [z2' (D), 1, s, 0..]
#pragma unroll
          for (int d = 0; d < D; ++d) ext[d] = zcol[d] - pk[PP::MU + d];
          ext[D + 1] = packed_quad<D>(pk + PP::P2, ext);
          ext[D] = 1.0;
          double* cb = colB + b * FBUF + tid * 4;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int q = 0; q < 4; q += 2)
              *reinterpret_cast<double2*>(cb + ks * T * 4 + q) = make_double2(ext[ks * 4 + q], ext[ks * 4 + q + 1]);
          wgt[b * WBUF + T + tid] = bcol;
        }
        handover_fence();
        named_bar_arrive<BAR_FULL>(b, NT);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: warp = 8-row strip of the tile
  const int strip = warp - NP;
  const int row = strip * 8 + (lane >> 2);           // this lane's row of the tile
  const int cpair = 2 * (lane & 3);                  // its column pair inside an 8-column group
  const double* ct = Ct + row * LDC + cpair;
  const unsigned etab_lane = (unsigned)__cvta_generic_to_shared(etab + (lane & (CF::REP - 1)));
  while (contract_next_item<D, T, NP, NC>(p, Ct, &s_item, it)) {
    const bool diag = it.diag;
#pragma unroll 2
    for (int k = 0; k < it.K; ++k) {
      const int b = k & 1;                    // (unrolled by two: a compile-time constant in each copy)
      named_bar_sync<BAR_FULL>(b, NT);
      const double* ra = rowA + b * FBUF + strip * 32 + lane;
      const double* cb = colB + b * FBUF + lane;
      // contraction weights of this lane's row: C_a[i][j..] (diagonal pairs) or beta_b[j..]; both advance 8 doubles per column group
      const double* wsrc = diag ? ct : wgt + b * WBUF + T + cpair;
      double a[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) a[ks] = ra[ks * T * 4];
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
      for (int cg = 0; cg < T / 8; cg += 2) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          dmma_m8n8k4(t[0], t[1], a[ks], cb[ks * T * 4 + cg * 32]);
          dmma_m8n8k4(t[2], t[3], a[ks], cb[ks * T * 4 + cg * 32 + 32]);
        }
        exp_tab_contract<4, CF::REP>(t, etab_lane);
        const double2 w0 = *reinterpret_cast<const double2*>(wsrc + cg * 8);
        const double2 w1 = *reinterpret_cast<const double2*>(wsrc + cg * 8 + 8);
        acc0 = fma(t[0], w0.x, acc0);
        acc1 = fma(t[2], w1.x, acc1);
        acc0 = fma(t[1], w0.y, acc0);
        acc1 = fma(t[3], w1.y, acc1);
      }
      double total = acc0 + acc1;
      if (!diag) total *= wgt[b * WBUF + row];
      red[b * DBUF + strip * 32 + lane] = total;
      handover_fence();
      named_bar_arrive<BAR_EMPTY>(b, NT);
    }
  }
}

}  // namespace gpp
