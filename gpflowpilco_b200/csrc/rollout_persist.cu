// Persistent moment-matched rollout: the whole H-step sweep of gpp_rollout_mm_fwd / gpp_rollout_mm_bwd in ONE cooperative launch
// per direction — the device-side counterpart of upstream's compiled scan (gpflow_pilco/dynamics/solvers.py:84-105 driven by
// gpflow_pilco/loops/pilco.py:207-217; reverse sweep = tape.gradient through it, utils/optimizers.py:52-56).
//
// Every CTA (one per SM) is warp specialised:
//   warps 0..3   "scalar group": the per-rollout stages of a step (rollout_mm_common.cuh, predict_kernels.cuh, finalize.cuh,
//                rollout_mm_bwd_common.cuh, predict_bwd_kernels.cuh — the very code the one-launch-per-stage path runs)
//   other warps  the Psi2 contraction: forward = the DMMA producer/consumer pipeline of contract_kernel.cuh on a 64 x 64 tile of C
//                that stays in shared memory for the whole sweep when the model has no more tiles than the GPU has SMs (M = 256,
//                L = 4: 136 tiles); backward = contract_grad_item work items pulled from a ticket counter.
// Rollout n's step t is a chain  scalar(n,t) -> contraction(n,t; all tiles) -> scalar(n,t+1) ...; the two roles of all CTAs meet
// only through per-rollout release/acquire counters (persist_common.cuh), never through a grid-wide barrier, so with many rollouts
// in flight the serial stages of one rollout hide behind the contractions of the others: N independent rollouts advance as a
// skewed pipeline and the FP64 pipe stays busy.
#include <algorithm>
#include <cstdlib>

#include "contract_kernel.cuh"
#include "finalize.cuh"
#include "persist_common.cuh"
#include "predict_bwd_kernels.cuh"
#include "predict_kernels.cuh"
#include "rollout_mm_bwd_common.cuh"
#include "rollout_persist.h"

namespace gpp {

constexpr int kPersistTile = 64;        // tile edge of the forward contraction
// forward CTA = 4 scalar + 4 producer + NC consumer warps; NC = 8 (512 threads, scalar warps at 224 registers) is what runs.

struct PersistSaved {                   // the per-step block of gpp_rollout_mm_fwd_save (RolloutSaved offsets), or base == nullptr
  double* base;
  size_t per_step, md, Sd, Sxd, cross, pre;
};

struct PersistFwdParams {
  RolloutMMParams r;                    // r.md / r.Sd / r.Sxd / r.cross: workspace buffers used when nothing is saved
  PersistSaved sv;
  int H;
  const double *m0, *S0;
  double *m_final, *S_final;
  // dynamics model
  const double *Z, *ell, *var, *beta, *C, *mean, *W;
  int M, Lm, Pm, model_uncertainty;
  const gpp_slot* slots;
  const int* pair_start;
  const int* pair_ab;
  int npairs, nslots;
  // workspace
  double *f1lat, *crosslat;
  unsigned long long* packs_ll;         // [N, npairs, PairPack::SIZE, 2] coefficient packs as tagged words (tag = step + 1); zero at launch
  unsigned long long* part_ll;          // [N, nslots, 2] tile partials, same protocol
  unsigned* part_hint;                  // [N] number of tile partials published so far (hint counter, persist_common.cuh)
  int debug;                            // developer switch (env GPP_PERSIST_DEBUG): 1 = drain the pipeline every step
};

__device__ __forceinline__ void persist_step_pointers(RolloutMMParams& p, const PersistSaved& sv, const RolloutMMParams& base, int t) {
  if (sv.base) {
    double* b = sv.base + (size_t)t * sv.per_step;
    p.md = b + sv.md; p.Sd = b + sv.Sd; p.Sxd = b + sv.Sxd; p.cross = b + sv.cross; p.pre = b + sv.pre;
  } else {
    p.md = base.md; p.Sd = base.Sd; p.Sxd = base.Sxd; p.cross = base.cross; p.pre = nullptr;
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward: scalar group
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ void persist_fwd_scalar(const PersistFwdParams& P) {
  constexpr int DP = D - 1;
  __shared__ PreShared<DP> sh;
  const int tid = threadIdx.x, G = gridDim.x;
  RolloutMMParams p = P.r;
  const int N = p.N, Dx = p.Dx;
  const int n0 = G - 1 - (int)blockIdx.x;       // rollouts are dealt from the last CTA down: CTAs without a tile take them first
  for (int n = n0; n < N; n += G) {
    for (int i = tid; i < Dx + Dx * Dx; i += kGroupThreads) {
      if (i < Dx) {
        const double v = P.m0[(size_t)n * Dx + i];
        p.m[(size_t)n * Dx + i] = v;
        if (p.traj_m) p.traj_m[(size_t)n * Dx + i] = v;
      } else {
        const double v = P.S0[(size_t)n * Dx * Dx + (i - Dx)];
        p.S[(size_t)n * Dx * Dx + (i - Dx)] = v;
        if (p.traj_S) p.traj_S[(size_t)n * Dx * Dx + (i - Dx)] = v;
      }
    }
    if (tid == 0) p.loss[n] = 0.0;
  }
  group_sync();
  FinalizeParams fp;
  fp.part = nullptr; fp.part_ll = P.part_ll; fp.ll_tag = 0; fp.slots = P.slots; fp.pair_start = P.pair_start; fp.pair_ab = P.pair_ab;
  fp.f1lat = P.f1lat; fp.crosslat = P.crosslat; fp.var = P.var; fp.mean = P.mean; fp.W = P.W;
  fp.f1 = p.f1; fp.Sff = p.Sff; fp.cross = nullptr;
  fp.N = N; fp.L = P.Lm; fp.P = P.Pm; fp.D = D; fp.npairs = P.npairs; fp.nslots = P.nslots; fp.full_cov = 1;
  fp.model_uncertainty = P.model_uncertainty; fp.jitter = 0.0;
  fp.post.m = p.m; fp.post.S = p.S; fp.post.Dx = Dx; fp.post.ring_m = nullptr; fp.post.ring_S = nullptr;
  for (int t = 0; t <= P.H; ++t) {
    for (int n = n0; n < N; n += G) {
      if (t > 0) {
        // step t-1 of rollout n: wait for its tile partials (tag t) -> outputs of the GP predict, Euler update, trajectory slice t
        if (tid == 0 && !(P.debug & 2)) hint_wait_ge(P.part_hint + n, (unsigned)t * (unsigned)P.nslots);
        group_sync();
        persist_step_pointers(p, P.sv, P.r, t - 1);
        fp.ll_tag = (unsigned)t;
        fp.cross = p.cross;
        fp.post.Sxd = p.Sxd;
        fp.post.traj_m = p.traj_m ? p.traj_m + (size_t)t * N * Dx : nullptr;
        fp.post.traj_S = p.traj_S ? p.traj_S + (size_t)t * N * Dx * Dx : nullptr;
        finalize_body<true>(fp, n);
        group_sync();
      }
      if (t == P.H) {
        if (P.H > 0) {                         // cost of the final state (the loss callback runs after every step)
          step_pre_encode<DP>(p, n, sh);
          if (tid == 64) p.loss[n] += expected_cost<double>(DP, sh.me, sh.See, p.target, p.W);
        }
        for (int i = tid; i < Dx + Dx * Dx; i += kGroupThreads) {
          if (i < Dx) { if (P.m_final) P.m_final[(size_t)n * Dx + i] = p.m[(size_t)n * Dx + i]; }
          else if (P.S_final) P.S_final[(size_t)n * Dx * Dx + (i - Dx)] = p.S[(size_t)n * Dx * Dx + (i - Dx)];
        }
        group_sync();
        continue;
      }
      persist_step_pointers(p, P.sv, P.r, t);
      double cost = 0.0;
      step_pre_forward<DP>(p, n, sh, t > 0 ? &cost : nullptr);
      if (t > 0 && tid == 64) p.loss[n] += cost;            // thread 64 is the only writer of loss[n]: step order, fixed
      step_pre_write<DP>(p, n, sh);
      if (p.pre) {
        constexpr int PS = PreSharedSize<DP>::value;
        const double* src = reinterpret_cast<const double*>(&sh);
        double* dst = p.pre + (size_t)n * PS;
        for (int i = tid; i < PS; i += kGroupThreads) dst[i] = src[i];
      }
      group_sync();                                          // md / Sd of rollout n are visible to the group
      // Psi2 coefficient packs of the kernel pairs, as tagged words: the contraction CTAs pick them up as they arrive
      publish_packs_tagged<D>(p.md, p.Sd, n, P.npairs,
                              [&](int q, int& a, int& b, int& dst) { a = P.pair_ab[2 * q]; b = P.pair_ab[2 * q + 1]; dst = n * P.npairs + q; },
                              P.ell, P.var, P.packs_ll, nullptr, (unsigned)(t + 1), p.info);
      psi1_body<D>(n, p.md, p.Sd, N, P.Lm, P.M, P.Z, P.ell, P.var, P.beta, P.f1lat, P.crosslat, p.info, nullptr);
      group_sync();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward: contraction warps (the k_contract pipeline with the item loop replaced by a static (step, tile, input) order)
// ---------------------------------------------------------------------------------------------------------
constexpr int kPersistFwdBuffers = 2;    // vector buffers between the producer and consumer warps (PersistFwdCfg::NBUF)

template <int D, int NC_>
struct PersistFwdCfg {
  static constexpr int T = kPersistTile, NP = 4, NC = NC_;              // producer warps 0,1: rows; 2,3: columns
  static_assert(NC == 8 || NC == 16, "one or two consumer warps per 8-row strip");
  static constexpr int HALVES = NC / 8, COLS = T / HALVES;             // columns of the tile a consumer warp covers
  static constexpr int THREADS = kGroupThreads + 32 * (NP + NC);
  // 192 / 104 (NC = 8): 74.4 us per step at 64 rollouts against 76.3 with 224 / 96; 160 / 112 gains another 0.7 us there but costs
  // 1.3 us per step at 1..16 rollouts, where the scalar stages are the critical path
  static constexpr int REGS_SCALAR = NC == 8 ? 192 : 112, REGS_CONTRACT = NC == 8 ? 104 : 72;
  static constexpr int PT = 32 * NP, CT = 32 * NC, NTC = PT + CT;      // producer / consumer / contraction threads
  static constexpr int KS = ExtLayout<D>::KS, LDC = T + 8;
  static constexpr int REP = KS <= 2 ? 16 : 8;                          // replication of the exp table
  static constexpr int FBUF = KS * T * 4, WBUF = 2 * T, DBUF = NC * 32, NRED = 4;
  static constexpr int NBUF = kPersistFwdBuffers;                       // vector buffers in flight (3 measured slower than 2: 80 vs 78 us per step at 64 rollouts)
  static_assert(NRED >= NBUF + 1, "lane partials: item j - NBUF is reduced while the consumers may already write item j");
  static constexpr int BAR_PROD = 5, BAR_TILE = 7;                      // named barriers; 6 is the scalar group's
  // The vector buffers change hands through two pairs of mbarriers (full[b]: 4 producer warps arrive, empty[b]: NC consumer warps
  // arrive) instead of counted named barriers: a bar.sync over producers + consumers also made every consumer warp wait for the
  // SLOWEST consumer warp of the previous input before it could start the next one (ncu: a quarter of the consumers' samples sat at
  // that barrier while the producers were idle too).  With mbarriers a consumer warp only waits for the producers.
  // shared memory (doubles)
  static constexpr int S_CT = 0;                                        // [T][LDC]        tile of C_a (diagonal pairs)
  static constexpr int S_COL = S_CT + T * LDC;                          // [NBUF][KS][T][4] extended column vectors B_j
  static constexpr int S_ROW = S_COL + NBUF * FBUF;                     // [NBUF][KS][T][4] extended row vectors A_i
  static constexpr int S_WGT = S_ROW + NBUF * FBUF;                     // [NBUF][2][T]    beta of the rows / columns (off-diagonal pairs)
  static constexpr int S_RED = S_WGT + NBUF * WBUF;                       // [NRED][NC][32]  lane partials of the last NRED inputs
  static constexpr int S_ETAB = S_RED + NRED * DBUF;                    // [256][REP]
  static constexpr int S_PKBUF = S_ETAB + 256 * REP;                    // [NBUF][PairPack<D>::SIZE]
  static constexpr int NSTAGE = 3;                                      // tagged words of the packs in flight (cp.async landing zone)
  static constexpr int S_STAGE = (S_PKBUF + NBUF * PairPack<D>::SIZE + 1) & ~1;     // [NSTAGE][PairPack<D>::SIZE][2] 64-bit words, 16-byte aligned
  static constexpr int S_MBAR = S_STAGE + NSTAGE * 2 * PairPack<D>::SIZE;   // full[NBUF], empty[NBUF] (mbarriers of the vector buffers)
  static constexpr int S_TOTAL = S_MBAR + 2 * NBUF;
  static_assert(S_STAGE % 2 == 0, "staging area must be 16-byte aligned");
};

// both contraction roles: (re)load the C tile of a diagonal pair; `ctid` = index among the NTC contraction threads
template <int D, int NC>
__device__ __forceinline__ void persist_load_tile(const PersistFwdParams& P, const gpp_slot& sl, double* Ct, int ctid) {
  using F = PersistFwdCfg<D, NC>;
  role_bar_sync<F::BAR_TILE, F::NTC>();
  const double* Ca = P.C + (size_t)sl.a * P.M * P.M;
  for (int idx = ctid; idx < F::T * F::T; idx += F::NTC) {
    const int ii = idx / F::T, jj = idx % F::T;
    const int ig = sl.ti * F::T + ii, jg = sl.tj * F::T + jj;
    Ct[ii * F::LDC + jj] = (jg < P.M && ig < P.M) ? Ca[(size_t)ig * P.M + jg] : 0.0;
  }
  role_bar_sync<F::BAR_TILE, F::NTC>();
}

// The contraction warps walk their work as SEGMENTS of inputs that share a C tile: a CTA that owns one tile (the usual case: no more
// tiles than SMs) has a single segment of H * N (step, rollout) items and its pipeline never drains; a CTA that owns several tiles
// has one segment of N items per (step, tile) and re-loads the tile in between.
struct PersistSegments {
  int nmy, nseg, seg_items;
  bool cont;
  __device__ PersistSegments(int nslots, int N, int H, int debug = 0) {
    const int G = gridDim.x, c = blockIdx.x;
    nmy = nslots > c ? (nslots - 1 - c) / G + 1 : 0;
    // Item j's pack needs the partial of item j - N (same rollout, previous step), and the pipeline publishes item j - NBUF before it
    // waits for pack j (few rollouts; with many it publishes later in iteration j, persist_fwd_producer): continuous operation
    // needs N >= kPersistFwdBuffers.  Fewer rollouts drain every step instead.
    cont = nmy == 1 && N >= kPersistFwdBuffers && !(debug & 1);
    nseg = cont ? 1 : H * nmy;
    seg_items = cont ? H * N : N;
    if (H == 0) nseg = 0;
  }
  __device__ int slot(int seg) const { return blockIdx.x + (cont ? 0 : seg % nmy) * gridDim.x; }
  __device__ int first_step(int seg) const { return cont ? 0 : seg / nmy; }
};

template <int D, int NC>
__device__ void persist_fwd_producer(const PersistFwdParams& P, double* smem) {
  using F = PersistFwdCfg<D, NC>;
  using PP = PairPack<D>;
  constexpr int T = F::T, PT = F::PT, KS = F::KS, NTC = F::NTC;
  constexpr int NPV = (PP::SIZE + PT - 1) / PT;
  double* Ct = smem + F::S_CT;
  double* colB = smem + F::S_COL;
  double* rowA = smem + F::S_ROW;
  double* wgt = smem + F::S_WGT;
  const double* red = smem + F::S_RED;
  double* pkbuf = smem + F::S_PKBUF;
  const int ptid = threadIdx.x - kGroupThreads, lane = ptid & 31;
  const bool is_row = ptid < T;                       // threads 0..63: row ptid of the tile; 64..127: column ptid - 64
  const int idx = is_row ? ptid : ptid - T;
  const int N = P.r.N;
  const size_t pk_stride = (size_t)P.npairs * PP::SIZE;
  const PersistSegments segs(P.nslots, N, P.H, P.debug);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + F::S_MBAR);
  uint64_t* empty = full + F::NBUF;
  unsigned pe = 0;                                     // bit b: parity of the next phase of empty[b] (carried across segments)
  int resident = -1;
  for (int seg = 0; seg < segs.nseg; ++seg) {
    const int slot = segs.slot(seg), t0 = segs.first_step(seg);
    const gpp_slot sl = P.slots[slot];
    const bool diag = sl.a == sl.b;
    if (slot != resident) {
      if (diag) persist_load_tile<D, NC>(P, sl, Ct, ptid);
      resident = slot;
    }
    // this thread's centre and weight do not change over the segment
    const int lat = is_row ? sl.a : sl.b;
    const int g = (is_row ? sl.ti : sl.tj) * T + idx;
    double z[D];
#pragma unroll
    for (int d = 0; d < D; ++d) z[d] = g < P.M ? P.Z[((size_t)lat * P.M + g) * D + d] : 0.0;
    const double bw = (!diag && g < P.M) ? P.beta[(size_t)lat * P.M + g] : 0.0;
    const unsigned long long* pk0 = P.packs_ll + 2 * (size_t)sl.pair * PP::SIZE;
    unsigned long long* out = P.part_ll + 2 * (size_t)slot;
    // coefficient pack of item j = tagged words written by rollout (j % N)'s scalar group for step t0 + j / N.  They are fetched TWO
    // items ahead with cp.async into a shared-memory landing zone (16 bytes = one tagged value per thread): a register prefetch
    // would hold a scoreboard slot across iterations and make every shared-memory read of the loop wait for the ~1 us L2 round trip
    // (measured: +0.2 us per item).  Speculative: the words may not have arrived yet; tags are tested when the item's turn comes.
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(smem + F::S_STAGE);
    // (rollout, step, landing slot) of items j, j - 2 and j + 2 are carried as counters: a division per use costs ~20 dependent
    // instructions on the hand-over path
    auto next_n = [&](int n) { return n + 1 == N ? 0 : n + 1; };
    auto issue = [&](int n, int slot3) {
      const unsigned long long* src = pk0 + 2 * (size_t)n * pk_stride;
      unsigned long long* dst = stage + (size_t)slot3 * 2 * PP::SIZE;
#pragma unroll
      for (int q = 0; q < NPV; ++q)
        if (ptid + q * PT < PP::SIZE) cp_async16(dst + 2 * (ptid + q * PT), src + 2 * (ptid + q * PT));
      cp_async_commit();
    };
    cp_async_wait<0>();                      // nothing of a previous segment is still landing
    issue(0, 0);
    issue(next_n(0), 1);                     // (an item index past the segment only fetches an existing pack a second time)
    int cn = 0, ct = t0, cs = 0;             // item j:      rollout, step, landing slot
    int rn = 0, rt = t0;                     // item j - NBUF:  rollout, step
    int in2 = next_n(next_n(0));             // item j + 2:  rollout (its landing slot is (cs + 2) % 3)
    // With many rollouts in flight the partial of item j - 2 is not urgent (its reader is the scalar stage of the NEXT step, N items
    // away): its reduction stays off the path EMPTY -> pack -> vectors -> FULL that the consumers wait for, and the two COLUMN warps
    // take turns at it — a column vector costs half the flops of a row vector (no R^T z), so those warps have the slack (ncu,
    // before: warp 0 reduced ahead of the pack wait = 43 % of its time, with the other three idle at the producers' barrier).
    // With few rollouts the partial IS the critical path and is published first.
    // Fixed order: over the consumer warps, then the lanes; tagged words, no fence (persist_common.cuh).
    const bool urgent = !segs.cont || N < 16;
    auto reduce_and_publish = [&](int item) {
      if ((ptid >> 5) != 2 + (urgent ? 0 : (item & 1))) return;
      const double* rp = red + (item & (F::NRED - 1)) * F::DBUF + lane;
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < F::NC; ++w) sum += rp[w * 32];
      sum = warp_sum(sum);
      if (lane == 0) {
        ll_store(out + 2 * (size_t)rn * P.nslots, sum, (unsigned)(rt + 1));
        if (!(P.debug & 2)) hint_add(P.part_hint + rn);
      }
    };
    static_assert(F::NBUF == 2, "b = j & 1 below");
#pragma unroll 2
    for (int j = 0; j < segs.seg_items + F::NBUF; ++j) {
      const int b = j & 1;                   // (unrolled by two: a compile-time constant in each copy)
      if (j >= F::NBUF) {                    // consumers are done with item j-NBUF: its vector buffers are free, its partial is complete
        mbar_wait_bounded(&empty[b], (pe >> b) & 1u);
        pe ^= 1u << b;
        if (urgent) reduce_and_publish(j - F::NBUF);
      }
      if (j < segs.seg_items) {
        double* pk = pkbuf + b * PP::SIZE;
        {
          // every thread waits for ITS words only (they go to the pack buffer next; the barrier after that store is the only one needed)
          const unsigned tag = (unsigned)(ct + 1);
          cp_async_wait<1>();                // the group of item j has landed (item j + 1's may still be in flight)
          const unsigned long long* mine = stage + (size_t)cs * 2 * PP::SIZE;
          const unsigned long long* src = pk0 + 2 * (size_t)cn * pk_stride;
#pragma unroll
          for (int q = 0; q < NPV; ++q) {
            const int e = ptid + q * PT;
            if (e < PP::SIZE) {
              unsigned long long w0 = mine[2 * e], w1 = mine[2 * e + 1];
              if (!ll_ready(w0, w1, tag)) pk[e] = ll_load(src + 2 * e, tag);   // not there yet: the rollout's scalar stage is still running
              else pk[e] = ll_value(w0, w1);
            }
          }
        }
        named_bar_sync_imm<F::BAR_PROD>(PT);   // pack j visible to the producers; pack j-NBUF (same buffer) no longer read
        issue(in2, cs == 0 ? 2 : cs - 1);      // (lands in the slot item j - 1 used; always committed so that the group count stays in step)
        double ext[4 * KS], zc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zc[d] = z[d] - pk[PP::MU + d];
        if (is_row) {                          // A_i = [R^T z1' (D), c0 + z1'^T P1 z1', 1, 0..]
#pragma unroll
          for (int e = 0; e < D; ++e) {
            double tt = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) tt = fma(zc[d], pk[PP::R + d * D + e], tt);
            ext[e] = tt;
          }
          ext[D] = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zc);
          ext[D + 1] = 1.0;
        } else {                               // B_j = [z2' (D), 1, z2'^T P2 z2', 0..]
#pragma unroll
          for (int d = 0; d < D; ++d) ext[d] = zc[d];
          ext[D] = 1.0;
          ext[D + 1] = packed_quad<D>(pk + PP::P2, zc);
        }
#pragma unroll
        for (int e = D + 2; e < 4 * KS; ++e) ext[e] = 0.0;
        double* dst = (is_row ? rowA : colB) + b * F::FBUF + idx * 4;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int q = 0; q < 4; q += 2)
            *reinterpret_cast<double2*>(dst + ks * T * 4 + q) = make_double2(ext[ks * 4 + q], ext[ks * 4 + q + 1]);
        wgt[b * F::WBUF + (is_row ? 0 : T) + idx] = bw;
        __syncwarp();                          // the warp's stores are ordered before lane 0's arrive (release at CTA scope)
        if (lane == 0) mbar_arrive(&full[b]);
        cn = next_n(cn);
        if (cn == 0) ++ct;
        cs = cs == 2 ? 0 : cs + 1;
        in2 = next_n(in2);
      }
      if (j >= F::NBUF) {
        if (!urgent) reduce_and_publish(j - F::NBUF);
        rn = next_n(rn);
        if (rn == 0) ++rt;
      }
    }
  }
}

template <int D, int NC>
__device__ void persist_fwd_consumer(const PersistFwdParams& P, double* smem) {
  using F = PersistFwdCfg<D, NC>;
  constexpr int T = F::T, KS = F::KS, LDC = F::LDC, NTC = F::NTC;
  double* Ct = smem + F::S_CT;
  double* colB = smem + F::S_COL;
  double* rowA = smem + F::S_ROW;
  double* wgt = smem + F::S_WGT;
  double* red = smem + F::S_RED;
  double* etab = smem + F::S_ETAB;
  const int ctid = threadIdx.x - 2 * kGroupThreads, lane = ctid & 31, cwarp = ctid >> 5;
  const int strip = cwarp & 7, col0 = (cwarp >> 3) * F::COLS;       // 8-row strip and first column of this warp's part of the tile
  const int row = strip * 8 + (lane >> 2);
  const int cpair = 2 * (lane & 3);
  const double* ct = Ct + row * LDC + col0 + cpair;
  const unsigned etab_lane = (unsigned)__cvta_generic_to_shared(etab + (lane & (F::REP - 1)));
  const PersistSegments segs(P.nslots, P.r.N, P.H, P.debug);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + F::S_MBAR);
  uint64_t* empty = full + F::NBUF;
  unsigned pf = 0;                                     // bit b: parity of the next phase of full[b]
  int resident = -1;
  for (int seg = 0; seg < segs.nseg; ++seg) {
    const int slot = segs.slot(seg);
    const gpp_slot sl = P.slots[slot];
    const bool diag = sl.a == sl.b;
    if (slot != resident) {
      if (diag) persist_load_tile<D, NC>(P, sl, Ct, F::PT + ctid);
      resident = slot;
    }
    static_assert(F::NBUF == 2, "b = j & 1 below");
    // unrolled by two items: the buffer index becomes a compile-time constant in each copy (shared-memory addresses fold)
#pragma unroll 2
    for (int j = 0; j < segs.seg_items; ++j) {
      const int b = j & 1;
      mbar_wait_bounded(&full[b], (pf >> b) & 1u);
      pf ^= 1u << b;
      const double* ra = rowA + b * F::FBUF + strip * 32 + lane;
      const double* cb = colB + b * F::FBUF + col0 * 4 + lane;
      const double* wsrc = diag ? ct : wgt + b * F::WBUF + T + col0 + cpair;
      double a[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) a[ks] = ra[ks * T * 4];
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
      for (int cg = 0; cg < F::COLS / 8; cg += 2) {
        double tt[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          dmma_m8n8k4(tt[0], tt[1], a[ks], cb[ks * T * 4 + cg * 32]);
          dmma_m8n8k4(tt[2], tt[3], a[ks], cb[ks * T * 4 + cg * 32 + 32]);
        }
        exp_tab_contract<4, F::REP>(tt, etab_lane);
        const double2 w0 = *reinterpret_cast<const double2*>(wsrc + cg * 8);
        const double2 w1 = *reinterpret_cast<const double2*>(wsrc + cg * 8 + 8);
        acc0 = fma(tt[0], w0.x, acc0);
        acc1 = fma(tt[2], w1.x, acc1);
        acc0 = fma(tt[1], w0.y, acc0);
        acc1 = fma(tt[3], w1.y, acc1);
      }
      double total = acc0 + acc1;
      if (!diag) total *= wgt[b * F::WBUF + row];
      red[(j & (F::NRED - 1)) * F::DBUF + cwarp * 32 + lane] = total;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }
  }
}

template <int D, int NC>
__global__ void __launch_bounds__(PersistFwdCfg<D, NC>::THREADS, 1) k_rollout_fwd_persist(const PersistFwdParams P) {
  using F = PersistFwdCfg<D, NC>;
  constexpr int kFwdThreads = F::THREADS;
  extern __shared__ __align__(16) double smem[];
  double* etab = smem + F::S_ETAB;
  for (int i = threadIdx.x; i < kContractTab * F::REP; i += kFwdThreads) etab[i] = kExp2Tab256[i / F::REP];
  if (threadIdx.x == 0) {
    uint64_t* mb = reinterpret_cast<uint64_t*>(smem + F::S_MBAR);
    for (int i = 0; i < F::NBUF; ++i) {
      mbar_init(mb + i, F::NP);
      mbar_init(mb + F::NBUF + i, F::NC);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  // Register reallocation, then a CTA-wide barrier BEFORE any role starts: contraction warps of a CTA that owns no tile return at
  // once, and a warp that exits while the registers it released are still unclaimed in the CTA pool takes them with it — the scalar
  // warps' allocation then never completes (observed as a hard hang, no time-out involved, only when the exit wins the race).
  // (The barrier is repeated inside each branch: ptxas budgets registers for the code DOMINATED by a setmaxnreg, so each role must
  //  stay in the branch of its own setmaxnreg — behind a common merge point every role would be compiled for the launch count.)
  if (warp < 4) {
    warpgroup_reg_inc<F::REGS_SCALAR>();
    role_bar_sync<8, kFwdThreads>();
    persist_fwd_scalar<D>(P);
  } else if (warp < 8) {
    warpgroup_reg_dec<F::REGS_CONTRACT>();
    role_bar_sync<8, kFwdThreads>();
    persist_fwd_producer<D, NC>(P, smem);
  } else {
    warpgroup_reg_dec<F::REGS_CONTRACT>();
    role_bar_sync<8, kFwdThreads>();
    persist_fwd_consumer<D, NC>(P, smem);
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward: the reverse sweep.  Per step t = H-1 .. 0 and rollout n (same order as gpp_rollout_mm_bwd's one-launch-per-stage path):
//   scalar group   bwd_post (cost gradient of state t+1, adjoint of the Euler update)  ->  prologue of the predict's adjoint
//                  (coefficient packs of the unordered kernel pairs, Psi1 forward + adjoint, un-mixing)  ->  publish ready[n]
//   contraction    L (L+1)/2 x nrb work items (contract_grad_item) per (step, rollout), drawn from one ticket counter in
//                  (step, rollout) order by the 16 contraction warps of every CTA  ->  statistics as tagged words
//   scalar group   bwd_finalize (D x D algebra per pair) -> bwd_pre (joint assembly, squashing link, policy, encoder adjoints)
// ---------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = kGroupThreads + kGradThreads;    // 4 scalar + 16 contraction warps

struct PersistBwdParams {
  RolloutMMParams r;
  RolloutBwdBuffers bw;
  PersistSaved sv;
  int H;
  const double *traj_m, *traj_S, *loss_bar;
  const double* cg;                     // [H,N,ndir] cost gradients of the states 1..H (k_cost_grad_ring, one launch)
  // dynamics model
  const double *Z, *ell, *var, *beta, *C, *W;
  int M, Lm, Pm;
  // workspace of the predict's adjoint
  double *packs, *Gs, *stats, *f1lat, *crosslat, *f1lat_bar, *crosslat_bar, *omega, *gm, *gS;
  int nrb;
  unsigned *ready, *ticket;             // [N], [1]; zero at launch (so is `stats`: tagged words, tag = sweep index + 1)
  unsigned* stat_hint;                  // [N] work items finished so far (hint counter, persist_common.cuh)
};

template <int D>
__device__ void persist_bwd_scalar(const PersistBwdParams& P, double* fsm) {
  constexpr int DP = D - 1;
  const int tid = threadIdx.x, G = gridDim.x;
  RolloutMMParams p = P.r;
  const RolloutBwdBuffers& bw = P.bw;
  const int N = p.N, Dx = p.Dx, Lm = P.Lm;
  const int ndir = Dx + Dx * (Dx + 1) / 2;
  const size_t sm = (size_t)N * Dx, sS = (size_t)N * Dx * Dx;
  const int items = Lm * (Lm + 1) / 2 * P.nrb;
  const int n0 = G - 1 - (int)blockIdx.x;
  __shared__ double li_sm[GPP_MAX_L * (D * D + 1)];
  BwdPrepareParams bp;
  bp.f1_bar = bw.f1_bar; bp.Sff_bar = bw.Sff_bar; bp.cross_bar = bw.cross_bar; bp.f1lat = P.f1lat; bp.W = P.W;
  bp.f1lat_bar = P.f1lat_bar; bp.crosslat_bar = P.crosslat_bar; bp.omega = P.omega;
  bp.N = N; bp.L = Lm; bp.P = P.Pm; bp.D = D; bp.full_cov = 1;
  BwdFinalizeParams fp;
  fp.m = nullptr; fp.S = nullptr; fp.ell = P.ell; fp.stats = P.stats; fp.omega = P.omega; fp.Gs = P.Gs; fp.gm = P.gm; fp.gS = P.gS;
  fp.m_bar = bw.md_bar; fp.S_bar = bw.Sd_bar; fp.N = N; fp.L = Lm; fp.nrb = P.nrb;
  auto set_step = [&](int t) {
    p.m = const_cast<double*>(P.traj_m) + (size_t)t * sm;       // read-only
    p.S = const_cast<double*>(P.traj_S) + (size_t)t * sS;
    persist_step_pointers(p, P.sv, P.r, t);
  };
  for (int t = P.H - 1; t >= -1; --t) {
    for (int n = n0; n < N; n += G) {
      if (t < P.H - 1) {
        // step t+1 of rollout n: wait for its statistics (tagged words) -> adjoint of the joint moments, then of the pre stage
        if (tid == 0) hint_wait_ge(P.stat_hint + n, (unsigned)(P.H - 1 - t) * (unsigned)items);
        group_sync();
        set_step(t + 1);
        fp.ll_tag = (unsigned)(P.H - 1 - t);
        bwd_finalize_body<D>(fp, n, fsm);
        group_sync();
        bwd_pre_body<DP>(p, bw, n);
        group_sync();
      }
      if (t < 0) continue;
      set_step(t);
      if (tid < 32) bwd_post_body(p, n, P.cg + ((size_t)t * N) * ndir, P.loss_bar, bw);
      group_sync();
      {
        // coefficient packs (and (Sigma + V_ab)^-1) of the unordered kernel pairs: pair u on lane u / 4 of warp u % 4
        const int u = (tid & 31) * 4 + (tid >> 5);
        if (u < Lm * (Lm + 1) / 2) {
          int a = 0, b = u;
          while (b >= Lm - a) { b -= Lm - a; ++a; }
          b += a;
          pack_body<D>(n * Lm * Lm + a * Lm + b, p.md, p.Sd, N, P.ell, P.var, nullptr, Lm * Lm, Lm, P.packs, P.Gs, p.info);
        }
      }
      psi1_body<D>(n, p.md, p.Sd, N, Lm, P.M, P.Z, P.ell, P.var, P.beta, P.f1lat, P.crosslat, p.info, li_sm);
      group_sync();
      bwd_prepare_input(bp, n);
      psi1_bwd_body<D>(n, p.md, p.Sd, Lm, P.M, P.Z, P.ell, P.var, P.beta, P.f1lat_bar, P.crosslat_bar, P.gm, P.gS, li_sm);
      __threadfence();
      group_sync();
      if (tid == 0) st_release_u32(P.ready + n, (unsigned)(P.H - t));
    }
  }
}

template <int D>
__device__ void persist_bwd_contract(const PersistBwdParams& P, double* smem) {
  __shared__ unsigned s_ticket[2];
  const int gtid = threadIdx.x - kGroupThreads;
  const int N = P.r.N, Lm = P.Lm;
  const int items = Lm * (Lm + 1) / 2 * P.nrb;
  const unsigned per_step = (unsigned)N * (unsigned)items, total = per_step * (unsigned)P.H;
  unsigned next = gtid == 0 ? atomicAdd(P.ticket, 1u) : 0u;
  for (int iter = 0;; ++iter) {
    if (gtid == 0) {
      const unsigned k = next;
      if (k < total) {
        spin_wait_ge(P.ready + (k % per_step) / items, k / per_step + 1u);
        next = atomicAdd(P.ticket, 1u);     // the following item's ticket: the round trip hides behind this item's contraction
      }
      s_ticket[iter & 1] = k;
    }
    grad_sync<true>();
    const unsigned k = s_ticket[iter & 1];
    if (k >= total) break;
    const int rem = (int)(k % per_step);
    const int n = rem / items, it = rem % items;
    // the statistics leave as tagged words (tag = sweep index + 1): nothing to fence, nothing to count
    contract_grad_item<D, true>(smem, gtid, n, it / P.nrb, it % P.nrb, P.Z, P.beta, P.C, P.packs, P.omega, P.stats, P.M, Lm, P.nrb, false,
                                k / per_step + 1u);
    if (gtid == 0) hint_add(P.stat_hint + n);        // issued by one thread, possibly before the other warps' words: only a hint
  }
}

template <int D>
__global__ void __launch_bounds__(kBwdThreads, 1) k_rollout_bwd_persist(const PersistBwdParams P, const int fsm_offset) {
  using CF = GradCfg<D>;
  extern __shared__ __align__(16) double smem[];
  double* etab = smem + CF::ETAB;
  for (int i = threadIdx.x; i < 256 * CF::REP; i += kBwdThreads) etab[i] = kExp2Tab256[i / CF::REP];
  __syncthreads();
  // (reallocate, then synchronise the CTA before any warp can return: see k_rollout_fwd_persist)
  // 192 / 72 registers: the contraction main loop is register-starved at 64 (measured 148 -> 145 us per step at 64 rollouts; 160 / 80
  // loses it again in the scalar stages: 146.5, and one rollout goes from 86 to 89.5 us per step)
  if (threadIdx.x < kGroupThreads) {
    warpgroup_reg_inc<192>();
    role_bar_sync<8, kBwdThreads>();
    persist_bwd_scalar<D>(P, smem + fsm_offset);
  } else {
    warpgroup_reg_dec<72>();
    role_bar_sync<8, kBwdThreads>();
    persist_bwd_contract<D>(P, smem);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int g_rollout_mode = -1;   // -1: not read yet; GPP_ROLLOUT_AUTO / _LEGACY / _PERSIST

int rollout_mode() {
  if (g_rollout_mode < 0) {
    const char* e = std::getenv("GPP_ROLLOUT_MODE");
    g_rollout_mode = e ? std::atoi(e) : GPP_ROLLOUT_AUTO;
    if (g_rollout_mode < GPP_ROLLOUT_AUTO || g_rollout_mode > GPP_ROLLOUT_PERSIST) g_rollout_mode = GPP_ROLLOUT_AUTO;
  }
  return g_rollout_mode;
}

static bool cooperative_ok() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
    ok = v ? 1 : 0;
  }
  return ok == 1;
}

static size_t pack_doubles(int D) {
  switch (D) {
#define GPP_CASE(d) case d: return PairPack<d>::SIZE;
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: return 0;
  }
}

bool persist_fwd_supported(const gpp_gp_model* dyn, int N, int Dx) {
  (void)N; (void)Dx;
  if (!dyn || dyn->D < 2 || dyn->D > 8 || !cooperative_ok()) return false;
  const gpp_gp_model::SlotTable& tab = dyn->tables[0][0];
  // a CTA re-loads its C tile whenever it owns more than one: fine for a few, the multi-launch path amortises better beyond that
  return tab.nslots <= 2 * num_sms();
}

PersistFwdLayout persist_fwd_layout(const gpp_gp_model* dyn, int N) {
  PersistFwdLayout lo{};
  const gpp_gp_model::SlotTable& tab = dyn->tables[0][0];
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  lo.packs = take(sizeof(unsigned long long) * 2 * pack_doubles(dyn->D) * tab.npairs * N);
  lo.part = take(sizeof(unsigned long long) * 2 * (size_t)tab.nslots * N);
  lo.f1lat = take(sizeof(double) * (size_t)dyn->L * N);
  lo.crosslat = take(sizeof(double) * (size_t)dyn->L * dyn->D * N);
  lo.flags = take(sizeof(unsigned) * (size_t)N);
  lo.total = off;
  return lo;
}

template <int D, int NC>
static int launch_fwd_persist_nc(PersistFwdParams& P, int grid, cudaStream_t stream) {
  using F = PersistFwdCfg<D, NC>;
  const size_t smem = sizeof(double) * F::S_TOTAL;
  GPP_CUDA_OK(cudaFuncSetAttribute(k_rollout_fwd_persist<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&P};
  GPP_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_rollout_fwd_persist<D, NC>, dim3(grid), dim3(F::THREADS), args, smem, stream));
  count_launch();
  return GPP_OK;
}

template <int D>
static int launch_fwd_persist(PersistFwdParams& P, int grid, cudaStream_t stream) {
  // NC = 16 (two consumer warps per strip, scalar group squeezed to 112 registers) was measured and is SLOWER (64 restarts: 110 vs 87 us
  // per step): with half the columns per warp the per-input hand-over dominates.  It stays a template parameter, not an instantiation.
  return launch_fwd_persist_nc<D, 8>(P, grid, stream);
}

int rollout_mm_fwd_persist(const gpp_gp_model* dyn, const RolloutMMParams& r, int H, const double* m0, const double* S0, double* m_final,
                           double* S_final, double* saved, char* ws_persist, cudaStream_t stream) {
  const int N = r.N;
  const gpp_gp_model::SlotTable& tab = dyn->tables[0][0];
  const PersistFwdLayout lo = persist_fwd_layout(dyn, N);
  PersistFwdParams P{};
  P.r = r;
  P.sv.base = saved;
  if (saved) {
    const RolloutSaved sv(N, r.Dx, r.D, r.L);
    P.sv.per_step = sv.per_step; P.sv.md = sv.md; P.sv.Sd = sv.Sd; P.sv.Sxd = sv.Sxd; P.sv.cross = sv.cross; P.sv.pre = sv.pre;
  }
  P.H = H; P.m0 = m0; P.S0 = S0; P.m_final = m_final; P.S_final = S_final;
  P.Z = dyn->Z; P.ell = dyn->ell; P.var = dyn->var; P.beta = dyn->beta; P.C = dyn->C; P.mean = dyn->mean; P.W = dyn->W;
  P.M = dyn->M; P.Lm = dyn->L; P.Pm = dyn->P; P.model_uncertainty = dyn->model_uncertainty;
  P.slots = tab.d_slots; P.pair_start = tab.d_pair_start; P.pair_ab = tab.d_pair_ab; P.npairs = tab.npairs; P.nslots = tab.nslots;
  {
    const char* e = std::getenv("GPP_PERSIST_DEBUG");
    P.debug = e ? std::atoi(e) : 0;
  }
  P.packs_ll = (unsigned long long*)(ws_persist + lo.packs); P.part_ll = (unsigned long long*)(ws_persist + lo.part);
  P.f1lat = (double*)(ws_persist + lo.f1lat); P.crosslat = (double*)(ws_persist + lo.crosslat);
  P.part_hint = (unsigned*)(ws_persist + lo.flags);
  GPP_CUDA_OK(cudaMemsetAsync(ws_persist + lo.packs, 0, lo.f1lat - lo.packs, stream));   // tags of the packs and the partials
  GPP_CUDA_OK(cudaMemsetAsync(ws_persist + lo.flags, 0, sizeof(unsigned) * (size_t)N, stream));
  const int grid = std::min(num_sms(), tab.nslots + N);
  switch (dyn->D) {
#define GPP_CASE(d) case d: return launch_fwd_persist<d>(P, grid, stream);
    GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default:
      set_error("persistent rollout: unsupported dynamics input dimension %d", dyn->D);
      return GPP_ERR_UNSUPPORTED;
  }
}

// ---- backward ---------------------------------------------------------------------------------------------
bool persist_bwd_supported(const gpp_gp_model* dyn, int N, int Dx) {
  (void)N;
  if (!dyn || dyn->D < 2 || dyn->D > 8 || !cooperative_ok()) return false;
  if (Dx + Dx * (Dx + 1) / 2 > 128) return false;
  const int ncb = (dyn->M + kGradCols - 1) / kGradCols;
  const int npairs = dyn->L * (dyn->L + 1) / 2;
  size_t fin = 0, fixed = 0;
  switch (dyn->D) {
#define GPP_CASE(d) case d: fin = (size_t)npairs * FinalizeSmem<d>::PER_PAIR; fixed = GradCfg<d>::FIXED; break;
    GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
  }
  return sizeof(double) * (fixed + (size_t)ncb * kGradCols + fin) + 16 * 1024 <= 227 * 1024;   // + the scalar stages' static arrays
}

PersistBwdLayout persist_bwd_layout(const gpp_gp_model* dyn, int N, int Dx, int H) {
  PersistBwdLayout lo{};
  const int D = dyn->D, L = dyn->L;
  size_t stat_doubles = 0;
  switch (D) {
#define GPP_CASE(d) case d: stat_doubles = GradStats<d>::SIZE; break;
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
  }
  lo.nrb = (dyn->M + kGradRows - 1) / kGradRows;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  lo.packs = take(sizeof(double) * pack_doubles(D) * L * L * N);
  lo.Gs = take(sizeof(double) * (size_t)D * D * L * L * N);
  lo.stats = take(2 * sizeof(double) * stat_doubles * L * L * lo.nrb * N);   // tagged words
  lo.f1lat = take(sizeof(double) * (size_t)L * N);
  lo.crosslat = take(sizeof(double) * (size_t)L * D * N);
  lo.f1lat_bar = take(sizeof(double) * (size_t)L * N);
  lo.crosslat_bar = take(sizeof(double) * (size_t)L * D * N);
  lo.omega = take(sizeof(double) * (size_t)L * L * N);
  lo.gm = take(sizeof(double) * (size_t)L * D * N);
  lo.gS = take(sizeof(double) * (size_t)L * D * D * N);
  lo.cg = take(sizeof(double) * (size_t)std::max(H, 1) * N * (Dx + Dx * (Dx + 1) / 2));
  lo.flags = take(sizeof(unsigned) * (2 * (size_t)N + 8));
  lo.total = off;
  return lo;
}

template <int D>
static int launch_bwd_persist(PersistBwdParams& P, int grid, cudaStream_t stream) {
  const int ncb = (P.M + kGradCols - 1) / kGradCols;
  const int npairs = P.Lm * (P.Lm + 1) / 2;
  int fsm_offset = GradCfg<D>::FIXED + ncb * kGradCols;
  fsm_offset = (fsm_offset + 1) & ~1;
  const size_t smem = sizeof(double) * ((size_t)fsm_offset + (size_t)npairs * FinalizeSmem<D>::PER_PAIR);
  GPP_CUDA_OK(cudaFuncSetAttribute(k_rollout_bwd_persist<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&P, (void*)&fsm_offset};
  GPP_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_rollout_bwd_persist<D>, dim3(grid), dim3(kBwdThreads), args, smem, stream));
  count_launch();
  return GPP_OK;
}

int rollout_mm_bwd_persist(const gpp_gp_model* dyn, const RolloutMMParams& r, const RolloutBwdBuffers& bw, int H, const double* traj_m,
                           const double* traj_S, const double* saved, const double* loss_bar, char* ws_persist, cudaStream_t stream) {
  if (H <= 0) return GPP_OK;
  const int N = r.N, Dx = r.Dx;
  const PersistBwdLayout lo = persist_bwd_layout(dyn, N, Dx, H);
  auto D_ = [&](size_t off) { return (double*)(ws_persist + off); };
  PersistBwdParams P{};
  P.r = r; P.bw = bw;
  const RolloutSaved sv(N, Dx, r.D, r.L);
  P.sv.base = const_cast<double*>(saved);
  P.sv.per_step = sv.per_step; P.sv.md = sv.md; P.sv.Sd = sv.Sd; P.sv.Sxd = sv.Sxd; P.sv.cross = sv.cross; P.sv.pre = sv.pre;
  P.H = H; P.traj_m = traj_m; P.traj_S = traj_S; P.loss_bar = loss_bar;
  P.Z = dyn->Z; P.ell = dyn->ell; P.var = dyn->var; P.beta = dyn->beta; P.C = dyn->C; P.W = dyn->W;
  P.M = dyn->M; P.Lm = dyn->L; P.Pm = dyn->P;
  P.packs = D_(lo.packs); P.Gs = D_(lo.Gs); P.stats = D_(lo.stats); P.f1lat = D_(lo.f1lat); P.crosslat = D_(lo.crosslat);
  P.f1lat_bar = D_(lo.f1lat_bar); P.crosslat_bar = D_(lo.crosslat_bar); P.omega = D_(lo.omega); P.gm = D_(lo.gm); P.gS = D_(lo.gS);
  P.nrb = lo.nrb;
  P.ready = (unsigned*)(ws_persist + lo.flags); P.stat_hint = P.ready + N; P.ticket = P.ready + 2 * N;
  GPP_CUDA_OK(cudaMemsetAsync(ws_persist + lo.flags, 0, sizeof(unsigned) * (2 * (size_t)N + 8), stream));
  GPP_CUDA_OK(cudaMemsetAsync(ws_persist + lo.stats, 0, lo.f1lat - lo.stats, stream));
  // cost gradients of all H trajectory states in one launch (dual numbers through the encoder and expected-cost rules)
  const int ndir = Dx + Dx * (Dx + 1) / 2;
  double* cg = D_(lo.cg);
  {
    const long long total = (long long)H * N * ndir;
    k_cost_grad_ring<<<(unsigned)((total + 63) / 64), 64, 0, stream>>>(r, traj_m + (size_t)N * Dx, traj_S + (size_t)N * Dx * Dx, H, ndir, cg);
    count_launch();
  }
  P.cg = cg;
  const int items = P.Lm * (P.Lm + 1) / 2 * lo.nrb;
  const int grid = (int)std::min<long long>(num_sms(), (long long)N * items + N);
  switch (dyn->D) {
#define GPP_CASE(d) case d: return launch_bwd_persist<d>(P, grid, stream);
    GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default:
      set_error("persistent rollout: unsupported dynamics input dimension %d", dyn->D);
      return GPP_ERR_UNSUPPORTED;
  }
}

}  // namespace gpp

extern "C" {

int gpp_rollout_mm_set_mode(int mode) {
  GPP_REQUIRE(mode >= GPP_ROLLOUT_AUTO && mode <= GPP_ROLLOUT_PERSIST, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_set_mode: unknown mode %d", mode);
  gpp::g_rollout_mode = mode;
  return GPP_OK;
}

int gpp_rollout_mm_get_mode(void) { return gpp::rollout_mode(); }

}  // extern "C"
