// Fused exact moment matching through a (multi-output) sparse / exact GP  —  the hot kernel of the path.
//
// Replaces gpflow_pilco/moment_matching/models.py:200-299 (and :44-197 as special cases).  The reference
// materialises eKuffu [N,L,M,L,M] (gpflow_pilco/utils/kernel_expectation.py:217-247) and runs two batched
// triangular solves over it (models.py:224-226, O(N L^2 M^3)); here Psi2 is never written:
//
//   k_pack_psi1 (one launch):
//   pack       one thread per (input n, kernel pair ab): D x D Cholesky, coefficients of
//              log Q_ij = r_i + s_j + z1'_i^T R z2'_j  (common.cuh, PairPack)
//   psi1       one warp per (n, latent): Psi1 contracted with beta -> latent mean and pre-inverted cross term
//   k_contract persistent CTAs pull (pair, tile, input-chunk) items; a CTA keeps one T x T tile of C_a
//              (= beta beta^T - B, diagonal pairs) or the beta vectors (off-diagonal pairs) on chip and streams
//              inputs through it: per 8 x 8 block of entries 2 DMMA (exponents), then per entry 8 FP64 ops (table exp)
//              + 1 DFMA (contraction).  Diagonal pairs use the symmetry Q_aa = Q_aa^T (upper tiles only, weight 2).
//   k_finalize per input: deterministic fixed-order sum of the tile partials, Sff = f2 - f1 f1^T + diag(var),
//              optional W mixing (LinearCoregionalization, models.py:279-286), mean constant, jitter.
//
// FP64 throughout (the reference is float64; 1e-6 relative parity target).  The only GEMM-shaped piece — the exponent of
// every entry as an inner product of extended row / column vectors — runs on the FP64 tensor path (mma.sync.m8n8k4.f64, SASS
// DMMA); the contraction itself is a Hadamard-weighted reduction of an elementwise exp, not a GEMM.
#include <algorithm>

#include "model.cuh"
#include "contract_kernel.cuh"
#include "predict_kernels.cuh"

namespace gpp {

// ---------------------------------------------------------------------------------------------------------
// k_finalize
// ---------------------------------------------------------------------------------------------------------
struct FinalizeParams {
  const double* part;
  const gpp_slot* slots;
  const int* pair_start;
  const int* pair_ab;
  const double* f1lat;      // [N,L]
  const double* crosslat;   // [N,D,L]
  const double* var;        // [L]
  const double* mean;       // [P]
  const double* W;          // [P,L] or null
  double* f1;               // [N,P]
  double* Sff;              // [N,P,P]
  double* cross;            // [N,D,P]
  int N, L, P, D, npairs, nslots, full_cov, model_uncertainty;
  double jitter;
  EulerPost post;           // used by k_finalize<true> only
};

template <bool POST>
__global__ void __launch_bounds__(128) k_finalize(FinalizeParams p) {
  __shared__ double f2[GPP_MAX_L * GPP_MAX_L];
  __shared__ double SffL[GPP_MAX_L * GPP_MAX_L];
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, P = p.P;
  for (int t = threadIdx.x; t < L * L; t += blockDim.x) f2[t] = 0.0;
  __syncthreads();
  for (int pr = warp; pr < p.npairs; pr += 4) {
    double s = 0.0;
    for (int k = p.pair_start[pr] + lane; k < p.pair_start[pr + 1]; k += 32)
      s = fma(p.slots[k].weight, p.part[(size_t)n * p.nslots + k], s);
    s = warp_sum(s);
    if (lane == 0) {
      int a = p.pair_ab[2 * pr], b = p.pair_ab[2 * pr + 1];
      f2[a * L + b] = s;
      f2[b * L + a] = s;
    }
  }
  __syncthreads();
  const double* f1l = p.f1lat + (size_t)n * L;
  for (int t = threadIdx.x; t < L * L; t += blockDim.x) {
    int a = t / L, b = t % L;
    double v = f2[t] - f1l[a] * f1l[b];
    if (a == b && p.model_uncertainty) v += p.var[a];
    SffL[t] = v;
  }
  __syncthreads();
  // outputs (optionally mixed by W)
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    double v = p.mean[t];
    if (p.W) {
      for (int l = 0; l < L; ++l) v = fma(p.W[t * L + l], f1l[l], v);
    } else {
      v += f1l[t];
    }
    p.f1[(size_t)n * P + t] = v;
  }
  for (int t = threadIdx.x; t < P * P; t += blockDim.x) {
    int a = t / P, b = t % P;
    double v;
    if (p.W) {
      v = 0.0;
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < L; ++k) v = fma(p.W[a * L + l] * p.W[b * L + k], SffL[l * L + k], v);
    } else {
      v = SffL[t];
    }
    if (a == b) v += p.jitter;
    if (!p.full_cov && a != b) v = 0.0;
    p.Sff[(size_t)n * P * P + t] = v;
  }
  for (int t = threadIdx.x; t < p.D * P; t += blockDim.x) {
    int d = t / P, o = t % P;
    const double* cl = p.crosslat + ((size_t)n * p.D + d) * L;
    double v;
    if (p.W) {
      v = 0.0;
      for (int l = 0; l < L; ++l) v = fma(p.W[o * L + l], cl[l], v);
    } else {
      v = cl[o];
    }
    p.cross[((size_t)n * p.D + d) * P + o] = v;
  }
  if (POST) {
    __syncthreads();                         // this block's f1 / Sff / cross are visible to all its threads
    const EulerPost& e = p.post;
    const int Dx = e.Dx, D = p.D;
    const double* Sxd = e.Sxd + (size_t)n * Dx * D;
    const double* cr = p.cross + (size_t)n * D * P;
    for (int t = threadIdx.x; t < Dx * Dx + Dx; t += blockDim.x) {
      if (t < Dx * Dx) {
        const int i = t / Dx, j = t % Dx;
        double sij = 0.0, sji = 0.0;
        for (int b = 0; b < D; ++b) {
          sij = fma(Sxd[i * D + b], cr[b * P + j], sij);
          sji = fma(Sxd[j * D + b], cr[b * P + i], sji);
        }
        const double v = e.S[(size_t)n * Dx * Dx + t] + sij + sji + p.Sff[(size_t)n * P * P + i * P + j];
        e.S[(size_t)n * Dx * Dx + t] = v;
        if (e.traj_S) e.traj_S[(size_t)n * Dx * Dx + t] = v;
        if (e.ring_S) e.ring_S[(size_t)n * Dx * Dx + t] = v;
      } else {
        const int i = t - Dx * Dx;
        const double v = e.m[(size_t)n * Dx + i] + p.f1[(size_t)n * P + i];
        e.m[(size_t)n * Dx + i] = v;
        if (e.traj_m) e.traj_m[(size_t)n * Dx + i] = v;
        if (e.ring_m) e.ring_m[(size_t)n * Dx + i] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct Plan {
  int tile_idx;      // 0: T=64, 1: T=128
  int diag_only;
  int nchunks, chunk;
  size_t off_packs, off_part, off_f1lat, off_crosslat, off_counter, total;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Plan make_plan(const gpp_gp_model* m, int N, int full_output_cov) {
  Plan pl{};
  pl.diag_only = (!full_output_cov && !m->coreg) ? 1 : 0;
  const int sms = num_sms();
  // big tiles once there is enough work to fill the machine with them, small tiles otherwise
  const gpp_gp_model::SlotTable& big = m->tables[1][pl.diag_only];
  pl.tile_idx = ((long long)big.nslots * N >= 8LL * sms && m->M >= 128) ? 1 : 0;
  const gpp_gp_model::SlotTable& tab = m->tables[pl.tile_idx][pl.diag_only];
  // aim for ~64 items per SM (dynamic scheduling tail <= ~1.5%), but keep chunks >= 4 inputs to amortise the tile load
  long long want_items = 64LL * sms;
  const int max_chunks = (N + 3) / 4;
  int nchunks = (int)std::max(1LL, (want_items + tab.nslots - 1) / tab.nslots);
  if (nchunks >= max_chunks) {
    // small batches: too little work for 64 items per SM.  Pick the split that minimises  waves x (inputs per item + ~2 inputs'
    // worth of per-item set-up: tile of C, table, pipeline fill)
    long long best = -1;
    for (int nc = 1; nc <= max_chunks; ++nc) {
      const long long waves = ((long long)tab.nslots * nc + sms - 1) / sms;
      const long long cost = waves * ((N + nc - 1) / nc + 2);
      if (best < 0 || cost < best) { best = cost; nchunks = nc; }
    }
  }
  pl.chunk = (N + nchunks - 1) / nchunks;
  pl.nchunks = (N + pl.chunk - 1) / pl.chunk;
  const int D = m->D;
  size_t pack_doubles = 0;
  switch (D) {
#define GPP_CASE(d) case d: pack_doubles = PairPack<d>::SIZE; break;
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
  }
  size_t off = 0;
  pl.off_packs = off;    off = align_up(off + sizeof(double) * pack_doubles * tab.npairs * N, 256);
  pl.off_part = off;     off = align_up(off + sizeof(double) * (size_t)tab.nslots * N, 256);
  pl.off_f1lat = off;    off = align_up(off + sizeof(double) * (size_t)m->L * N, 256);
  pl.off_crosslat = off; off = align_up(off + sizeof(double) * (size_t)m->L * D * N, 256);
  pl.off_counter = off;  off = align_up(off + 256, 256);
  pl.total = off;
  return pl;
}

template <int D, int T, int NP, int NC>
static int launch_contract(const ContractParams& cp, cudaStream_t stream) {
  constexpr int NT = ContractCfg<D, T, NP, NC>::NT;
  size_t smem = sizeof(double) * ContractCfg<D, T, NP, NC>::TOTAL;
  static bool configured = false;
  if (!configured) {
    GPP_CUDA_OK(cudaFuncSetAttribute(k_contract<D, T, NP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int per_sm = 1;
  GPP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_contract<D, T, NP, NC>, NT, smem));
  per_sm = std::max(per_sm, 1);
  int grid = std::min(num_sms() * per_sm, cp.nslots * cp.nchunks);
  profile_begin(stream);
  k_contract<D, T, NP, NC><<<grid, NT, smem, stream>>>(cp);
  profile_end(stream);
  count_launch();
  return GPP_OK;
}

template <int D>
static int predict_fwd(const gpp_gp_model* m, const double* mu, const double* S, int N, double* f1, double* Sff,
                       double* cross, int full_output_cov, double jitter, char* ws, const Plan& pl, int* info,
                       cudaStream_t stream, const EulerPost* post) {
  using PP = PairPack<D>;
  const gpp_gp_model::SlotTable& tab = m->tables[pl.tile_idx][pl.diag_only];
  double* packs = (double*)(ws + pl.off_packs);
  double* part = (double*)(ws + pl.off_part);
  double* f1lat = (double*)(ws + pl.off_f1lat);
  double* crosslat = (double*)(ws + pl.off_crosslat);
  unsigned* counter = (unsigned*)(ws + pl.off_counter);
  PackPsi1Params pp;
  pp.m = mu; pp.S = S; pp.Z = m->Z; pp.ell = m->ell; pp.var = m->var; pp.beta = m->beta; pp.pair_ab = tab.d_pair_ab;
  pp.packs = packs; pp.f1lat = f1lat; pp.crosslat = crosslat; pp.Gs = nullptr; pp.counter = counter; pp.info = info;
  pp.N = N; pp.L = m->L; pp.M = m->M; pp.npairs = tab.npairs;
  launch_pack_psi1<D>(pp, stream);
  count_launch();
  ContractParams cp;
  cp.Z = m->Z; cp.beta = m->beta; cp.C = m->C; cp.packs = packs; cp.part = part; cp.slots = tab.d_slots;
  cp.counter = counter; cp.N = N; cp.M = m->M; cp.L = m->L; cp.npairs = tab.npairs; cp.nslots = tab.nslots;
  cp.nchunks = pl.nchunks; cp.chunk = pl.chunk;
  int rc = pl.tile_idx ? launch_contract<D, 128, 4, 16>(cp, stream) : launch_contract<D, 64, 2, 8>(cp, stream);
  if (rc != GPP_OK) return rc;
  FinalizeParams fp;
  fp.part = part; fp.slots = tab.d_slots; fp.pair_start = tab.d_pair_start; fp.pair_ab = tab.d_pair_ab;
  fp.f1lat = f1lat; fp.crosslat = crosslat; fp.var = m->var; fp.mean = m->mean; fp.W = m->W;
  fp.f1 = f1; fp.Sff = Sff; fp.cross = cross; fp.N = N; fp.L = m->L; fp.P = m->P; fp.D = D;
  fp.npairs = tab.npairs; fp.nslots = tab.nslots; fp.full_cov = full_output_cov;
  fp.model_uncertainty = m->model_uncertainty; fp.jitter = jitter;
  if (post) {
    fp.post = *post;
    k_finalize<true><<<N, 128, 0, stream>>>(fp);
  } else {
    k_finalize<false><<<N, 128, 0, stream>>>(fp);
  }
  count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  (void)sizeof(PP);
  return GPP_OK;
}

}  // namespace gpp

extern "C" {

size_t gpp_mm_gp_predict_workspace_bytes(const gpp_gp_model* model, int N) {
  if (!model || N <= 0) return 0;
  // worst case over the two covariance modes so one workspace serves both
  return std::max(gpp::make_plan(model, N, 1).total, gpp::make_plan(model, N, 0).total);
}

int gpp_mm_gp_predict_fwd(const gpp_gp_model* model, const double* m, const double* S, int N, double* f1, double* Sff,
                          double* cross, int full_output_cov, double jitter, void* workspace, size_t workspace_bytes,
                          int* info, void* stream_) {
  return gpp::mm_predict_enqueue(model, m, S, N, f1, Sff, cross, full_output_cov, jitter, workspace, workspace_bytes, info,
                                 (cudaStream_t)stream_, nullptr);
}

}  // extern "C"

int gpp::mm_predict_enqueue(const gpp_gp_model* model, const double* m, const double* S, int N, double* f1, double* Sff,
                            double* cross, int full_output_cov, double jitter, void* workspace, size_t workspace_bytes,
                            int* info, cudaStream_t stream, const gpp::EulerPost* post) {
  GPP_REQUIRE(model && m && S && f1 && Sff && cross && workspace, GPP_ERR_NULL, "gpp_mm_gp_predict_fwd: null argument");
  GPP_REQUIRE(N >= 1, GPP_ERR_BAD_SHAPE, "gpp_mm_gp_predict_fwd: N=%d", N);
  GPP_REQUIRE(!post || post->Dx == model->P, GPP_ERR_BAD_SHAPE, "rollout: the dynamics model has %d outputs for a %d-dimensional state",
              model->P, post ? post->Dx : 0);
  gpp::Plan pl = gpp::make_plan(model, N, full_output_cov);
  GPP_REQUIRE(workspace_bytes >= pl.total, GPP_ERR_WORKSPACE, "gpp_mm_gp_predict_fwd: workspace %zu < required %zu",
              workspace_bytes, pl.total);
  char* ws = (char*)workspace;
  switch (model->D) {
#define GPP_CASE(d) \
  case d: return gpp::predict_fwd<d>(model, m, S, N, f1, Sff, cross, full_output_cov, jitter, ws, pl, info, stream, post);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default:
      gpp::set_error("gpp_mm_gp_predict_fwd: unsupported D=%d", model->D);
      return GPP_ERR_UNSUPPORTED;
  }
}
