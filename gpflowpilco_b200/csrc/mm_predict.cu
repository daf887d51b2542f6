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
#include "finalize.cuh"

namespace gpp {

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
  static PerDeviceSmemOptIn configured;
  if (configured.raise(smem)) {
    GPP_CUDA_OK(cudaFuncSetAttribute(k_contract<D, T, NP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  fp.part = part; fp.part_ll = nullptr; fp.ll_tag = 0; fp.slots = tab.d_slots; fp.pair_start = tab.d_pair_start; fp.pair_ab = tab.d_pair_ab;
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
  GPP_NVTX_RANGE();
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
