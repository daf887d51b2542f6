// Internal definition of the opaque GP model handle.
#pragma once
#include <vector>

#include "common.cuh"

struct gpp_slot {        // one (kernel pair, tile) work unit of the Psi2 contraction
  int pair;              // index into the pair list
  int a, b;              // latent indices (a <= b)
  int ti, tj;            // tile coordinates
  double weight;         // 2 for strictly-upper tiles of a diagonal pair (symmetry), else 1
};

struct gpp_gp_model {
  int L = 0, M = 0, D = 0, P = 0;
  int whiten = 0, model_uncertainty = 1, coreg = 0;
  // device arrays owned by the handle
  double* Z = nullptr;        // [L,M,D]
  double* ell = nullptr;      // [L,D]
  double* var = nullptr;      // [L]
  double* beta = nullptr;     // [L,M]
  double* C = nullptr;        // [L,M,M]   beta beta^T - B
  double* mean = nullptr;     // [P] (zeros if none)
  double* W = nullptr;        // [P,L] or null
  double* Luu = nullptr;      // [L,M,M]   row-major lower Cholesky of Kuu (kept for the pathwise update)
  double* q_mu = nullptr;     // [M,L]     copies of the variational parameters (pathwise u-samples)
  double* q_sqrt = nullptr;   // [L,M,M]   lower triangle only (zeros if none)
  std::vector<double> h_jitter;   // [L] diagonal added to Kuu
  void* blas = nullptr;       // lazily created cublasHandle_t used by gpp_pathwise_generate
  // host copies of small parameters
  std::vector<double> h_ell, h_var;
  // slot tables (device) for the two tile sizes x {all pairs, diagonal pairs only}
  struct SlotTable {
    int tile = 0, npairs = 0, nslots = 0;
    gpp_slot* d_slots = nullptr;
    int* d_pair_start = nullptr;   // [npairs+1] slot range per pair
    int* d_pair_ab = nullptr;      // [npairs,2]
    std::vector<gpp_slot> h_slots;
  };
  SlotTable tables[2][2];     // [tile 64|128][full|diag-only]
};

namespace gpp {

// Optional rollout epilogue of the forward predict's k_finalize (block n, after f1 / Sff / cross of input n are written):
// the Euler step of the moment-matched state, x' = x + f with dt = 1 (upstream dynamics/solvers.py MomentMatchingEuler,
// forward_sde.py:112-124):  m' = m + f1,  S' = S + Sxf + Sxf^T + Sff,  Sxf = Sxd cross.  Requires P == Dx.
struct EulerPost {
  double *m, *S;              // [N,Dx], [N,Dx,Dx]  current state, updated in place
  const double* Sxd;          // [N,Dx,D]           Cov(x, d) of the pre stage
  double *traj_m, *traj_S;    // optional: slice of step t+1, [N,Dx], [N,Dx,Dx]
  double *ring_m, *ring_S;    // optional: cost-ring slot of this step, same shapes
  int Dx;
};

// the forward predict as the rollouts call it (mm_predict.cu); gpp_mm_gp_predict_fwd is the same without an epilogue
int mm_predict_enqueue(const gpp_gp_model* model, const double* m, const double* S, int N, double* f1, double* Sff,
                       double* cross, int full_output_cov, double jitter, void* workspace, size_t workspace_bytes,
                       int* info, cudaStream_t stream, const EulerPost* post = nullptr);

}  // namespace gpp
