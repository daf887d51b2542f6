// Host-side interface of the persistent rollout kernels (rollout_persist.cu) used by rollout_mm.cu / rollout_mm_bwd.cu.
#pragma once
#include "model.cuh"

namespace gpp {

struct RolloutMMParams;
struct RolloutBwdBuffers;

int rollout_mode();   // GPP_ROLLOUT_AUTO / _LEGACY / _PERSIST (gpp_rollout_mm_set_mode, env GPP_ROLLOUT_MODE)

struct PersistFwdLayout {
  size_t packs, part, f1lat, crosslat, flags, total;
};
bool persist_fwd_supported(const gpp_gp_model* dyn, int N, int Dx);
PersistFwdLayout persist_fwd_layout(const gpp_gp_model* dyn, int N);
// r: the rollout's parameter block with its workspace buffers set (m, S, md, Sd, Sxd, f1, Sff, cross, loss, traj_*)
int rollout_mm_fwd_persist(const gpp_gp_model* dyn, const RolloutMMParams& r, int H, const double* m0, const double* S0, double* m_final,
                           double* S_final, double* saved, char* ws_persist, cudaStream_t stream);

struct PersistBwdLayout {
  size_t packs, Gs, stats, f1lat, crosslat, f1lat_bar, crosslat_bar, omega, gm, gS, cg, flags, total;
  int nrb;
};
bool persist_bwd_supported(const gpp_gp_model* dyn, int N, int Dx);
PersistBwdLayout persist_bwd_layout(const gpp_gp_model* dyn, int N, int Dx, int H);
// r: parameter block (policy, encoder, cost); bw: adjoint buffers (zeroed by the caller); saved / traj_*: what gpp_rollout_mm_fwd_save kept
int rollout_mm_bwd_persist(const gpp_gp_model* dyn, const RolloutMMParams& r, const RolloutBwdBuffers& bw, int H, const double* traj_m,
                           const double* traj_S, const double* saved, const double* loss_bar, char* ws_persist, cudaStream_t stream);

}  // namespace gpp
