#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native GPflowPILCO hot path.

Workload (BASELINE.json configs[1]): batched exact moment-matched GP prediction — N Gaussian input states per GPU
pushed through E=4 independent exact SE-ARD GPs on 1000 training points (D=6), full 4x4 output covariance and
input-output cross-covariance, FP64.  One "step" = one pass of the hot path over the N inputs of every rank;
metric = Gaussian states moment-matched through the GP per second, whole job.  The rollout proper (config #1 / #5: encoder ->
policy -> GP dynamics -> Euler -> cost over a horizon, forward + backward) is reported under "policy_opt_step", the pathwise half of
BASELINE's metric under "pathwise"; at N > 1 rank 0 also checks the sharded closures against single-process results ("multirank_check").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--inputs N_per_gpu]

Prints ONE JSON line (rank 0).  `--impl reference` times the line-by-line CPU restatement of the upstream
algorithm (oracle/, triangular-solve form of gpflow_pilco/moment_matching/models.py:200-299) on the host cores,
on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mm_gp_predict_states_per_s"
UNIT = "gaussian_states/s"
M_TRAIN, D_IN, E_OUT = 1000, 6, 4
FLOP_PER_ENTRY = 2 * D_IN + 26          # D FMAs + 2 adds + exp(=20) + 2 contraction FMAs
ENTRIES_PER_INPUT = E_OUT * (E_OUT + 1) // 2 * M_TRAIN * M_TRAIN


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--inputs", type=int, default=8192, help="Gaussian inputs per GPU (weak scaling)")
  ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-pathwise", action="store_true")
  ap.add_argument("--pathwise-particles", type=int, default=148 * 512, help="particles per GPU per launch (one wave of 512-particle CTAs)")
  ap.add_argument("--pathwise-horizon", type=int, default=100)
  ap.add_argument("--pathwise-bases", type=int, default=4096)
  ap.add_argument("--no-pathwise-full", dest="pathwise_full", action="store_false",
                  help="skip the run of all 2^20 particles of config #4 (about 4 s on one GPU, split over the ranks)")
  ap.add_argument("--no-policy-opt", action="store_true")
  ap.add_argument("--no-psi2", action="store_true")
  ap.add_argument("--restarts-total", type=int, default=512,
                  help="policy restarts of config #5, FIXED TOTAL sharded over the ranks (strong scaling: 512 on 1 GPU, 64 per GPU on 8)")
  ap.add_argument("--restart-horizon", type=int, default=100)
  return ap.parse_args()


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), "measured (MEASURED_PEAKS.json)"
  return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
  FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index: int):
    self.gpu_index = gpu_index
    self.lines = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                    "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._pump, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in self.lines:
      parts = [p.strip() for p in ln.split(",")]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1]))
        mx.append(float(parts[2]))
      except ValueError:
        continue
      for name, val in zip(names, parts[5:9]):
        if val.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------
def build_workload(n_inputs: int, rank: int):
  from gpflowpilco_b200 import synthetic
  cfg = synthetic.config2_batched_mm_predict(N=n_inputs, M=M_TRAIN, D=D_IN, E=E_OUT, seed=0)
  if rank:   # every rank shares the model (seed 0) but owns different inputs
    rng = np.random.default_rng(1000 + rank)
    cfg["mu"] = rng.random((n_inputs, D_IN))
    cfg["cov"] = synthetic.generate_covariance(rng, D_IN, n_inputs, 0.1)
  return cfg


def oracle_model(cfg):
  import torch
  from oracle import gp_models as gm
  from oracle import psi_stats as ps
  ks = [ps.SEKernel(float(cfg["variance"][e]), torch.as_tensor(cfg["lengthscales"][e])) for e in range(E_OUT)]
  X = torch.as_tensor(cfg["X"])
  # exact GP as SVGP(q_mu = Y - c, q_sqrt = 0, whiten=False, Kuu jitter = noise variance)  (upstream models.py:44-111)
  return gm.SVGPModel(ks, [X] * E_OUT, torch.as_tensor(cfg["Y"] - cfg["mean_const"]),
                      torch.zeros(E_OUT, M_TRAIN, M_TRAIN, dtype=torch.float64), whiten=False,
                      mean_const=torch.as_tensor(cfg["mean_const"]))


def oracle_predict(model, cfg, idx, reference_form: bool):
  import torch
  import oracle.gp_models as gm
  from oracle.moments import GaussianMoments
  x = GaussianMoments(torch.as_tensor(cfg["mu"][idx]), torch.as_tensor(cfg["cov"][idx]), True)
  old = gm.Kuu.__defaults__
  gm.Kuu.__defaults__ = (float(cfg["noise_variance"][0]),)
  try:
    fn = gm.mm_svgp_mo if reference_form else gm.mm_sparse_reassociated
    return fn(x, model)
  finally:
    gm.Kuu.__defaults__ = old


def time_cpu(cfg, seconds: float, reference_form: bool, steps: int = 1):
  """Times the oracle on a bounded sample; returns (states_per_s, sample_description, per_step_ms)."""
  import torch
  model = oracle_model(cfg)
  t0 = time.perf_counter()
  oracle_predict(model, cfg, slice(0, 1), reference_form)
  t1 = time.perf_counter() - t0
  # memory bound: the upstream form materialises eKuffu [n,4,1000,4,1000] (128 MB per input) plus ~6 temporaries of it
  cap = 6 if reference_form else 48
  n = int(max(1, min(cfg["mu"].shape[0], cap, seconds / max(t1, 1e-3) / max(steps, 1))))
  times = []
  for _ in range(steps):
    t0 = time.perf_counter()
    oracle_predict(model, cfg, slice(0, n), reference_form)
    times.append(time.perf_counter() - t0)
  best = min(times)
  form = "upstream triangular-solve form" if reference_form else "O(M^2) re-associated form"
  return n / best, f"{n} of the {cfg['mu'].shape[0]} inputs of config #2 per step, {form}, torch float64, {torch.get_num_threads()} threads", 1e3 * float(np.mean(times))


def time_cpu_rollout(cfg, H):
  """CPU restatement of config #1 (oracle, upstream triangular-solve form, forward only): seconds per rollout, best of 2."""
  import torch
  from oracle import gp_models as gm
  from oracle import moments as mo
  from oracle import psi_stats as ps
  from oracle import rollout as ro

  def model(pp):
    L = pp["Z"].shape[0]
    ks = [ps.SEKernel(float(pp["variance"][l]), torch.as_tensor(pp["lengthscales"][l])) for l in range(L)]
    return gm.SVGPModel(ks, [torch.as_tensor(pp["Z"][l]) for l in range(L)], torch.as_tensor(pp["q_mu"]), torch.as_tensor(pp["q_sqrt"]),
                        whiten=bool(pp["whiten"]), mean_const=torch.as_tensor(pp["mean_const"]))
  dyn, pol = model(cfg["dynamics"]), model(cfg["policy"])
  enc = mo.TrigonometricEncoder(cfg["active_dims"])
  obj = mo.GaussianObjective(cfg["target"], cfg["W"])
  best, loss = 1e9, None
  for _ in range(2):
    t0 = time.perf_counter()
    loss = ro.mm_rollout(torch.as_tensor(cfg["m0"]), torch.as_tensor(cfg["S0"]), H, lambda s: gm.mm_svgp(s, dyn),
                         lambda s: gm.mm_policy(s, pol, cfg["squash_scale"], cfg["squash_shift"]), enc, obj)
    best = min(best, time.perf_counter() - t0)
  return {"forward_ms": 1e3 * best, "rollout_steps_per_s_forward": H / best, "loss": float(loss[0]), "kind": "port",
          "cores": torch.get_num_threads(), "sample": "the whole config #1 rollout, oracle in upstream's triangular-solve form"}


ARGS = None
_emit = print


def pathwise_section(dev, lib, pk, world):
  """Second half of BASELINE's metric: pathwise trajectory-steps/s (config #4 shapes: L=4, M=256, D=6, F=4096, H=100).
  Particles are sharded over ranks by global index (no data-path collective); value is the whole-job aggregate."""
  import ctypes
  import torch
  import torch.distributed as dist
  from gpflowpilco_b200 import ops, synthetic
  from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths, rollout_pathwise, rollout_pathwise_chunked
  from gpflowpilco_b200.rollouts import PolicyParams
  args = ARGS
  rank = int(os.environ.get("RANK", "0"))
  S, H, F = args.pathwise_particles, args.pathwise_horizon, args.pathwise_bases
  cfg = synthetic.config1_cartpole()
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True,
                             mean_const=T(d["mean_const"]))
  policy = PolicyParams(T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), whiten=True,
                        squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  t0 = time.perf_counter()
  paths = generate_paths(handle, S, F, seed=0, first_particle=rank * S)
  x0 = draw_initial_states(T(cfg["m0"][0]), T(cfg["S0"][0]), 0, rank * S, S)
  torch.cuda.synchronize()
  gen_s = time.perf_counter() - t0
  beta = policy.beta()
  target, W = T(cfg["target"]), T(cfg["W"])
  lib.gpp_profile_enable(1)
  times = []
  sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
  sampler.start()
  for it in range(4):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    loss, _, _ = rollout_pathwise(paths, policy, x0, H, cfg["active_dims"], target, W, beta=beta)
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    if it:
      times.append(ms.value)
  pw_clocks = sampler.stop()
  # mixed-precision variant (FP32 Fourier weights + FP32 cosine polynomial, gpp_rollout_pathwise_fwd_mixed): reported beside the
  # FP64 figure, never instead of it
  mx_times = []
  paths.weights_f32()
  sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
  sampler.start()
  for it in range(3):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    loss_mx, _, _ = rollout_pathwise(paths, policy, x0, H, cfg["active_dims"], target, W, beta=beta, mixed_precision=True)
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    if it:
      mx_times.append(ms.value)
  mx_clocks = sampler.stop()
  mx_diff = float((loss_mx.mean() - loss.mean()).abs() / loss.mean().abs())
  lib.gpp_profile_enable(0)
  # gradient mode: forward with per-step Jacobians + reverse sweep (policy gradient), events around the pair
  from gpflowpilco_b200.autograd import rollout_pathwise_loss
  Zg, eg, qg = [x.clone().requires_grad_(True) for x in (policy.Z, policy.lengthscales, policy.q_mu)]
  gtimes = []
  for it in range(3):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gl = rollout_pathwise_loss(paths, Zg, eg, policy.variance, qg, x0, H, cfg["active_dims"], target, W, squash_scale=cfg["squash_scale"],
                               squash_shift=cfg["squash_shift"])
    gl.sum().backward()
    e1.record()
    e1.synchronize()
    if it:
      gtimes.append(e0.elapsed_time(e1))
    Zg.grad = eg.grad = qg.grad = None
  # the whole of config #4: 2^20 particles over all ranks, in chunks of one wave each; path generation (Philox draws + the
  # update-weight solve, on the device) and the rollout of every chunk are inside the timed region
  full = None
  if args.pathwise_full:
    total_particles = 2 ** 20
    per_rank = total_particles // world
    first = rank * per_rank
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # chunks of one wave of CTAs; the generation of chunk k+1 overlaps the rollout of chunk k on a second stream (pathwise.py)
    acc_loss = rollout_pathwise_chunked(handle, policy, T(cfg["m0"][0]), T(cfg["S0"][0]), per_rank, F, 0, H, cfg["active_dims"], target, W,
                                        first_particle=first, particles_per_launch=S, beta=beta)
    if world > 1:
      dist.all_reduce(acc_loss)                       # the one collective of the pathwise closure: the summed cost
    torch.cuda.synchronize()
    tf_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
    full = {"particles": total_particles, "horizon": H, "seconds": float(tf_[0]), "particle_steps_per_s": total_particles * H / float(tf_[0]),
            "mean_loss": float(acc_loss) / total_particles, "includes": "device-side path generation of every chunk + rollouts + cost all-reduce; the generation of chunk k+1 is issued on a second stream but does not overlap measurably (the rollout kernel holds every SM's register file)"}
    if world > 1:
      # 1 -> N curve of this collective-bearing path: rank 0 repeats ITS share alone (the others wait at the barrier); one GPU needs
      # `world` such shares for the whole job, so efficiency = T(share, alone) / T(sharded job, all ranks + all-reduce)
      dist.barrier()
      alone = None
      if rank == 0:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rollout_pathwise_chunked(handle, policy, T(cfg["m0"][0]), T(cfg["S0"][0]), per_rank, F, 0, H, cfg["active_dims"], target, W,
                                 first_particle=0, particles_per_launch=S, beta=beta)
        torch.cuda.synchronize()
        alone = time.perf_counter() - t0
      dist.barrier()
      if rank == 0:
        full["share_alone_seconds"] = alone
        full["efficiency_vs_n1"] = alone / float(tf_[0])
        full["efficiency_note"] = "T(one rank's share of the particles, run alone on rank 0) / T(sharded job incl. all-reduce); 1-GPU job time = world x share"
    else:
      full["efficiency_vs_n1"] = 1.0
  t = torch.tensor([float(np.mean(times)), float(np.mean(gtimes)), float(np.mean(mx_times))], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  sec, gsec, msec = float(t[0]) * 1e-3, float(t[1]) * 1e-3, float(t[2]) * 1e-3
  L, M, D = 4, d["Z"].shape[1], 6
  bytes_per_pstep = 8 * L * (F + M) + 2 * 8 * 4
  bytes_per_pstep_mx = 4 * L * F + 8 * L * M + 2 * 8 * 4
  flop_per_pstep = L * F * (2 * D + 2 + 20) + L * M * (2 * D + 2 + 20) + 30 * (2 * 5 + 22)
  psteps = S * H
  return {
      "metric": "pathwise_particle_steps_per_s", "value": world * psteps / sec, "unit": "particle_steps/s",
      "config": {"workload": "config#4 pathwise cart-pole rollouts", "particles_per_gpu_per_launch": S, "bases": F, "horizon": H,
                 "latents": L, "inducing": M, "weights": "streamed from HBM (generated on device beforehand, Philox by global particle index)",
                 "note": "1M particles = ceil(2^20 / particles_per_launch) identical launches per GPU",
                 "parity": "UNPINNED: gpflow_sampling (upstream's path sampler) is not in the reference tree; the kernels are checked "
                           "against oracle/pathwise.py, a restatement of the published algorithm (DESIGN section 2)"},
      "ms_per_launch": 1e3 * sec, "generation_s": gen_s, "mean_loss": float(loss.mean()), "clocks": pw_clocks,
      "full_config4": full,
      "with_policy_gradient": {"value": world * psteps / gsec, "unit": "particle_steps/s (gradient-mode forward + reverse sweep)",
                               "ms": 1e3 * gsec, "hbm_frac": bytes_per_pstep * psteps / gsec / 1e9 / pk.get("hbm_gbs")},
      "mixed_precision": {
          "value": world * psteps / msec, "unit": "particle_steps/s", "ms_per_launch": 1e3 * msec, "speedup_vs_fp64": sec / msec,
          "what": "gpp_rollout_pathwise_fwd_mixed: FP32 Fourier weights and cosine polynomial; phases (DMMA), quarter-turn reduction, "
                  "canonical-basis part, policy, cost and state update FP64; FP32 partial sums folded into FP64 every 32 feature tiles",
          "tolerance": "one-step drift within 5e-6 of the FP64 kernel (tests/test_gpu_pathwise.py); NOT the headline dtype",
          "mean_loss_rel_diff_vs_fp64": mx_diff, "clocks": mx_clocks,
          "roofline": {"bound": "hbm", "achieved": bytes_per_pstep_mx * psteps / msec / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                       "frac": bytes_per_pstep_mx * psteps / msec / 1e9 / pk.get("hbm_gbs"),
                       "algorithmic": f"{bytes_per_pstep_mx} B per particle-step (FP32 Fourier weights)"}},
      "roofline": {"bound": "hbm", "achieved": bytes_per_pstep * psteps / sec / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                   "frac": bytes_per_pstep * psteps / sec / 1e9 / pk.get("hbm_gbs"), "traffic": 1.0001 * bytes_per_pstep * psteps,
                   "traffic_source": "ncu --set full at H=4 (profiles/r2f_pathwise_dmma_full.txt): dram read 42.236 GB vs 42.232 GB algorithmic; scaled to this launch",
                   "kernel": "k_pathwise_rollout",
                   "algorithmic": f"{bytes_per_pstep} B and {flop_per_pstep} flop per particle-step",
                   "fp64_achieved_tflops": flop_per_pstep * psteps / sec / 1e12},
  }


def psi2_section(dev, lib, pk):
  """BASELINE config #3 (kernel-expectation stress): materialised eKzxKxz, N = 1024 Gaussian inputs, M = 2048 inducing points, D = 8 —
  34.4 GB of output.  The C ABI takes at most the inputs the caller has output memory for: the whole config is run as 4 calls of 256
  inputs into one caller-owned [1024, 2048, 2048] tensor, back to back on the stream, timed by one event pair (whole job); the
  kernel-only figure comes from the library's profile events around one launch.  Sub-case (ii) two kernels / two inducing sets (the
  branch upstream's test exercises) is the headline, sub-case (i) same kernel / same inducing set is reported beside it."""
  import torch
  from gpflowpilco_b200 import ops, synthetic
  N_ALL, CH = 1024, 256
  c3 = synthetic.config3_psi2_stress(N=N_ALL)
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  mu3, cov3 = T(c3["mu"]), T(c3["cov"])
  two = (T(c3["Z1"]), T(c3["lengthscales1"]), c3["variance1"], T(c3["Z2"]), T(c3["lengthscales2"]), c3["variance2"])
  same = (T(c3["Z1"]), T(c3["lengthscales1"]), c3["variance1"])
  M = c3["Z1"].shape[0]
  out = torch.empty(N_ALL, M, M, dtype=torch.float64, device=dev)    # caller-owned, as the C ABI has it (34.4 GB)
  nbytes = out.numel() * 8

  def whole(model_args):
    for c in range(0, N_ALL, CH):
      ops.ekzxkxz(mu3[c:c + CH], cov3[c:c + CH], *model_args, check=False, out=out[c:c + CH])

  def timed(model_args):
    ts = []
    for it in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize()
      e0.record()
      whole(model_args)
      e1.record()
      torch.cuda.synchronize()
      if it:
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)) * 1e-3
  t_two = timed(two)
  t_same = timed(same)
  lib.gpp_profile_enable(1)
  kms = []
  for it in range(3):
    ops.ekzxkxz(mu3[:CH], cov3[:CH], *two, check=False, out=out[:CH])
    torch.cuda.synchronize()
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    if it:
      kms.append(ms.value)
  lib.gpp_profile_enable(0)
  del out
  ks = float(np.mean(kms)) * 1e-3
  kbytes = nbytes * CH / N_ALL
  return {"metric": "psi2_entries_per_s", "value": nbytes / 8 / t_two, "unit": "entries/s (whole config #3: 4 calls, coefficient packs + column vectors + main kernel)",
          "config": {"workload": "config#3 Psi2 stress, sub-case (ii) two kernels / two inducing sets", "inputs": N_ALL, "inputs_per_call": CH,
                     "inducing": M, "dims": 8, "output_gb": nbytes / 1e9},
          "whole_config_ms": 1e3 * t_two, "whole_config_gbs": nbytes / t_two / 1e9, "whole_config_frac": nbytes / t_two / 1e9 / pk.get("hbm_gbs"),
          "same_kernel_same_features": {"whole_config_ms": 1e3 * t_same, "whole_config_gbs": nbytes / t_same / 1e9,
                                        "frac": nbytes / t_same / 1e9 / pk.get("hbm_gbs"),
                                        "note": "sub-case (i): symmetric in (i, j); every entry is evaluated and written (no triangle reuse)"},
          "roofline": {"bound": "hbm", "achieved": kbytes / ks / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s", "frac": kbytes / ks / 1e9 / pk.get("hbm_gbs"),
                       "traffic": 8.54e9, "traffic_source": "ncu --set full (profiles/r1b_psi2_final_full.txt): dram write 8.54 GB per launch = algorithmic 8.59 GB",
                       "kernel": "k_ekzxkxz", "kernel_ms": 1e3 * ks, "algorithmic": "8 B written per entry"}}


def policy_opt_section(dev, lib, world, fp64_peak):
  """BASELINE config #5 (full PILCO policy-optimisation step): R policy restarts per GPU, cart-pole models of config #1
  (M=256 dynamics, 30 policy centres), horizon H, forward + backward of the moment-matched rollout; restarts are sharded over
  ranks, every rank receives loss[R_total] by all-gather.  Reports rollout-steps/s (forward + backward), whole job."""
  import torch
  import torch.distributed as dist
  from gpflowpilco_b200 import distributed as gd
  from gpflowpilco_b200 import ops, synthetic
  args = ARGS
  Rt, H = args.restarts_total, args.restart_horizon
  if Rt % world:
    raise ValueError("--restarts-total must be a multiple of the number of GPUs")
  R = Rt // world
  cfg = synthetic.config1_cartpole()
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True,
                             mean_const=T(d["mean_const"]))
  g = torch.Generator().manual_seed(5)
  Z = T(p["Z"]).repeat(Rt, 1, 1) + 0.3 * torch.randn(Rt, *p["Z"].shape[1:], dtype=torch.float64, generator=g).to(dev)
  ell = T(p["lengthscales"]).repeat(Rt, 1) * torch.exp(torch.empty(Rt, 1, dtype=torch.float64).uniform_(-0.7, 0.7, generator=g)).to(dev)
  q = 1e-3 * torch.randn(Rt, p["Z"].shape[1], dtype=torch.float64, generator=g).to(dev)
  var = T(p["variance"]).repeat(Rt)
  def step():
    return gd.mm_restart_losses_and_grads(handle, Z, ell, var, q, T(cfg["m0"]), T(cfg["S0"]), H, cfg["active_dims"], T(cfg["target"]),
                                          T(cfg["W"]), cfg["squash_scale"], cfg["squash_shift"])
  step()
  times = []
  launches0 = lib.gpp_launch_count()
  for _ in range(2):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses, _, grads = step()
    e1.record()
    e1.synchronize()
    times.append(e0.elapsed_time(e1))
  launches = (lib.gpp_launch_count() - launches0) // 2
  t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  sec = float(t[0]) * 1e-3
  # BASELINE config #1 (the reference's own CPU-runnable case): one cart-pole rollout, N = 1, H = 30, forward and forward+backward
  from gpflowpilco_b200.autograd import rollout_mm_loss
  from gpflowpilco_b200 import rollouts as gp_rollouts
  H1 = cfg["horizon"]
  Z1 = T(p["Z"]).clone().requires_grad_(True)
  q1 = T(p["q_mu"][:, 0][None]).clone().requires_grad_(True)
  e1_ = T(p["lengthscales"]).clone().requires_grad_(True)
  c1 = {"fwd_ms": [], "fwd_bwd_ms": []}
  launches1 = 0
  for it in range(4):
    l0_ = lib.gpp_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    l1 = rollout_mm_loss(handle, Z1, e1_, T(p["variance"]), q1, T(cfg["m0"]), T(cfg["S0"]), H1, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]),
                         squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"], check="defer")
    ev[1].record()
    l1.sum().backward()
    ev[2].record()
    gp_rollouts.raise_deferred()          # the three not-positive-definite flags, one synchronisation (rollouts.py)
    torch.cuda.synchronize()
    launches1 = lib.gpp_launch_count() - l0_
    if it:
      c1["fwd_ms"].append(ev[0].elapsed_time(ev[1]))
      c1["fwd_bwd_ms"].append(ev[0].elapsed_time(ev[2]))
    Z1.grad = q1.grad = e1_.grad = None
  config1 = {"workload": "config#1 cart-pole MM rollout, N=1, H=%d, M=256 dynamics, 30 policy centres" % H1,
             "forward_ms": float(np.mean(c1["fwd_ms"])), "forward_backward_ms": float(np.mean(c1["fwd_bwd_ms"])),
             "rollout_steps_per_s_forward": H1 / float(np.mean(c1["fwd_ms"])) * 1e3, "loss": float(l1.detach()[0])}
  if world == 1 and not args.no_cpu_baseline:
    config1["cpu"] = time_cpu_rollout(cfg, H1)
  # the same evaluations replayed from a captured CUDA graph (gpflowpilco_b200/graphs.py): forward + backward + policy adjoint
  from gpflowpilco_b200.graphs import GraphedMMPolicyGradient
  def graph_ms(Zg, eg, vg, qg, m0g, S0g, Hg):
    gr = GraphedMMPolicyGradient(handle, Zg, eg, vg, qg, m0g, S0g, Hg, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]),
                                 squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
    ts = []
    for _ in range(4):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      gl, _ = gr(Zg, eg, qg, check=False)
      e1.record()
      e1.synchronize()
      ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[1:])), gl.clone()
  g1_ms, g1_loss = graph_ms(Z1.detach(), e1_.detach(), T(p["variance"]), q1.detach(), T(cfg["m0"]), T(cfg["S0"]), H1)
  config1["graph_forward_backward_ms"] = g1_ms
  config1["graph_loss_equal"] = bool(torch.equal(g1_loss, l1.detach()))
  lo, Rl = gd.shard_range(Rt, int(os.environ.get("RANK", "0")), world)
  hi = lo + Rl
  g5_ms, _ = graph_ms(Z[lo:hi].contiguous(), ell[lo:hi].contiguous(), var[lo:hi].contiguous(), q[lo:hi].contiguous(),
                      T(cfg["m0"]).expand(Rl, -1).contiguous(), T(cfg["S0"]).expand(Rl, -1, -1).contiguous(), H)
  M, L, D = d["Z"].shape[1], 4, 6
  flop_per_step = 4 * (L * (L + 1) // 2) * M * M * (2 * D + 26)      # fwd + bwd counted as 4 x forward
  ach5 = Rt * H * flop_per_step / sec / 1e12 / world
  ach1 = H1 * flop_per_step / (config1["forward_backward_ms"] * 1e-3) / 1e12
  config1["roofline"] = {"bound": "fp64 (latency-bound at one rollout: a chain of serial D x D stages per step)", "achieved": ach1, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": ach1 / fp64_peak}
  config1["gpu_launches_forward_backward"] = int(launches1)
  return {"metric": "mm_policy_opt_rollout_steps_per_s", "value": Rt * H / sec, "unit": "rollout_steps/s (forward+backward)",
          "config": {"workload": "config#5 policy-optimisation step", "restarts_total": Rt, "restarts_per_gpu": R, "scaling": "strong (fixed total)",
                     "horizon": H, "dynamics_inducing": M, "policy_centres": int(p["Z"].shape[1]),
                     "collective": "all-gather of loss[R] (uneven-safe) inside the timed region, gradients stay sharded",
                     "h_loop": "on the device: one persistent cooperative kernel per sweep direction (csrc/rollout_persist.cu)"},
          "ms_per_opt_step": 1e3 * sec, "graph_ms_per_opt_step_local_share": g5_ms,
          "gpu_launches_per_opt_step": int(launches), "mean_loss": float(losses.mean()),
          "config1_rollout": config1,
          "grad_norm": float(grads[0].norm()),
          "roofline": {"bound": "fp64", "achieved": ach5, "peak": fp64_peak, "unit": "TFLOP/s per GPU", "frac": ach5 / fp64_peak,
                       "algorithmic": f"{flop_per_step} flop per rollout-step (4 x forward, SURVEY 8d)",
                       "kernels": "k_rollout_fwd_persist + k_rollout_bwd_persist (all of a sweep in one launch each)"}}


def multirank_check(dev, world):
  """N > 1 only: the sharded closures (particles by global index + cost/gradient all-reduce; policy restarts + loss all-gather)
  against the same closures evaluated by rank 0 alone, on small cases.  What tests/test_gpu_multirank.py asserts on a 2-GPU box,
  executed here on whatever the driver launched; rank 0 reports the worst relative differences."""
  import torch
  import torch.distributed as dist
  from gpflowpilco_b200 import distributed as gd
  from gpflowpilco_b200 import ops, synthetic
  rank = dist.get_rank()
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  cfg["policy"]["q_mu"] = 100.0 * cfg["policy"]["q_mu"]
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
  pw_args = (handle, T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), T(cfg["m0"][0]), T(cfg["S0"][0]))
  pw_kw = dict(total_particles=1000, num_bases=64, seed=4, horizon=3, active_dims=cfg["active_dims"], cost_target=T(cfg["target"]),
               cost_W=T(cfg["W"]), squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  R = 2 * world + 1                                           # uneven shards on purpose
  g = torch.Generator().manual_seed(0)
  Z = T(p["Z"]).repeat(R, 1, 1) + 0.1 * torch.randn(R, *p["Z"].shape[1:], dtype=torch.float64, generator=g).to(dev)
  ell, q, var = T(p["lengthscales"]).repeat(R, 1), T(p["q_mu"][:, 0][None]).repeat(R, 1), T(p["variance"]).repeat(R)
  mm_args = (handle, Z, ell, var, q, T(cfg["m0"]), T(cfg["S0"]), 4, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), cfg["squash_scale"],
             cfg["squash_shift"])
  # sharded over all ranks
  loss, grads = gd.pathwise_policy_loss_and_grad(*pw_args, **pw_kw)
  losses, (start, count), rg = gd.mm_restart_losses_and_grads(*mm_args)
  # gradients of every rank's restart block, gathered in global order for the comparison
  rg_all = [gd.allgather_concat(x, R) for x in rg]
  out = None
  # rank 0 alone: a one-member group makes the helpers take their single-process path
  g0 = dist.new_group([0])
  if rank == 0:
    loss1, grads1 = gd.pathwise_policy_loss_and_grad(*pw_args, **pw_kw, group=g0)
    losses1, _, rg1 = gd.mm_restart_losses_and_grads(*mm_args, group=g0)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    out = {"pathwise_mean_loss_rel": abs(float(loss) - float(loss1)) / abs(float(loss1)),
           "pathwise_grad_rel": max(rel(a, b) for a, b in zip(grads, grads1)),
           "mm_restart_loss_rel": rel(losses, losses1),
           "mm_restart_grad_rel": max(rel(a, b) for a, b in zip(rg_all, rg1)),
           "cases": f"pathwise: 1000 particles, 64 bases, H=3, one all-reduce of [1+P]; MM: {R} restarts (uneven shards), H=4, all-gather of loss"}
    out["ok"] = bool(out["pathwise_mean_loss_rel"] < 1e-12 and out["pathwise_grad_rel"] < 1e-9 and out["mm_restart_loss_rel"] < 1e-12
                     and out["mm_restart_grad_rel"] < 1e-9)
    if not out["ok"]:
      raise RuntimeError(f"sharded closures disagree with the single-process result: {out}")
  dist.barrier()
  return out


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch
  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core this process may run on
  torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
  cfg = build_workload(min(args.inputs, 256), 0)
  total = args.steps + args.warmup
  budget = 150.0
  rate, sample, ms = time_cpu(cfg, budget / max(total, 1), reference_form=True, steps=total)
  line = {
      "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f64", "data": "synthetic",
      "config": {"workload": "config#2 batched exact MM GP predict (M=1000, D=6, E=4, full 4x4 cov + cross)",
                 "inputs_per_gpu": args.inputs},
      "cpu_baseline": {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
      "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "note": "upstream needs tensorflow/gpflow (absent): timed is oracle/, its line-by-line CPU restatement",
  }
  _emit(json.dumps(line))


def run_b200(args):
  import torch
  import torch.distributed as dist
  from gpflowpilco_b200 import _lib, models
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise RuntimeError("bench.py --impl b200 needs a CUDA device: the hot path has no CPU implementation")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  lib = _lib.load()
  pk, peak_src = peaks()

  cfg = build_workload(args.inputs, rank)
  N = args.inputs
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  # the reference-facing objects: E independent exact GPs == multi-output model with shared inputs
  kern = models.SeparateIndependent([models.SquaredExponential(cfg["variance"][e], T(cfg["lengthscales"][e])) for e in range(E_OUT)])
  for k in kern.kernels:
    k.variance = k.variance.to(dev)
  model = models.SVGP(kern, models.SharedIndependentInducingVariables(models.InducingPoints(T(cfg["X"]))),
                      q_mu=T(cfg["Y"] - cfg["mean_const"]), q_sqrt=torch.zeros(E_OUT, M_TRAIN, M_TRAIN, dtype=torch.float64, device=dev),
                      whiten=False, mean_function=models.Constant(T(cfg["mean_const"])))
  model.kuu_jitter = [float(v) for v in cfg["noise_variance"]]     # exact-GP noise on the diagonal (GPR semantics)
  mu_d, cov_d = T(cfg["mu"]), T(cfg["cov"])
  mu_h = torch.as_tensor(cfg["mu"]).pin_memory()
  cov_h = torch.as_tensor(cfg["cov"]).pin_memory()
  out_h = [torch.empty(N, E_OUT, dtype=torch.float64).pin_memory(), torch.empty(N, E_OUT, E_OUT, dtype=torch.float64).pin_memory(),
           torch.empty(N, D_IN, E_OUT, dtype=torch.float64).pin_memory()]
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

  def step_resident():
    return moment_matching(GaussianMoments((mu_d, cov_d), True), model, check=False)

  def step_e2e():
    m = mu_h.to(dev, non_blocking=True)
    S = cov_h.to(dev, non_blocking=True)
    match = moment_matching(GaussianMoments((m, S), True), model, check=False)
    out_h[0].copy_(match.y.mean(), non_blocking=True)
    out_h[1].copy_(match.y.covariance(), non_blocking=True)
    out_h[2].copy_(match.cross[0], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return match

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # FP64 pipe roofline denominator, measured in this run
  sink = torch.empty(148 * 8 * 256, dtype=torch.float64, device=dev)
  fp64_peak = 0.0
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.gpp_microbench_fp64(148 * 8, 256, 20000, ctypes.c_void_p(sink.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record()
    torch.cuda.synchronize()
    fp64_peak = max(fp64_peak, 148 * 8 * 256 * 20000 * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)

  for _ in range(args.warmup):
    step_resident()
  lib.gpp_profile_enable(1)
  sampler = ClockSampler(local)
  barrier()
  sampler.start()
  launches0 = lib.gpp_launch_count()
  step_ms, kern_ms = [], []
  for _ in range(args.steps):
    flush.fill_(1)                         # L2 flush between timed iterations (outside the event pairs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step_resident()
    e1.record()
    e1.synchronize()
    step_ms.append(e0.elapsed_time(e1))
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    kern_ms.append(ms.value)
  launches = lib.gpp_launch_count() - launches0
  barrier()
  clocks = sampler.stop()
  lib.gpp_profile_enable(0)
  total_s = sum(step_ms) * 1e-3

  # end to end: host buffers in, host buffers out, through the reference-facing API
  for _ in range(2):
    step_e2e()
  barrier()
  t0 = time.perf_counter()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    step_e2e()
  e1.record()
  e1.synchronize()
  e2e_s = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)

  t = torch.tensor([total_s, e2e_s], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  total_s, e2e_s = float(t[0]), float(t[1])

  # second half of the metric; every rank takes part (its particles are sharded by global index)
  pathwise = None if args.no_pathwise else pathwise_section(dev, lib, pk, world)
  policy_opt = None if args.no_policy_opt else policy_opt_section(dev, lib, world, fp64_peak)
  psi2 = psi2_section(dev, lib, pk) if (rank == 0 and not args.no_psi2) else None
  mr_check = multirank_check(dev, world) if world > 1 else None

  if rank == 0:
    kern_s = float(np.mean(kern_ms)) * 1e-3
    achieved = N * ENTRIES_PER_INPUT * FLOP_PER_ENTRY / kern_s / 1e12
    line = {
        "metric": METRIC, "value": world * N * args.steps / total_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config#2 batched exact MM GP predict (M=1000, D=6, E=4, full 4x4 cov + cross)",
                   "inputs_per_gpu": N, "l2": "flushed between timed steps (256 MB write)", "timing": "CUDA events per step, max over ranks"},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "traffic": 96.25e6 if N == 8192 else None,
                     "traffic_source": "ncu --set full of this command at N=8192 (profiles/r2m_contract_final_bench_full.txt): dram read 78.41 MB + write 17.84 MB per launch",
                     "kernel": "k_contract", "kernel_ms": 1e3 * kern_s,
                     "kernel_share_of_step": kern_s / (total_s / args.steps),
                     "algorithmic": f"{FLOP_PER_ENTRY} flop/entry x {ENTRIES_PER_INPUT} entries/input x {N} inputs",
                     "peak_source": "in-run DFMA microbenchmark (gpp_microbench_fp64); MEASURED_PEAKS.json has no FP64 figure",
                     "peak_nominal": 148 * 64 * 2 * 1.965e-3, "frac_of_nominal": achieved / (148 * 64 * 2 * 1.965e-3),
                     "entries_evaluated_per_input": (E_OUT * (M_TRAIN // 128 + (1 if M_TRAIN % 128 else 0)) * ((M_TRAIN // 128 + (1 if M_TRAIN % 128 else 0)) + 1) // 2
                                                     + E_OUT * (E_OUT - 1) // 2 * (M_TRAIN // 128 + (1 if M_TRAIN % 128 else 0)) ** 2) * 128 * 128,
                     "entries_note": "algorithmic count (SURVEY 8d) = E(E+1)/2 M^2 = 1.0e7 per input; the kernel evaluates 128 x 128 tiles, upper tiles only for the E diagonal pairs",
                     "hbm_peak_gbs": pk.get("hbm_gbs"), "hbm_peak_source": peak_src},
        "e2e": {"value": world * N * args.steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(mu_h.numel() * 8 + cov_h.numel() * 8), "d2h_bytes_per_step": int(sum(o.numel() * 8 for o in out_h))},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if pathwise is not None:
      line["pathwise"] = pathwise
    if policy_opt is not None:
      line["policy_opt_step"] = policy_opt
    if psi2 is not None:
      line["psi2"] = psi2
    if mr_check is not None:
      line["multirank_check"] = mr_check
    if not args.no_cpu_baseline and world == 1:      # reported at N=1 only (rank 0), on a bounded sample
      small = {k: (v[:64] if k in ("mu", "cov") else v) for k, v in cfg.items()}
      rate, sample, _ = time_cpu(small, args.cpu_baseline_seconds, reference_form=True)
      rate2, sample2, _ = time_cpu(small, args.cpu_baseline_seconds / 2, reference_form=False)
      import torch as _t
      line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port", "sample": sample,
                              "reassociated_form": {"value": rate2, "sample": sample2}}
    _emit(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  global ARGS, _emit
  args = parse_args()
  ARGS = args
  # the contract is ONE JSON line on stdout: libraries that print to fd 1 (NCCL prints its version there) are sent to stderr,
  # and the line itself is written to the saved descriptor
  sys.stdout.flush()
  real_stdout = os.dup(1)
  os.dup2(2, 1)

  def _emit(line: str):
    os.write(real_stdout, (line + "\n").encode())
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
