"""Quick device timing of the hot kernels (developer tool; bench.py is the contract)."""
import ctypes
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gpflowpilco_b200 import _lib, ops, synthetic

lib = _lib.load()
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)


def dfma_peak():
  sink = torch.empty(148 * 8 * 256, dtype=torch.float64, device=dev)
  best = 0.0
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20000
    e0.record()
    _lib.check(lib.gpp_microbench_fp64(148 * 8, 256, iters, ctypes.c_void_p(sink.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = max(best, 148 * 8 * 256 * iters * 16 / (ms * 1e-3) / 1e12)
  return best


def main():
  N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
  print("DFMA peak TFLOP/s:", dfma_peak())
  cfg = synthetic.config2_batched_mm_predict(N=N)
  E = cfg["Y"].shape[1]
  t0 = time.time()
  h = ops.GPModelHandle(T(np.broadcast_to(cfg["X"], (E,) + cfg["X"].shape).copy()), T(cfg["lengthscales"]), T(cfg["variance"]),
                        T(cfg["Y"] - cfg["mean_const"]), None, whiten=False, mean_const=T(cfg["mean_const"]),
                        kuu_jitter=list(cfg["noise_variance"]))
  torch.cuda.synchronize()
  print("model prepare s:", time.time() - t0)
  mu, cov = T(cfg["mu"]), T(cfg["cov"])
  lib.gpp_profile_enable(1)
  for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f1, Sff, cross = h.predict(mu, cov, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    tot = e0.elapsed_time(e1)
    entries = N * 10 * 1000 * 1000
    print(f"iter {it}: total {tot:.2f} ms, contract {ms.value:.2f} ms, {N / tot * 1e3:.0f} inputs/s, "
          f"{entries * 38 / (ms.value * 1e-3) / 1e12:.2f} algorithmic TFLOP/s")
  print("f1[0]", f1[0].tolist(), "Sff[0,0]", Sff[0, 0].tolist())

  c3 = synthetic.config3_psi2_stress(N=256)
  mu3, cov3 = T(c3["mu"]), T(c3["cov"])
  for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.ekzxkxz(mu3, cov3, T(c3["Z1"]), T(c3["lengthscales1"]), c3["variance1"], T(c3["Z2"]), T(c3["lengthscales2"]),
                      c3["variance2"], check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = ctypes.c_float()
    lib.gpp_profile_last_ms(ctypes.byref(ms))
    print(f"psi2 N=256: total {e0.elapsed_time(e1):.2f} ms kernel {ms.value:.2f} ms -> {out.numel() * 8 / (ms.value * 1e-3) / 1e9:.0f} GB/s written")
    del out


if __name__ == "__main__":
  main()
