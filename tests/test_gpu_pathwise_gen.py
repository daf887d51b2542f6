"""GPU parity for the device-side path generation: Philox words bit-exact vs oracle/philox.py, normal/uniform draws,
initial states, and generated function draws vs oracle/pathwise.py::generate_paths (same seed, same global particle
indices, including a sharded generation that must reproduce the unsharded one)."""
import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import pathwise as pw
from oracle import philox
from tests.helpers import DTYPE, cuda_handle, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


@pytest.mark.parametrize("first,stream,seed", [(0, 0, 0), (12345, 2, 0xDEADBEEFCAFE), (2 ** 33 + 7, 5, 2 ** 63 + 11)])
def test_philox_words_bit_exact(first, stream, seed):
  from gpflowpilco_b200.pathwise import philox_raw
  n = 1000
  got = philox_raw(first, n, stream, seed).cpu().numpy().astype(np.uint32)
  ref = philox.philox4x32(np.arange(first, first + n, dtype=np.uint64), stream, seed)
  assert np.array_equal(got, ref)          # bit-exact sample indexing


def test_basis_and_initial_state_draws():
  from gpflowpilco_b200.pathwise import draw_basis, draw_initial_states
  L, F, D, seed = 3, 64, 6, 99
  omega, phase = draw_basis(L, F, D, seed)
  o_ref, p_ref = pw.draw_basis(L, F, D, seed)
  scaled_close(omega, o_ref, 1e-14, "omega")
  scaled_close(phase, p_ref, 1e-15, "phase")
  m0 = torch.tensor([0.0, np.pi, 0.0, 0.0], dtype=DTYPE)
  S0 = torch.diag(torch.tensor([0.01, 0.02, 0.03, 0.04], dtype=DTYPE)) + 0.002
  x0 = draw_initial_states(_dev(m0), _dev(S0), seed, 1000, 50)
  x_ref = pw.draw_initial_states(m0, torch.linalg.cholesky(S0), seed, 1000, 50)
  scaled_close(x0, x_ref, 1e-14, "x0")


@pytest.mark.parametrize("whiten", [True, False])
def test_generate_paths_matches_oracle_and_is_sharding_invariant(whiten):
  from gpflowpilco_b200.pathwise import generate_paths
  L, M, D, F, S, seed = 2, 24, 3, 64, 40, 7
  P = synthetic.random_svgp(L=L, M=M, D=D, seed=5, whiten=whiten, z_scale=2.0)
  model = oracle_svgp(P)
  ref = pw.generate_paths(model, F, seed, 0, S)
  h = cuda_handle(P)
  got = generate_paths(h, S, F, seed, first_particle=0)
  scaled_close(got.w[:, :, :S].permute(2, 0, 1), ref.w, 1e-13, "prior weights")
  # v = Kuu^-1 (...) amplifies round-off by cond(Kuu); compare the function values it defines (well conditioned) tightly
  # and the raw weights loosely
  scaled_close(got.v[:, :M, :S].permute(2, 0, 1), ref.v, 1e-6, "update weights")
  g = torch.Generator().manual_seed(0)
  x = torch.randn(S, D, dtype=DTYPE, generator=g)
  f_ref = pw.evaluate_paths(model, ref, x)
  packed_ref = pw.Paths(ref.omega, ref.phase, got.w[:, :, :S].permute(2, 0, 1).cpu(), got.v[:, :M, :S].permute(2, 0, 1).cpu())
  scaled_close(pw.evaluate_paths(model, packed_ref, x), f_ref, 1e-9, "function values of the generated draws")
  # two shards of 25 + 15 particles reproduce the same draws
  a = generate_paths(h, 25, F, seed, first_particle=0)
  b = generate_paths(h, 15, F, seed, first_particle=25)
  assert torch.equal(a.w[:, :, :25], got.w[:, :, :25]) and torch.equal(b.w[:, :, :15], got.w[:, :, 25:40])
  scaled_close(torch.cat([a.v[:, :M, :25], b.v[:, :M, :15]], -1), got.v[:, :M, :S].cpu(), 1e-9, "sharded update weights")
