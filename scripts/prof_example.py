"""cProfile of the façade's policy-optimisation loop (examples/cartpole_policy_optimisation.py)."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "--steps", "30"]
import examples.cartpole_policy_optimisation as ex
ex.main()
pr = cProfile.Profile(); pr.enable(); ex.main(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
