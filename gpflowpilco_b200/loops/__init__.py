from gpflowpilco_b200.loops.pilco import (AbstractPILCO, EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO,  # noqa: F401
                                          PathwisePILCO)
