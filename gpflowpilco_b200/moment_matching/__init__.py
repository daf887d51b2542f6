from gpflowpilco_b200.moment_matching.core import *          # noqa: F401,F403
from gpflowpilco_b200.moment_matching import models           # noqa: F401  (registers the GP rules)
from gpflowpilco_b200.moment_matching import rules            # noqa: F401  (maths / bijector / encoder rules)
