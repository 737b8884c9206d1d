"""bwgr_b200 -- B200-native (sm_100a) marker-effect update loop of bWGR behind the reference's own
function names.  The product path is CUDA only: importing works anywhere, calling needs a B200."""
from .api import (PATH_AUTO, PATH_BLOCKED, PATH_GRID, PATH_SMALL_N, STORE_2BIT, STORE_F32, STORE_I8, BayesA, BayesB, BayesC, BayesRR,  # noqa: F401
                  EmStepper, Genotypes, KMUP, KMUP2, em_fit, emBA, emBB, emBC, emBL, emEN, emRR, gibbs_fit, wgr, MRR3, MRR3F, mrr, mrr_float,
                  emDE, emML, emBCpi, lasso, BayesL, BayesCpi, BayesDpi, emCV, mcmcCV, GSRR, GSFLM, emML2, BayesA2, BayesB2, BayesRR2)
from ._lib import BwgrError, LIB_PATH, SYMBOLS  # noqa: F401
from .api import trim  # noqa: F401
