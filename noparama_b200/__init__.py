"""noparama_b200 -- B200-native (sm_100a) Gibbs reassignment for Dirichlet-process mixtures.

One hot path of mrquincle/noparama (Neal Algorithm 8 and the machinery it shares with the split-merge samplers),
behind the reference's sampler seam.  All compute is in libnpb200.so (CUDA); there is no CPU path.
"""
from .api import (ALG8, ALG2, ALG2_CONJUGATE, JAIN_NEAL, TRIADIC, UPDATE_POSTERIOR_DRAW, UPDATE_POSTERIOR_MEAN, BUGCOMPAT_DEFAULT, BUGCOMPAT_DEGENERATE_IW, BUGCOMPAT_UNDERFLOW,
                  Chains, Comm, Context, Dataset, MCMC, MultivariateNormal, NealAlgorithm8, NealAlgorithm2, NealAlgorithm2Conjugate, JainNealAlgorithm, TriadicAlgorithm,
                  NormalInverseWishart, NormalInverseGamma, ScalarNoiseNormal, FAMILY_REGRESSION, FAMILY_ANGULAR, NpbError,
                  SweepStats, load_library, replay_alg8, replay_split_merge, scan_order, LIB_PATH, EXPORTS)
from . import synthetic, diagnostics

__all__ = ["ALG8", "ALG2", "ALG2_CONJUGATE", "JAIN_NEAL", "TRIADIC", "Chains", "Context", "Dataset", "MCMC", "MultivariateNormal",
           "NealAlgorithm8", "NealAlgorithm2", "NealAlgorithm2Conjugate", "JainNealAlgorithm", "TriadicAlgorithm", "NormalInverseWishart", "NormalInverseGamma", "ScalarNoiseNormal", "FAMILY_REGRESSION", "FAMILY_ANGULAR", "NpbError", "SweepStats", "load_library", "synthetic"]
