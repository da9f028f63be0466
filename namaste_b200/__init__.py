"""namaste_b200 -- B200-native implementation of Namaste's single-transit MCMC likelihood path.

Public surface mirrors the reference (namaste/namaste.py) for that path only:
MonotransitModel, MonoLnPrior, MonoOnlyLogProb, MonoOnlyNegLogProb, CalcTdur, CalcVel, the prior
construction (calcmaxvel, monoPriors, initLD, BuildMeanPriors) and RunMCMC's set-up / trimming helpers;
plus the staged-ensemble objects in :mod:`namaste_b200.core`.  All arithmetic is in the
CUDA library (namaste_b200/csrc, C ABI in include/namaste_b200.h); there is no CPU path.
"""
from .namaste import (MonotransitModel, MonoLnPrior, MonoOnlyLogProb, MonoOnlyNegLogProb, CalcTdur, CalcVel,
                      stage_lightcurve, clear_cache, calcminp, calcmaxvel, monoPriors, initLD, BuildMeanPriors,
                      transit_window_mask, nonan, AnomCutDiff, initial_walkers, trim_samples, RunMCMC, Optimize_mono, VelToOrbit, PlanetRtoM, MonoFinalPars)
from .sampler import EnsembleSampler
from .distributed import CandidateBatchSampler
from .gp import GP, RealTerm, JitterTerm, RotationTerm, MonoLogProb, MonoNegLogProb
from .core import LightCurve, occultquad, lnprior, prior_table, measure_fp64_peak, launch_count, PARAM_NAMES

__version__ = "0.1.0"
