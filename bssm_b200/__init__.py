"""bssm_b200 -- B200-native particle-filter engine behind bssm's R surface.

Only the particle-filter hot path of helske/bssm (bootstrap filter, psi-APF,
particle smoother, IS-MCMC post-correction); see DESIGN.md for scope.
"""
from .models import Model, ar1_lg, ar1_ng, bsm_lg, bsm_ng, ssm_ulg, init_mode, nlg_register, sde_register, ssm_mng, ssm_nlg, ssm_sde, ssm_ung, svm  # noqa: F401
from .api import (FilterJob, PostcorrJob, bootstrap_filter, ekpf_filter, importance_sample, logLik, particle_smoother,  # noqa: F401
                  post_correct, stratified_sample, suggest_N, gaussian_approx_mode, gaussian_psi_smoother,
                  synthetic_svm_mcmc_output)
from ._lib import EngineError, lib  # noqa: F401
