"""B200-native APR / BPR-MF hot path of feay1234/Adversarial-Collaborative-Filtering.

The directory name carries the reference's name and is not a Python identifier; import it as ``apr_b200`` (the shim
package next to it) -- ``from apr_b200.APR import MF, training, sampling, shuffle``.
Modules mirror the reference files of the path: APR, utils, evaluation, evaluation_adv, Dataset, Recommender,
FastAdversarialMF, run_adv, run_adv_ori.  The math lives in csrc/ (hand-written sm_100a CUDA) behind the C ABI of
include/apr_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
