"""vgpmp_b200: B200-native (sm_100a) implementation of vgpmp's per-iteration ELBO hot path.

Module layout mirrors the reference package `gpflow_vgpmp` for the hot path only:
models/, likelihoods/, kernels/, inducing_variables/, kernel_conditioning/, covariances/, kullback_leiblers/,
derivatives/ (the fused reverse pass), utils/{sampler, sdf_utils, robot, miscellaneous}.  All arithmetic is in
csrc/*.cu behind the C-ABI of include/vgpmp_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
