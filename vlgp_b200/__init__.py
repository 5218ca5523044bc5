"""vlgp_b200 -- B200-native variational-EM engine behind the vLGP entry points of catniplab/vlgp.

    import vlgp_b200 as vlgp
    result = vlgp.fit(trials, n_factors=3)        # {"trials", "params", "config"}

The arithmetic runs in libvlgp_b200.so (hand-written sm_100a CUDA, ctypes C ABI); there is no CPU fallback.
"""
from .api import fit, posterior_cov, sample_posterior, transform  # noqa: F401

__all__ = ["fit", "posterior_cov", "sample_posterior", "transform"]
__version__ = "0.1.0"
