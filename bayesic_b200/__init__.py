"""bayesic_b200: B200-native evaluation of Bayesic's einsum plans (sufficient
statistics, mixture responsibilities, ELBO terms).  See DESIGN.md."""
__version__ = '0.1.0'
