"""clustering_b200 -- B200-native `clustering density` hot path of moldyn/Clustering.

The product is the C-ABI library ``libdcb200.so`` (include/dcb200.h; sources in clustering_b200/csrc) and
the C++ shim with the reference's own signatures (include/dcb200/density_cuda.hpp).  This Python package is the
test / benchmark harness around it: ctypes bindings (``lib``), a mirror of the reference's density
operators on numpy arrays (``density``), the one-process-per-GPU driver (``dist``) and the synthetic
trajectories of SURVEY.md section 8d (``synth``).  There is no CPU fallback anywhere in here.
"""
