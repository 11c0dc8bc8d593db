"""glomap_b200 -- B200-native numeric core for global SfM (rotation averaging,
BATA global positioning, bundle adjustment) behind the estimator interfaces of
colmap/glomap.  The product path is hand-written sm_100a CUDA in
``csrc/`` reached through the C ABI declared in ``include/b200sfm.h``;
this package only holds the host-side mirror of the reference estimator
classes (``estimators.py``), the flat problem containers and synthetic scene
generators.  There is NO CPU fallback: importing the estimators without the
built shared library raises.
"""
__all__ = ["estimators", "synthetic", "geometry"]
