"""
shot_fpfh_b200 — the B200 (sm_100a) hot path of aubin-tchoi/shot-fpfh behind the reference's own Python API:
fixed-radius neighbour search -> SHOT / FPFH descriptors -> descriptor nearest-neighbour matching.

    from shot_fpfh_b200.descriptors import ShotMultiprocessor, compute_fpfh_descriptor
    from shot_fpfh_b200.matching import basic_matching, match_descriptors, double_matching_with_rejects

`shot_fpfh_b200.dropin.install()` rebinds those names inside an importable reference `shot_fpfh` package so that
its `RegistrationPipeline` and `register_point_clouds` script run unchanged on top of the CUDA kernels.

Importing the compute modules loads csrc/libshotfpfh_b200.so and raises if it has not been built; there is no
CPU fallback. `shot_fpfh_b200.synthetic` and `shot_fpfh_b200.subsampling` are pure NumPy helpers and import
without the library.
"""

__version__ = "0.1.0"
