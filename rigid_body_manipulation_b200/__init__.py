"""B200-native batched rigid-body inverse dynamics behind the call surface of barikata1984/rigid-body-manipulation.

  engine          device-resident Model + batched entry points (C ABI of include/rbm_b200.h through ctypes)
  model           MuJoCo-free model front-end (MJCF subset / CAD numbers -> kernel constants)
  identification  Gram pack -> inertial parameters (host 10x10 solve), reference score
  distributed     sharding + the one collective (Gram all-reduce)
  dropin/         `dynamics`, `transformations` packages with the reference's names and signatures
  lie             SE3 / SO3 value types (liegroups' own when installed)

No module here imports anything under oracle/ (test infrastructure), and nothing computes on the CPU when the CUDA
library or device is missing: those calls raise.
"""
__version__ = "0.1.0"
