"""plancklens_b200: B200-native (sm_100a) replacement for the spherical-harmonic hot path of plancklens.

The package mirrors the reference's module layout for that path (`shts`, `utils_spin`, `qcinv.*`, `filt.*`,
`qest`, `utils_qe`) and routes every transform through hand-written CUDA behind the C ABI declared in
`include/plk.h` (library `plancklens_b200/csrc/libplk_b200.so`).  There is no CPU fallback: importing a compute
module without the built library, or calling it without a CUDA device, raises.
"""
__version__ = "0.1.0"
