"""Drop-in `figaroh` package exposing the dynamic-identification hot path on B200.

Only figaroh.tools.{regressor,qrdecomposition,robot} and the two preprocessing functions of
figaroh.identification.identification_tools exist here: they mirror the
reference modules of the same names (same functions, arguments, return values and
error behaviour) and run on the CUDA library in ../lib through figaroh_b200.
"""
__version__ = "0.1.0"
