"""B200-native field-solve + event-selection hot path of DeviceKMC (libkmc_b200.so + host mirror).

Import via importlib (the directory name mirrors the reference repository's name):

    import importlib
    kmc = importlib.import_module(
        "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200")
"""
from .api import *  # noqa: F401,F403
from .api import load_library, LIB_PATH, SIGNATURES, Params, Context, KMatrix, Events, Structure, DeviceKMC  # noqa: F401
