"""larndsim_b200 — B200-native (sm_100a) implementation of the larnd-sim-jax detector-simulation hot path.

Module map (reference module -> this package):
    larndsim.consts_jax -> larndsim_b200.consts     larndsim.sim_jax    -> larndsim_b200.sim
    larndsim.detsim_jax -> larndsim_b200.detsim     larndsim.fee_jax    -> larndsim_b200.fee
    larndsim.losses_jax -> larndsim_b200.losses (adc2charge / mmd / mse_adc / params_loss)
    larndsim.quenching_jax -> larndsim_b200.quenching     larndsim.drifting_jax -> larndsim_b200.drifting
    optimize.dataio (chop_tracks, pad_batch) -> larndsim_b200.dataio     jax.random (key/split/normal) -> larndsim_b200.jrandom
All heavy work is done by hand-written CUDA kernels in csrc/, reached through the C ABI of
include/larnd_b200.h; there is no CPU fallback.
"""
from ._lib import LarndError, build_library, get_lib  # noqa: F401
from .consts import (RecombinationMode, build_params_class, get_vdrift, load_detector_properties,  # noqa: F401
                     load_geometry_json, load_lut)
from .sim import (pad_size, shift_tracks, simulate_hits, simulate_parametrized, simulate_probabilistic,  # noqa: F401
                  simulate_stochastic, simulate_wfs)
