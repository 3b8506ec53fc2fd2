"""moldyn_b200 — B200-native (sm_100a) implementation of AndrewChe7/moldyn's `solve` step loop.

The product is the C-ABI shared library built from moldyn_b200/csrc (include/moldyn_b200.h); this package is the
thin host-side mirror of the reference's solver interface on top of it.  No CPU fallback exists.
"""
from .solver import (Barostat, Integrator, K_B, MdError, MultiState, Potential, PotentialsDatabase, Solver, State, Thermostat,
                     get_center_of_mass_velocity, get_kinetic_energy, get_momentum_of_system, get_potential_energy,
                     get_pressure, get_temperature, get_thermal_energy, update_force)

__all__ = [
    "Barostat", "Integrator", "K_B", "MdError", "MultiState", "Potential", "PotentialsDatabase", "Solver", "State", "Thermostat",
    "get_center_of_mass_velocity", "get_kinetic_energy", "get_momentum_of_system", "get_potential_energy",
    "get_pressure", "get_temperature", "get_thermal_energy", "update_force",
]
