"""ctypes binding of the collision part of the CPU oracle (oracle/rumdeed_oracle_collisions.c).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py): restates src/mod_collisions.F90 and
src/mod_polynomialroots.F90 of the reference; nothing under rumdeed_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .oracle import Oracle, Rng, _PD, _PI

REMOVE_TOP, REMOVE_RECOM = 1, 3


class CrossTables(C.Structure):
    _fields_ = [("n_tot", C.c_int), ("n_ion", C.c_int),
                ("tot_energy", _PD), ("tot_data", _PD), ("ion_energy", _PD), ("ion_data", _PD)]


class CollConstants(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("R_inf", "Ryd", "N_n", "N_bind", "Z_eff", "Z_eff2")]


class RecombEvent(C.Structure):
    _fields_ = [("step", C.c_int), ("ion_pos", C.c_double * 3),
                ("elec_speed", C.c_double), ("dist", C.c_double), ("recom_rad", C.c_double),
                ("elec_slot", C.c_int), ("ion_slot", C.c_int), ("elec_emit", C.c_int), ("ion_life", C.c_int),
                ("t", C.c_double)]


class IonizationEvent(C.Structure):
    _fields_ = [("step", C.c_int), ("in_slot", C.c_int), ("pos", C.c_double * 3),
                ("in_speed", C.c_double), ("out_speed", C.c_double), ("new_speed", C.c_double),
                ("new_vel", C.c_double * 3), ("ejec_pos", C.c_double * 3), ("ejec_vel", C.c_double * 3),
                ("ion_pos", C.c_double * 3), ("E1", C.c_double), ("collE", C.c_double), ("ejecE", C.c_double),
                ("elec_emit", C.c_int)]


def _d(a):
    return a.ctypes.data_as(_PD)


def _i(a):
    return a.ctypes.data_as(_PI)


class Collisions:
    """Oracle collision routines on plain arrays; `tables` = (tot_energy, tot_data, ion_energy, ion_data)."""

    def __init__(self, orc: Oracle | None = None, tables=None, seed: int = 1):
        self.orc = orc or Oracle()
        lib = self.lib = self.orc.lib
        D = C.c_double
        PT = C.POINTER(CrossTables)
        lib.orc_coll_get_constants.argtypes = [C.POINTER(CollConstants)]
        for n in ("orc_normal_dist", "orc_folded_normal_dist"):
            getattr(lib, n).argtypes = [D, D, D]; getattr(lib, n).restype = D
        lib.orc_folded_normal_max.argtypes = [D, D]; lib.orc_folded_normal_max.restype = D
        lib.orc_kramers_cross_section.argtypes = [D]; lib.orc_kramers_cross_section.restype = D
        lib.orc_binary_search.argtypes = [_PD, C.c_int, D, _PI, _PI]; lib.orc_binary_search.restype = C.c_int
        for n in ("orc_find_cross_tot_data", "orc_find_cross_ion_data"):
            getattr(lib, n).argtypes = [PT, D]; getattr(lib, n).restype = D
        lib.orc_update_collision_data.argtypes = [PT, _PD, _PD]
        lib.orc_solve_polynomial.argtypes = [D, D, D, D, D, _PI, _PD]
        lib.orc_recombination_pair.argtypes = [_PD, _PD, _PD, _PD, D, D, _PD, _PD]
        lib.orc_recombination_pair.restype = C.c_int
        lib.orc_discrete_recombination_ots.argtypes = [C.c_int, _PD, _PD, _PD, _PI, _PI, _PI, _PI, _PI, _PD, C.c_int, D,
                                                       C.POINTER(RecombEvent), C.c_int, _PI, _PI]
        lib.orc_discrete_recombination_ots.restype = C.c_int
        PR = C.POINTER(Rng)
        lib.orc_get_injected_vec.argtypes = [PR, D, _PD, _PD]
        lib.orc_get_ejected_vec.argtypes = [PR, D, D, _PD, _PD]
        lib.orc_continuous_ionization_ots.argtypes = [PR, PT, C.c_int, _PD, _PD, _PD, _PI, _PI, _PI, D, D, C.c_int,
                                                      C.POINTER(IonizationEvent), C.c_int, _PI]
        lib.orc_continuous_ionization_ots.restype = C.c_int
        lib.orc_rng_seed.argtypes = [PR, C.c_uint64]
        self.k = CollConstants()
        lib.orc_coll_get_constants(C.byref(self.k))
        self.rng = Rng()
        lib.orc_rng_seed(C.byref(self.rng), seed)
        self.tables = None
        if tables is not None:
            self.set_tables(*tables)

    def set_tables(self, tot_energy, tot_data, ion_energy, ion_data):
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (tot_energy, tot_data, ion_energy, ion_data)]
        t = CrossTables()
        t.n_tot, t.n_ion = len(self._keep[0]), len(self._keep[2])
        t.tot_energy, t.tot_data, t.ion_energy, t.ion_data = (_d(a) for a in self._keep)
        self.tables = t

    # -- scalar helpers -------------------------------------------------------------------
    def normal_dist(self, mu, sigma, x): return self.lib.orc_normal_dist(mu, sigma, x)
    def folded_normal_dist(self, mu, sigma, x): return self.lib.orc_folded_normal_dist(mu, sigma, x)
    def folded_normal_max(self, mu, sigma): return self.lib.orc_folded_normal_max(mu, sigma)
    def kramers(self, energy): return self.lib.orc_kramers_cross_section(energy)
    def cross_tot(self, energy): return self.lib.orc_find_cross_tot_data(C.byref(self.tables), energy)
    def cross_ion(self, energy): return self.lib.orc_find_cross_ion_data(C.byref(self.tables), energy)

    def collision_data(self, vel):
        """Update_Collision_Data per row: columns energy, ion_cross_sec, ion_cross_rad, recom_cross_rad, tot_cross_sec."""
        vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(vel), 5))
        for r in range(len(vel)):
            self.lib.orc_update_collision_data(C.byref(self.tables), _d(vel[r]), _d(out[r]))
        return out

    def solve_polynomial(self, a, b, c, d, e):
        code = C.c_int(0)
        z = np.empty(8)
        self.lib.orc_solve_polynomial(a, b, c, d, e, C.byref(code), _d(z))
        return code.value, z[0::2] + 1j * z[1::2]

    def recombination_pair(self, ion_pos, elec_pos, elec_vel, elec_acc, recom_rad, dt):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (ion_pos, elec_pos, elec_vel, elec_acc)]
        t, dist = C.c_double(0), C.c_double(0)
        hit = self.lib.orc_recombination_pair(_d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]), recom_rad, dt, C.byref(t), C.byref(dist))
        return bool(hit), t.value, dist.value

    def discrete_recombination(self, pos, vel, acc, species, mask, life, step_born, emitter, recom_rad, step, dt,
                               max_events=None):
        """Returns (nrRecombinations, n_expired, events, mask_after, reason)."""
        n = len(species)
        pos, vel, acc, recom_rad = (np.ascontiguousarray(a, dtype=np.float64) for a in (pos, vel, acc, recom_rad))
        species, life, step_born, emitter = (np.ascontiguousarray(a, dtype=np.int32) for a in (species, life, step_born, emitter))
        mask = np.array(mask, dtype=np.int32)
        reason = np.zeros(n, dtype=np.int32)
        cap = max_events if max_events is not None else max(n, 1)
        ev = (RecombEvent * cap)()
        nexp = C.c_int(0)
        nr = self.lib.orc_discrete_recombination_ots(n, _d(pos), _d(vel), _d(acc), _i(species), _i(mask), _i(life),
                                                     _i(step_born), _i(emitter), _d(recom_rad), step, dt, ev, cap,
                                                     _i(reason), C.byref(nexp))
        return nr, nexp.value, [ev[k] for k in range(min(nr, cap))], mask, reason

    def injected_vec(self, T, par_vel):
        v = np.ascontiguousarray(par_vel, dtype=np.float64); out = np.empty(3)
        self.lib.orc_get_injected_vec(C.byref(self.rng), T, _d(v), _d(out))
        return out

    def ejected_vec(self, W, T, par_vel):
        v = np.ascontiguousarray(par_vel, dtype=np.float64); out = np.empty(3)
        self.lib.orc_get_ejected_vec(C.byref(self.rng), W, T, _d(v), _d(out))
        return out

    def continuous_ionization(self, pos, prev_pos, vel, species, mask, emitter, n_d, cyl_radius, step, max_events=None):
        """Returns (nrIonizations, nrCollisions, events, vel_after, emitter_after)."""
        n = len(species)
        pos, prev_pos = (np.ascontiguousarray(a, dtype=np.float64) for a in (pos, prev_pos))
        vel = np.array(vel, dtype=np.float64)
        species, mask = (np.ascontiguousarray(a, dtype=np.int32) for a in (species, mask))
        emitter = np.array(emitter, dtype=np.int32)
        cap = max_events if max_events is not None else max(n, 1)
        ev = (IonizationEvent * cap)()
        ncoll = C.c_int(0)
        nr = self.lib.orc_continuous_ionization_ots(C.byref(self.rng), C.byref(self.tables), n, _d(pos), _d(prev_pos), _d(vel),
                                                    _i(species), _i(mask), _i(emitter), n_d, cyl_radius, step, ev, cap,
                                                    C.byref(ncoll))
        return nr, ncoll.value, [ev[k] for k in range(min(nr, cap))], vel, emitter
