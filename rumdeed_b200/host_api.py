"""ctypes binding of librumdeed_host.so, the C++ mirror of the RUMDEED host around the hot path
(input namelist, emission plugins init / do-emission / clean-up, main loop, writers).

`Simulation` drives whole runs from a deck directory (`input`, `work`, `laser`) or from an in-memory
set-up, and exposes the samplers so the parity tests can check them against the CPU checker.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import Rb2Error, load_library

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "librumdeed_host.so")
RUN_EXE_PATH = os.path.join(_HERE, "rumdeed_b200_run")

SUPPLY_FE, SUPPLY_GTF = 1, 2
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)


class Setup(C.Structure):
    _fields_ = [
        ("emission_mode", C.c_int), ("V_s", C.c_double), ("box_dim", C.c_double * 3), ("time_step", C.c_double),
        ("image_charge", C.c_int), ("N_ic_max", C.c_int),
        ("emitters_pos", C.c_double * 3), ("emitters_dim", C.c_double * 3),
        ("emitters_type", C.c_int), ("emitters_delay", C.c_int), ("T_temp", C.c_double), ("mh_batch", C.c_int),
        ("planes_N", C.c_int), ("planes_z", C.c_double * 10),
        ("cuba_epsabs", C.c_double), ("cuba_epsrel", C.c_double), ("cuba_mineval", C.c_int), ("cuba_maxeval", C.c_int),
        ("work_y_num", C.c_int), ("work_x_num", C.c_int), ("work_w_theta", _PD),
        ("laser_gauss_mode", C.c_int), ("laser_mode", C.c_int), ("photon_mode", C.c_int),
        ("laser_energy", C.c_double), ("laser_variation", C.c_double),
        ("gauss_center", C.c_double), ("gauss_width", C.c_double), ("gauss_amplitude", C.c_double),
        ("max_particles", C.c_int), ("seed", C.c_ulonglong), ("ramo_sections", C.c_int),
    ]


class State(C.Structure):
    _fields_ = [
        ("step", C.c_int), ("nrPart", C.c_int), ("nrElec", C.c_int), ("nrIon", C.c_int), ("nrID", C.c_int),
        ("nrElecEmit", C.c_int),
        ("nrEmitted_total", C.c_longlong), ("nrAbsorbed_top", C.c_longlong), ("nrAbsorbed_bot", C.c_longlong),
        ("N_sup", C.c_double), ("df_avg", C.c_double), ("a_rate", C.c_double), ("MH_std", C.c_double), ("MH_std_tip", C.c_double),
        ("F_avg", C.c_double * 3), ("neval", C.c_int), ("fail", C.c_int), ("integral_error", C.c_double),
        ("ramo_current", C.c_double * 4), ("ramo_total", C.c_double), ("ramo_integral", C.c_double),
        ("avg_elec_vel", C.c_double * 3), ("accel_ms", C.c_float), ("step_ms", C.c_float),
        ("t_dev_step", C.c_double), ("t_dev_accel", C.c_double),
        ("t_emission", C.c_double), ("t_md_step", C.c_double), ("t_remove", C.c_double), ("t_io", C.c_double),
        ("nrIonizations_total", C.c_longlong), ("nrRecombinations_total", C.c_longlong),
        ("t_collisions", C.c_double), ("t_dev_collisions", C.c_double),
        ("t_em_quad", C.c_double), ("t_em_mh", C.c_double), ("t_em_add", C.c_double), ("n_candidates_total", C.c_longlong),
    ]


_hlib = None


def load_host_library():
    global _hlib
    if _hlib is not None:
        return _hlib
    load_library()  # the device library must be loadable first (no CPU fallback)
    if not os.path.exists(HOST_LIB_PATH):
        raise Rb2Error(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
    lib = C.CDLL(HOST_LIB_PATH)
    V = C.c_void_p
    lib.rh_create_from_dir.argtypes = [C.c_char_p, C.c_int, C.c_ulonglong, C.c_int]; lib.rh_create_from_dir.restype = V
    lib.rh_create.argtypes = [C.POINTER(Setup)]; lib.rh_create.restype = V
    lib.rh_init.argtypes = [V]
    lib.rh_step.argtypes = [V, C.c_int]
    lib.rh_run.argtypes = [V, C.c_int, C.c_int]
    lib.rh_get_state.argtypes = [V, C.POINTER(State)]
    lib.rh_steps_in_input.argtypes = [V]
    lib.rh_get_ramo_sections.argtypes = [V, C.c_int, _PD]
    lib.rh_set_option.argtypes = [V, C.c_char_p, C.c_double]
    lib.rh_destroy.argtypes = [V]; lib.rh_destroy.restype = None
    lib.rh_last_error.argtypes = [V]; lib.rh_last_error.restype = C.c_char_p
    lib.rh_cuba_integrate.argtypes = [V, C.c_int, _PD, _PD, _PI, _PI]
    lib.rh_mh_rectangle_J.argtypes = [V, _PD, _PD, _PD]
    lib.rh_mh_rectangle_J_batch.argtypes = [V, C.c_int, _PD, _PD, _PD]
    lib.rh_mh_rectangle_J_thermo.argtypes = [V, _PD]
    lib.rh_mh_rectangle_J_thermo_batch.argtypes = [V, C.c_int, _PD, _PI]
    lib.rh_metro_algo_tip_v3.argtypes = [V, C.c_int, _PD, _PD, _PD, _PD, _PD]
    lib.rh_metro_algo_tip_v3_batch.argtypes = [V, C.c_int, C.c_int, _PD, _PD, _PD]
    lib.rh_tip_supply_grid.argtypes = [V, C.c_int, C.c_int, _PD, _PD]
    lib.rh_do_emission.argtypes = [V, C.c_int, _PI]
    lib.rh_w_theta_xy.argtypes = [V, _PD, _PI]; lib.rh_w_theta_xy.restype = C.c_double
    lib.rh_kevin_jgtf_v2.argtypes = [C.c_double] * 3; lib.rh_kevin_jgtf_v2.restype = C.c_double
    _hlib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_PD)


class Simulation:
    """program RUMDEED (src/main.F90) on top of the device hot path."""

    def __init__(self, deck_dir: str | None = None, write_files=False, seed=0, max_particles=0, init=True, **setup):
        self.lib = load_host_library()
        self._keep = None
        if deck_dir is not None:
            self.ptr = self.lib.rh_create_from_dir(os.fsencode(deck_dir), int(write_files), seed, max_particles)
        else:
            u = Setup()
            u.emission_mode = setup.get("emission_mode", 10)
            u.V_s = setup["V_s"]
            u.box_dim[:] = setup["box_dim"]
            u.time_step = setup["time_step"]
            u.image_charge = int(setup.get("image_charge", True))
            u.N_ic_max = setup.get("N_ic_max", 1)
            u.emitters_pos[:] = setup["emitters_pos"]
            u.emitters_dim[:] = setup["emitters_dim"]
            u.emitters_type = setup.get("emitters_type", 2)
            u.emitters_delay = setup.get("emitters_delay", 0)
            u.T_temp = setup.get("T_temp", 293.15)
            u.mh_batch = int(setup.get("mh_batch", False))
            pz = list(setup.get("planes_z", ()))
            u.planes_N = len(pz)
            for k, z in enumerate(pz):
                u.planes_z[k] = z
            u.cuba_epsabs = setup.get("cuba_epsabs", 0.0)
            u.cuba_epsrel = setup.get("cuba_epsrel", 0.0)
            u.cuba_mineval = setup.get("cuba_mineval", 0)
            u.cuba_maxeval = setup.get("cuba_maxeval", 0)
            w = np.ascontiguousarray(np.atleast_2d(np.asarray(setup.get("w_theta", ((2.0,),)), dtype=np.float64)))
            self._keep = w
            u.work_y_num, u.work_x_num = w.shape
            u.work_w_theta = _d(w)
            laser = setup.get("laser")
            if laser:
                u.laser_gauss_mode, u.laser_mode, u.photon_mode = laser["gauss_mode"], laser["laser_mode"], laser["photon_mode"]
                u.laser_energy, u.laser_variation = laser["energy"], laser.get("variation", 0.0)
                u.gauss_center, u.gauss_width, u.gauss_amplitude = laser.get("center", 0.0), laser.get("width", 1.0), laser.get("amplitude", 0.0)
            u.max_particles = setup.get("max_particles", max_particles or 200000)
            u.seed = seed
            u.ramo_sections = int(setup.get("ramo_sections", 0))
            self.ptr = self.lib.rh_create(C.byref(u))
        if not self.ptr:
            raise Rb2Error("rh_create failed")
        self.open = True
        if init:  # allocates the device store: needs the GPU (no CPU fallback)
            self._check(self.lib.rh_init(self.ptr))

    def _check(self, rc):
        if rc != 0:
            raise Rb2Error("rumdeed host error: " + self.lib.rh_last_error(self.ptr).decode())

    def close(self):
        if getattr(self, "open", False):
            self.lib.rh_destroy(self.ptr)
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def steps_in_input(self):
        return self.lib.rh_steps_in_input(self.ptr)

    def step(self, i) -> State:
        self._check(self.lib.rh_step(self.ptr, int(i)))
        return self.state()

    def run(self, first, n):
        self._check(self.lib.rh_run(self.ptr, int(first), int(n)))
        return self.state()

    def state(self) -> State:
        st = State()
        self.lib.rh_get_state(self.ptr, C.byref(st))
        return st

    def set_option(self, name, value):
        self._check(self.lib.rh_set_option(self.ptr, name.encode(), float(value)))

    def ramo_current_emit(self, n_sec):
        out = np.zeros(n_sec)
        self._check(self.lib.rh_get_ramo_sections(self.ptr, int(n_sec), _d(out)))
        return out

    # -- samplers / quadrature --------------------------------------------------------------------------
    def Cuba_Integrate(self, kind):
        i, e, n, f = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        self._check(self.lib.rh_cuba_integrate(self.ptr, kind, C.byref(i), C.byref(e), C.byref(n), C.byref(f)))
        return i.value, e.value, n.value, f.value

    def Metropolis_Hastings_rectangle_J(self):
        df, F, pos = np.zeros(1), np.zeros(1), np.zeros(3)
        rc = self.lib.rh_mh_rectangle_J(self.ptr, _d(df), _d(F), _d(pos))
        if rc == -2:
            self._check(rc)
        return rc, df[0], F[0], pos

    def Metropolis_Hastings_rectangle_J_batch(self, M):
        df, F, pos = np.zeros(M), np.zeros(M), np.zeros((M, 3))
        self._check(self.lib.rh_mh_rectangle_J_batch(self.ptr, M, _d(df), _d(F), _d(pos)))
        return df, F, pos

    def Metropolis_Hastings_rectangle_J_thermo_batch(self, M):
        """mh_batch=2: M thermal-field chains in lock-step on the device.  Returns (pos, ok)."""
        pos, ok = np.zeros((M, 3)), np.zeros(M, dtype=np.int32)
        self._check(self.lib.rh_mh_rectangle_J_thermo_batch(self.ptr, M, _d(pos), ok.ctypes.data_as(_PI)))
        return pos, ok

    def Metropolis_Hastings_rectangle_J_thermo(self):
        pos = np.zeros(3)
        rc = self.lib.rh_mh_rectangle_J_thermo(self.ptr, _d(pos))
        if rc == -2:
            self._check(rc)
        return rc, pos

    def Metro_algo_tip_v3(self, ndim=80):
        xi, phi, eta_f, df, pos = np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(3)
        rc = self.lib.rh_metro_algo_tip_v3(self.ptr, ndim, _d(xi), _d(phi), _d(eta_f), _d(df), _d(pos))
        if rc == -2:
            self._check(rc)
        return rc, xi[0], phi[0], eta_f[0], df[0], pos

    def Metro_algo_tip_v3_batch(self, M, ndim=80):
        eta_f, df, pos = np.zeros(M), np.zeros(M), np.zeros((M, 3))
        self._check(self.lib.rh_metro_algo_tip_v3_batch(self.ptr, M, ndim, _d(eta_f), _d(df), _d(pos)))
        return eta_f, df, pos

    def Tip_Supply_Grid(self, nr_xi=100, nr_phi=100):
        n_s, fa = np.zeros(1), np.zeros(1)
        self._check(self.lib.rh_tip_supply_grid(self.ptr, nr_xi, nr_phi, _d(n_s), _d(fa)))
        return n_s[0], fa[0]

    def Do_Emission(self, step):
        n = C.c_int()
        self._check(self.lib.rh_do_emission(self.ptr, int(step), C.byref(n)))
        return n.value

    def w_theta_xy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        sec = C.c_int()
        return self.lib.rh_w_theta_xy(self.ptr, _d(pos), C.byref(sec)), sec.value


def kevin_jgtf_v2(F, T, w):
    return load_host_library().rh_kevin_jgtf_v2(F, T, w)
