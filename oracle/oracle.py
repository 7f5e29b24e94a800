"""ctypes binding of the CPU oracle (oracle/rumdeed_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / --impl reference legs.  Nothing under rumdeed_b200/
may import this module.

The oracle restates the reference Fortran (file:line cited per function in
rumdeed_oracle.c) and is pinned against the golden vectors of the reference's
own src/mod_tests.F90 in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SPECIES_ELEC, SPECIES_ION, SPECIES_ATOM = 1, 2, 3
REMOVE_TOP, REMOVE_BOT = 1, 2
GEOM_PLANAR, GEOM_TIP = 1, 2
PLANES_MAX = 10
MAX_LIFE_TIME = 1000


class Constants(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "pi", "h", "k_b", "c", "mu_0", "epsilon_0", "m_u", "h_bar", "m_0", "q_0",
        "m_N2", "m_N2p", "length_scale", "time_scale", "div_fac_c", "a_FN", "b_FN", "l_const")]


class Params(C.Structure):
    _fields_ = [
        ("geometry", C.c_int), ("image_charge", C.c_int), ("N_ic_max", C.c_int), ("planes_N", C.c_int),
        ("V_s", C.c_double), ("d", C.c_double), ("E_z", C.c_double),
        ("box_dim", C.c_double * 3), ("time_step", C.c_double),
        ("planes_z", C.c_double * PLANES_MAX),
        ("d_tip", C.c_double), ("R_base", C.c_double), ("h_tip", C.c_double),
        ("a_foci", C.c_double), ("eta_1", C.c_double), ("theta_tip", C.c_double), ("r_tip", C.c_double),
        ("max_xi", C.c_double), ("shift_z", C.c_double),
        ("pre_fac_E_tip", C.c_double), ("pre_fac_E_tip_unit_voltage", C.c_double),
    ]

    def set_planes(self, zs):
        zs = list(zs)
        self.planes_N = len(zs)
        for k, z in enumerate(zs):
            self.planes_z[k] = z


class Event(C.Structure):
    _fields_ = [("kind", C.c_int), ("plane", C.c_int), ("index", C.c_int),
                ("x", C.c_double), ("y", C.c_double),
                ("vx", C.c_double), ("vy", C.c_double), ("vz", C.c_double),
                ("emit", C.c_int), ("sec", C.c_int), ("id", C.c_int)]


_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)


class StoreStruct(C.Structure):
    _fields_ = [
        ("capacity", C.c_int),
        ("pos", _PD), ("prev_pos", _PD), ("vel", _PD), ("acc", _PD), ("acc_prev", _PD), ("acc_prev2", _PD),
        ("charge", _PD), ("mass", _PD),
        ("species", _PI), ("step", _PI), ("emitter", _PI), ("section", _PI), ("life", _PI), ("id", _PI), ("mask", _PI),
        ("nrPart", C.c_int), ("nrElec", C.c_int), ("nrIon", C.c_int), ("nrAtom", C.c_int), ("nrID", C.c_int),
        ("nrPart_dropped", C.c_int),
        ("nrPart_remove", C.c_int), ("nrElec_remove", C.c_int), ("nrIon_remove", C.c_int), ("nrAtom_remove", C.c_int),
        ("nrPart_remove_top", C.c_int), ("nrPart_remove_bot", C.c_int),
        ("nrElec_remove_top", C.c_int), ("nrElec_remove_bot", C.c_int),
        ("nrIon_remove_top", C.c_int), ("nrIon_remove_bot", C.c_int),
        ("charge_rev", C.c_int),
        ("life_time", (C.c_longlong * 4) * (MAX_LIFE_TIME + 1)),
        ("ramo_current", C.c_double * 4),
        ("avg_part_vel", C.c_double * 3), ("avg_elec_vel", C.c_double * 3), ("avg_ion_vel", C.c_double * 3),
        ("events", C.POINTER(Event)), ("n_events", C.c_int), ("cap_events", C.c_int),
        ("ramo_current_emit", _PD),
    ]


def build(force: bool = False, march: str | None = None, out_dir: str | None = None) -> None:
    """Compile liboracle.so / liboracle_fast.so with the committed Makefile."""
    out = out_dir or _HERE
    need = force or not (os.path.exists(os.path.join(out, "liboracle.so"))
                         and os.path.exists(os.path.join(out, "liboracle_fast.so")))
    if not need:
        src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in (
            "rumdeed_oracle.c", "rumdeed_oracle.h", "rumdeed_oracle_emission.c", "rumdeed_oracle_emission.h",
            "rumdeed_oracle_collisions.c", "rumdeed_oracle_collisions.h"))
        need = src_m > min(os.path.getmtime(os.path.join(out, f)) for f in ("liboracle.so", "liboracle_fast.so"))
    if need:
        cmd = ["make", "-C", _HERE, "-B", f"OUT={out}"]
        if march:
            cmd.append(f"MARCH={march}")
        subprocess.run(cmd, check=True, capture_output=True)


def _d(a):
    return a.ctypes.data_as(_PD)


def _i(a):
    return a.ctypes.data_as(_PI) if a is not None else None


def _f64(a, shape_last3=False):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Oracle:
    """Loaded oracle library.  `fast=True` loads the -O3 baseline build."""

    def __init__(self, fast: bool = False, path: str | None = None):
        if path is None:
            build()
            path = os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so")
        self.lib = lib = C.CDLL(path)
        PP = C.POINTER(Params)
        lib.orc_get_constants.argtypes = [C.POINTER(Constants)]
        lib.orc_params_planar.argtypes = [PP, C.c_double, C.c_double, _PD, C.c_double, C.c_int, C.c_int]
        lib.orc_params_tip.argtypes = [PP, C.c_double, C.c_double, C.c_double, C.c_double, _PD, C.c_double, C.c_int]
        for name in ("orc_force_image_charges_v2", "orc_sphere_ic_field", "orc_image_charge_effect"):
            getattr(lib, name).argtypes = [PP, _PD, _PD, _PD]
        for name in ("orc_field_E_planar", "orc_field_E_hyperboloid", "orc_field_E", "orc_E_zunit", "orc_surface_normal"):
            getattr(lib, name).argtypes = [PP, _PD, _PD]
        lib.orc_xi_coor.argtypes = [PP, C.c_double, C.c_double, C.c_double]; lib.orc_xi_coor.restype = C.c_double
        lib.orc_eta_coor.argtypes = [PP, C.c_double, C.c_double, C.c_double]; lib.orc_eta_coor.restype = C.c_double
        lib.orc_phi_coor.argtypes = [C.c_double, C.c_double]; lib.orc_phi_coor.restype = C.c_double
        lib.orc_xyz_corr.argtypes = [PP, C.c_double, C.c_double, C.c_double, _PD]
        lib.orc_field_normal.argtypes = [PP, _PD, _PD]; lib.orc_field_normal.restype = C.c_double
        lib.orc_tip_area.argtypes = [PP] + [C.c_double] * 4; lib.orc_tip_area.restype = C.c_double
        lib.orc_accel_generic.argtypes = [PP, C.c_int, _PD, _PD, _PD, _PI, _PD]
        lib.orc_accel_planar.argtypes = [PP, C.c_int, _PD, _PD, _PD, _PI, _PD]
        lib.orc_accel_planar_rows.argtypes = [PP, C.c_int, _PD, _PD, _PD, _PI, _PD, C.c_int, C.c_int, C.c_int]
        lib.orc_accel_planar_rows.restype = C.c_longlong
        lib.orc_accel_gather.argtypes = [PP, C.c_int, _PD, _PD, _PD, _PD]
        lib.orc_accel_gather_ld_rows.argtypes = [PP, C.c_int, _PD, _PD, _PD, C.c_int, _PI, _PD]
        lib.orc_accel_gather_ld.argtypes = [PP, C.c_int, _PD, _PD, _PD, C.c_int, C.c_int, _PD]
        lib.orc_calc_field_at.argtypes = [PP, C.c_int, _PD, _PD, _PI, _PD, _PD]
        lib.orc_calc_field_at_ld.argtypes = [PP, C.c_int, _PD, _PD, _PI, _PD, _PD]
        lib.orc_calc_field_at_batch.argtypes = [PP, C.c_int, _PD, _PD, _PI, C.c_int, _PD, _PD]
        PS = C.POINTER(StoreStruct)
        lib.orc_store_new.argtypes = [C.c_int]; lib.orc_store_new.restype = PS
        lib.orc_store_free.argtypes = [PS]
        lib.orc_store_clear_events.argtypes = [PS]
        lib.orc_add_particle.argtypes = [PS, PP, _PD, _PD, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_add_particle.restype = C.c_int
        lib.orc_mark_particle_remove.argtypes = [PS, C.c_int, C.c_int]
        lib.orc_remove_particles.argtypes = [PS, C.c_int]
        for name in ("orc_update_position", "orc_update_acceleration", "orc_update_velocity", "orc_step"):
            getattr(lib, name).argtypes = [PS, PP]
        for name in ("orc_fn_v_y", "orc_fn_t_y", "orc_fn_escape_prob_log", "orc_fn_elec_supply_log",
                     "orc_fn_elec_supply_v2", "orc_tip_v_y", "orc_tip_t_y", "orc_tip_escape_prob"):
            getattr(lib, name).argtypes = [PP, C.c_double, C.c_double]
            getattr(lib, name).restype = C.c_double
        lib.orc_tip_elec_supply.argtypes = [PP, C.c_double, C.c_double, C.c_double]
        lib.orc_tip_elec_supply.restype = C.c_double
        lib.orc_max_threads.restype = C.c_int
        lib.orc_nearest_elec.argtypes = [C.c_int, _PD, _PI, _PD, _PI]
        self.k = Constants()
        lib.orc_get_constants(C.byref(self.k))

    # -- parameters -------------------------------------------------------------
    def params_planar(self, V_s, d, box_dim, time_step, image_charge, N_ic_max) -> Params:
        p = Params()
        bd = (C.c_double * 3)(*box_dim)
        self.lib.orc_params_planar(C.byref(p), V_s, d, bd, time_step, int(bool(image_charge)), int(N_ic_max))
        return p

    def params_tip(self, V_s, d_tip, R_base, h_tip, box_dim, time_step, image_charge) -> Params:
        p = Params()
        bd = (C.c_double * 3)(*box_dim)
        self.lib.orc_params_tip(C.byref(p), V_s, d_tip, R_base, h_tip, bd, time_step, int(bool(image_charge)))
        return p

    # -- small vector helpers -----------------------------------------------------
    def _v3(self, fn, p, *vecs):
        out = np.zeros(3)
        args = [_d(np.ascontiguousarray(v, dtype=np.float64)) for v in vecs]
        # keep the temporaries alive for the duration of the call
        keep = [np.ascontiguousarray(v, dtype=np.float64) for v in vecs]
        args = [_d(v) for v in keep]
        fn(C.byref(p), *args, _d(out))
        return out

    def force_image_charges_v2(self, p, pos_1, pos_2):
        return self._v3(self.lib.orc_force_image_charges_v2, p, pos_1, pos_2)

    def sphere_ic_field(self, p, pos_1, pos_2):
        return self._v3(self.lib.orc_sphere_ic_field, p, pos_1, pos_2)

    def image_charge_effect(self, p, pos_1, pos_2):
        return self._v3(self.lib.orc_image_charge_effect, p, pos_1, pos_2)

    def field_E(self, p, pos):
        return self._v3(self.lib.orc_field_E, p, pos)

    def field_E_hyperboloid(self, p, pos):
        return self._v3(self.lib.orc_field_E_hyperboloid, p, pos)

    def E_zunit(self, p, pos):
        return self._v3(self.lib.orc_E_zunit, p, pos)

    def surface_normal(self, p, pos):
        return self._v3(self.lib.orc_surface_normal, p, pos)

    def xyz_corr(self, p, xi, eta, phi):
        out = np.zeros(3)
        self.lib.orc_xyz_corr(C.byref(p), xi, eta, phi, _d(out))
        return out

    # -- accelerations ----------------------------------------------------------
    @staticmethod
    def _prep(pos, q, m, species=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = pos.shape[0]
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(n)
        m = np.ascontiguousarray(m, dtype=np.float64).reshape(n) if m is not None else None
        sp = np.ascontiguousarray(species, dtype=np.int32).reshape(n) if species is not None else None
        return n, pos, q, m, sp

    def accel_generic(self, p, pos, q, m, species=None, acc0=None):
        n, pos, q, m, sp = self._prep(pos, q, m, species)
        acc = np.zeros((n, 3)) if acc0 is None else np.array(acc0, dtype=np.float64).reshape(n, 3).copy()
        self.lib.orc_accel_generic(C.byref(p), n, _d(pos), _d(q), _d(m), _i(sp), _d(acc))
        return acc

    def accel_planar(self, p, pos, q, m, species=None, acc0=None):
        n, pos, q, m, sp = self._prep(pos, q, m, species)
        acc = np.zeros((n, 3)) if acc0 is None else np.array(acc0, dtype=np.float64).reshape(n, 3).copy()
        self.lib.orc_accel_planar(C.byref(p), n, _d(pos), _d(q), _d(m), _i(sp), _d(acc))
        return acc

    def accel_planar_rows(self, p, pos, q, m, i0, i1, stride=1, species=None):
        n, pos, q, m, sp = self._prep(pos, q, m, species)
        acc = np.zeros((n, 3))
        pairs = self.lib.orc_accel_planar_rows(C.byref(p), n, _d(pos), _d(q), _d(m), _i(sp), _d(acc), i0, i1, stride)
        return acc, int(pairs)

    def accel_gather(self, p, pos, q, m):
        n, pos, q, m, _ = self._prep(pos, q, m)
        acc = np.zeros((n, 3))
        self.lib.orc_accel_gather(C.byref(p), n, _d(pos), _d(q), _d(m), _d(acc))
        return acc

    def accel_gather_ld(self, p, pos, q, m, i0=0, i1=None):
        n, pos, q, m, _ = self._prep(pos, q, m)
        i1 = n if i1 is None else i1
        acc = np.zeros((i1 - i0, 3))
        self.lib.orc_accel_gather_ld(C.byref(p), n, _d(pos), _d(q), _d(m), i0, i1, _d(acc))
        return acc

    def accel_gather_ld_rows(self, p, pos, q, m, rows):
        """Long-double truth for the listed rows (OpenMP over the list)."""
        n, pos, q, m, _ = self._prep(pos, q, m)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        acc = np.zeros((rows.size, 3))
        self.lib.orc_accel_gather_ld_rows(C.byref(p), n, _d(pos), _d(q), _d(m), rows.size, _i(rows), _d(acc))
        return acc

    # -- fields -------------------------------------------------------------------
    def calc_field_at(self, p, pos, q, pt, species=None, ld=False):
        n, pos, q, _, sp = self._prep(pos, q, None, species)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out = np.zeros(3)
        fn = self.lib.orc_calc_field_at_ld if ld else self.lib.orc_calc_field_at
        fn(C.byref(p), n, _d(pos), _d(q), _i(sp), _d(pt), _d(out))
        return out

    def calc_field_at_batch(self, p, pos, q, pts, species=None):
        n, pos, q, _, sp = self._prep(pos, q, None, species)
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        out = np.zeros_like(pts)
        self.lib.orc_calc_field_at_batch(C.byref(p), n, _d(pos), _d(q), _i(sp), pts.shape[0], _d(pts), _d(out))
        return out

    def nearest_elec(self, pos, species):
        """Sample_Elec_Position (mod_pair.F90:975-1037): nearest other electron of every electron."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        sp = np.ascontiguousarray(species, dtype=np.int32)
        n = pos.shape[0]
        dist = np.empty(n)
        idx = np.empty(n, dtype=np.int32)
        self.lib.orc_nearest_elec(n, _d(pos), _i(sp), _d(dist), _i(idx))
        return dist, idx

    # -- FN helpers ---------------------------------------------------------------
    def fn_v_y(self, p, F, w): return self.lib.orc_fn_v_y(C.byref(p), F, w)
    def fn_t_y(self, p, F, w): return self.lib.orc_fn_t_y(C.byref(p), F, w)
    def fn_escape_prob_log(self, p, F, w): return self.lib.orc_fn_escape_prob_log(C.byref(p), F, w)
    def fn_elec_supply_log(self, p, F, w): return self.lib.orc_fn_elec_supply_log(C.byref(p), F, w)
    def fn_elec_supply_v2(self, p, F, w): return self.lib.orc_fn_elec_supply_v2(C.byref(p), F, w)
    def tip_escape_prob(self, p, F, w): return self.lib.orc_tip_escape_prob(C.byref(p), F, w)
    def tip_elec_supply(self, p, A, F, w): return self.lib.orc_tip_elec_supply(C.byref(p), A, F, w)

    def set_threads(self, n: int) -> None:
        self.lib.orc_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    def store(self, capacity: int) -> "Store":
        return Store(self, capacity)


class Store:
    """The reference's particle arrays + bookkeeping (mod_global / mod_pair)."""

    def __init__(self, orc: Oracle, capacity: int):
        self.orc = orc
        self.ptr = orc.lib.orc_store_new(capacity)
        self.s = self.ptr.contents
        self.capacity = capacity

    def __del__(self):
        try:
            self.orc.lib.orc_store_free(self.ptr)
        except Exception:
            pass

    def add(self, p, pos, vel, species, step, emit, life=-1, sec=1):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        pp = C.byref(p) if p is not None else None
        return self.orc.lib.orc_add_particle(self.ptr, pp, _d(pos), _d(vel), species, step, emit, life, sec)

    def mark(self, i, reason): self.orc.lib.orc_mark_particle_remove(self.ptr, i, reason)
    def remove(self, step): self.orc.lib.orc_remove_particles(self.ptr, step)
    def update_position(self, p): self.orc.lib.orc_update_position(self.ptr, C.byref(p))
    def update_acceleration(self, p): self.orc.lib.orc_update_acceleration(self.ptr, C.byref(p))
    def update_velocity(self, p): self.orc.lib.orc_update_velocity(self.ptr, C.byref(p))
    def step(self, p): self.orc.lib.orc_step(self.ptr, C.byref(p))
    def clear_events(self): self.orc.lib.orc_store_clear_events(self.ptr)

    def _arr(self, ptr, n, cols=None, dtype=np.float64):
        cnt = n * (cols or 1)
        if cnt == 0:
            return np.zeros((0, cols) if cols else (0,), dtype=dtype)
        a = np.ctypeslib.as_array(ptr, shape=(cnt,))
        return a.reshape(n, cols) if cols else a

    @property
    def n(self): return self.s.nrPart
    @property
    def pos(self): return self._arr(self.s.pos, self.n, 3)
    @property
    def prev_pos(self): return self._arr(self.s.prev_pos, self.n, 3)
    @property
    def vel(self): return self._arr(self.s.vel, self.n, 3)
    @property
    def acc(self): return self._arr(self.s.acc, self.n, 3)
    @property
    def acc_prev(self): return self._arr(self.s.acc_prev, self.n, 3)
    @property
    def acc_prev2(self): return self._arr(self.s.acc_prev2, self.n, 3)
    @property
    def charge(self): return self._arr(self.s.charge, self.n)
    @property
    def mass(self): return self._arr(self.s.mass, self.n)
    @property
    def species(self): return self._arr(self.s.species, self.n, dtype=np.int32)
    @property
    def step_born(self): return self._arr(self.s.step, self.n, dtype=np.int32)
    @property
    def emitter(self): return self._arr(self.s.emitter, self.n, dtype=np.int32)
    @property
    def section(self): return self._arr(self.s.section, self.n, dtype=np.int32)
    @property
    def life(self): return self._arr(self.s.life, self.n, dtype=np.int32)
    @property
    def ids(self): return self._arr(self.s.id, self.n, dtype=np.int32)

    def mask(self, n=None):
        return self._arr(self.s.mask, self.n if n is None else n, dtype=np.int32)

    def events(self):
        out = []
        for k in range(self.s.n_events):
            e = self.s.events[k]
            out.append(dict(kind=e.kind, plane=e.plane, index=e.index, x=e.x, y=e.y,
                            vx=e.vx, vy=e.vy, vz=e.vz, emit=e.emit, sec=e.sec, id=e.id))
        return out

    def life_time(self, lt, species):
        return int(self.s.life_time[lt][species])

    def ramo_current_emit(self, n_sec=96 * 96):
        """ramo_current_emit(1:n_sec, 1) of the last velocity update (src/mod_verlet.F90:489-492)."""
        return np.ctypeslib.as_array(self.s.ramo_current_emit, shape=(96 * 96,))[:n_sec].copy()


# ------------------------------------------------------------------------------------------------------
# emission samplers (oracle/rumdeed_oracle_emission.c)
SUPPLY_FE, SUPPLY_GTF = 1, 2


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class EmissionStruct(C.Structure):
    _fields_ = [("p", C.POINTER(Params)), ("store", C.POINTER(StoreStruct)),
                ("emit_pos", C.c_double * 3), ("emit_dim", C.c_double * 3),
                ("y_num", C.c_int), ("x_num", C.c_int), ("w_theta_arr", _PD),
                ("T_temp", C.c_double), ("a_rate", C.c_double), ("MH_std", C.c_double),
                ("MH_std_tip", C.c_double), ("residual", C.c_double)]


class Emission:
    """One rectangular emitter with a checkerboard work function on top of an oracle Store."""

    def __init__(self, orc: Oracle, p: Params, store: Store, emit_pos, emit_dim, w_theta=((2.0,),), T_temp=293.15, seed=1):
        self.orc, self.p, self.store = orc, p, store
        lib = orc.lib
        self.w = np.ascontiguousarray(np.atleast_2d(np.asarray(w_theta, dtype=np.float64)))
        e = self.e = EmissionStruct()
        e.p = C.pointer(p)
        e.store = store.ptr
        e.emit_pos[:] = list(emit_pos)
        e.emit_dim[:] = list(emit_dim)
        e.y_num, e.x_num = self.w.shape
        e.w_theta_arr = _d(self.w)
        e.T_temp = T_temp
        e.a_rate, e.MH_std, e.MH_std_tip, e.residual = 1.0, 0.0125, 1.0, 0.0
        self.rng = Rng()
        PE, PR = C.POINTER(EmissionStruct), C.POINTER(Rng)
        if not getattr(lib, "_emission_bound", False):
            lib.orc_rng_seed.argtypes = [PR, C.c_uint64]
            lib.orc_rng_uniform.argtypes = [PR]; lib.orc_rng_uniform.restype = C.c_double
            lib.orc_box_muller.argtypes = [PR, _PD, _PD, _PD]
            lib.orc_rand_poisson.argtypes = [PR, C.c_double]; lib.orc_rand_poisson.restype = C.c_int
            lib.orc_get_mb_velocity.argtypes = [PR, C.c_double, _PD]
            lib.orc_w_theta_xy.argtypes = [PE, _PD, _PI]; lib.orc_w_theta_xy.restype = C.c_double
            lib.orc_kevin_jgtf_v2.argtypes = [C.c_double] * 3; lib.orc_kevin_jgtf_v2.restype = C.c_double
            lib.orc_supply_integrand.argtypes = [PE, C.c_int, _PD, _PD]; lib.orc_supply_integrand.restype = C.c_double
            lib.orc_supply_grid.argtypes = [PE, C.c_int, C.c_int, _PD]; lib.orc_supply_grid.restype = C.c_double
            lib.orc_mh_rectangle_J.argtypes = [PE, PR, _PD, _PD, _PD]; lib.orc_mh_rectangle_J.restype = C.c_int
            lib.orc_mh_rectangle_J_batch.argtypes = [PE, PR, C.c_int, _PD, _PD, _PD]
            lib.orc_do_field_emission_planar.argtypes = [PE, PR, C.c_int, C.c_double, C.c_int, _PD]
            lib.orc_do_field_emission_planar.restype = C.c_int
            lib.orc_mh_rectangle_J_thermo.argtypes = [PE, PR, _PD]; lib.orc_mh_rectangle_J_thermo.restype = C.c_int
            lib.orc_do_field_thermo_emission_planar.argtypes = [PE, PR, C.c_int, C.c_double]
            lib.orc_do_field_thermo_emission_planar.restype = C.c_int
            lib.orc_do_photo_emission_rectangle.argtypes = [PE, PR, C.c_int, C.c_double, C.c_int, C.c_int]
            lib.orc_do_photo_emission_rectangle.restype = C.c_int
            lib.orc_get_laser_energy.argtypes = [PR, C.c_double, C.c_double]; lib.orc_get_laser_energy.restype = C.c_double
            lib.orc_tip_supply_grid.argtypes = [PE, C.c_int, C.c_int, _PD]; lib.orc_tip_supply_grid.restype = C.c_double
            lib.orc_metro_algo_tip_v3.argtypes = [PE, PR, C.c_int, _PD, _PD, _PD, _PD, _PD]
            lib.orc_metro_algo_tip_v3.restype = C.c_int
            lib.orc_do_field_emission_tip.argtypes = [PE, PR, C.c_int, C.c_double]; lib.orc_do_field_emission_tip.restype = C.c_int
            lib._emission_bound = True
        lib.orc_rng_seed(C.byref(self.rng), seed)

    def _E(self):
        return C.byref(self.e)

    def _R(self):
        return C.byref(self.rng)

    def uniform(self):
        return self.orc.lib.orc_rng_uniform(self._R())

    def w_theta_xy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        sec = C.c_int(0)
        w = self.orc.lib.orc_w_theta_xy(self._E(), _d(pos), C.byref(sec))
        return w, sec.value

    def kevin_jgtf_v2(self, F, T, w):
        return self.orc.lib.orc_kevin_jgtf_v2(F, T, w)

    def supply_integrand(self, kind, xx):
        xx = np.ascontiguousarray(xx, dtype=np.float64)
        f = np.zeros(3)
        return self.orc.lib.orc_supply_integrand(self._E(), kind, _d(xx), _d(f)), f

    def supply_grid(self, kind, n):
        f = np.zeros(3)
        return self.orc.lib.orc_supply_grid(self._E(), kind, n, _d(f)), f

    def mh_rectangle_J(self):
        df, F, pos = np.zeros(1), np.zeros(1), np.zeros(3)
        rc = self.orc.lib.orc_mh_rectangle_J(self._E(), self._R(), _d(df), _d(F), _d(pos))
        return rc, df[0], F[0], pos

    def mh_rectangle_J_batch(self, M):
        df, F, pos = np.zeros(M), np.zeros(M), np.zeros((M, 3))
        self.orc.lib.orc_mh_rectangle_J_batch(self._E(), self._R(), M, _d(df), _d(F), _d(pos))
        return df, F, pos

    def do_field_emission_planar(self, step, N_sup, mh_batch=False):
        dfa = np.zeros(1)
        n = self.orc.lib.orc_do_field_emission_planar(self._E(), self._R(), step, N_sup, int(mh_batch), _d(dfa))
        return n, dfa[0]

    def mh_rectangle_J_thermo(self):
        pos = np.zeros(3)
        rc = self.orc.lib.orc_mh_rectangle_J_thermo(self._E(), self._R(), _d(pos))
        return rc, pos

    def do_field_thermo_emission_planar(self, step, N_sup):
        return self.orc.lib.orc_do_field_thermo_emission_planar(self._E(), self._R(), step, N_sup)

    def do_photo_emission_rectangle(self, step, p_eV, photon_mode=1, max_elec_emit=-1):
        return self.orc.lib.orc_do_photo_emission_rectangle(self._E(), self._R(), step, p_eV, photon_mode, max_elec_emit)

    def get_laser_energy(self, laser_energy, laser_variation):
        return self.orc.lib.orc_get_laser_energy(self._R(), laser_energy, laser_variation)

    def tip_supply_grid(self, nr_xi=100, nr_phi=100):
        fa = np.zeros(1)
        return self.orc.lib.orc_tip_supply_grid(self._E(), nr_xi, nr_phi, _d(fa)), fa[0]

    def metro_algo_tip_v3(self, ndim=80):
        xi, phi, eta_f, df, pos = np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(3)
        rc = self.orc.lib.orc_metro_algo_tip_v3(self._E(), self._R(), ndim, _d(xi), _d(phi), _d(eta_f), _d(df), _d(pos))
        return rc, xi[0], phi[0], eta_f[0], df[0], pos

    def do_field_emission_tip(self, step, n_s):
        return self.orc.lib.orc_do_field_emission_tip(self._E(), self._R(), step, n_s)
