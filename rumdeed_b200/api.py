"""ctypes binding of librumdeed_b200.so and a host-side mirror of the reference interface.

`HotPath` exposes the reference's own entry points for the per-timestep hot path under
their Fortran names (Add_Particle, Mark_Particles_Remove, Remove_Particles,
Update_Position, Calculate_Acceleration_Particles, Calc_Field_at, Calc_Field_at_Batch,
Particles_To_Device, Release_Device_Particles) so that the parity tests read like the
reference's src/mod_tests.F90.  Everything runs through the C ABI declared in
include/rumdeed_b200.h; there is no CPU fallback: if the CUDA library is missing or no
sm_100 device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RB2_LIB_PATH") or os.path.join(_HERE, "librumdeed_b200.so")  # RB2_LIB_PATH: kernel-variant builds (tools/build_variants.sh)

GEOM_PLANAR, GEOM_TIP = 1, 2
SPECIES_ELEC, SPECIES_ION, SPECIES_ATOM = 1, 2, 3
REMOVE_TOP, REMOVE_BOT, REMOVE_RECOM, REMOVE_ION = 1, 2, 3, 4
PLANES_MAX = 10
MAX_LIFE_TIME = 1000

# physical constants of the reference (src/mod_global.F90:26-75); epsilon_0 is derived
PI = 3.141592653589793238462643383279502884197169399375105820974944592307816406286
MU_0 = 1.25663706212e-6
C_LIGHT = 299792458.0
EPSILON_0 = 1.0 / (MU_0 * C_LIGHT ** 2)
Q_0 = 1.602176634e-19
M_0 = 9.1093837015e-31
M_U = 1.66053906660e-27
M_N2 = 28.0134 * M_U
M_N2P = M_N2 - M_0
LENGTH_SCALE = 1.0e-9
TIME_SCALE = 1.0e-12
DIV_FAC_C = 1.0 / (4.0 * PI * EPSILON_0 * 1.0)


class Rb2Error(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("geometry", C.c_int), ("image_charge", C.c_int), ("N_ic_max", C.c_int), ("planes_N", C.c_int),
        ("V_s", C.c_double), ("d", C.c_double), ("E_z", C.c_double),
        ("box_dim", C.c_double * 3), ("time_step", C.c_double),
        ("planes_z", C.c_double * PLANES_MAX),
        ("a_foci", C.c_double), ("eta_1", C.c_double), ("shift_z", C.c_double),
        ("pre_fac_E_tip", C.c_double), ("pre_fac_E_tip_unit_voltage", C.c_double),
        ("h_tip", C.c_double), ("r_tip", C.c_double), ("max_xi", C.c_double),
        ("capacity", C.c_int), ("device", C.c_int),
    ]


class Counts(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nrPart", "nrElec", "nrIon", "nrAtom", "nrID", "nrPart_dropped",
        "nrPart_remove", "nrElec_remove", "nrIon_remove", "nrAtom_remove",
        "nrPart_remove_top", "nrPart_remove_bot", "nrElec_remove_top", "nrElec_remove_bot",
        "nrIon_remove_top", "nrIon_remove_bot", "nrPart_remove_ion", "nrElec_remove_ion", "nrAtom_remove_ion")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Event(C.Structure):
    _fields_ = [("kind", C.c_int), ("plane", C.c_int), ("index", C.c_int),
                ("x", C.c_double), ("y", C.c_double),
                ("vx", C.c_double), ("vy", C.c_double), ("vz", C.c_double),
                ("emit", C.c_int), ("sec", C.c_int), ("id", C.c_int)]


class StepResult(C.Structure):
    _fields_ = [("ramo_current", C.c_double * 4),
                ("avg_part_vel", C.c_double * 3), ("avg_elec_vel", C.c_double * 3), ("avg_ion_vel", C.c_double * 3),
                ("n_events", C.c_int), ("counts", Counts),
                ("accel_ms", C.c_float), ("step_ms", C.c_float)]


class MhConfig(C.Structure):
    """rb2_mh_config."""
    _fields_ = [("kind", C.c_int), ("ndim", C.c_int), ("ndim_first", C.c_int), ("image_charge", C.c_int),
                ("y_num", C.c_int), ("x_num", C.c_int),
                ("emit_pos", C.c_double * 2), ("emit_dim", C.c_double * 2), ("T_temp", C.c_double),
                ("init_std", C.c_double), ("target_rate", C.c_double), ("std_gain", C.c_double),
                ("std_min", C.c_double), ("std_max", C.c_double)]


class CollisionConfig(C.Structure):
    """rb2_collision_config."""
    _fields_ = [("collision_mode", C.c_int), ("ion_life_time", C.c_int), ("n_d", C.c_double), ("cyl_radius", C.c_double),
                ("n_tot", C.c_int), ("n_ion", C.c_int),
                ("tot_energy", C.POINTER(C.c_double)), ("tot_data", C.POINTER(C.c_double)),
                ("ion_energy", C.POINTER(C.c_double)), ("ion_data", C.POINTER(C.c_double))]


class RecombRecord(C.Structure):
    """rb2_recomb_record."""
    _fields_ = [(n, C.c_int) for n in ("step", "elec_slot", "ion_slot", "elec_emit", "ion_life",
                                      "elec_sec", "elec_id", "ion_emit", "ion_sec", "ion_id")] + \
               [("ion_pos", C.c_double * 3), ("elec_pos", C.c_double * 3),
                ("elec_speed", C.c_double), ("dist", C.c_double), ("recom_rad", C.c_double), ("t", C.c_double)]


class IonizationRecord(C.Structure):
    """rb2_ionization_record."""
    _fields_ = [(n, C.c_int) for n in ("step", "in_slot", "new_id", "ion_id", "elec_emit", "pad")] + \
               [("pos", C.c_double * 3), ("in_speed", C.c_double), ("out_speed", C.c_double), ("new_speed", C.c_double),
                ("new_vel", C.c_double * 3), ("ejec_pos", C.c_double * 3), ("ejec_vel", C.c_double * 3),
                ("ion_pos", C.c_double * 3), ("E1", C.c_double), ("collE", C.c_double), ("ejecE", C.c_double)]


class CollisionResult(C.Structure):
    """rb2_collision_result."""
    _fields_ = [(n, C.c_int) for n in ("nrCollisions", "nrIonizations", "nrRecombinations", "nrIonsExpired",
                                      "nrPart_remove_recom", "nrElec_remove_recom", "nrIon_remove_recom",
                                      "n_candidates")] + [("counts", Counts), ("ms", C.c_float)]


P2P_HANDLE_BYTES = 64

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)

# every symbol include/rumdeed_b200.h declares (checked by tests/test_abi_symbols.py)
EXPORTS = (
    "rb2_init", "rb2_finalize", "rb2_update_config", "rb2_last_error_string", "rb2_device_available",
    "rb2_upload_particles", "rb2_download_particles", "rb2_get_counts",
    "rb2_add_particles", "rb2_capacity_left", "rb2_mark_remove", "rb2_remove_marked", "rb2_get_life_time",
    "rb2_step", "rb2_update_position", "rb2_accel_only", "rb2_update_velocity", "rb2_get_events", "rb2_get_ramo_sections", "rb2_accel_host",
    "rb2_field_batch", "rb2_field_batch_delta", "rb2_field_window_open", "rb2_field_window_close", "rb2_field_surface_z", "rb2_mh_planar", "rb2_mh_planar_serial", "rb2_mh_tip", "rb2_tip_supply_set_grid", "rb2_tip_supply", "rb2_planar_supply_level", "rb2_sym_plan_probe",
    "rb2_set_partition", "rb2_set_pair_rank", "rb2_accel_partial", "rb2_accel_finalize", "rb2_set_option", "rb2_device_buffer", "rb2_synchronize", "rb2_stream",
    "rb2_p2p_export", "rb2_p2p_attach", "rb2_p2p_detach", "rb2_set_devices", "rb2_nearest_electron",
    "rb2_collisions_init", "rb2_collision_data", "rb2_continuous_ionization", "rb2_discrete_recombination",
    "rb2_do_collisions", "rb2_get_recombination_records", "rb2_get_ionization_records", "rb2_probe_quartic_roots",
    "rb2_fp64_peak", "rb2_launch_count", "rb2_get_stat", "rb2_last_accel_info",
)

_lib = None


def load_library(path: str | None = None):
    """Load librumdeed_b200.so and declare the prototypes.  Fails loudly when absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise Rb2Error(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(p)
    PC = C.POINTER(Config)
    lib.rb2_init.argtypes = [PC]
    lib.rb2_update_config.argtypes = [PC]
    lib.rb2_last_error_string.restype = C.c_char_p
    lib.rb2_upload_particles.argtypes = [C.c_int] + [_PD] * 8 + [_PI] * 6 + [C.c_int]
    lib.rb2_download_particles.argtypes = [_PD] * 8 + [_PI] * 7
    lib.rb2_get_counts.argtypes = [C.POINTER(Counts)]
    lib.rb2_add_particles.argtypes = [C.c_int, _PD, _PD, _PI, C.c_int, _PI, _PI, _PI]
    lib.rb2_capacity_left.argtypes = [_PI]
    lib.rb2_set_devices.argtypes = [C.c_int, _PI]
    lib.rb2_get_stat.argtypes = [C.c_char_p, _PD]
    lib.rb2_mark_remove.argtypes = [C.c_int, _PI, _PI]
    lib.rb2_remove_marked.argtypes = [C.c_int, C.POINTER(Counts)]
    lib.rb2_get_life_time.argtypes = [C.POINTER(C.c_longlong)]
    lib.rb2_step.argtypes = [C.c_int, C.POINTER(StepResult)]
    lib.rb2_update_position.argtypes = [C.c_int]
    lib.rb2_update_velocity.argtypes = [C.POINTER(StepResult)]
    lib.rb2_get_events.argtypes = [C.c_int, C.POINTER(Event), _PI]
    lib.rb2_get_ramo_sections.argtypes = [C.c_int, C.c_int, _PD]
    lib.rb2_accel_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rb2_field_batch.argtypes = [C.c_int, _PD, _PD]
    lib.rb2_field_batch_delta.argtypes = [C.c_int, _PD, C.c_int, _PD, _PD, _PD]
    lib.rb2_field_surface_z.argtypes = [C.c_int, _PD, _PD]
    lib.rb2_mh_planar.argtypes = [C.POINTER(MhConfig), _PD, C.c_int, C.c_ulonglong, _PD, _PD, _PD, _PD, _PD]
    lib.rb2_mh_planar_serial.argtypes = [C.POINTER(MhConfig), _PD, C.c_int, C.c_ulonglong, _PD, _PD, _PD, _PI, _PD, _PD]
    lib.rb2_mh_tip.argtypes = [C.c_int, C.c_int, C.c_ulonglong, _PD, _PD, _PD, _PD, _PD]
    lib.rb2_sym_plan_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                       C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_ulonglong)]
    lib.rb2_tip_supply_set_grid.argtypes = [C.c_int, _PD, _PD, _PD]
    lib.rb2_tip_supply.argtypes = [_PD, _PD]
    lib.rb2_planar_supply_level.argtypes = [C.POINTER(MhConfig), _PD, C.c_int, C.c_int, _PD, C.c_int, C.c_int, _PD, _PD]
    lib.rb2_set_partition.argtypes = [C.c_int, C.c_int]
    lib.rb2_set_pair_rank.argtypes = [C.c_int, C.c_int]
    lib.rb2_set_option.argtypes = [C.c_char_p, C.c_double]
    lib.rb2_device_buffer.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.rb2_stream.argtypes = [C.POINTER(C.c_void_p)]
    lib.rb2_p2p_export.argtypes = [C.c_int, C.c_void_p]
    lib.rb2_nearest_electron.argtypes = [_PD, _PI]
    lib.rb2_p2p_attach.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.rb2_collisions_init.argtypes = [C.POINTER(CollisionConfig)]
    lib.rb2_collision_data.argtypes = [_PD]
    lib.rb2_continuous_ionization.argtypes = [C.c_int, C.c_ulonglong, C.POINTER(CollisionResult)]
    lib.rb2_discrete_recombination.argtypes = [C.c_int, C.POINTER(CollisionResult)]
    lib.rb2_do_collisions.argtypes = [C.c_int, C.c_ulonglong, C.POINTER(CollisionResult)]
    lib.rb2_get_recombination_records.argtypes = [C.c_int, C.POINTER(RecombRecord), _PI]
    lib.rb2_get_ionization_records.argtypes = [C.c_int, C.POINTER(IonizationRecord), _PI]
    lib.rb2_probe_quartic_roots.argtypes = [C.c_int, _PD, _PI, _PD]
    lib.rb2_fp64_peak.argtypes = [C.c_double, _PD, C.POINTER(C.c_float)]
    lib.rb2_launch_count.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    lib.rb2_last_accel_info.argtypes = [C.POINTER(C.c_float)] + [_PI] * 4
    for name in EXPORTS:
        if name != "rb2_last_error_string":
            getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_PD) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_PI) if a is not None else None


def planar_config(V_s, d, box_dim, time_step, image_charge=True, N_ic_max=1, capacity=1 << 20, planes_z=(), device=-1):
    """Config for the planar diode: what Init_Field_Emission_v2 / Init leave in mod_global."""
    c = Config()
    c.geometry = GEOM_PLANAR
    c.image_charge = int(bool(image_charge))
    c.N_ic_max = int(N_ic_max)
    c.V_s = V_s
    c.d = d
    c.E_z = -1.0 * V_s / d  # Set_Voltage, src/mod_verlet.F90:2050-2052
    c.box_dim[:] = box_dim
    c.time_step = time_step
    planes_z = list(planes_z)
    c.planes_N = len(planes_z)
    for k, z in enumerate(planes_z):
        c.planes_z[k] = z
    c.capacity = int(capacity)
    c.device = device
    return c


def tip_config(V_s, d_tip, R_base, h_tip, box_dim, time_step, image_charge=True, capacity=1 << 16, planes_z=(), device=-1):
    """Config for the hyperboloid tip: Init_Emission_Tip, src/mod_emission_tip.f90:105-125."""
    c = Config()
    c.geometry = GEOM_TIP
    c.image_charge = int(bool(image_charge))
    c.N_ic_max = 0
    c.V_s = V_s
    c.d = d_tip + h_tip
    c.E_z = -1.0 * V_s / c.d
    c.box_dim[:] = box_dim
    c.time_step = time_step
    eta_2 = 0.0
    c.max_xi = h_tip / d_tip + 1.0
    c.a_foci = math.sqrt(d_tip ** 2 * R_base ** 2 / (h_tip ** 2 + 2 * d_tip * h_tip) + d_tip ** 2)
    c.eta_1 = -1.0 * d_tip / c.a_foci
    theta = math.acos(d_tip / c.a_foci)
    c.r_tip = c.a_foci * math.sin(theta) * math.tan(theta)
    c.shift_z = abs(c.a_foci * c.eta_1 * c.max_xi)
    lg = math.log((1.0 + c.eta_1) / (1.0 - c.eta_1) * (1.0 - eta_2) / (1.0 + eta_2))
    c.pre_fac_E_tip_unit_voltage = 2.0 * 1.0 / (c.a_foci * lg)
    c.pre_fac_E_tip = 2.0 * V_s / (c.a_foci * lg)
    c.h_tip = h_tip
    planes_z = list(planes_z)
    c.planes_N = len(planes_z)
    for k, z in enumerate(planes_z):
        c.planes_z[k] = z
    c.capacity = int(capacity)
    c.device = device
    return c


class HotPath:
    """The device-resident hot path behind the reference's own procedure names."""

    def __init__(self, config: Config, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self.cfg = config
        self._check(self.lib.rb2_init(C.byref(config)))
        self.closed = False

    @classmethod
    def attach(cls):
        """View on the store another host (e.g. host_api.Simulation) has already initialised."""
        obj = cls.__new__(cls)
        obj.lib = load_library()
        obj.cfg = None
        obj.closed = True  # never finalises somebody else's state
        return obj

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise Rb2Error(f"rb2 error {rc}: {self.lib.rb2_last_error_string().decode()}")

    def close(self):
        if not self.closed:
            self.lib.rb2_finalize()
            self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def update_config(self, config: Config):
        self.cfg = config
        self._check(self.lib.rb2_update_config(C.byref(config)))

    # -- particle state ----------------------------------------------------------------------------
    def upload(self, pos, charge, mass, vel=None, acc=None, acc_prev=None, acc_prev2=None, prev_pos=None,
               species=None, step=None, emitter=None, section=None, life=None, ids=None, nrID=-1):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        g = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        pos = f(pos).reshape(-1, 3)
        n = pos.shape[0]
        keep = [pos, f(prev_pos), f(vel), f(acc), f(acc_prev), f(acc_prev2), f(charge), f(mass)]
        ints = [g(species), g(step), g(emitter), g(section), g(life), g(ids)]
        self._check(self.lib.rb2_upload_particles(n, *[_d(a) for a in keep], *[_i(a) for a in ints], nrID))
        return n

    def download(self, what=("pos", "vel", "acc", "charge")):
        n = self.counts().nrPart
        names = ("pos", "prev_pos", "vel", "acc", "acc_prev", "acc_prev2", "charge", "mass")
        inames = ("species", "step", "emitter", "section", "life", "id", "mask")
        out = {}
        dargs, iargs = [], []
        for nm in names:
            if nm in what:
                out[nm] = np.zeros((n, 3) if nm not in ("charge", "mass") else (n,))
                dargs.append(_d(out[nm]))
            else:
                dargs.append(None)
        for nm in inames:
            if nm in what:
                out[nm] = np.zeros(n, dtype=np.int32)
                iargs.append(_i(out[nm]))
            else:
                iargs.append(None)
        self._check(self.lib.rb2_download_particles(*dargs, *iargs))
        return out

    def counts(self) -> Counts:
        k = Counts()
        self._check(self.lib.rb2_get_counts(C.byref(k)))
        return k

    # -- mod_pair ---------------------------------------------------------------------------------
    def Add_Particle(self, par_pos, par_vel, par_species, step, emit, life=-1, opt_sec=1):
        """src/mod_pair.F90:29 (one particle)."""
        self.Add_Particles([par_pos], [par_vel], [par_species], step, [emit], [opt_sec], [life])

    def Add_Particles(self, pos, vel, species, step, emit=None, sec=None, life=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        k = pos.shape[0]
        vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(k, 3)
        species = np.ascontiguousarray(species, dtype=np.int32).reshape(k)
        g = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32).reshape(k)
        emit, sec, life = g(emit), g(sec), g(life)
        self._check(self.lib.rb2_add_particles(k, _d(pos), _d(vel), _i(species), int(step), _i(emit), _i(sec), _i(life)))

    def Mark_Particles_Remove(self, i, m):
        """src/mod_pair.F90:169; i is 0-based here."""
        idx = np.atleast_1d(np.asarray(i, dtype=np.int32))
        rs = np.broadcast_to(np.asarray(m, dtype=np.int32), idx.shape).copy()
        self._check(self.lib.rb2_mark_remove(idx.size, _i(idx), _i(rs)))

    def Remove_Particles(self, step) -> Counts:
        """src/mod_pair.F90:352."""
        k = Counts()
        self._check(self.lib.rb2_remove_marked(int(step), C.byref(k)))
        return k

    def life_time(self):
        out = np.zeros((MAX_LIFE_TIME + 1, 4), dtype=np.int64)
        self._check(self.lib.rb2_get_life_time(out.ctypes.data_as(C.POINTER(C.c_longlong))))
        return out

    # -- mod_verlet ------------------------------------------------------------------------------
    def Update_Position(self, step) -> StepResult:
        """src/mod_verlet.F90:115 -> Velocity_Verlet: position, acceleration, velocity."""
        r = StepResult()
        self._check(self.lib.rb2_step(int(step), C.byref(r)))
        return r

    def Update_Particle_Position(self, step):
        self._check(self.lib.rb2_update_position(int(step)))

    def Calculate_Acceleration_Particles(self):
        """src/mod_verlet.F90:597 (overwrites particles_cur_accel like the OpenACC path)."""
        self._check(self.lib.rb2_accel_only())

    def Update_Particle_Velocity(self) -> StepResult:
        r = StepResult()
        self._check(self.lib.rb2_update_velocity(C.byref(r)))
        return r

    def events(self):
        n = C.c_int(0)
        self._check(self.lib.rb2_get_events(0, None, C.byref(n)))
        if n.value == 0:
            return []
        buf = (Event * n.value)()
        self._check(self.lib.rb2_get_events(n.value, buf, C.byref(n)))
        return [dict(kind=e.kind, plane=e.plane, index=e.index, x=e.x, y=e.y, vx=e.vx, vy=e.vy, vz=e.vz,
                     emit=e.emit, sec=e.sec, id=e.id) for e in buf]

    def ramo_current_emit(self, n_sec, n_emit=1):
        """ramo_current_emit(1:n_sec, 1:n_emit) of the last velocity update (src/mod_verlet.F90:489-492);
        needs set_option("ramo_sections", S) beforehand."""
        out = np.zeros((n_emit, n_sec))
        self._check(self.lib.rb2_get_ramo_sections(int(n_sec), int(n_emit), _d(out)))
        return out[0] if n_emit == 1 else out

    def accel_host(self, pos, charge, mass, out=None):
        """Stateless host-buffer form (upload, kernel, copy-out), src/mod_verlet.F90:1254-1340."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = pos.shape[0]
        charge = np.ascontiguousarray(charge, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        if out is None:
            out = np.zeros((n, 3))
        self._check(self.lib.rb2_accel_host(n, pos.ctypes.data, charge.ctypes.data, mass.ctypes.data, out.ctypes.data))
        return out

    def accel_host_ptr(self, n, pos_ptr, charge_ptr, mass_ptr, out_ptr):
        """Same with raw host pointers (e.g. pinned torch tensors)."""
        self._check(self.lib.rb2_accel_host(int(n), pos_ptr, charge_ptr, mass_ptr, out_ptr))

    # -- field -----------------------------------------------------------------------------------
    def Calc_Field_at_Batch(self, pos_in):
        """src/mod_verlet.F90:1635."""
        pts = np.ascontiguousarray(pos_in, dtype=np.float64).reshape(-1, 3)
        out = np.zeros_like(pts)
        if pts.shape[0]:
            self._check(self.lib.rb2_field_batch(pts.shape[0], _d(pts), _d(out)))
        return out

    def Calc_Field_at(self, pos_xyz):
        """src/mod_verlet.F90:1466."""
        return self.Calc_Field_at_Batch(np.asarray(pos_xyz, dtype=np.float64).reshape(1, 3))[0]

    def Calc_Field_at_Batch_delta(self, pos_in, new_pos, new_charge):
        pts = np.ascontiguousarray(pos_in, dtype=np.float64).reshape(-1, 3)
        npos = np.ascontiguousarray(new_pos, dtype=np.float64).reshape(-1, 3)
        nq = np.ascontiguousarray(new_charge, dtype=np.float64).reshape(-1)
        out = np.zeros_like(pts)
        if pts.shape[0]:
            self._check(self.lib.rb2_field_batch_delta(pts.shape[0], _d(pts), npos.shape[0], _d(npos), _d(nq), _d(out)))
        return out

    def field_surface_z(self, pos_in):
        """E_z at points of the cathode plane z = 0 (planar geometry), see rb2_field_surface_z."""
        pts = np.ascontiguousarray(pos_in, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(pts.shape[0])
        if pts.shape[0]:
            self._check(self.lib.rb2_field_surface_z(pts.shape[0], _d(pts), _d(out)))
        return out

    def mh_planar(self, M, emit_pos, emit_dim, w_theta, seed, kind=1, image_charge=True, T_temp=293.15,
                  a_rate=1.0, MH_std=0.0125, ndim=None, ndim_first=None):
        """Device-resident lock-step chains (src/mod_field_emission_v2.F90:1284-1458).
        Returns (df, F, pos, a_rate, MH_std)."""
        w = np.ascontiguousarray(w_theta, dtype=np.float64)
        if w.ndim != 2:
            w = w.reshape(1, -1)
        c = MhConfig()
        c.kind = kind
        c.ndim = (25 if kind == 2 else 200) if ndim is None else ndim
        c.ndim_first = (0 if kind == 2 else int(round(c.ndim * 0.25))) if ndim_first is None else ndim_first
        c.image_charge = int(bool(image_charge))
        c.y_num, c.x_num = w.shape
        c.emit_pos[:] = list(emit_pos)[:2]
        c.emit_dim[:] = list(emit_dim)[:2]
        c.T_temp = T_temp
        c.init_std, c.target_rate, c.std_gain = 0.10, 0.35, 0.025
        c.std_min, c.std_max = (0.005 if kind == 2 else 0.00005), 0.1250
        df, F, pos = np.zeros(M), np.zeros(M), np.zeros((M, 3))
        ar, sd = C.c_double(a_rate), C.c_double(MH_std)
        self._check(self.lib.rb2_mh_planar(C.byref(c), _d(w), M, seed, _d(df), _d(F), _d(pos), C.byref(ar), C.byref(sd)))
        return df, F, pos, ar.value, sd.value

    def mh_planar_serial(self, M, emit_pos, emit_dim, w_theta, seed, kind=1, image_charge=True, T_temp=293.15,
                         a_rate=1.0, MH_std=0.0125, ndim=None, ndim_first=None):
        """The reference's default serial chains (mh_batch = .false.) of one time step in one kernel: chain s sees the
        electrons emitted by chains < s.  Returns (df, F, pos, emitted, a_rate, MH_std)."""
        w = np.ascontiguousarray(w_theta, dtype=np.float64)
        if w.ndim != 2:
            w = w.reshape(1, -1)
        c = MhConfig()
        c.kind = kind
        c.ndim = (25 if kind == 2 else 200) if ndim is None else ndim
        c.ndim_first = (0 if kind == 2 else int(round(c.ndim * 0.25))) if ndim_first is None else ndim_first
        c.image_charge = int(bool(image_charge))
        c.y_num, c.x_num = w.shape
        c.emit_pos[:] = list(emit_pos)[:2]
        c.emit_dim[:] = list(emit_dim)[:2]
        c.T_temp = T_temp
        c.init_std, c.target_rate, c.std_gain = 0.10, 0.35, 0.025
        c.std_min, c.std_max = (0.005 if kind == 2 else 0.00005), 0.1250
        df, F, pos, em = np.zeros(M), np.zeros(M), np.zeros((M, 3)), np.zeros(M, dtype=np.int32)
        ar, sd = C.c_double(a_rate), C.c_double(MH_std)
        self._check(self.lib.rb2_mh_planar_serial(C.byref(c), _d(w), M, seed, _d(df), _d(F), _d(pos), _i(em), C.byref(ar), C.byref(sd)))
        return df, F, pos, em, ar.value, sd.value

    def mh_tip(self, M, seed, ndim=80, a_rate=0.5, MH_std=1.0):
        """Device-resident lock-step tip chains (Metro_algo_tip_v3, src/mod_emission_tip.f90:1241-1390).
        Returns (eta_f, df, pos, a_rate, MH_std)."""
        ef, df, pos = np.zeros(M), np.zeros(M), np.zeros((M, 3))
        ar, sd = C.c_double(a_rate), C.c_double(MH_std)
        self._check(self.lib.rb2_mh_tip(M, ndim, seed, _d(ef), _d(df), _d(pos), C.byref(ar), C.byref(sd)))
        return ef, df, pos, ar.value, sd.value

    def planar_supply_level(self, emit_pos, emit_dim, w_theta, shifts, n_done, n_new, kind=1, T_temp=293.15):
        """One level of the planar supply quadrature on the device (rb2_planar_supply_level): per-shift sums of the
        integrand over lattice nodes n_done+1 .. n_done+n_new and the sum of E_z over them."""
        w = np.ascontiguousarray(w_theta, dtype=np.float64)
        if w.ndim != 2:
            w = w.reshape(1, -1)
        c = MhConfig()
        c.kind = kind
        c.y_num, c.x_num = w.shape
        c.emit_pos[:] = list(emit_pos)[:2]
        c.emit_dim[:] = list(emit_dim)[:2]
        c.T_temp = T_temp
        sh = np.ascontiguousarray(shifts, dtype=np.float64).reshape(-1, 2)
        sums = np.zeros(sh.shape[0])
        ez = C.c_double(0.0)
        self._check(self.lib.rb2_planar_supply_level(C.byref(c), _d(w), kind, sh.shape[0], _d(sh), int(n_done), int(n_new), _d(sums), C.byref(ez)))
        return sums, ez.value

    def tip_supply_set_grid(self, pts, normals, area):
        """Nodes (M,3), unit normals (M,3) and patch areas (M,) of the tip's supply grid; kept on the device."""
        pts = np.ascontiguousarray(pts, dtype=np.float64); normals = np.ascontiguousarray(normals, dtype=np.float64)
        area = np.ascontiguousarray(area, dtype=np.float64)
        self._check(self.lib.rb2_tip_supply_set_grid(int(area.shape[0]), _d(pts), _d(normals), _d(area)))

    def tip_supply(self):
        """(n_s, sum of the normal field over the nodes): Elec_Supply summed over the resident grid, src/mod_emission_tip.f90:431-481."""
        ns, fs = C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.rb2_tip_supply(C.byref(ns), C.byref(fs)))
        return ns.value, fs.value

    def Particles_To_Device(self):
        self._check(self.lib.rb2_field_window_open())

    def Release_Device_Particles(self):
        self._check(self.lib.rb2_field_window_close())

    # -- multi-GPU / measurement ---------------------------------------------------------------------
    def set_devices(self, devices):
        """One process, several GPUs: replicas of the store on every listed device, the pair work split over them."""
        d = np.ascontiguousarray(devices, dtype=np.int32)
        self._check(self.lib.rb2_set_devices(d.size, _i(d)))

    def set_partition(self, i_begin, i_end):
        self._check(self.lib.rb2_set_partition(int(i_begin), int(i_end)))

    def set_option(self, name: str, value: float):
        self._check(self.lib.rb2_set_option(name.encode(), float(value)))

    def set_pair_rank(self, rank, world):
        self._check(self.lib.rb2_set_pair_rank(int(rank), int(world)))

    def accel_partial(self):
        self._check(self.lib.rb2_accel_partial())

    def accel_finalize(self):
        self._check(self.lib.rb2_accel_finalize())

    def Sample_Elec_Position(self):
        """The sweep of Sample_Elec_Position (mod_pair.F90:975-1037): (nearest distance, 0-based index of the nearest
        other electron) for every particle; rows that are not electrons hold (1000.0, -1)."""
        n = self.counts().nrPart
        dist = np.empty(n)
        idx = np.empty(n, dtype=np.int32)
        if n > 0:
            self._check(self.lib.rb2_nearest_electron(_d(dist), _i(idx)))
        return dist, idx

    # -- mod_collisions ------------------------------------------------------------------------------
    def Init_Collisions(self, collision_mode, tot_energy, tot_data, ion_energy, ion_data, n_d, cyl_radius,
                        ion_life_time=100000000):
        """Read_Cross_Section + the namelist values Do_Electron_Atom_Collisions reads (src/mod_collisions.F90)."""
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (tot_energy, tot_data, ion_energy, ion_data)]
        c = CollisionConfig()
        c.collision_mode, c.ion_life_time, c.n_d, c.cyl_radius = int(collision_mode), int(ion_life_time), n_d, cyl_radius
        c.n_tot, c.n_ion = len(keep[0]), len(keep[2])
        c.tot_energy, c.tot_data, c.ion_energy, c.ion_data = (_d(a) for a in keep)
        self._check(self.lib.rb2_collisions_init(C.byref(c)))

    def Update_Collision_Data_All(self):
        """src/mod_collisions.F90:2155: (n, 5) = energy, ion_cross_sec, ion_cross_rad, recom_cross_rad, tot_cross_sec."""
        n = self.counts().nrPart
        out = np.zeros((n, 5))
        if n:
            self._check(self.lib.rb2_collision_data(_d(out)))
        return out

    def Do_Continuous_Ionization(self, step, seed) -> CollisionResult:
        r = CollisionResult()
        self._check(self.lib.rb2_continuous_ionization(int(step), int(seed), C.byref(r)))
        return r

    def Do_Discrete_Recombination(self, step) -> CollisionResult:
        r = CollisionResult()
        self._check(self.lib.rb2_discrete_recombination(int(step), C.byref(r)))
        return r

    def Do_Electron_Atom_Collisions(self, step, seed) -> CollisionResult:
        """src/mod_collisions.F90:30."""
        r = CollisionResult()
        self._check(self.lib.rb2_do_collisions(int(step), int(seed), C.byref(r)))
        return r

    def recombination_records(self):
        n = C.c_int(0)
        self._check(self.lib.rb2_get_recombination_records(0, None, C.byref(n)))
        buf = (RecombRecord * max(n.value, 1))()
        self._check(self.lib.rb2_get_recombination_records(n.value, buf, C.byref(n)))
        return [buf[k] for k in range(n.value)]

    def ionization_records(self):
        n = C.c_int(0)
        self._check(self.lib.rb2_get_ionization_records(0, None, C.byref(n)))
        buf = (IonizationRecord * max(n.value, 1))()
        self._check(self.lib.rb2_get_ionization_records(n.value, buf, C.byref(n)))
        return [buf[k] for k in range(n.value)]

    def probe_quartic_roots(self, coeffs):
        """Device QuarticRoots on rows {quartic, cubic, quadratic, linear, constant}: (codes[n], roots[n, 4] complex)."""
        co = np.ascontiguousarray(coeffs, dtype=np.float64).reshape(-1, 5)
        n = co.shape[0]
        codes = np.zeros(n, dtype=np.int32)
        roots = np.zeros((n, 8))
        self._check(self.lib.rb2_probe_quartic_roots(n, _d(co), _i(codes), _d(roots)))
        return codes, roots[:, 0::2] + 1j * roots[:, 1::2]

    def p2p_export(self, n_max) -> bytes:
        """This process's exchange block (partial pair sums + flags) as a CUDA IPC handle."""
        buf = C.create_string_buffer(P2P_HANDLE_BYTES)
        self._check(self.lib.rb2_p2p_export(int(n_max), buf))
        return buf.raw

    def p2p_attach(self, world, rank, handles):
        """Map every rank's exchange block (handles: the `world` exported handles in rank order).  From here on
        the acceleration evaluation exchanges the partial sums over peer memory inside its finalise kernel."""
        blob = b"".join(handles)
        assert len(blob) == world * P2P_HANDLE_BYTES
        self._check(self.lib.rb2_p2p_attach(int(world), int(rank), C.c_char_p(blob)))

    def p2p_detach(self):
        self._check(self.lib.rb2_p2p_detach())

    def device_buffer(self, name: str):
        p = C.c_void_p()
        b = C.c_size_t()
        self._check(self.lib.rb2_device_buffer(name.encode(), C.byref(p), C.byref(b)))
        return p.value, b.value

    def stream(self):
        p = C.c_void_p()
        self._check(self.lib.rb2_stream(C.byref(p)))
        return p.value

    def synchronize(self):
        self._check(self.lib.rb2_synchronize())

    def fp64_peak(self, ms_target=50.0):
        t = C.c_double()
        ms = C.c_float()
        self._check(self.lib.rb2_fp64_peak(ms_target, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def launch_count(self, reset=False):
        v = C.c_longlong()
        self._check(self.lib.rb2_launch_count(C.byref(v), int(reset)))
        return v.value

    def stat(self, name: str) -> float:
        v = C.c_double()
        self._check(self.lib.rb2_get_stat(name.encode(), C.byref(v)))
        return v.value

    def last_accel_info(self):
        ms = C.c_float()
        gx, gy, bl, js = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.rb2_last_accel_info(C.byref(ms), C.byref(gx), C.byref(gy), C.byref(bl), C.byref(js)))
        return dict(ms=ms.value, grid_x=gx.value, grid_y=gy.value, block=bl.value, j_chunk=js.value)


def sym_plan_probe(n, world=1, rank=0, tpl=0, sm_count=148, waves=64.0, kmax=12, gmax=24, budget_mb=2048.0):
    """Host-only planning of the pair-symmetric work (rb2_sym_plan_probe; needs no GPU): returns a dict with the unit shape
    (T, K, G, band width, tiles, superblocks), the cost dealt to every rank, this rank's units (band, set, group) in launch
    order and the hash of the owner table."""
    lib = load_library()
    shape = (C.c_int * 6)()
    cost = (C.c_longlong * world)()
    n_units = C.c_int(0)
    h = C.c_ulonglong(0)
    rc = lib.rb2_sym_plan_probe(n, tpl, world, rank, sm_count, waves, kmax, gmax, budget_mb, shape, cost, None, 0, C.byref(n_units), C.byref(h))
    if rc:
        raise Rb2Error(f"rb2 error {rc}: {lib.rb2_last_error_string().decode()}")
    cap = max(1, n_units.value)
    units = (C.c_int * (3 * cap))()
    rc = lib.rb2_sym_plan_probe(n, tpl, world, rank, sm_count, waves, kmax, gmax, budget_mb, shape, cost, units, cap, C.byref(n_units), C.byref(h))
    if rc:
        raise Rb2Error(f"rb2 error {rc}: {lib.rb2_last_error_string().decode()}")
    T, K, G, Wb, nsb, nIb = list(shape)
    return dict(T=T, K=K, G=G, Wb=Wb, nsb=nsb, nIb=nIb, rank_cost=list(cost), table_hash=h.value,
                units=np.array(list(units), dtype=np.int64).reshape(-1, 3)[: n_units.value])


def device_available() -> bool:
    try:
        return bool(load_library().rb2_device_available())
    except Rb2Error:
        return False
