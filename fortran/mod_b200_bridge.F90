!-------------------------------------------------------------------------------
! mod_b200_bridge -- ISO_C_BINDING layer between the RUMDEED Fortran host and
! librumdeed_b200.so (C ABI: include/rumdeed_b200.h).
!
! NOTE: no Fortran compiler exists in the image this repository is built in, so
! this file has never been compiled there.  It is kept deliberately mechanical:
! one interface block per C entry point plus thin wrappers with the names and
! argument conventions of the reference routines they replace.  The same C
! symbols are exercised by tests/ through ctypes with Fortran-layout arrays.
!
! Build (with the reference's makefile):  add mod_b200_bridge.o after
! mod_hyperboloid_tip.o, compile with -DRUMDEED_B200 and link
! -L<repo>/rumdeed_b200 -lrumdeed_b200.  See INTEGRATION.md for the five call
! sites in mod_verlet.F90 / mod_pair.F90 / main.F90 that dispatch here.
!-------------------------------------------------------------------------------
module mod_b200_bridge
  use, intrinsic :: iso_c_binding
  use mod_global
  use mod_hyperboloid_tip, only: a_foci, eta_1, shift_z, pre_fac_E_tip, &
                               & pre_fac_E_tip_unit_voltage, h_tip, r_tip, max_xi
  implicit none
  private

  public :: B200_Init, B200_Finalize, B200_Sync_Config
  public :: B200_Upload_Particles, B200_Download_Particles
  public :: B200_Add_Particle, B200_Mark_Particles_Remove, B200_Remove_Particles
  public :: B200_Update_Position, B200_Calculate_Acceleration_Particles
  public :: B200_Calc_Field_at, B200_Calc_Field_at_Batch
  public :: B200_Particles_To_Device, B200_Release_Device_Particles
  public :: B200_Init_Collisions, B200_Do_Electron_Atom_Collisions

  integer, parameter :: RB2_GEOM_PLANAR = 1, RB2_GEOM_TIP = 2
  integer, parameter :: RB2_PLANES_MAX = 10

  ! struct rb2_config (include/rumdeed_b200.h)
  type, bind(C) :: rb2_config
    integer(c_int) :: geometry, image_charge, N_ic_max, planes_N
    real(c_double) :: V_s, d, E_z
    real(c_double) :: box_dim(3)
    real(c_double) :: time_step
    real(c_double) :: planes_z(RB2_PLANES_MAX)
    real(c_double) :: a_foci, eta_1, shift_z, pre_fac_E_tip, pre_fac_E_tip_unit_voltage, h_tip, r_tip, max_xi
    integer(c_int) :: capacity, device
  end type rb2_config

  type, bind(C) :: rb2_counts
    integer(c_int) :: nrPart, nrElec, nrIon, nrAtom, nrID, nrPart_dropped
    integer(c_int) :: nrPart_remove, nrElec_remove, nrIon_remove, nrAtom_remove
    integer(c_int) :: nrPart_remove_top, nrPart_remove_bot
    integer(c_int) :: nrElec_remove_top, nrElec_remove_bot
    integer(c_int) :: nrIon_remove_top, nrIon_remove_bot
    integer(c_int) :: nrPart_remove_ion, nrElec_remove_ion, nrAtom_remove_ion
  end type rb2_counts

  type, bind(C) :: rb2_event
    integer(c_int) :: kind, plane, index
    real(c_double) :: x, y, vx, vy, vz
    integer(c_int) :: emit, sec, id
  end type rb2_event

  type, bind(C) :: rb2_mh_config
    integer(c_int) :: kind, ndim, ndim_first, image_charge, y_num, x_num
    real(c_double) :: emit_pos(2), emit_dim(2)
    real(c_double) :: T_temp
    real(c_double) :: init_std, target_rate, std_gain, std_min, std_max
  end type rb2_mh_config

  ! collisions (COLLISION_MODE 1, 2): struct rb2_collision_config / _result / records
  type, bind(C) :: rb2_collision_config
    integer(c_int) :: collision_mode, ion_life_time
    real(c_double) :: n_d, cyl_radius
    integer(c_int) :: n_tot, n_ion
    type(c_ptr)    :: tot_energy, tot_data, ion_energy, ion_data
  end type rb2_collision_config

  type, bind(C) :: rb2_recomb_record
    integer(c_int) :: step, elec_slot, ion_slot, elec_emit, ion_life
    integer(c_int) :: elec_sec, elec_id, ion_emit, ion_sec, ion_id
    real(c_double) :: ion_pos(3), elec_pos(3)
    real(c_double) :: elec_speed, dist, recom_rad, t
  end type rb2_recomb_record

  type, bind(C) :: rb2_ionization_record
    integer(c_int) :: step, in_slot, new_id, ion_id, elec_emit, pad
    real(c_double) :: pos(3)
    real(c_double) :: in_speed, out_speed, new_speed
    real(c_double) :: new_vel(3), ejec_pos(3), ejec_vel(3), ion_pos(3)
    real(c_double) :: E1, collE, ejecE
  end type rb2_ionization_record

  type, bind(C) :: rb2_step_result
    real(c_double) :: ramo_current(4)
    real(c_double) :: avg_part_vel(3), avg_elec_vel(3), avg_ion_vel(3)
    integer(c_int) :: n_events
    type(rb2_counts) :: counts
    real(c_float)  :: accel_ms, step_ms
  end type rb2_step_result

  type, bind(C) :: rb2_collision_result
    integer(c_int) :: nrCollisions, nrIonizations, nrRecombinations, nrIonsExpired
    integer(c_int) :: nrPart_remove_recom, nrElec_remove_recom, nrIon_remove_recom, n_candidates
    type(rb2_counts) :: counts
    real(c_float)  :: ms
  end type rb2_collision_result

  interface
    integer(c_int) function rb2_collisions_init(cfg) bind(C, name='rb2_collisions_init')
      import :: c_int, rb2_collision_config
      type(rb2_collision_config), intent(in) :: cfg
    end function
    integer(c_int) function rb2_do_collisions(step, seed, res) bind(C, name='rb2_do_collisions')
      import :: c_int, c_long_long, rb2_collision_result
      integer(c_int), value :: step
      integer(c_long_long), value :: seed
      type(rb2_collision_result), intent(out) :: res
    end function
    integer(c_int) function rb2_get_recombination_records(max_records, recs, n_out) bind(C, name='rb2_get_recombination_records')
      import :: c_int, rb2_recomb_record
      integer(c_int), value :: max_records
      type(rb2_recomb_record), intent(out) :: recs(*)
      integer(c_int), intent(out) :: n_out
    end function
    integer(c_int) function rb2_get_ionization_records(max_records, recs, n_out) bind(C, name='rb2_get_ionization_records')
      import :: c_int, rb2_ionization_record
      integer(c_int), value :: max_records
      type(rb2_ionization_record), intent(out) :: recs(*)
      integer(c_int), intent(out) :: n_out
    end function
  end interface

  interface
    integer(c_int) function rb2_init(cfg) bind(C, name='rb2_init')
      import :: c_int, rb2_config
      type(rb2_config), intent(in) :: cfg
    end function
    integer(c_int) function rb2_finalize() bind(C, name='rb2_finalize')
      import :: c_int
    end function
    integer(c_int) function rb2_update_config(cfg) bind(C, name='rb2_update_config')
      import :: c_int, rb2_config
      type(rb2_config), intent(in) :: cfg
    end function
    integer(c_int) function rb2_upload_particles(n, pos, prev_pos, vel, acc, acc_prev, acc_prev2, charge, mass, &
                          & species, step, emitter, section, life, id, nrID) bind(C, name='rb2_upload_particles')
      import :: c_int, c_double
      integer(c_int), value :: n, nrID
      real(c_double), intent(in) :: pos(3, *), prev_pos(3, *), vel(3, *), acc(3, *), acc_prev(3, *), acc_prev2(3, *)
      real(c_double), intent(in) :: charge(*), mass(*)
      integer(c_int), intent(in) :: species(*), step(*), emitter(*), section(*), life(*), id(*)
    end function
    integer(c_int) function rb2_download_particles(pos, prev_pos, vel, acc, acc_prev, acc_prev2, charge, mass, &
                          & species, step, emitter, section, life, id, mask) bind(C, name='rb2_download_particles')
      import :: c_int, c_double, c_ptr
      real(c_double), intent(out) :: pos(3, *), prev_pos(3, *), vel(3, *), acc(3, *), acc_prev(3, *), acc_prev2(3, *)
      real(c_double), intent(out) :: charge(*), mass(*)
      integer(c_int), intent(out) :: species(*), step(*), emitter(*), section(*), life(*), id(*)
      type(c_ptr), value :: mask ! pass c_null_ptr: the Fortran masks are all .true. after Remove_Particles
    end function
    integer(c_int) function rb2_add_particles(k, pos, vel, species, step, emit, sec, life) bind(C, name='rb2_add_particles')
      import :: c_int, c_double
      integer(c_int), value :: k, step
      real(c_double), intent(in) :: pos(3, *), vel(3, *)
      integer(c_int), intent(in) :: species(*), emit(*), sec(*), life(*)
    end function
    integer(c_int) function rb2_mark_remove(k, index, reason) bind(C, name='rb2_mark_remove')
      import :: c_int
      integer(c_int), value :: k
      integer(c_int), intent(in) :: index(*), reason(*)
    end function
    integer(c_int) function rb2_remove_marked(step, counts) bind(C, name='rb2_remove_marked')
      import :: c_int, rb2_counts
      integer(c_int), value :: step
      type(rb2_counts), intent(out) :: counts
    end function
    integer(c_int) function rb2_step(step, res) bind(C, name='rb2_step')
      import :: c_int, rb2_step_result
      integer(c_int), value :: step
      type(rb2_step_result), intent(out) :: res
    end function
    integer(c_int) function rb2_accel_only() bind(C, name='rb2_accel_only')
      import :: c_int
    end function
    integer(c_int) function rb2_get_events(max_events, events, n_out) bind(C, name='rb2_get_events')
      import :: c_int, rb2_event
      integer(c_int), value :: max_events
      type(rb2_event), intent(out) :: events(*)
      integer(c_int), intent(out) :: n_out
    end function
    integer(c_int) function rb2_field_batch(M, pos_in, field_out) bind(C, name='rb2_field_batch')
      import :: c_int, c_double
      integer(c_int), value :: M
      real(c_double), intent(in)  :: pos_in(3, *)
      real(c_double), intent(out) :: field_out(3, *)
    end function
    integer(c_int) function rb2_field_surface_z(M, pos_in, Ez_out) bind(C, name='rb2_field_surface_z')
      import :: c_int, c_double
      integer(c_int), value :: M
      real(c_double), intent(in)  :: pos_in(3, *)
      real(c_double), intent(out) :: Ez_out(*)
    end function
    integer(c_int) function rb2_mh_planar(cfg, w_theta, M, seed, df_out, F_out, pos_out, a_rate_io, mh_std_io) &
        bind(C, name='rb2_mh_planar')
      import :: c_int, c_double, c_long_long, rb2_mh_config
      type(rb2_mh_config), intent(in) :: cfg
      real(c_double), intent(in) :: w_theta(*)          ! transpose(w_theta_arr): rows of the `work` file
      integer(c_int), value :: M
      integer(c_long_long), value :: seed
      real(c_double), intent(out) :: df_out(*), F_out(*), pos_out(3, *)
      real(c_double), intent(inout) :: a_rate_io, mh_std_io
    end function
    integer(c_int) function rb2_field_window_open() bind(C, name='rb2_field_window_open')
      import :: c_int
    end function
    integer(c_int) function rb2_field_window_close() bind(C, name='rb2_field_window_close')
      import :: c_int
    end function
    ! lock-step tip chains of a time step on the device (replaces the per-candidate Metro_algo_tip_v3 calls)
    integer(c_int) function rb2_mh_tip(M, ndim, seed, eta_f_out, df_out, pos_out, a_rate_io, mh_std_io) bind(C, name='rb2_mh_tip')
      import :: c_int, c_long_long, c_double
      integer(c_int), value :: M, ndim
      integer(c_long_long), value :: seed
      real(c_double), intent(out) :: eta_f_out(*), df_out(*), pos_out(3, *)
      real(c_double), intent(inout) :: a_rate_io, mh_std_io
    end function
    ! one level of the planar supply quadrature on the device (nodes of K shifted lattices, field, integrand, per-shift sums)
    integer(c_int) function rb2_planar_supply_level(cfg, w_theta, kind, K, shifts, n_done, n_new, sums_out, ez_sum_out) &
        bind(C, name='rb2_planar_supply_level')
      import :: c_int, c_double, rb2_mh_config
      type(rb2_mh_config), intent(in) :: cfg
      real(c_double), intent(in) :: w_theta(*), shifts(2, *)
      integer(c_int), value :: kind, K, n_done, n_new
      real(c_double), intent(out) :: sums_out(*), ez_sum_out
    end function
    ! supply sum over the tip's (xi, phi) grid on the device (nodes, normals, areas handed over once)
    integer(c_int) function rb2_tip_supply_set_grid(M, pts, normals, area) bind(C, name='rb2_tip_supply_set_grid')
      import :: c_int, c_double
      integer(c_int), value :: M
      real(c_double), intent(in) :: pts(3, *), normals(3, *), area(*)
    end function
    integer(c_int) function rb2_tip_supply(n_s_out, F_sum_out) bind(C, name='rb2_tip_supply')
      import :: c_int, c_double
      real(c_double), intent(out) :: n_s_out, F_sum_out
    end function
    integer(c_int) function rb2_nearest_electron(dist_out, id_out) bind(C, name='rb2_nearest_electron')
      import :: c_int, c_double
      real(c_double), intent(out) :: dist_out(*)
      integer(c_int), intent(out) :: id_out(*)          ! 0-based index of the nearest other electron, -1: none
    end function
    ! multi-GPU (one process per GPU): export / gather / attach the exchange blocks, see INTEGRATION.md section 4
    integer(c_int) function rb2_p2p_export(n_max, handle_out) bind(C, name='rb2_p2p_export')
      import :: c_int, c_char
      integer(c_int), value :: n_max
      character(kind=c_char), intent(out) :: handle_out(64)
    end function
    integer(c_int) function rb2_p2p_attach(world, rank, handles) bind(C, name='rb2_p2p_attach')
      import :: c_int, c_char
      integer(c_int), value :: world, rank
      character(kind=c_char), intent(in) :: handles(64, *)   ! handles(:, r+1) = what rank r exported
    end function
    integer(c_int) function rb2_p2p_detach() bind(C, name='rb2_p2p_detach')
      import :: c_int
    end function
    ! ONE process, several GPUs (this program is a single process): after B200_Init, before the first particle
    integer(c_int) function rb2_set_devices(n_devices, devices) bind(C, name='rb2_set_devices')
      import :: c_int
      integer(c_int), value :: n_devices
      integer(c_int), intent(in) :: devices(*)           ! devices(1) = the device of rb2_init
    end function
    integer(c_int) function rb2_set_option(name, value) bind(C, name='rb2_set_option')
      import :: c_int, c_char, c_double
      character(kind=c_char), intent(in) :: name(*)      ! null terminated
      real(c_double), value :: value
    end function
    ! ramo_current_emit(1:n_sec, 1:n_emit) of the last velocity update (mod_verlet.F90:489-492)
    integer(c_int) function rb2_get_ramo_sections(n_sec, n_emit, ramo_out) bind(C, name='rb2_get_ramo_sections')
      import :: c_int, c_double
      integer(c_int), value :: n_sec, n_emit
      real(c_double), intent(out) :: ramo_out(n_sec, *)
    end function
    integer(c_int) function rb2_capacity_left(room) bind(C, name='rb2_capacity_left')
      import :: c_int
      integer(c_int), intent(out) :: room
    end function
    ! the default sampler (mh_batch = .false.): the serial chains of a time step in one kernel
    integer(c_int) function rb2_mh_planar_serial(cfg, w_theta, M, seed, df_out, F_out, pos_out, emit_out, a_rate_io, mh_std_io) &
        bind(C, name='rb2_mh_planar_serial')
      import :: c_int, c_double, c_long_long, rb2_mh_config
      type(rb2_mh_config), intent(in) :: cfg
      real(c_double), intent(in) :: w_theta(*)
      integer(c_int), value :: M
      integer(c_long_long), value :: seed
      real(c_double), intent(out) :: df_out(*), F_out(*), pos_out(3, *)
      integer(c_int), intent(out) :: emit_out(*)         ! 1: candidate emitted (insert it with B200_Add_Particle, in order)
      real(c_double), intent(inout) :: a_rate_io, mh_std_io
    end function
    function rb2_last_error_string() bind(C, name='rb2_last_error_string') result(p)
      import :: c_ptr
      type(c_ptr) :: p
    end function
  end interface

  type(rb2_event), allocatable :: ev_buf(:)

contains

  subroutine Check(rc, where)
    integer(c_int), intent(in)   :: rc
    character(len=*), intent(in) :: where
    if (rc /= 0) then
      print '(a, a, a, i0)', 'RUMDEED: librumdeed_b200 error in ', where, ', code ', rc
      error stop 2
    end if
  end subroutine Check

  ! Geometry id: the same test ACC_Geometry() does (mod_verlet.F90:1168); 0 = not on the device.
  subroutine Fill_Config(cfg, geom)
    type(rb2_config), intent(out) :: cfg
    integer, intent(in)           :: geom
    cfg%geometry = geom
    cfg%image_charge = merge(1, 0, image_charge)
    cfg%N_ic_max = N_ic_max
    cfg%planes_N = planes_N
    cfg%V_s = V_s
    cfg%d = d
    cfg%E_z = E_z
    cfg%box_dim = box_dim
    cfg%time_step = time_step
    cfg%planes_z = planes_z
    cfg%a_foci = a_foci;  cfg%eta_1 = eta_1;  cfg%shift_z = shift_z
    cfg%pre_fac_E_tip = pre_fac_E_tip
    cfg%pre_fac_E_tip_unit_voltage = pre_fac_E_tip_unit_voltage
    cfg%h_tip = h_tip;  cfg%r_tip = r_tip;  cfg%max_xi = max_xi
    cfg%capacity = MAX_PARTICLES
    cfg%device = -1
  end subroutine Fill_Config

  ! Call at the end of Init_* (main.F90:146), once the procedure pointers are bound.
  subroutine B200_Init(geom)
    integer, intent(in) :: geom
    type(rb2_config)    :: cfg
    call Fill_Config(cfg, geom)
    call Check(rb2_init(cfg), 'rb2_init')
    ! ramo_current_emit(sec, emit) is only accumulated on request (write_ramo_sec, mod_global.F90:352)
    if (write_ramo_sec) call Check(rb2_set_option('ramo_sections'//c_null_char, real(MAX_SECTIONS, c_double)), 'rb2_set_option')
  end subroutine B200_Init

  ! Optional: let this (single) process drive several GPUs.  devices(1) must be the current device.
  subroutine B200_Set_Devices(devices)
    integer, intent(in) :: devices(:)
    call Check(rb2_set_devices(int(size(devices), c_int), int(devices, c_int)), 'rb2_set_devices')
  end subroutine B200_Set_Devices

  subroutine B200_Finalize()
    call Check(rb2_finalize(), 'rb2_finalize')
  end subroutine B200_Finalize

  ! After Set_Voltage or any change of image_charge / N_ic_max (the unit tests do that).
  subroutine B200_Sync_Config(geom)
    integer, intent(in) :: geom
    type(rb2_config)    :: cfg
    call Fill_Config(cfg, geom)
    call Check(rb2_update_config(cfg), 'rb2_update_config')
  end subroutine B200_Sync_Config

  ! One-off synchronisation points (e.g. before Write_Position or a host-only module runs).
  subroutine B200_Upload_Particles()
    if (nrPart > 0) then
      call Check(rb2_upload_particles(int(nrPart, c_int), particles_cur_pos, particles_prev_pos, particles_cur_vel, &
               & particles_cur_accel, particles_prev_accel, particles_prev2_accel, particles_charge, particles_mass, &
               & particles_species, particles_step, particles_emitter, particles_section, particles_life, particles_id, &
               & int(nrID, c_int)), 'rb2_upload_particles')
    end if
  end subroutine B200_Upload_Particles

  subroutine B200_Download_Particles()
    if (nrPart > 0) then
      call Check(rb2_download_particles(particles_cur_pos, particles_prev_pos, particles_cur_vel, particles_cur_accel, &
               & particles_prev_accel, particles_prev2_accel, particles_charge, particles_mass, particles_species, &
               & particles_step, particles_emitter, particles_section, particles_life, particles_id, c_null_ptr), &
               & 'rb2_download_particles')
    end if
  end subroutine B200_Download_Particles

  ! Add_Particle (mod_pair.F90:29): the host keeps doing its own bookkeeping and file
  ! writes; this mirrors the new particle into the device store (slot nrPart, id nrID).
  subroutine B200_Add_Particle(par_pos, par_vel, par_species, step, emit, life, sec)
    double precision, dimension(1:3), intent(in) :: par_pos, par_vel
    integer, intent(in)                          :: par_species, step, emit, life, sec
    real(c_double) :: p(3, 1), v(3, 1)
    integer(c_int) :: s(1), e(1), c(1), l(1)
    p(:, 1) = par_pos;  v(:, 1) = par_vel
    s(1) = par_species;  e(1) = emit;  c(1) = sec;  l(1) = life
    call Check(rb2_add_particles(1_c_int, p, v, s, int(step, c_int), e, c, l), 'rb2_add_particles')
  end subroutine B200_Add_Particle

  ! Mark_Particles_Remove (mod_pair.F90:169) for marks decided on the host (collisions).
  subroutine B200_Mark_Particles_Remove(i, m)
    integer, intent(in) :: i, m
    integer(c_int)      :: idx(1), why(1)
    idx(1) = i - 1 ! 0-based on the C side
    why(1) = m
    call Check(rb2_mark_remove(1_c_int, idx, why), 'rb2_mark_remove')
  end subroutine B200_Mark_Particles_Remove

  ! Remove_Particles (mod_pair.F90:352): compaction on the device, counters back to mod_global.
  subroutine B200_Remove_Particles(step)
    integer, intent(in) :: step
    type(rb2_counts)    :: k
    call Check(rb2_remove_marked(int(step, c_int), k), 'rb2_remove_marked')
    nrPart = k%nrPart;  nrElec = k%nrElec;  nrIon = k%nrIon;  nrAtom = k%nrAtom
    nrElecIon = nrElec + nrIon
    nrPart_remove = 0;  nrElec_remove = 0;  nrIon_remove = 0;  nrAtom_remove = 0
    nrPart_remove_top = 0;  nrPart_remove_bot = 0
    nrElec_remove_top = 0;  nrElec_remove_bot = 0
    nrIon_remove_top = 0;  nrIon_remove_bot = 0
    nrPart_remove_ion = 0;  nrElec_remove_ion = 0;  nrAtom_remove_ion = 0
  end subroutine B200_Remove_Particles

  ! Update_Position(step) (main.F90:190): one fused device step.  Fills ramo_current, the
  ! velocity averages and the remove counters, and writes the absorb / plane records in the
  ! serial order (ascending particle index) with the units the reference uses.
  subroutine B200_Update_Position(step)
    integer, intent(in)   :: step
    type(rb2_step_result) :: r
    integer(c_int)        :: n_ev
    integer               :: k
    call Check(rb2_step(int(step, c_int), r), 'rb2_step')
    ramo_current(1:nrSpecies) = r%ramo_current(2:nrSpecies+1) ! C index = species id
    avg_part_vel = r%avg_part_vel;  avg_elec_vel = r%avg_elec_vel;  avg_ion_vel = r%avg_ion_vel
    nrPart_remove = r%counts%nrPart_remove;  nrElec_remove = r%counts%nrElec_remove
    nrIon_remove = r%counts%nrIon_remove;    nrAtom_remove = r%counts%nrAtom_remove
    nrPart_remove_top = r%counts%nrPart_remove_top;  nrPart_remove_bot = r%counts%nrPart_remove_bot
    nrElec_remove_top = r%counts%nrElec_remove_top;  nrElec_remove_bot = r%counts%nrElec_remove_bot
    nrIon_remove_top = r%counts%nrIon_remove_top;    nrIon_remove_bot = r%counts%nrIon_remove_bot
    nrPart_remove_ion = r%counts%nrPart_remove_ion;  nrElec_remove_ion = r%counts%nrElec_remove_ion
    nrAtom_remove_ion = r%counts%nrAtom_remove_ion
    if (write_ramo_sec) then ! mod_verlet.F90:489-492; written by Write_Ramo_Current, mod_pair.F90:822-826
      call Check(rb2_get_ramo_sections(int(MAX_SECTIONS, c_int), int(MAX_EMITTERS, c_int), ramo_current_emit), 'rb2_get_ramo_sections')
    end if
    if (r%n_events > 0) then
      if (.not. allocated(ev_buf)) allocate(ev_buf(max(1024, int(r%n_events))))
      if (size(ev_buf) < r%n_events) then
        deallocate(ev_buf);  allocate(ev_buf(2*int(r%n_events)))
      end if
      call Check(rb2_get_events(int(size(ev_buf), c_int), ev_buf, n_ev), 'rb2_get_events')
      do k = 1, n_ev
        select case (ev_buf(k)%kind)
        case (1) ! mod_pair.F90:243-245
          write(unit=ud_density_absorb_top) ev_buf(k)%x, ev_buf(k)%y, ev_buf(k)%vx, ev_buf(k)%vy, ev_buf(k)%vz, &
                                          & ev_buf(k)%emit, ev_buf(k)%sec, ev_buf(k)%id
        case (2) ! mod_pair.F90:255-256
          write(unit=ud_density_absorb_bot) ev_buf(k)%x, ev_buf(k)%y, ev_buf(k)%emit, ev_buf(k)%sec, ev_buf(k)%id
        case (3) ! mod_verlet.F90:360-362
          write(unit=planes_ud(ev_buf(k)%plane + 1)) ev_buf(k)%x, ev_buf(k)%y, ev_buf(k)%vx, ev_buf(k)%vy, ev_buf(k)%vz, &
                                          & ev_buf(k)%emit, ev_buf(k)%sec, ev_buf(k)%id
        end select
      end do
    end if
  end subroutine B200_Update_Position

  ! Calculate_Acceleration_Particles (mod_verlet.F90:597), used directly by the unit tests.
  subroutine B200_Calculate_Acceleration_Particles()
    call Check(rb2_accel_only(), 'rb2_accel_only')
  end subroutine B200_Calculate_Acceleration_Particles

  ! Calc_Field_at_Batch (mod_verlet.F90:1635)
  subroutine B200_Calc_Field_at_Batch(M, pos_in, field_out)
    integer, intent(in)                                :: M
    double precision, dimension(1:3, 1:M), intent(in)  :: pos_in
    double precision, dimension(1:3, 1:M), intent(out) :: field_out
    if (M < 1) return
    call Check(rb2_field_batch(int(M, c_int), pos_in, field_out), 'rb2_field_batch')
  end subroutine B200_Calc_Field_at_Batch

  ! Calc_Field_at (mod_verlet.F90:1466)
  function B200_Calc_Field_at(pos_xyz) result(field)
    double precision, dimension(1:3), intent(in) :: pos_xyz
    double precision, dimension(1:3)             :: field
    double precision                             :: p(3, 1), f(3, 1)
    p(:, 1) = pos_xyz
    call Check(rb2_field_batch(1_c_int, p, f), 'rb2_field_batch')
    field = f(:, 1)
  end function B200_Calc_Field_at

  ! The sweep of Sample_Elec_Position (mod_pair.F90:990-1011); the caller keeps its file writer (:1014-1035).
  subroutine B200_Sample_Elec_Nearest()
    integer(c_int), allocatable :: id0(:)
    allocate(id0(max(nrPart, 1)))
    particles_nearest_dist = 1000.0d0
    if (nrPart > 0) then
      call Check(rb2_nearest_electron(particles_nearest_dist, id0), 'rb2_nearest_electron')
      where (id0(1:nrPart) >= 0) particles_nearest_id(1:nrPart) = id0(1:nrPart) + 1
    end if
  end subroutine B200_Sample_Elec_Nearest

  subroutine B200_Particles_To_Device()
    call Check(rb2_field_window_open(), 'rb2_field_window_open')
  end subroutine B200_Particles_To_Device

  subroutine B200_Release_Device_Particles()
    call Check(rb2_field_window_close(), 'rb2_field_window_close')
  end subroutine B200_Release_Device_Particles

  !-----------------------------------------------------------------------------
  ! Collisions (mod_collisions.F90).  B200_Init_Collisions is called once after Read_Cross_Section (main.F90:397);
  ! B200_Do_Electron_Atom_Collisions replaces the body of Do_Electron_Atom_Collisions (mod_collisions.F90:30-76)
  ! for collision_mode 1 and 2 and feeds the reference's own writers.
  subroutine B200_Init_Collisions(cross_tot_energy, cross_tot_data, cross_ion_energy, cross_ion_data)
    double precision, dimension(:), contiguous, target, intent(in) :: cross_tot_energy, cross_tot_data
    double precision, dimension(:), contiguous, target, intent(in) :: cross_ion_energy, cross_ion_data
    type(rb2_collision_config) :: c
    c%collision_mode = collision_mode
    c%ion_life_time  = ion_life_time
    c%n_d            = n_d
    c%cyl_radius     = emitters_dim(1, 1)
    c%n_tot = size(cross_tot_energy); c%n_ion = size(cross_ion_energy)
    c%tot_energy = c_loc(cross_tot_energy); c%tot_data = c_loc(cross_tot_data)
    c%ion_energy = c_loc(cross_ion_energy); c%ion_data = c_loc(cross_ion_data)
    call Check(rb2_collisions_init(c), 'rb2_collisions_init')
  end subroutine B200_Init_Collisions

  subroutine B200_Do_Electron_Atom_Collisions(step, nrCollisions, nrIonizations, nrRecombinations)
    integer, intent(in)  :: step
    integer, intent(out) :: nrCollisions, nrIonizations, nrRecombinations
    type(rb2_collision_result) :: r
    type(rb2_recomb_record), allocatable :: rec(:)
    type(rb2_ionization_record), allocatable :: ion(:)
    double precision :: rnd
    integer(c_long_long) :: seed
    integer :: k, n, IFAIL
    call random_number(rnd)                       ! the device generator is keyed by one host draw per step
    seed = int(rnd*9.0d18, c_long_long)
    call Check(rb2_do_collisions(step, seed, r), 'rb2_do_collisions')
    nrCollisions = r%nrCollisions; nrIonizations = r%nrIonizations; nrRecombinations = r%nrRecombinations
    nrPart = r%counts%nrPart;  nrElec = r%counts%nrElec;  nrIon = r%counts%nrIon;  nrAtom = r%counts%nrAtom
    nrID = r%counts%nrID
    nrPart_remove = r%counts%nrPart_remove;  nrElec_remove = r%counts%nrElec_remove
    nrIon_remove = r%counts%nrIon_remove;    nrAtom_remove = r%counts%nrAtom_remove
    nrPart_remove_top = r%counts%nrPart_remove_top;  nrIon_remove_top = r%counts%nrIon_remove_top
    nrPart_remove_recom = r%nrPart_remove_recom
    nrElec_remove_recom = r%nrElec_remove_recom
    nrIon_remove_recom  = r%nrIon_remove_recom
    if (nrIonizations > 0) then
      allocate(ion(nrIonizations))
      call Check(rb2_get_ionization_records(nrIonizations, ion, n), 'rb2_get_ionization_records')
      do k = 1, min(n, nrIonizations)             ! Write_Ionization_Data, mod_pair.F90:930-935
        write(unit=ud_ionization_data) ion(k)%step, ion(k)%pos, ion(k)%in_speed, ion(k)%out_speed, ion(k)%new_speed, &
          & 0.0d0, 0.0d0, ion(k)%in_slot + 1, ion(k)%new_id, ion(k)%ion_id, ion(k)%elec_emit
      end do
    end if
    if (nrRecombinations > 0) then
      allocate(rec(nrRecombinations))
      call Check(rb2_get_recombination_records(nrRecombinations, rec, n), 'rb2_get_recombination_records')
      do k = 1, min(n, nrRecombinations)
        ! what the two Mark_Particles_Remove(.., remove_recom) calls write (mod_pair.F90:265-272, :308-315)
        write(unit=ud_density_absorb_recom) rec(k)%ion_pos, rec(k)%ion_emit, rec(k)%ion_sec, rec(k)%ion_id, species_ion, &
          & cur_time/time_step*time_scale
        write(unit=ud_density_absorb_recom) rec(k)%elec_pos, rec(k)%elec_emit, rec(k)%elec_sec, rec(k)%elec_id, species_elec, &
          & cur_time/time_step*time_scale
        ! Write_Recombination_Data, mod_pair.F90:919-926
        write(unit=ud_recombination_data) rec(k)%step, rec(k)%ion_pos, rec(k)%elec_speed, rec(k)%dist, rec(k)%recom_rad, &
          & rec(k)%elec_slot + 1, rec(k)%ion_slot + 1, rec(k)%elec_emit, rec(k)%ion_life
      end do
    end if
    write(ud_coll, '(i6,tr2,i6,tr2,i6,tr2,i6)', iostat=IFAIL) step, nrCollisions, nrIonizations, nrRecombinations
  end subroutine B200_Do_Electron_Atom_Collisions

end module mod_b200_bridge
