! mskc_bind.f90 -- ISO_C_BINDING interface of libmskc.so (include/mskc.h) for the Fortran host.
!
! This is the binding a MulSKIPS maintainer adds to mulskips-source/ so that main.f calls the CUDA event
! loop instead of mcevent/*.f: one module, arrays passed with c_loc exactly as the socket shim already does
! (mulskips-source-LaserAnnealing/sockets/fsockets.f:52-64,165-171).  See INTEGRATION.md for the edit of
! main.f:297-379.  Not compiled in this repository's image (no Fortran compiler there); kept in step with
! include/mskc.h by tests/test_abi.py::test_fortran_binding_names.
module mskc_bind
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int), parameter :: MSKC_STOP_BUDGET = 0, MSKC_STOP_MAXITER = 1, MSKC_STOP_MAXTIME = 2, &
                               MSKC_STOP_FRAME = 3, MSKC_STOP_FEWMC = 4, MSKC_STOP_RESOLIDIFIED = 5, MSKC_STOP_ERROR = -1
  integer(c_int), parameter :: MSKC_TREE_SRC = 0, MSKC_TREE_LA = 1

  type, bind(C) :: mskc_parity_args
    integer(c_int64_t) :: max_iter
    real(c_double)     :: max_time
    integer(c_int32_t) :: exit_on_time
    integer(c_int32_t) :: pad
    integer(c_int64_t) :: out_every
    real(c_double)     :: out_time
    integer(c_int64_t) :: framecounter
    integer(c_int64_t) :: event_budget
  end type

  type, bind(C) :: mskc_parity_result
    integer(c_int64_t) :: iter, iter_out
    real(c_double)     :: time
    integer(c_int32_t) :: last_site(3)
    integer(c_int32_t) :: stop_reason
    integer(c_int32_t) :: num_atoms, num_voids, num_adatom, pad
    integer(c_int64_t) :: events_done, site_updates
    real(c_double)     :: tree_root
    real(c_float)      :: kernel_ms
    integer(c_int32_t) :: launches
  end type

  type, bind(C) :: mskc_parallel_args
    integer(c_int64_t) :: n_sweeps
    real(c_double)     :: events_per_tile, tau_fixed
    integer(c_int64_t) :: seed
    integer(c_int32_t) :: tile, flags
  end type

  type, bind(C) :: mskc_parallel_result
    integer(c_int64_t) :: events_done, site_updates, sweeps_done
    real(c_double)     :: time, total_rate
    integer(c_int64_t) :: n_active
    integer(c_int64_t) :: count_dep(3), count_evap, count_abs, count_des
    real(c_float)      :: kernel_ms
    integer(c_int32_t) :: launches
    real(c_float)      :: sweep_kernel_ms
    integer(c_int32_t) :: stop_reason
    integer(c_int32_t) :: sampler_launches, max_tile_events
    real(c_double) :: tau
  end type

  type, bind(C) :: mskc_counters
    integer(c_int32_t) :: num_atoms, num_voids, num_adatom, num_adatom_occ
    integer(c_int32_t) :: CountCrystal(3), CountCov(3)
    integer(c_int64_t) :: iter
    real(c_double)     :: time
  end type

  interface
    ! replaces the static arrays of modules/defsystem.f:27-68 and mcevent/AllocateArrays.f
    integer(c_int) function mskc_create(lenx, leny, lenz, ncrystal, ncov, listcrystal, listcov, device_id, nummcmax, h) bind(C, name="mskc_create")
      import :: c_int, c_ptr
      integer(c_int), value :: lenx, leny, lenz, ncrystal, ncov, device_id, nummcmax
      type(c_ptr), value :: listcrystal, listcov
      type(c_ptr) :: h
    end function
    integer(c_int) function mskc_destroy(h) bind(C, name="mskc_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    type(c_ptr) function mskc_last_error(h) bind(C, name="mskc_last_error")
      import :: c_ptr
      type(c_ptr), value :: h
    end function
    ! after CALL SetProbability() (main.f:218): the module arrays PtransE, PtransD, PtransAbs, PtransDes as they are
    integer(c_int) function mskc_set_rates(h, PtransE, PtransD, PtransAbs, PtransDes, PtransZig) bind(C, name="mskc_set_rates")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h, PtransE, PtransD, PtransAbs, PtransDes
      real(c_double), value :: PtransZig
    end function
    integer(c_int) function mskc_set_la(h, enable, Tm, P0, sigma, Tflex, tmin, tmax) bind(C, name="mskc_set_la")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: enable, tmin, tmax
      real(c_double), value :: Tm, P0, sigma, Tflex
    end function
    ! InitSt 'ST' (main.f:215-216): after SetProbabilityFromField, instead of mskc_set_rates.  The ref* module arrays have
    ! the temperature index FIRST (defsystem.f:77-81); the C side wants one contiguous block per temperature, so pass
    ! reshape(refPtransE(1:ntemp,:,:,:,:,:,:,:), order=...) copies or fill a (12288,ntemp) buffer block by block.
    integer(c_int) function mskc_set_rate_field(h, ntemp, reftemp, refPE, refPD, refPAbs, refPDes, PtransZig, tmin, tmax, &
                                                fixed_temperature) bind(C, name="mskc_set_rate_field")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h, reftemp, refPE, refPD, refPAbs, refPDes
      integer(c_int), value :: ntemp, tmin, tmax, fixed_temperature
      real(c_double), value :: PtransZig
    end function
    ! mulskips-source-LaserAnnealing/main.f: once after mskc_create (species at bits 28-29, event codes 1-2/3/4/4+k) ...
    integer(c_int) function mskc_set_tree_variant(h, variant) bind(C, name="mskc_set_tree_variant")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: variant
    end function
    ! ... and after SetCAD (la/main.f:578-611): Tm, P0, sigma, Tflex(:NCrystal), InitCountCrystal, NumOcc0, NumAdAtomOcc
    integer(c_int) function mskc_set_la_fork(h, enable, Tm, P0, sigma, Tflex, tmin, tmax, InitCountCrystal, num_occ0, num_adatom_occ) &
        bind(C, name="mskc_set_la_fork")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, Tm, P0, sigma, Tflex, InitCountCrystal
      integer(c_int), value :: enable, tmin, tmax, num_occ0, num_adatom_occ
    end function
    ! after the Set* builder (main.f:225-295): LattCoo and the three lists flattened to int32 arrays
    integer(c_int) function mskc_upload_state(h, LattCoo, n_mc, mc_xyz, mc_nn, mc_event, mc_rate, n_bulk, bulk_xyz, bulk_nn, &
                                              n_void, void_xyz, void_nn) bind(C, name="mskc_upload_state")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, LattCoo, mc_xyz, mc_nn, mc_event, mc_rate, bulk_xyz, bulk_nn, void_xyz, void_nn
      integer(c_int), value :: n_mc, n_bulk, n_void
    end function
    integer(c_int) function mskc_init_flat(h, len1, walls, two_species) bind(C, name="mskc_init_flat")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: len1, walls, two_species
    end function
    integer(c_int) function mskc_seed(h, idum, runtype, counter0, ran3_state) bind(C, name="mskc_seed")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, ran3_state
      integer(c_int), value :: idum, runtype
      integer(c_int64_t), value :: counter0
    end function
    integer(c_int) function mskc_get_rng(h, ran3_state, counter, idum) bind(C, name="mskc_get_rng")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, ran3_state, counter, idum
    end function
    integer(c_int) function mskc_set_clock(h, iter, time, CountCrystal, CountCov) bind(C, name="mskc_set_clock")
      import :: c_int, c_ptr, c_int64_t, c_double
      type(c_ptr), value :: h, CountCrystal, CountCov
      integer(c_int64_t), value :: iter
      real(c_double), value :: time
    end function
    ! the DO loop main.f:304-371 between two output frames
    integer(c_int) function mskc_run_parity(h, a, r) bind(C, name="mskc_run_parity")
      import :: c_int, c_ptr, mskc_parity_args, mskc_parity_result
      type(c_ptr), value :: h
      type(mskc_parity_args), intent(in) :: a
      type(mskc_parity_result), intent(out) :: r
    end function
    integer(c_int) function mskc_run_parallel(h, a, r) bind(C, name="mskc_run_parallel")
      import :: c_int, c_ptr, mskc_parallel_args, mskc_parallel_result
      type(c_ptr), value :: h
      type(mskc_parallel_args), intent(in) :: a
      type(mskc_parallel_result), intent(out) :: r
    end function
    integer(c_int) function mskc_rerate_all(h, n_sites, total_rate, kernel_ms) bind(C, name="mskc_rerate_all")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, n_sites, total_rate, kernel_ms
    end function
    ! SetT (initial/SetT.f:25-50, la/main.f:618): b(LenZ, LenX*LenY) int32 kelvin, 10 = wall
    integer(c_int) function mskc_set_temperature(h, T) bind(C, name="mskc_set_temperature")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, T
    end function
    ! what WriteMolMolXYZFile, the checkpoint writer (main.f:385-403) and the Coor dump (main.f:407-430) read
    integer(c_int) function mskc_download_lattice(h, LattCoo) bind(C, name="mskc_download_lattice")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, LattCoo
    end function
    integer(c_int) function mskc_patch_lattice(h, n, cell, word) bind(C, name="mskc_patch_lattice")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, cell, word
      integer(c_int), value :: n
    end function
    integer(c_int) function mskc_get_counters(h, c) bind(C, name="mskc_get_counters")
      import :: c_int, c_ptr, mskc_counters
      type(c_ptr), value :: h
      type(mskc_counters), intent(out) :: c
    end function
    integer(c_int) function mskc_download_lists(h, mc_xyz, mc_nn, mc_event, mc_rate, void_xyz, void_nn, bulk_xyz, bulk_nn, &
                                                bulk_cap, n_bulk_found) bind(C, name="mskc_download_lists")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, mc_xyz, mc_nn, mc_event, mc_rate, void_xyz, void_nn, bulk_xyz, bulk_nn, n_bulk_found
      integer(c_int64_t), value :: bulk_cap
    end function
    integer(c_int) function mskc_state_hash(h, hash) bind(C, name="mskc_state_hash")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, hash
    end function
    integer(c_int) function mskc_census_lattice(h, c, kernel_ms) bind(C, name="mskc_census_lattice")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, c, kernel_ms
    end function
    integer(c_int) function mskc_set_temperature_u16(h, T) bind(C, name="mskc_set_temperature_u16")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, T
    end function
    integer(c_int) function mskc_download_phases(h, phases) bind(C, name="mskc_download_phases")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, phases
    end function
    integer(c_int) function mskc_last_dense_ms(h, ms) bind(C, name="mskc_last_dense_ms")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, ms
    end function
    integer(c_int) function mskc_replay_teams(h, n, site_xyz, event, arm, tree_root, n_iterations, n_shared) bind(C, name="mskc_replay_teams")
      import :: c_int, c_ptr, c_int64_t, c_double
      type(c_ptr), value :: h, site_xyz, event, arm, n_iterations, n_shared
      integer(c_int64_t), value :: n
      real(c_double), value :: tree_root
    end function
    integer(c_int) function mskc_check_flags(h, n_sites, n_flag_mismatch, n_coor_mismatch, kernel_ms) bind(C, name="mskc_check_flags")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, n_sites, n_flag_mismatch, n_coor_mismatch, kernel_ms
    end function
    integer(c_int) function mskc_debug_tile_stats(h, out, max_tiles, ntiles) bind(C, name="mskc_debug_tile_stats")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, out, ntiles
      integer(c_int64_t), value :: max_tiles
    end function
    integer(c_int) function mskc_replay_events(h, n, site_xyz, event, arm, tree_root) bind(C, name="mskc_replay_events")
      import :: c_int, c_ptr, c_int64_t, c_double
      type(c_ptr), value :: h, site_xyz, event, arm
      integer(c_int64_t), value :: n
      real(c_double), value :: tree_root
    end function
    integer(c_int) function mskc_download_active(h, cap, xyz, event, rate, n) bind(C, name="mskc_download_active")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, xyz, event, rate, n
      integer(c_int64_t), value :: cap
    end function
    integer(c_int) function mskc_download_bonds(h, n, xyz, nn) bind(C, name="mskc_download_bonds")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, xyz, nn
      integer(c_int64_t), value :: n
    end function
    ! several GPUs: x-decomposed slabs (no counterpart in the reference)
    integer(c_int) function mskc_create_slab(global_lenx, leny, lenz, ncrystal, ncov, listcrystal, listcov, rank, nranks, device_id, h) &
        bind(C, name="mskc_create_slab")
      import :: c_int, c_ptr
      integer(c_int), value :: global_lenx, leny, lenz, ncrystal, ncov, rank, nranks, device_id
      type(c_ptr), value :: listcrystal, listcov
      type(c_ptr) :: h
    end function
    integer(c_int) function mskc_comm_unique_id(id128) bind(C, name="mskc_comm_unique_id")
      import :: c_int, c_ptr
      type(c_ptr), value :: id128
    end function
    integer(c_int) function mskc_comm_connect(h, id128) bind(C, name="mskc_comm_connect")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, id128
    end function
    integer(c_int) function mskc_extract_frame(h, cap, cell, word, n, kernel_ms) bind(C, name="mskc_extract_frame")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, cell, word, n, kernel_ms
      integer(c_int64_t), value :: cap
    end function
    integer(c_int) function mskc_extract_frame_all(h, cap, cell, word, n, kernel_ms) bind(C, name="mskc_extract_frame_all")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, cell, word, n, kernel_ms
      integer(c_int64_t), value :: cap
    end function
    integer(c_int) function mskc_extract_antisites(h, nunlike, cap, cell, word, n, kernel_ms) bind(C, name="mskc_extract_antisites")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, cell, word, n, kernel_ms
      integer(c_int), value :: nunlike     ! 0 or 3 (the C prototype calls it target)
      integer(c_int64_t), value :: cap
    end function
    integer(c_int) function mskc_gather_lattice(h, n, cell, word) bind(C, name="mskc_gather_lattice")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h, cell, word
      integer(c_int64_t), value :: n
    end function
  end interface
end module mskc_bind
