!-------------------------------------------------------------------------------
! maniac_gpu_iface.f90 -- ISO_C_BINDING view of include/maniac_gpu.h
!
! Drop this file into MANIAC's src/ directory.  It declares the C ABI of the
! sm_100a energy engine (libmaniac_gpu) and the small marshalling helpers the
! replacement bodies in energy_glue_gpu.f90 use.  Nothing here computes an
! energy: with the library missing the link fails, and with no CUDA device
! mgpu_init returns non-zero and the run aborts through abort_run -- there is
! no CPU fallback on this path.
!
! NOTE: the build image of this repository has no Fortran compiler, so this
! file is shipped as source (never compiled here) and its calling conventions are exercised by tests/c_abi_driver.c
! (built and run by tests/test_c_abi_driver.py),
! which makes the same calls with the same argument conventions (by-value
! scalars, contiguous column-major arrays, 0-based ids after the shim's -1).
!-------------------------------------------------------------------------------
module maniac_gpu_iface

    use, intrinsic :: iso_c_binding
    use, intrinsic :: iso_fortran_env, only: real64
    implicit none

    integer(c_int), parameter :: MGPU_MAX_RES = 8, MGPU_MAX_SITES = 16
    integer(c_int), parameter :: MGPU_KIND_MOVE = 0, MGPU_KIND_CREATE = 1, MGPU_KIND_DELETE = 2, MGPU_KIND_SWAP = 3

    ! mgpu_residue, include/maniac_gpu.h:61-73
    type, bind(C) :: mgpu_residue
        integer(c_int32_t) :: natom, is_active, nmol, capacity
        type(c_ptr) :: charges, types, com, offset
        real(c_double) :: mass, fugacity, chemical_potential
    end type mgpu_residue

    ! mgpu_system, include/maniac_gpu.h:77-93
    type, bind(C) :: mgpu_system
        real(c_double) :: matrix(9), lo(3)
        integer(c_int32_t) :: nres
        type(c_ptr) :: residues
        integer(c_int32_t) :: ntypes
        type(c_ptr) :: epsilon, sigma
        real(c_double) :: temperature, ewald_tolerance, real_space_cutoff
        real(c_double) :: translation_step, rotation_step_angle
        real(c_double) :: p_translation, p_rotation, p_swap, p_insertion_deletion, p_widom
        integer(c_int32_t) :: n_walkers, device
    end type mgpu_system

    interface
        integer(c_int) function mgpu_init(sys) bind(C, name="mgpu_init")
            import :: c_int, mgpu_system
            type(mgpu_system), intent(in) :: sys
        end function
        subroutine mgpu_finalize() bind(C, name="mgpu_finalize")
        end subroutine
        type(c_ptr) function mgpu_last_error() bind(C, name="mgpu_last_error")
            import :: c_ptr
        end function
        integer(c_int) function mgpu_set_molecule(walker, res, mol, com, offset) bind(C, name="mgpu_set_molecule")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res, mol
            real(c_double), intent(in) :: com(3), offset(*)
        end function
        integer(c_int) function mgpu_set_count(walker, res, n) bind(C, name="mgpu_set_count")
            import :: c_int, c_int32_t
            integer(c_int32_t), value :: walker, res, n
        end function
        integer(c_int) function mgpu_total_energy(walker, out) bind(C, name="mgpu_total_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker
            real(c_double), intent(out) :: out(6)
        end function
        integer(c_int) function mgpu_pairwise_energy_for_molecule(walker, res, mol, skip_ordering_check, com, offset, &
                e_non_coulomb, e_coulomb) bind(C, name="mgpu_pairwise_energy_for_molecule")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res, mol, skip_ordering_check
            real(c_double), intent(in) :: com(3), offset(*)
            real(c_double), intent(out) :: e_non_coulomb, e_coulomb
        end function
        integer(c_int) function mgpu_ewald_self_energy_single_mol(res, e) bind(C, name="mgpu_ewald_self_energy_single_mol")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: res
            real(c_double), intent(out) :: e
        end function
        integer(c_int) function mgpu_intra_res_real_coulomb_energy(walker, res, mol, com, offset, e) &
                bind(C, name="mgpu_intra_res_real_coulomb_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res, mol
            real(c_double), intent(in) :: com(3), offset(*)
            real(c_double), intent(out) :: e
        end function
        integer(c_int) function mgpu_reciprocal_ewald_energy(walker, e) bind(C, name="mgpu_reciprocal_ewald_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker
            real(c_double), intent(out) :: e
        end function
        integer(c_int) function mgpu_old_energy(walker, res, mol, kind, out) bind(C, name="mgpu_old_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res, mol, kind
            real(c_double), intent(out) :: out(6)
        end function
        integer(c_int) function mgpu_new_energy(walker, res, mol, kind, com, offset, out) bind(C, name="mgpu_new_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res, mol, kind
            real(c_double), intent(in) :: com(3), offset(*)
            real(c_double), intent(out) :: out(6)
        end function
        ! attempt_swap_move (src/swapping.f90:59-88): both energy calls of the swap in one
        integer(c_int) function mgpu_swap_energy(walker, res_old, mol_old, res_new, com, offset, e_old, e_new) &
                bind(C, name="mgpu_swap_energy")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker, res_old, mol_old, res_new
            real(c_double), intent(in) :: com(3), offset(*)
            real(c_double), intent(out) :: e_old(6), e_new(6)
        end function
        ! one block of monte_carlo_loop for many walkers, host records in / out (src/monte_carlo.f90:40-118)
        integer(c_int) function mgpu_block(first_walker, n_walkers, n_steps, blob_in, offsets_in, &
                blob_out, capacity_out, offsets_out) bind(C, name="mgpu_block")
            import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
            integer(c_int32_t), value :: first_walker, n_walkers
            integer(c_int64_t), value :: n_steps, capacity_out
            type(c_ptr), value :: blob_in, offsets_in       ! c_null_ptr = continue from the device-resident state
            real(c_double), intent(out) :: blob_out(*)
            integer(c_int64_t), intent(out) :: offsets_out(*)
        end function
        integer(c_int) function mgpu_adjust_move_step_sizes(first_walker, n_walkers) bind(C, name="mgpu_adjust_move_step_sizes")
            import :: c_int, c_int32_t
            integer(c_int32_t), value :: first_walker, n_walkers
        end function
        integer(c_int) function mgpu_commit(walker) bind(C, name="mgpu_commit")
            import :: c_int, c_int32_t
            integer(c_int32_t), value :: walker
        end function
        integer(c_int) function mgpu_rollback(walker) bind(C, name="mgpu_rollback")
            import :: c_int, c_int32_t
            integer(c_int32_t), value :: walker
        end function
        integer(c_int) function mgpu_get_Ak(walker, re_im) bind(C, name="mgpu_get_Ak")
            import :: c_int, c_int32_t, c_double
            integer(c_int32_t), value :: walker
            real(c_double), intent(out) :: re_im(*)
        end function
        integer(c_int) function mgpu_widom_batch(walker, res, first_id, n, seed, dE_out, sum_w, n_ok) &
                bind(C, name="mgpu_widom_batch")
            import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
            integer(c_int32_t), value :: walker, res
            integer(c_int64_t), value :: first_id, n, seed
            type(c_ptr), value :: dE_out            ! c_null_ptr = do not return per-insertion dE
            real(c_double), intent(out) :: sum_w
            integer(c_int64_t), intent(out) :: n_ok
        end function
        integer(c_int) function mgpu_sweep(first_walker, n_walkers, n_steps, trace_walker, trace) bind(C, name="mgpu_sweep")
            import :: c_int, c_int32_t, c_int64_t, c_ptr
            integer(c_int32_t), value :: first_walker, n_walkers, trace_walker
            integer(c_int64_t), value :: n_steps
            type(c_ptr), value :: trace
        end function
        integer(c_int) function mgpu_seed(seed) bind(C, name="mgpu_seed")
            import :: c_int, c_int64_t
            integer(c_int64_t), value :: seed
        end function
    end interface

contains

    !> Map a non-zero return code onto the reference's error convention
    !> (abort_run, src/output_utils.f90:581-605: banner + `stop code`).
    subroutine mgpu_check(rc)
        use output_utils, only: abort_run
        integer(c_int), intent(in) :: rc
        character(kind=c_char), pointer :: cmsg(:)
        character(len=512) :: msg
        integer :: i
        if (rc == 0) return
        call c_f_pointer(mgpu_last_error(), cmsg, [512])
        msg = ""
        do i = 1, 512
            if (cmsg(i) == c_null_char) exit
            msg(i:i) = cmsg(i)
        end do
        call abort_run(trim(msg), int(rc))
    end subroutine mgpu_check

end module maniac_gpu_iface
