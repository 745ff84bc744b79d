!-------------------------------------------------------------------------------
! energy_glue_gpu.f90 -- GPU bodies for the energy seam of MANIAC.
!
! The move drivers (src/translation.f90, rotation.f90, creation.f90,
! deletion.f90, swapping.f90, widom.f90) keep calling
!     compute_old_energy / compute_new_energy        (monte_carlo_utils.f90:300-423)
!     accept_* / reject_* / save_molecule_state      (monte_carlo_utils.f90:429-505,579-591)
!     update_system_energy                           (energy_utils.f90:22-39)
! with their current signatures.  A maintainer replaces the BODIES of those
! routines by the ones below (or renames the originals *_cpu and `use`s this
! module first).  The CPU energy routines are no longer called on this path.
!
! Division of labour
!   host (Fortran): RNG, proposals (guest%com / guest%offset are still mutated
!                   by propose_translation_move, apply_random_rotation,
!                   insert_and_orient_molecule), acceptance rule, counters, I/O
!   device        : every energy term, S(k) (= ewald%Ak), commit / rollback
!
! The device holds the last COMMITTED configuration; a proposal never touches
! it.  compute_new_energy therefore passes the proposed geometry explicitly
! and leaves a pending trial; accept -> mgpu_commit, reject -> mgpu_rollback.
! save_single_mol_fourier_terms / restore_single_mol_fourier /
! replace_fourier_terms_single_mol (ewald_phase.f90:17-175) become no-ops.
!-------------------------------------------------------------------------------
module energy_glue_gpu

    use, intrinsic :: iso_c_binding
    use, intrinsic :: iso_fortran_env, only: real64
    use simulation_state
    use maniac_gpu_iface
    implicit none

    integer(c_int32_t), parameter :: WALKER0 = 0   ! the Fortran host drives one walker

contains

    pure integer(c_int32_t) function kind_of(is_creation, is_deletion)
        logical, intent(in), optional :: is_creation, is_deletion
        kind_of = MGPU_KIND_MOVE
        if (present(is_creation)) then
            if (is_creation) kind_of = MGPU_KIND_CREATE
        end if
        if (present(is_deletion)) then
            if (is_deletion) kind_of = MGPU_KIND_DELETE
        end if
    end function kind_of

    subroutine unpack_energy(buf, e)
        real(c_double), intent(in) :: buf(6)
        type(energy_type), intent(out) :: e
        ! field order of energy_type, src/simulation_state.f90:61-69
        e%non_coulomb = buf(1); e%coulomb = buf(2); e%recip_coulomb = buf(3)
        e%ewald_self = buf(4); e%intra_coulomb = buf(5); e%total = buf(6)
    end subroutine unpack_energy

    !> body of compute_old_energy (monte_carlo_utils.f90:367-423)
    subroutine compute_old_energy(res_type, mol_index, is_creation, is_deletion)
        integer, intent(in) :: res_type, mol_index
        logical, intent(in), optional :: is_creation, is_deletion
        real(c_double) :: buf(6)
        call mgpu_check(mgpu_old_energy(WALKER0, int(res_type - 1, c_int32_t), int(mol_index - 1, c_int32_t), &
                                        kind_of(is_creation, is_deletion), buf))
        call unpack_energy(buf, old)
    end subroutine compute_old_energy

    !> body of compute_new_energy (monte_carlo_utils.f90:300-361)
    subroutine compute_new_energy(res_type, mol_index, is_creation, is_deletion)
        integer, intent(in) :: res_type, mol_index
        logical, intent(in), optional :: is_creation, is_deletion
        real(c_double) :: buf(6), com(3), off(3, MGPU_MAX_SITES)
        integer :: n
        n = res%atom(res_type)
        com = guest%com(:, res_type, mol_index)
        off = 0.0_real64
        off(:, 1:n) = guest%offset(:, res_type, mol_index, 1:n)     ! contiguous temporary, atom-major
        call mgpu_check(mgpu_new_energy(WALKER0, int(res_type - 1, c_int32_t), int(mol_index - 1, c_int32_t), &
                                        kind_of(is_creation, is_deletion), com, off, buf))
        call unpack_energy(buf, new)
    end subroutine compute_new_energy

    !> tail of accept_molecule_move / accept_creation_move / accept_deletion_move:
    !> the running totals stay on the host exactly as in the reference; the device
    !> applies the same update to its copy (coordinates, S(k), counts, energies).
    subroutine gpu_accept()
        call mgpu_check(mgpu_commit(WALKER0))
    end subroutine gpu_accept

    !> tail of reject_molecule_move / reject_creation_move / reject_deletion_move
    !> (replaces restore_single_mol_fourier, ewald_phase.f90:74-125)
    subroutine gpu_reject()
        call mgpu_check(mgpu_rollback(WALKER0))
    end subroutine gpu_reject

    !> body of update_system_energy (energy_utils.f90:22-39)
    subroutine update_system_energy(box)
        type(type_box), intent(inout) :: box
        real(c_double) :: buf(6)
        call mgpu_check(mgpu_total_energy(WALKER0, buf))
        call unpack_energy(buf, energy)
    end subroutine update_system_energy

    !> attempt_swap_move (swapping.f90:59-88): the two energy calls of a swap and what lies between them.  Called in
    !> swapping.f90 after insert_and_orient_molecule(residue_type_bis, molecule_index_bis, ..., place_random_com=.false.)
    !> has written the new molecule; replaces compute_old_energy(..., is_deletion) + remove_molecule + the count updates
    !> + compute_new_energy(..., is_creation).  accept_swap_move -> gpu_accept, reject_swap_move -> gpu_reject.
    subroutine compute_swap_energies(res_type, mol_index, res_type_bis, mol_index_bis)
        integer, intent(in) :: res_type, mol_index, res_type_bis, mol_index_bis
        real(c_double) :: buf_old(6), buf_new(6), com(3), off(3, MGPU_MAX_SITES)
        integer :: n
        n = res%atom(res_type_bis)
        com = guest%com(:, res_type_bis, mol_index_bis)
        off = 0.0_real64
        off(:, 1:n) = guest%offset(:, res_type_bis, mol_index_bis, 1:n)
        call mgpu_check(mgpu_swap_energy(WALKER0, int(res_type - 1, c_int32_t), int(mol_index - 1, c_int32_t), &
                                         int(res_type_bis - 1, c_int32_t), com, off, buf_old, buf_new))
        call unpack_energy(buf_old, old)
        call unpack_energy(buf_new, new)
    end subroutine compute_swap_energies

    !> one block of monte_carlo_loop (monte_carlo.f90:40-118) for n_walkers device-resident walkers: nb_step MC steps with
    !> the device-side drivers, then the walkers' records (counts, energies, counters, coordinates, ewald%Ak) in HOST memory
    !> for the per-block writers (write_utils.f90:16-34), then adjust_move_step_sizes (monte_carlo_utils.f90:98-134).
    subroutine gpu_block(n_walkers, nb_step, blob, capacity, offsets)
        integer, intent(in) :: n_walkers, nb_step
        integer(c_int64_t), intent(in) :: capacity
        real(c_double), intent(out) :: blob(*)
        integer(c_int64_t), intent(out) :: offsets(*)
        call mgpu_check(mgpu_block(0_c_int32_t, int(n_walkers, c_int32_t), int(nb_step, c_int64_t), c_null_ptr, c_null_ptr, &
                                   blob, capacity, offsets))
        if (mc_input%recalibrate_moves) call mgpu_check(mgpu_adjust_move_step_sizes(0_c_int32_t, int(n_walkers, c_int32_t)))
    end subroutine gpu_block

    !> widom.f90: n independent test insertions in one launch instead of n calls
    !> of widom_trial; accumulates statistic%weight / statistic%sample (:74-92).
    subroutine widom_batch(res_type, n, seed)
        integer, intent(in) :: res_type
        integer(c_int64_t), intent(in) :: n, seed
        real(c_double) :: sum_w
        integer(c_int64_t) :: n_ok
        call mgpu_check(mgpu_widom_batch(WALKER0, int(res_type - 1, c_int32_t), 0_c_int64_t, n, seed, c_null_ptr, sum_w, n_ok))
        statistic%weight(res_type) = statistic%weight(res_type) + sum_w
        statistic%sample(res_type) = statistic%sample(res_type) + int(n)
    end subroutine widom_batch

end module energy_glue_gpu
