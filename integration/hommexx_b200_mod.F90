! hommexx_b200_mod — the Fortran side of include/hommexx_b200.h section B: what a HOMME build adds next to
! prim_cxx_driver_mod.F90 to run on libhommexx_b200 instead of the Kokkos functors. Every other binding
! (init_*_c, prim_run_subcycle_c, cxx_push_results_to_f90, f90_push_forcing_to_cxx, ...) is the reference's own
! interface block, unchanged (src/prim_main.F90:55-82, src/share/prim_driver_mod.F90:63,143,586-697,1328-1345,
! src/share/prim_cxx_driver_mod.F90:27-110).
!
! NOT compiled in this repository (the build image has no Fortran compiler); INTEGRATION.md walks through it.
module hommexx_b200_mod
  use iso_c_binding, only: c_int, c_char, c_int64_t, c_double, c_ptr
  implicit none
  private
  public :: hommexx_b200_wire_gpus, hxx_held_suarez_forcing

  interface
    ! int hommexx_b200_nccl_unique_id(void* out128)
    function hommexx_b200_nccl_unique_id(id) bind(c, name="hommexx_b200_nccl_unique_id") result(ierr)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
      integer(kind=c_int) :: ierr
    end function
    ! void hommexx_b200_set_comm(int rank, int size, int device, const void* nccl_unique_id)
    subroutine hommexx_b200_set_comm(rank, nranks, device, id) bind(c, name="hommexx_b200_set_comm")
      import :: c_int, c_char
      integer(kind=c_int), value :: rank, nranks, device
      character(kind=c_char), intent(in) :: id(128)
    end subroutine
    ! int64_t hommexx_b200_launch_count(void)  — kernels launched so far (0 would mean "no GPU path")
    function hommexx_b200_launch_count() bind(c, name="hommexx_b200_launch_count") result(n)
      import :: c_int64_t
      integer(kind=c_int64_t) :: n
    end function
    ! void hxx_held_suarez_forcing(const double* lat, const double* hyam, const double* hybm)   (section C)
    ! Held-Suarez FM / FT evaluated on the device from the dycore's own state at n0, in place of hs_forcing +
    ! f90_push_forcing_to_cxx: call before prim_run_subcycle_c with ftype = 0. lat = elem(:)%spherep(:,:)%lat packed
    ! [np, np, nelemd] (read on the first call only), hyam / hybm = hvcoord%hyam, hvcoord%hybm.
    subroutine hxx_held_suarez_forcing(lat, hyam, hybm) bind(c, name="hxx_held_suarez_forcing")
      import :: c_double
      real(kind=c_double), intent(in) :: lat(*), hyam(*), hybm(*)
    end subroutine
  end interface

contains

  ! Call once per MPI rank BEFORE initialize_hommexx_session (prim_driver_mod.F90:143): replaces the GPU
  ! round-robin of ExecSpaceDefs.cpp:36-50 and gives the library the NCCL communicator its DSS halo uses.
  subroutine hommexx_b200_wire_gpus(rank, nranks, gpus_per_node, mpi_comm)
    integer, intent(in) :: rank, nranks, gpus_per_node, mpi_comm
    character(kind=c_char) :: nccl_id(128)
    integer :: ierr
    include 'mpif.h'
    nccl_id = c_char_'0'
    if (nranks > 1) then
      if (rank == 0) then
        if (hommexx_b200_nccl_unique_id(nccl_id) /= 0) stop 'hommexx_b200: ncclGetUniqueId failed'
      end if
      call MPI_Bcast(nccl_id, 128, MPI_BYTE, 0, mpi_comm, ierr)
    end if
    call hommexx_b200_set_comm(int(rank, c_int), int(nranks, c_int), int(mod(rank, gpus_per_node), c_int), nccl_id)
  end subroutine

end module hommexx_b200_mod
