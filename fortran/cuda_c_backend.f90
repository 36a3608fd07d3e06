!! cuda_c_backend.f90 -- the Fortran side of the drop-in boundary.
!!
!! A `cuda_c` backend for x3d2: `cuda_c_backend_t` extends the reference's abstract `base_backend_t`
!! (src/backend/backend.f90:13-62) and forwards every deferred procedure through `iso_c_binding` to the C ABI of
!! include/x3d2c.h (libx3d2c.so: hand-written CUDA C++ for sm_100a). With it solver.f90, vector_calculus.f90,
!! time_integrator.f90 and the cases run unchanged; the only edit to the reference is the backend selection in
!! src/xcompact.f90:15-22,87-107 (shown in INTEGRATION.md).
!!
!! Modelled on the reference's CUDA-Fortran backend (same roles, no device code on this side):
!!   cuda_c_field_t / cuda_c_allocator_t   <- src/backend/cuda/allocator.f90:9-90
!!   cuda_c_tdsops_t                       <- src/backend/cuda/tdsops.f90:9-90
!!   cuda_c_poisson_fft_t                  <- src/backend/cuda/poisson_fft.f90:37-95,183-399
!!   cuda_c_backend_t                      <- src/backend/cuda/backend.f90:42-152
!!
!! This image has no Fortran compiler (SURVEY.md F1), so this file cannot be compiled here; tests/test_fortran_shim.py
!! parses it and checks every `bind(c, name=...)` interface against include/x3d2c.h (name, number of arguments,
!! by-value vs by-reference). It uses only Fortran 2008 + iso_c_binding and the reference's public modules.
!!
!! Build (on a machine with gfortran or nvfortran + MPI):
!!   add this file to SRC in src/CMakeLists.txt, link x3d2 against libx3d2c.so, compile with -DCUDA_C.

module m_cuda_c_common
  implicit none
  !> Pencil-group width of the cuda_c layouts (X3D2C_SZ in include/x3d2c.h; role of src/backend/cuda/common.f90:4)
  integer, parameter :: SZ = 32
end module m_cuda_c_common

!=======================================================================================================================
module m_cuda_c_bindings
  !! Interfaces of the C ABI (include/x3d2c.h). Device fields, contexts and handles are opaque `type(c_ptr)` values.
  use iso_c_binding
  implicit none

  !> x3d2c_config (include/x3d2c.h)
  type, bind(c) :: x3d2c_config
    integer(c_int) :: dims_vert(3)
    integer(c_int) :: dims_cell(3)
    integer(c_int) :: dims_vert_global(3)
    integer(c_int) :: dims_cell_global(3)
    integer(c_int) :: nproc_dir(3)
    integer(c_int) :: nrank_dir(3)
    integer(c_int) :: n_offset(3)
    integer(c_int) :: pprev(3)
    integer(c_int) :: pnext(3)
    integer(c_int) :: periodic(3)
    integer(c_int) :: sz
    integer(c_int) :: rank
    integer(c_int) :: nproc
    integer(c_int) :: device
    integer(c_int) :: flags
    type(c_ptr) :: nccl_unique_id
  end type x3d2c_config

  interface
    function x3d2c_last_error() bind(c, name='x3d2c_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function x3d2c_last_error

    function x3d2c_nccl_unique_id(out128) bind(c, name='x3d2c_nccl_unique_id') result(ierr)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: out128(128)
      integer(c_int) :: ierr
    end function x3d2c_nccl_unique_id

    function x3d2c_create(cfg, ctx) bind(c, name='x3d2c_create') result(ierr)
      import :: c_int, c_ptr, x3d2c_config
      type(x3d2c_config), intent(in) :: cfg
      type(c_ptr), intent(out) :: ctx
      integer(c_int) :: ierr
    end function x3d2c_create

    function x3d2c_destroy(ctx) bind(c, name='x3d2c_destroy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: ierr
    end function x3d2c_destroy

    function x3d2c_sync(ctx) bind(c, name='x3d2c_sync') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: ierr
    end function x3d2c_sync

    function x3d2c_get_padded_dims(ctx, dims_padded, n_groups, ngrid) &
      bind(c, name='x3d2c_get_padded_dims') result(ierr)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr), value :: ctx
      integer(c_int), intent(out) :: dims_padded(3), n_groups(3)
      integer(c_long_long), intent(out) :: ngrid
      integer(c_int) :: ierr
    end function x3d2c_get_padded_dims

    ! ---- fields
    function x3d2c_field_alloc(ctx, dev) bind(c, name='x3d2c_field_alloc') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), intent(out) :: dev
      integer(c_int) :: ierr
    end function x3d2c_field_alloc

    function x3d2c_field_free(ctx, dev) bind(c, name='x3d2c_field_free') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, dev
      integer(c_int) :: ierr
    end function x3d2c_field_free

    function x3d2c_field_fill(ctx, dev, c) bind(c, name='x3d2c_field_fill') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, dev
      real(c_double), value :: c
      integer(c_int) :: ierr
    end function x3d2c_field_fill

    function x3d2c_copy_data_to_f(ctx, dev, host_data) bind(c, name='x3d2c_copy_data_to_f') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, dev
      real(c_double), intent(in) :: host_data(*)
      integer(c_int) :: ierr
    end function x3d2c_copy_data_to_f

    function x3d2c_copy_f_to_data(ctx, host_data, dev) bind(c, name='x3d2c_copy_f_to_data') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: host_data(*)
      type(c_ptr), value :: dev
      integer(c_int) :: ierr
    end function x3d2c_copy_f_to_data

    ! ---- tdsops
    function x3d2c_tdsops_create(ctx, n_tds, n_rhs, move, periodic, coeffs, coeffs_s, coeffs_e, dist_fw, dist_bw, &
                                 dist_sa, dist_sc, dist_af, stretch, stretch_correct, ops) &
      bind(c, name='x3d2c_tdsops_create') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: n_tds, n_rhs, move, periodic
      real(c_double), intent(in) :: coeffs(*), coeffs_s(*), coeffs_e(*)
      real(c_double), intent(in) :: dist_fw(*), dist_bw(*), dist_sa(*), dist_sc(*), dist_af(*)
      real(c_double), intent(in) :: stretch(*), stretch_correct(*)
      type(c_ptr), intent(out) :: ops
      integer(c_int) :: ierr
    end function x3d2c_tdsops_create

    function x3d2c_tdsops_destroy(ctx, ops) bind(c, name='x3d2c_tdsops_destroy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, ops
      integer(c_int) :: ierr
    end function x3d2c_tdsops_destroy

    ! ---- operators
    function x3d2c_transeq(ctx, dir, du, dv, dw, u, v, w, nu, der1st, der1st_sym, der2nd, der2nd_sym) &
      bind(c, name='x3d2c_transeq') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir
      type(c_ptr), value :: du, dv, dw, u, v, w
      real(c_double), value :: nu
      type(c_ptr), value :: der1st, der1st_sym, der2nd, der2nd_sym
      integer(c_int) :: ierr
    end function x3d2c_transeq

    function x3d2c_transeq_species(ctx, dir, dspec, uvw, spec, nu, der1st, der1st_sym, der2nd, sync) &
      bind(c, name='x3d2c_transeq_species') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir
      type(c_ptr), value :: dspec, uvw, spec
      real(c_double), value :: nu
      type(c_ptr), value :: der1st, der1st_sym, der2nd
      integer(c_int), value :: sync
      integer(c_int) :: ierr
    end function x3d2c_transeq_species

    function x3d2c_tds_solve(ctx, dir, du, u, ops) bind(c, name='x3d2c_tds_solve') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir
      type(c_ptr), value :: du, u, ops
      integer(c_int) :: ierr
    end function x3d2c_tds_solve

    function x3d2c_reorder(ctx, rdr, dst, src) bind(c, name='x3d2c_reorder') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: rdr
      type(c_ptr), value :: dst, src
      integer(c_int) :: ierr
    end function x3d2c_reorder

    function x3d2c_sum_yintox(ctx, u, u_y) bind(c, name='x3d2c_sum_yintox') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, u, u_y
      integer(c_int) :: ierr
    end function x3d2c_sum_yintox

    function x3d2c_sum_zintox(ctx, u, u_z) bind(c, name='x3d2c_sum_zintox') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, u, u_z
      integer(c_int) :: ierr
    end function x3d2c_sum_zintox

    function x3d2c_veccopy(ctx, dst, src) bind(c, name='x3d2c_veccopy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, dst, src
      integer(c_int) :: ierr
    end function x3d2c_veccopy

    function x3d2c_vecadd(ctx, a, x, b, y) bind(c, name='x3d2c_vecadd') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: a
      type(c_ptr), value :: x
      real(c_double), value :: b
      type(c_ptr), value :: y
      integer(c_int) :: ierr
    end function x3d2c_vecadd

    function x3d2c_vecmult(ctx, y, x) bind(c, name='x3d2c_vecmult') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, y, x
      integer(c_int) :: ierr
    end function x3d2c_vecmult

    function x3d2c_field_scale(ctx, f, a) bind(c, name='x3d2c_field_scale') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, f
      real(c_double), value :: a
      integer(c_int) :: ierr
    end function x3d2c_field_scale

    function x3d2c_field_shift(ctx, f, a) bind(c, name='x3d2c_field_shift') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, f
      real(c_double), value :: a
      integer(c_int) :: ierr
    end function x3d2c_field_shift

    function x3d2c_scalar_product(ctx, dir, data_loc, x, y, s) bind(c, name='x3d2c_scalar_product') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir, data_loc
      type(c_ptr), value :: x, y
      real(c_double), intent(out) :: s
      integer(c_int) :: ierr
    end function x3d2c_scalar_product

    function x3d2c_field_max_mean(ctx, dir, data_loc, f, max_val, mean_val) &
      bind(c, name='x3d2c_field_max_mean') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir, data_loc
      type(c_ptr), value :: f
      real(c_double), intent(out) :: max_val, mean_val
      integer(c_int) :: ierr
    end function x3d2c_field_max_mean

    function x3d2c_slice_max_sum(ctx, dir, data_loc, f, i_slice, max_val, sum_val) &
      bind(c, name='x3d2c_slice_max_sum') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: dir, data_loc
      type(c_ptr), value :: f
      integer(c_int), value :: i_slice
      real(c_double), intent(out) :: max_val, sum_val
      integer(c_int) :: ierr
    end function x3d2c_slice_max_sum

    function x3d2c_field_volume_integral(ctx, data_loc, f, s) bind(c, name='x3d2c_field_volume_integral') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: data_loc
      type(c_ptr), value :: f
      real(c_double), intent(out) :: s
      integer(c_int) :: ierr
    end function x3d2c_field_volume_integral

    function x3d2c_field_set_face(ctx, f, data_loc, c_start, c_end, face) &
      bind(c, name='x3d2c_field_set_face') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, f
      integer(c_int), value :: data_loc
      real(c_double), value :: c_start, c_end
      integer(c_int), value :: face
      integer(c_int) :: ierr
    end function x3d2c_field_set_face

    function x3d2c_field_set_face_from_field(ctx, f, f_start, data_loc, c_end, face, flow_rate_diff) &
      bind(c, name='x3d2c_field_set_face_from_field') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx, f, f_start
      integer(c_int), value :: data_loc
      real(c_double), value :: c_end
      integer(c_int), value :: face
      real(c_double), value :: flow_rate_diff
      integer(c_int) :: ierr
    end function x3d2c_field_set_face_from_field

    function x3d2c_compute_vorticity(ctx, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz) &
      bind(c, name='x3d2c_compute_vorticity') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz
      integer(c_int) :: ierr
    end function x3d2c_compute_vorticity

    function x3d2c_compute_qcriterion(ctx, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz) &
      bind(c, name='x3d2c_compute_qcriterion') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz
      integer(c_int) :: ierr
    end function x3d2c_compute_qcriterion

    ! ---- FFT Poisson
    function x3d2c_poisson_spec_layout(ctx, n_spec, n_sp_st) bind(c, name='x3d2c_poisson_spec_layout') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), intent(out) :: n_spec(3), n_sp_st(3)
      integer(c_int) :: ierr
    end function x3d2c_poisson_spec_layout

    function x3d2c_poisson_create(ctx, waves, ax, bx, ay, by, az, bz, p) &
      bind(c, name='x3d2c_poisson_create') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: waves(*), ax(*), bx(*), ay(*), by(*), az(*), bz(*)
      type(c_ptr), intent(out) :: p
      integer(c_int) :: ierr
    end function x3d2c_poisson_create

    function x3d2c_poisson_create_010(ctx, waves, ax, bx, ay, by, az, bz, stretched, a_odd_re, a_odd_im, &
                                      a_even_re, a_even_im, p) bind(c, name='x3d2c_poisson_create_010') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: waves(*), ax(*), bx(*), ay(*), by(*), az(*), bz(*)
      integer(c_int), value :: stretched
      type(c_ptr), value :: a_odd_re, a_odd_im, a_even_re, a_even_im
      type(c_ptr), intent(out) :: p
      integer(c_int) :: ierr
    end function x3d2c_poisson_create_010

    function x3d2c_poisson_destroy(ctx, p) bind(c, name='x3d2c_poisson_destroy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p
      integer(c_int) :: ierr
    end function x3d2c_poisson_destroy

    function x3d2c_fft_forward(ctx, p, f_c) bind(c, name='x3d2c_fft_forward') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_forward

    function x3d2c_fft_backward(ctx, p, f_c) bind(c, name='x3d2c_fft_backward') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_backward

    function x3d2c_fft_postprocess_000(ctx, p) bind(c, name='x3d2c_fft_postprocess_000') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p
      integer(c_int) :: ierr
    end function x3d2c_fft_postprocess_000

    function x3d2c_fft_postprocess_010(ctx, p) bind(c, name='x3d2c_fft_postprocess_010') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p
      integer(c_int) :: ierr
    end function x3d2c_fft_postprocess_010

    function x3d2c_fft_postprocess_100(ctx, p) bind(c, name='x3d2c_fft_postprocess_100') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p
      integer(c_int) :: ierr
    end function x3d2c_fft_postprocess_100

    function x3d2c_fft_postprocess_110(ctx, p) bind(c, name='x3d2c_fft_postprocess_110') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p
      integer(c_int) :: ierr
    end function x3d2c_fft_postprocess_110

    function x3d2c_fft_forward_100(ctx, p, f_c) bind(c, name='x3d2c_fft_forward_100') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_forward_100

    function x3d2c_fft_forward_110(ctx, p, f_c) bind(c, name='x3d2c_fft_forward_110') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_forward_110

    function x3d2c_fft_backward_100(ctx, p, f_c) bind(c, name='x3d2c_fft_backward_100') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_backward_100

    function x3d2c_fft_backward_110(ctx, p, f_c) bind(c, name='x3d2c_fft_backward_110') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_c
      integer(c_int) :: ierr
    end function x3d2c_fft_backward_110

    function x3d2c_enforce_periodicity_x(ctx, p, f_out, f_in) bind(c, name='x3d2c_enforce_periodicity_x') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_enforce_periodicity_x

    function x3d2c_undo_periodicity_x(ctx, p, f_out, f_in) bind(c, name='x3d2c_undo_periodicity_x') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_undo_periodicity_x

    function x3d2c_enforce_periodicity_y(ctx, p, f_out, f_in) bind(c, name='x3d2c_enforce_periodicity_y') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_enforce_periodicity_y

    function x3d2c_undo_periodicity_y(ctx, p, f_out, f_in) bind(c, name='x3d2c_undo_periodicity_y') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_undo_periodicity_y

    function x3d2c_enforce_periodicity_xy(ctx, p, f_out, f_in) bind(c, name='x3d2c_enforce_periodicity_xy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_enforce_periodicity_xy

    function x3d2c_undo_periodicity_xy(ctx, p, f_out, f_in) bind(c, name='x3d2c_undo_periodicity_xy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, p, f_out, f_in
      integer(c_int) :: ierr
    end function x3d2c_undo_periodicity_xy
  end interface

  !> Context of the (single) cuda_c backend of this rank. The allocator creates device blocks before the backend
  !> object exists (src/xcompact.f90:87-97), so the context is created by whoever needs it first.
  type(c_ptr), save :: x3d2c_ctx = c_null_ptr

contains

  subroutine x3d2c_check(ierr, where)
    !! The reference aborts on misuse (`error stop`); so does the shim, with the library's message.
    integer(c_int), intent(in) :: ierr
    character(*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: text
    integer :: i

    if (ierr == 0) return
    call c_f_pointer(x3d2c_last_error(), msg, [512])
    text = ''
    do i = 1, 512
      if (msg(i) == c_null_char) exit
      text(i:i) = msg(i)
    end do
    print *, 'cuda_c backend: ', where, ': ', trim(text)
    error stop 'cuda_c backend failed'
  end subroutine x3d2c_check

end module m_cuda_c_bindings

!=======================================================================================================================
module m_cuda_c_allocator
  !! Device blocks owned by libx3d2c.so (role of src/backend/cuda/allocator.f90). Pool semantics (get_block /
  !! release_block, src/allocator.f90:113-162) stay in the reference's allocator_t; only allocation and fill cross the ABI.
  use iso_c_binding
  use m_allocator, only: allocator_t
  use m_common, only: dp
  use m_field, only: field_t
  use m_cuda_c_bindings

  implicit none

  type, extends(field_t) :: cuda_c_field_t
    type(c_ptr) :: dev = c_null_ptr  !! device pointer of the padded block (ngrid doubles)
    integer :: shape3(3) = 0        !! the shape a host-side view would have (set_shape)
  contains
    procedure :: fill => fill_cuda_c
    procedure :: get_shape => get_shape_cuda_c
    procedure :: set_shape => set_shape_cuda_c
  end type cuda_c_field_t

  type, extends(allocator_t) :: cuda_c_allocator_t
  contains
    procedure :: create_block => create_cuda_c_block
  end type cuda_c_allocator_t

  interface cuda_c_allocator_t
    module procedure cuda_c_allocator_init
  end interface cuda_c_allocator_t

contains

  function cuda_c_allocator_init(dims, sz) result(allocator)
    integer, intent(in) :: dims(3), sz
    type(cuda_c_allocator_t) :: allocator

    allocator%allocator_t = allocator_t(dims, sz)
  end function cuda_c_allocator_init

  function create_cuda_c_block(self, next) result(ptr)
    !! src/backend/cuda/allocator.f90:81-90 with the storage taken from x3d2c_field_alloc
    class(cuda_c_allocator_t), intent(inout) :: self
    class(field_t), pointer, intent(in) :: next
    type(cuda_c_field_t), pointer :: newblock
    class(field_t), pointer :: ptr

    if (.not. c_associated(x3d2c_ctx)) then
      error stop 'cuda_c allocator: create the backend context (cuda_c_context_init) before requesting blocks'
    end if
    allocate (newblock)
    self%next_id = self%next_id + 1
    call x3d2c_check(x3d2c_field_alloc(x3d2c_ctx, newblock%dev), 'field_alloc')
    newblock%refcount = 0
    newblock%next => next
    newblock%id = self%next_id
    nullify (newblock%data)
    ptr => newblock
  end function create_cuda_c_block

  subroutine fill_cuda_c(self, c)
    class(cuda_c_field_t) :: self
    real(dp), intent(in) :: c

    call x3d2c_check(x3d2c_field_fill(x3d2c_ctx, self%dev, real(c, c_double)), 'field_fill')
  end subroutine fill_cuda_c

  function get_shape_cuda_c(self) result(dims)
    class(cuda_c_field_t) :: self
    integer :: dims(3)

    dims = self%shape3
  end function get_shape_cuda_c

  subroutine set_shape_cuda_c(self, dims)
    !! Directional layouts are private to the backend (SURVEY.md F2): the shape is bookkeeping only
    class(cuda_c_field_t) :: self
    integer, intent(in) :: dims(3)

    self%shape3 = dims
  end subroutine set_shape_cuda_c

  function dev_of(f) result(p)
    !! device pointer of a field handed to the backend (role of resolve_field_t, cuda/backend.f90:1279-1288)
    class(field_t), intent(in) :: f
    type(c_ptr) :: p

    p = c_null_ptr
    select type (f)
    type is (cuda_c_field_t)
      p = f%dev
    class default
      error stop 'cuda_c backend: the field was not created by cuda_c_allocator_t'
    end select
  end function dev_of

end module m_cuda_c_allocator

!=======================================================================================================================
module m_cuda_c_tdsops
  !! tdsops_t plus the handle of its device copy (role of src/backend/cuda/tdsops.f90:9-90)
  use iso_c_binding
  use m_common, only: dp
  use m_tdsops, only: tdsops_t, tdsops_init
  use m_cuda_c_bindings

  implicit none

  type, extends(tdsops_t) :: cuda_c_tdsops_t
    type(c_ptr) :: handle = c_null_ptr
  end type cuda_c_tdsops_t

  interface cuda_c_tdsops_t
    module procedure cuda_c_tdsops_init
  end interface cuda_c_tdsops_t

contains

  function cuda_c_tdsops_init( &
    n_tds, delta, operation, scheme, bc_start, bc_end, &
    stretch, stretch_correct, n_halo, from_to, sym, c_nu, nu0_nu &
    ) result(tdsops)
    type(cuda_c_tdsops_t) :: tdsops
    integer, intent(in) :: n_tds
    real(dp), intent(in) :: delta
    character(*), intent(in) :: operation, scheme
    integer, intent(in) :: bc_start, bc_end
    real(dp), optional, intent(in) :: stretch(:), stretch_correct(:)
    integer, optional, intent(in) :: n_halo
    character(*), optional, intent(in) :: from_to
    logical, optional, intent(in) :: sym
    real(dp), optional, intent(in) :: c_nu, nu0_nu

    real(c_double) :: cs(9, 4), ce(9, 4)
    integer(c_int) :: periodic

    ! the host tables are the reference's own (src/tdsops.f90:63-203) ...
    tdsops%tdsops_t = tdsops_init(n_tds, delta, operation, scheme, bc_start, &
                                  bc_end, stretch, stretch_correct, n_halo, &
                                  from_to, sym, c_nu, nu0_nu)
    if (tdsops%pentadiag) then
      error stop 'cuda_c backend: the pentadiagonal scheme is not reachable from tds_solve / transeq'
    end if
    if (tdsops%n_halo /= 4) error stop 'cuda_c backend: n_halo must be 4'
    ! ... and are uploaded once. coeffs_s(tap, row) is already [row][tap] in C order.
    cs = tdsops%coeffs_s
    ce = tdsops%coeffs_e
    periodic = merge(1_c_int, 0_c_int, tdsops%periodic)
    call x3d2c_check(x3d2c_tdsops_create( &
                     x3d2c_ctx, int(tdsops%n_tds, c_int), int(tdsops%n_rhs, c_int), &
                     int(tdsops%move, c_int), periodic, tdsops%coeffs, cs, ce, &
                     tdsops%dist_fw, tdsops%dist_bw, tdsops%dist_sa, tdsops%dist_sc, &
                     tdsops%dist_af, tdsops%stretch, tdsops%stretch_correct, &
                     tdsops%handle), 'tdsops_create')
  end function cuda_c_tdsops_init

  function ops_of(tdsops) result(h)
    class(tdsops_t), intent(in) :: tdsops
    type(c_ptr) :: h

    h = c_null_ptr
    select type (tdsops)
    type is (cuda_c_tdsops_t)
      h = tdsops%handle
    class default
      error stop 'cuda_c backend: the operator was not created by alloc_tdsops of this backend'
    end select
  end function ops_of

end module m_cuda_c_tdsops

!=======================================================================================================================
module m_cuda_c_poisson_fft
  !! poisson_fft_t over the C ABI (role of src/backend/cuda/poisson_fft.f90). base_init, waves_set and
  !! stretching_matrix are the reference's (src/poisson_fft.f90:120-204,275-831); the spectral buffer is hidden state
  !! of the library handle between fft_forward / fft_postprocess / fft_backward.
  use iso_c_binding
  use m_common, only: dp, CELL
  use m_field, only: field_t
  use m_mesh, only: mesh_t
  use m_poisson_fft, only: poisson_fft_t
  use m_tdsops, only: dirps_t
  use m_cuda_c_allocator, only: dev_of
  use m_cuda_c_bindings

  implicit none

  type, extends(poisson_fft_t) :: cuda_c_poisson_fft_t
    type(c_ptr) :: handle = c_null_ptr
  contains
    procedure :: fft_forward => fft_forward_cuda_c
    procedure :: fft_forward_010 => fft_forward_cuda_c
    procedure :: fft_forward_100 => fft_forward_100_cuda_c
    procedure :: fft_forward_110 => fft_forward_110_cuda_c
    procedure :: fft_backward => fft_backward_cuda_c
    procedure :: fft_backward_010 => fft_backward_cuda_c
    procedure :: fft_backward_100 => fft_backward_100_cuda_c
    procedure :: fft_backward_110 => fft_backward_110_cuda_c
    procedure :: fft_postprocess_000 => fft_postprocess_000_cuda_c
    procedure :: fft_postprocess_010 => fft_postprocess_010_cuda_c
    procedure :: fft_postprocess_100 => fft_postprocess_100_cuda_c
    procedure :: fft_postprocess_110 => fft_postprocess_110_cuda_c
    procedure :: enforce_periodicity_x => enforce_periodicity_x_cuda_c
    procedure :: undo_periodicity_x => undo_periodicity_x_cuda_c
    procedure :: enforce_periodicity_y => enforce_periodicity_y_cuda_c
    procedure :: undo_periodicity_y => undo_periodicity_y_cuda_c
    procedure :: enforce_periodicity_xy => enforce_periodicity_xy_cuda_c
    procedure :: undo_periodicity_xy => undo_periodicity_xy_cuda_c
  end type cuda_c_poisson_fft_t

  interface cuda_c_poisson_fft_t
    module procedure cuda_c_poisson_fft_init
  end interface cuda_c_poisson_fft_t

contains

  function cuda_c_poisson_fft_init(mesh, xdirps, ydirps, zdirps, lowmem) result(poisson_fft)
    type(mesh_t), intent(in) :: mesh
    type(dirps_t), intent(in) :: xdirps, ydirps, zdirps
    logical, optional, intent(in) :: lowmem
    type(cuda_c_poisson_fft_t) :: poisson_fft

    integer(c_int) :: n_spec(3), n_sp_st(3), stretched
    real(c_double), allocatable, target :: w(:)
    real(c_double), allocatable, target :: ao(:, :, :, :), ae(:, :, :, :)
    integer :: i, j, k, idx

    ! the layout of the spectral pencil is the backend's (z slabs in physical, y slabs in spectral space)
    call x3d2c_check(x3d2c_poisson_spec_layout(x3d2c_ctx, n_spec, n_sp_st), 'poisson_spec_layout')
    call poisson_fft%base_init(mesh, xdirps, ydirps, zdirps, int(n_spec), int(n_sp_st))
    if (present(lowmem)) poisson_fft%lowmem = lowmem  ! no effect: the tensors are factorised once on creation

    ! complex(dp) waves(nx_spec, ny_spec, nz_spec) as interleaved re / im
    allocate (w(2*size(poisson_fft%waves)))
    idx = 0
    do k = 1, poisson_fft%nz_spec
      do j = 1, poisson_fft%ny_spec
        do i = 1, poisson_fft%nx_spec
          w(idx + 1) = real(poisson_fft%waves(i, j, k), kind=dp)
          w(idx + 2) = aimag(poisson_fft%waves(i, j, k))
          idx = idx + 2
        end do
      end do
    end do

    if (poisson_fft%periodic_x .and. poisson_fft%periodic_y .and. poisson_fft%periodic_z) then
      call x3d2c_check(x3d2c_poisson_create( &
                       x3d2c_ctx, w, poisson_fft%ax, poisson_fft%bx, poisson_fft%ay, poisson_fft%by, &
                       poisson_fft%az, poisson_fft%bz, poisson_fft%handle), 'poisson_create')
    else if (poisson_fft%periodic_x .and. (.not. poisson_fft%periodic_y) .and. poisson_fft%periodic_z) then
      if (.not. poisson_fft%stretched_y) then
        stretched = 0
        call x3d2c_check(x3d2c_poisson_create_010( &
                         x3d2c_ctx, w, poisson_fft%ax, poisson_fft%bx, poisson_fft%ay, poisson_fft%by, &
                         poisson_fft%az, poisson_fft%bz, stretched, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, &
                         poisson_fft%handle), 'poisson_create_010')
      else if (poisson_fft%stretched_y_sym) then
        stretched = 1
        ! a_*_im hold the same numbers as a_*_re (every complex coefficient of poisson_fft.f90 is (1 + i) x); the
        ! library checks that and keeps one copy. Local target copies give c_loc something legal to point at.
        ao = poisson_fft%a_odd_re; ae = poisson_fft%a_even_re
        call x3d2c_check(x3d2c_poisson_create_010( &
                         x3d2c_ctx, w, poisson_fft%ax, poisson_fft%bx, poisson_fft%ay, poisson_fft%by, &
                         poisson_fft%az, poisson_fft%bz, stretched, c_loc(ao), c_loc(ao), c_loc(ae), c_loc(ae), &
                         poisson_fft%handle), 'poisson_create_010')
      else
        stretched = 2
        ao = poisson_fft%a_re
        call x3d2c_check(x3d2c_poisson_create_010( &
                         x3d2c_ctx, w, poisson_fft%ax, poisson_fft%bx, poisson_fft%ay, poisson_fft%by, &
                         poisson_fft%az, poisson_fft%bz, stretched, c_loc(ao), c_loc(ao), c_null_ptr, c_null_ptr, &
                         poisson_fft%handle), 'poisson_create_010')
      end if
    else
      error stop 'cuda_c backend: walls in x (100 / 110) are not implemented'
    end if
  end function cuda_c_poisson_fft_init

  subroutine fft_forward_cuda_c(self, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_fft_forward(x3d2c_ctx, self%handle, dev_of(f_in)), 'fft_forward')
  end subroutine fft_forward_cuda_c

  subroutine fft_backward_cuda_c(self, f_out)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    call x3d2c_check(x3d2c_fft_backward(x3d2c_ctx, self%handle, dev_of(f_out)), 'fft_backward')
  end subroutine fft_backward_cuda_c

  subroutine fft_forward_100_cuda_c(self, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_fft_forward_100(x3d2c_ctx, self%handle, dev_of(f_in)), 'fft_forward_100')
  end subroutine fft_forward_100_cuda_c

  subroutine fft_forward_110_cuda_c(self, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_fft_forward_110(x3d2c_ctx, self%handle, dev_of(f_in)), 'fft_forward_110')
  end subroutine fft_forward_110_cuda_c

  subroutine fft_backward_100_cuda_c(self, f_out)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    call x3d2c_check(x3d2c_fft_backward_100(x3d2c_ctx, self%handle, dev_of(f_out)), 'fft_backward_100')
  end subroutine fft_backward_100_cuda_c

  subroutine fft_backward_110_cuda_c(self, f_out)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    call x3d2c_check(x3d2c_fft_backward_110(x3d2c_ctx, self%handle, dev_of(f_out)), 'fft_backward_110')
  end subroutine fft_backward_110_cuda_c

  subroutine fft_postprocess_000_cuda_c(self)
    class(cuda_c_poisson_fft_t) :: self
    call x3d2c_check(x3d2c_fft_postprocess_000(x3d2c_ctx, self%handle), 'fft_postprocess_000')
  end subroutine fft_postprocess_000_cuda_c

  subroutine fft_postprocess_010_cuda_c(self)
    class(cuda_c_poisson_fft_t) :: self
    call x3d2c_check(x3d2c_fft_postprocess_010(x3d2c_ctx, self%handle), 'fft_postprocess_010')
  end subroutine fft_postprocess_010_cuda_c

  subroutine fft_postprocess_100_cuda_c(self)
    class(cuda_c_poisson_fft_t) :: self
    call x3d2c_check(x3d2c_fft_postprocess_100(x3d2c_ctx, self%handle), 'fft_postprocess_100')
  end subroutine fft_postprocess_100_cuda_c

  subroutine fft_postprocess_110_cuda_c(self)
    class(cuda_c_poisson_fft_t) :: self
    call x3d2c_check(x3d2c_fft_postprocess_110(x3d2c_ctx, self%handle), 'fft_postprocess_110')
  end subroutine fft_postprocess_110_cuda_c

  subroutine enforce_periodicity_x_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_enforce_periodicity_x(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'enforce_periodicity_x')
  end subroutine enforce_periodicity_x_cuda_c

  subroutine undo_periodicity_x_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_undo_periodicity_x(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'undo_periodicity_x')
  end subroutine undo_periodicity_x_cuda_c

  subroutine enforce_periodicity_y_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_enforce_periodicity_y(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'enforce_periodicity_y')
  end subroutine enforce_periodicity_y_cuda_c

  subroutine undo_periodicity_y_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_undo_periodicity_y(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'undo_periodicity_y')
  end subroutine undo_periodicity_y_cuda_c

  subroutine enforce_periodicity_xy_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_enforce_periodicity_xy(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'enforce_periodicity_xy')
  end subroutine enforce_periodicity_xy_cuda_c

  subroutine undo_periodicity_xy_cuda_c(self, f_out, f_in)
    class(cuda_c_poisson_fft_t) :: self
    class(field_t), intent(inout) :: f_out
    class(field_t), intent(in) :: f_in
    call x3d2c_check(x3d2c_undo_periodicity_xy(x3d2c_ctx, self%handle, dev_of(f_out), dev_of(f_in)), &
                     'undo_periodicity_xy')
  end subroutine undo_periodicity_xy_cuda_c

end module m_cuda_c_poisson_fft

!=======================================================================================================================
module m_cuda_c_backend
  use iso_c_binding
  use mpi

  use m_allocator, only: allocator_t
  use m_base_backend, only: base_backend_t
  use m_common, only: dp, move_data_loc, DIR_X, DIR_Y, DIR_Z, DIR_C, VERT, NULL_LOC, &
                      X_FACE, Y_FACE, Z_FACE, BC_DIRICHLET
  use m_field, only: field_t
  use m_mesh, only: mesh_t
  use m_tdsops, only: dirps_t, tdsops_t

  use m_cuda_c_allocator, only: cuda_c_allocator_t, cuda_c_field_t, dev_of
  use m_cuda_c_bindings
  use m_cuda_c_common, only: SZ
  use m_cuda_c_poisson_fft, only: cuda_c_poisson_fft_t
  use m_cuda_c_tdsops, only: cuda_c_tdsops_t, ops_of

  implicit none

  type, extends(base_backend_t) :: cuda_c_backend_t
  contains
    procedure :: alloc_tdsops => alloc_cuda_c_tdsops
    procedure :: transeq_x => transeq_x_cuda_c
    procedure :: transeq_y => transeq_y_cuda_c
    procedure :: transeq_z => transeq_z_cuda_c
    procedure :: transeq_species => transeq_species_cuda_c
    procedure :: tds_solve => tds_solve_cuda_c
    procedure :: reorder => reorder_cuda_c
    procedure :: sum_yintox => sum_yintox_cuda_c
    procedure :: sum_zintox => sum_zintox_cuda_c
    procedure :: veccopy => veccopy_cuda_c
    procedure :: vecadd => vecadd_cuda_c
    procedure :: vecmult => vecmult_cuda_c
    procedure :: scalar_product => scalar_product_cuda_c
    procedure :: field_max_mean => field_max_mean_cuda_c
    procedure :: slice_max_sum => slice_max_sum_cuda_c
    procedure :: field_scale => field_scale_cuda_c
    procedure :: field_shift => field_shift_cuda_c
    procedure :: field_volume_integral => field_volume_integral_cuda_c
    procedure :: field_set_face => field_set_face_cuda_c
    procedure :: field_set_face_from_field => field_set_face_from_field_cuda_c
    procedure :: compute_vorticity => compute_vorticity_cuda_c
    procedure :: compute_qcriterion => compute_qcriterion_cuda_c
    procedure :: copy_data_to_f => copy_data_to_f_cuda_c
    procedure :: copy_f_to_data => copy_f_to_data_cuda_c
    procedure :: init_poisson_fft => init_cuda_c_poisson_fft
  end type cuda_c_backend_t

  interface cuda_c_backend_t
    module procedure init
  end interface cuda_c_backend_t

contains

  subroutine cuda_c_context_init(mesh, device)
    !! Creates the library context of this rank: what cuda_backend_t%init receives through mesh_t
    !! (src/backend/cuda/backend.f90:95-152). Call it once, before the allocator hands out blocks
    !! (src/xcompact.f90:87-97 builds the allocator before the backend).
    type(mesh_t), intent(in) :: mesh
    integer, intent(in) :: device !! CUDA device ordinal of this rank (src/xcompact.f90:57-60: mod(nrank, ndevs))

    type(x3d2c_config) :: cfg
    character(kind=c_char), target :: nccl_id(128)
    integer :: ierr

    if (c_associated(x3d2c_ctx)) return
    cfg%dims_vert = int(mesh%get_dims(VERT), c_int)
    cfg%dims_cell = int(mesh%grid%cell_dims, c_int)
    cfg%dims_vert_global = int(mesh%grid%global_vert_dims, c_int)
    cfg%dims_cell_global = int(mesh%grid%global_cell_dims, c_int)
    cfg%nproc_dir = int(mesh%par%nproc_dir, c_int)
    cfg%nrank_dir = int(mesh%par%nrank_dir, c_int)
    cfg%n_offset = int(mesh%par%n_offset, c_int)
    cfg%pprev = int(mesh%par%pprev, c_int)
    cfg%pnext = int(mesh%par%pnext, c_int)
    cfg%periodic = merge(1_c_int, 0_c_int, mesh%grid%periodic_BC)
    cfg%sz = int(SZ, c_int)
    cfg%rank = int(mesh%par%nrank, c_int)
    cfg%nproc = int(mesh%par%nproc, c_int)
    cfg%device = int(device, c_int)
    cfg%flags = 0_c_int
    cfg%nccl_unique_id = c_null_ptr
    if (mesh%par%nproc > 1) then
      ! one ncclUniqueId for the job: made by rank 0, broadcast over MPI, consumed by ncclCommInitRank in the library
      if (mesh%par%nrank == 0) call x3d2c_check(x3d2c_nccl_unique_id(nccl_id), 'nccl_unique_id')
      call MPI_Bcast(nccl_id, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr)
      cfg%nccl_unique_id = c_loc(nccl_id)
    end if
    call x3d2c_check(x3d2c_create(cfg, x3d2c_ctx), 'create')
  end subroutine cuda_c_context_init

  function init(mesh, allocator) result(backend)
    type(mesh_t), target, intent(inout) :: mesh
    class(allocator_t), target, intent(inout) :: allocator
    type(cuda_c_backend_t) :: backend

    call backend%base_init()
    if (.not. c_associated(x3d2c_ctx)) then
      error stop 'cuda_c backend: call cuda_c_context_init(mesh, device) before building the allocator and the backend'
    end if
    select type (allocator)
    type is (cuda_c_allocator_t)
      backend%allocator => allocator
    class default
      error stop 'cuda_c backend needs a cuda_c_allocator_t'
    end select
    backend%mesh => mesh
  end function init

  subroutine alloc_cuda_c_tdsops( &
    self, tdsops, n_tds, delta, operation, scheme, bc_start, bc_end, &
    stretch, stretch_correct, n_halo, from_to, sym, c_nu, nu0_nu &
    )
    class(cuda_c_backend_t) :: self
    class(tdsops_t), allocatable, intent(inout) :: tdsops
    integer, intent(in) :: n_tds
    real(dp), intent(in) :: delta
    character(*), intent(in) :: operation, scheme
    integer, intent(in) :: bc_start, bc_end
    real(dp), optional, intent(in) :: stretch(:), stretch_correct(:)
    integer, optional, intent(in) :: n_halo
    character(*), optional, intent(in) :: from_to
    logical, optional, intent(in) :: sym
    real(dp), optional, intent(in) :: c_nu, nu0_nu

    allocate (cuda_c_tdsops_t :: tdsops)
    select type (tdsops)
    type is (cuda_c_tdsops_t)
      tdsops = cuda_c_tdsops_t(n_tds, delta, operation, scheme, bc_start, &
                               bc_end, stretch, stretch_correct, n_halo, &
                               from_to, sym, c_nu, nu0_nu)
    end select
  end subroutine alloc_cuda_c_tdsops

  subroutine transeq_dir(du, dv, dw, u, v, w, nu, dirps)
    !! The library takes the fields in natural order and permutes them itself (omp/backend.f90:154,168,182)
    class(field_t), intent(inout) :: du, dv, dw
    class(field_t), intent(in) :: u, v, w
    real(dp), intent(in) :: nu
    type(dirps_t), intent(in) :: dirps

    call x3d2c_check(x3d2c_transeq( &
                     x3d2c_ctx, int(dirps%dir, c_int), dev_of(du), dev_of(dv), dev_of(dw), &
                     dev_of(u), dev_of(v), dev_of(w), real(nu, c_double), &
                     ops_of(dirps%der1st), ops_of(dirps%der1st_sym), &
                     ops_of(dirps%der2nd), ops_of(dirps%der2nd_sym)), 'transeq')
    du%data_loc = u%data_loc; dv%data_loc = v%data_loc; dw%data_loc = w%data_loc
  end subroutine transeq_dir

  subroutine transeq_x_cuda_c(self, du, dv, dw, u, v, w, nu, dirps)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: du, dv, dw
    class(field_t), intent(in) :: u, v, w
    real(dp), intent(in) :: nu
    type(dirps_t), intent(in) :: dirps
    call transeq_dir(du, dv, dw, u, v, w, nu, dirps)
  end subroutine transeq_x_cuda_c

  subroutine transeq_y_cuda_c(self, du, dv, dw, u, v, w, nu, dirps)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: du, dv, dw
    class(field_t), intent(in) :: u, v, w
    real(dp), intent(in) :: nu
    type(dirps_t), intent(in) :: dirps
    call transeq_dir(du, dv, dw, u, v, w, nu, dirps)
  end subroutine transeq_y_cuda_c

  subroutine transeq_z_cuda_c(self, du, dv, dw, u, v, w, nu, dirps)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: du, dv, dw
    class(field_t), intent(in) :: u, v, w
    real(dp), intent(in) :: nu
    type(dirps_t), intent(in) :: dirps
    call transeq_dir(du, dv, dw, u, v, w, nu, dirps)
  end subroutine transeq_z_cuda_c

  subroutine transeq_species_cuda_c(self, dspec, uvw, spec, nu, dirps, sync)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: dspec
    class(field_t), intent(in) :: uvw, spec
    real(dp), intent(in) :: nu
    type(dirps_t), intent(in) :: dirps
    logical, intent(in) :: sync

    call x3d2c_check(x3d2c_transeq_species( &
                     x3d2c_ctx, int(dirps%dir, c_int), dev_of(dspec), dev_of(uvw), dev_of(spec), &
                     real(nu, c_double), ops_of(dirps%der1st), ops_of(dirps%der1st_sym), &
                     ops_of(dirps%der2nd), merge(1_c_int, 0_c_int, sync)), 'transeq_species')
    dspec%data_loc = spec%data_loc
  end subroutine transeq_species_cuda_c

  subroutine tds_solve_cuda_c(self, du, u, tdsops)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: du
    class(field_t), intent(in) :: u
    class(tdsops_t), intent(in) :: tdsops

    ! cuda/backend.f90:449-470: direction check and data_loc bookkeeping stay on this side
    if (u%dir /= du%dir) error stop 'DIR mismatch between fields in tds_solve.'
    if (u%data_loc /= NULL_LOC) then
      du%data_loc = move_data_loc(u%data_loc, u%dir, tdsops%move)
    end if
    call x3d2c_check(x3d2c_tds_solve(x3d2c_ctx, int(u%dir, c_int), dev_of(du), dev_of(u), ops_of(tdsops)), &
                     'tds_solve')
  end subroutine tds_solve_cuda_c

  subroutine reorder_cuda_c(self, u_, u, direction)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: u_
    class(field_t), intent(in) :: u
    integer, intent(in) :: direction

    call x3d2c_check(x3d2c_reorder(x3d2c_ctx, int(direction, c_int), dev_of(u_), dev_of(u)), 'reorder')
    u_%data_loc = u%data_loc
  end subroutine reorder_cuda_c

  subroutine sum_yintox_cuda_c(self, u, u_)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: u
    class(field_t), intent(in) :: u_
    call x3d2c_check(x3d2c_sum_yintox(x3d2c_ctx, dev_of(u), dev_of(u_)), 'sum_yintox')
  end subroutine sum_yintox_cuda_c

  subroutine sum_zintox_cuda_c(self, u, u_)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: u
    class(field_t), intent(in) :: u_
    call x3d2c_check(x3d2c_sum_zintox(x3d2c_ctx, dev_of(u), dev_of(u_)), 'sum_zintox')
  end subroutine sum_zintox_cuda_c

  subroutine veccopy_cuda_c(self, dst, src)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: dst
    class(field_t), intent(in) :: src

    if (src%dir /= dst%dir) error stop 'Called vector copy with incompatible fields'
    if (dst%dir == DIR_C) error stop 'veccopy does not support DIR_C fields'
    call x3d2c_check(x3d2c_veccopy(x3d2c_ctx, dev_of(dst), dev_of(src)), 'veccopy')
  end subroutine veccopy_cuda_c

  subroutine vecadd_cuda_c(self, a, x, b, y)
    class(cuda_c_backend_t) :: self
    real(dp), intent(in) :: a
    class(field_t), intent(in) :: x
    real(dp), intent(in) :: b
    class(field_t), intent(inout) :: y

    if (x%dir /= y%dir) error stop 'Called vector add with incompatible fields'
    if (y%dir == DIR_C) error stop 'vecadd does not support DIR_C fields'
    call x3d2c_check(x3d2c_vecadd(x3d2c_ctx, real(a, c_double), dev_of(x), real(b, c_double), dev_of(y)), 'vecadd')
  end subroutine vecadd_cuda_c

  subroutine vecmult_cuda_c(self, y, x)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: y
    class(field_t), intent(in) :: x

    if (x%dir /= y%dir) error stop 'Called vector multiply with incompatible fields'
    call x3d2c_check(x3d2c_vecmult(x3d2c_ctx, dev_of(y), dev_of(x)), 'vecmult')
  end subroutine vecmult_cuda_c

  real(dp) function scalar_product_cuda_c(self, x, y) result(s)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(in) :: x, y
    real(c_double) :: s_c

    if ((x%data_loc == NULL_LOC) .or. (y%data_loc == NULL_LOC)) then
      error stop 'You must set the data_loc before calling scalar product'
    end if
    if ((x%data_loc /= y%data_loc) .or. (x%dir /= y%dir)) then
      error stop 'Called scalar product with incompatible fields'
    end if
    ! the library all-reduces over its NCCL communicator: the result is global, as MPI_Allreduce gives in
    ! cuda/backend.f90:864
    call x3d2c_check(x3d2c_scalar_product(x3d2c_ctx, int(x%dir, c_int), int(x%data_loc, c_int), &
                                          dev_of(x), dev_of(y), s_c), 'scalar_product')
    s = real(s_c, dp)
  end function scalar_product_cuda_c

  subroutine field_max_mean_cuda_c(self, max_val, mean_val, f, enforced_data_loc)
    class(cuda_c_backend_t) :: self
    real(dp), intent(out) :: max_val, mean_val
    class(field_t), intent(in) :: f
    integer, optional, intent(in) :: enforced_data_loc
    integer :: data_loc
    real(c_double) :: mx, mn

    if (f%data_loc == NULL_LOC .and. (.not. present(enforced_data_loc))) then
      error stop 'The input field to field_max_mean does not have a valid f%data_loc.'
    end if
    data_loc = f%data_loc
    if (present(enforced_data_loc)) data_loc = enforced_data_loc
    call x3d2c_check(x3d2c_field_max_mean(x3d2c_ctx, int(f%dir, c_int), int(data_loc, c_int), dev_of(f), mx, mn), &
                     'field_max_mean')
    max_val = real(mx, dp); mean_val = real(mn, dp)
  end subroutine field_max_mean_cuda_c

  subroutine slice_max_sum_cuda_c(self, max_val, sum_val, f, i_slice, enforced_data_loc)
    class(cuda_c_backend_t) :: self
    real(dp), intent(out) :: max_val, sum_val
    class(field_t), intent(in) :: f
    integer, intent(in) :: i_slice
    integer, optional, intent(in) :: enforced_data_loc
    integer :: data_loc
    real(c_double) :: mx, sm

    if (f%data_loc == NULL_LOC .and. (.not. present(enforced_data_loc))) then
      error stop 'The input field to slice_max_sum does not have a valid f%data_loc.'
    end if
    data_loc = f%data_loc
    if (present(enforced_data_loc)) data_loc = enforced_data_loc
    ! rank-local values; the caller reduces across ranks (src/backend/backend.f90:238-253)
    call x3d2c_check(x3d2c_slice_max_sum(x3d2c_ctx, int(f%dir, c_int), int(data_loc, c_int), dev_of(f), &
                                         int(i_slice, c_int), mx, sm), 'slice_max_sum')
    max_val = real(mx, dp); sum_val = real(sm, dp)
  end subroutine slice_max_sum_cuda_c

  subroutine field_scale_cuda_c(self, f, a)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(in) :: f
    real(dp), intent(in) :: a
    call x3d2c_check(x3d2c_field_scale(x3d2c_ctx, dev_of(f), real(a, c_double)), 'field_scale')
  end subroutine field_scale_cuda_c

  subroutine field_shift_cuda_c(self, f, a)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(in) :: f
    real(dp), intent(in) :: a
    call x3d2c_check(x3d2c_field_shift(x3d2c_ctx, dev_of(f), real(a, c_double)), 'field_shift')
  end subroutine field_shift_cuda_c

  real(dp) function field_volume_integral_cuda_c(self, f) result(s)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(in) :: f
    real(c_double) :: s_c

    if (f%data_loc == NULL_LOC) error stop 'You must set the data_loc before calling volume integral.'
    if (f%dir /= DIR_X) error stop 'Volume integral can only be called on DIR_X fields currently'
    call x3d2c_check(x3d2c_field_volume_integral(x3d2c_ctx, int(f%data_loc, c_int), dev_of(f), s_c), &
                     'field_volume_integral')
    s = real(s_c, dp)
  end function field_volume_integral_cuda_c

  subroutine field_set_face_cuda_c(self, f, c_start, c_end, face, bc_start, bc_end, flow_rate_diff)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: f
    real(dp), intent(in) :: c_start, c_end
    integer, intent(in) :: face
    integer, optional, intent(in) :: bc_start, bc_end
    real(dp), optional, intent(in) :: flow_rate_diff

    if (f%dir /= DIR_X) error stop 'Setting a field face is only supported for DIR_X fields.'
    if (f%data_loc == NULL_LOC) error stop 'field_set_face requires a valid data_loc.'
    ! Y_FACE (walls of the channel): Dirichlet values; X_FACE / Z_FACE: reported by the library as not supported
    ! (as omp/backend.f90:930-947); the optional BC arguments only matter for X_FACE
    call x3d2c_check(x3d2c_field_set_face(x3d2c_ctx, dev_of(f), int(f%data_loc, c_int), real(c_start, c_double), &
                                          real(c_end, c_double), int(face, c_int)), 'field_set_face')
  end subroutine field_set_face_cuda_c

  subroutine field_set_face_from_field_cuda_c(self, f, f_start, c_end, face, bc_start, bc_end, flow_rate_diff)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: f
    class(field_t), intent(in) :: f_start
    real(dp), intent(in) :: c_end
    integer, intent(in) :: face
    integer, optional, intent(in) :: bc_start, bc_end
    real(dp), optional, intent(in) :: flow_rate_diff
    real(c_double) :: frd

    if (f%dir /= DIR_X) error stop 'field_set_face_from_field: only supported for DIR_X fields.'
    if (f_start%dir /= DIR_X) error stop 'field_set_face_from_field: f_start must be DIR_X.'
    frd = 0._c_double
    if (present(flow_rate_diff)) frd = real(flow_rate_diff, c_double)
    call x3d2c_check(x3d2c_field_set_face_from_field(x3d2c_ctx, dev_of(f), dev_of(f_start), int(f%data_loc, c_int), &
                                                     real(c_end, c_double), int(face, c_int), frd), &
                     'field_set_face_from_field')
  end subroutine field_set_face_from_field_cuda_c

  subroutine compute_vorticity_cuda_c(self, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: field_out
    class(field_t), intent(in) :: dudx, dudy, dudz
    class(field_t), intent(in) :: dvdx, dvdy, dvdz
    class(field_t), intent(in) :: dwdx, dwdy, dwdz

    call x3d2c_check(x3d2c_compute_vorticity(x3d2c_ctx, dev_of(field_out), dev_of(dudx), dev_of(dudy), dev_of(dudz), &
                                             dev_of(dvdx), dev_of(dvdy), dev_of(dvdz), dev_of(dwdx), dev_of(dwdy), &
                                             dev_of(dwdz)), 'compute_vorticity')
  end subroutine compute_vorticity_cuda_c

  subroutine compute_qcriterion_cuda_c(self, field_out, dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz)
    class(cuda_c_backend_t) :: self
    class(field_t), intent(inout) :: field_out
    class(field_t), intent(in) :: dudx, dudy, dudz
    class(field_t), intent(in) :: dvdx, dvdy, dvdz
    class(field_t), intent(in) :: dwdx, dwdy, dwdz

    call x3d2c_check(x3d2c_compute_qcriterion(x3d2c_ctx, dev_of(field_out), dev_of(dudx), dev_of(dudy), dev_of(dudz), &
                                              dev_of(dvdx), dev_of(dvdy), dev_of(dvdz), dev_of(dwdx), dev_of(dwdy), &
                                              dev_of(dwdz)), 'compute_qcriterion')
  end subroutine compute_qcriterion_cuda_c

  subroutine copy_data_to_f_cuda_c(self, f, data)
    !! whole padded block, host -> device (cuda/backend.f90:1246-1252); `data` is contiguous (it is field_t%data of a
    !! host block in every caller: base_backend_t%set_field_data, src/backend/backend.f90:436-466)
    class(cuda_c_backend_t), intent(inout) :: self
    class(field_t), intent(inout) :: f
    real(dp), dimension(:, :, :), intent(in) :: data

    if (size(data) /= self%allocator%ngrid) error stop 'copy_data_to_f: data must be a whole padded block'
    call x3d2c_check(x3d2c_copy_data_to_f(x3d2c_ctx, dev_of(f), data), 'copy_data_to_f')
  end subroutine copy_data_to_f_cuda_c

  subroutine copy_f_to_data_cuda_c(self, data, f)
    class(cuda_c_backend_t), intent(inout) :: self
    real(dp), dimension(:, :, :), intent(out) :: data
    class(field_t), intent(in) :: f

    if (size(data) /= self%allocator%ngrid) error stop 'copy_f_to_data: data must be a whole padded block'
    call x3d2c_check(x3d2c_copy_f_to_data(x3d2c_ctx, data, dev_of(f)), 'copy_f_to_data')
  end subroutine copy_f_to_data_cuda_c

  subroutine init_cuda_c_poisson_fft(self, mesh, xdirps, ydirps, zdirps, lowmem)
    class(cuda_c_backend_t) :: self
    type(mesh_t), intent(in) :: mesh
    type(dirps_t), intent(in) :: xdirps, ydirps, zdirps
    logical, optional, intent(in) :: lowmem

    allocate (cuda_c_poisson_fft_t :: self%poisson_fft)
    select type (poisson_fft => self%poisson_fft)
    type is (cuda_c_poisson_fft_t)
      poisson_fft = cuda_c_poisson_fft_t(mesh, xdirps, ydirps, zdirps, lowmem)
    end select
  end subroutine init_cuda_c_poisson_fft

end module m_cuda_c_backend
