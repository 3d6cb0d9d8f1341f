! rays_b200_m.f90 -- ISO_C_BINDING face of librays_b200.so (include/rays_b200.h) and the drop-in
! replacement of   subroutine trace_rays   (RAYS_project/RAYS_lib/ray_tracing.f90:1-290).
!
! SOURCE ONLY: this image has no Fortran compiler, so this file has never been compiled here
! (tests/test_fortran_binding.py parses it instead: every bind(C) type is checked field by field against
! include/rays_b200.h, every interface against the C prototype, and every routine called is defined).  It is the
! binding a RAYS maintainer adds to RAYS_lib (see INTEGRATION.md): remove ray_tracing.f90 from
! RAYS_lib/CMakeLists.txt, add this file, link librays_b200.so.  Everything else of the host program --
! namelist input, the ray_init_m launchers, module selection, write_results_NC, post_process_RAYS -- is
! unchanged: trace_rays still has no arguments and still fills the ray_results_m module arrays.
!
! The derived types below are bind(C) mirrors of the structs in include/rays_b200.h; field order and
! kinds must not be changed (rays_b200_struct_sizes lets the host verify the layout at start-up).

module rays_b200_m

    use, intrinsic :: iso_c_binding
    implicit none

    integer, parameter :: RAYS_NSPECIES = 6, RAYS_FLAG_LEN = 60

    type, bind(C) :: rays_spline1d
        integer(c_int32_t) :: nx, pad_
        type(c_ptr) :: x_grid, fspl
    end type
    type, bind(C) :: rays_spline2d
        integer(c_int32_t) :: nx, ny
        type(c_ptr) :: x_grid, y_grid, fspl
    end type
    type, bind(C) :: rays_slab_eq
        real(c_double) :: xmin, xmax, ymin, ymax, zmin, zmax, rmaj, rmin, x0
        integer(c_int32_t) :: bx_prof_model, by_prof_model, bz_prof_model, dens_prof_model
        real(c_double) :: bx0, by0, bz0, LBy_shear_scale, LBz_scale, dBzdx, Ln_scale, dndx, alphan1, alphan2, n_min
        integer(c_int32_t) :: t_prof_model(RAYS_NSPECIES)
        real(c_double) :: LT_scale, dtdx, alphat1(RAYS_NSPECIES), alphat2(RAYS_NSPECIES), T_min(RAYS_NSPECIES)
    end type
    type, bind(C) :: rays_solovev_eq
        real(c_double) :: rmaj, kappa, bphi0, iota0, outer_bound, psiB, inner_bound, vert_bound, r_Zmax
        real(c_double) :: box_rmin, box_rmax, box_zmin, box_zmax
        integer(c_int32_t) :: dens_prof_model, t_prof_model(RAYS_NSPECIES), pad_
        real(c_double) :: alphan1, alphan2, alphat1(RAYS_NSPECIES), alphat2(RAYS_NSPECIES)
    end type
    type, bind(C) :: rays_axisym_eq
        integer(c_int32_t) :: magnetics_model, density_prof_model, temperature_prof_model(RAYS_NSPECIES)
        real(c_double) :: r_axis, z_axis, box_rmin, box_rmax, box_zmin, box_zmax
        real(c_double) :: inner_bound, outer_bound, upper_bound, lower_bound
        real(c_double) :: plasma_psi_limit, alphan1, alphan2, d_scrape_off, T_scrape_off
        real(c_double) :: alphat1(RAYS_NSPECIES), alphat2(RAYS_NSPECIES)
        real(c_double) :: sm_rmaj, sm_kappa, sm_bphi0, sm_iota0, sm_psiB
        real(c_double) :: sm_box_rmin, sm_box_rmax, sm_box_zmin, sm_box_zmax
        type(rays_spline1d) :: ne_spline, Te_spline, Ti_spline
        type(rays_spline2d) :: Psi_spline
        type(rays_spline1d) :: T_spline
        real(c_double) :: eq_psibound
    end type
    type, bind(C) :: rays_mirror_eq
        integer(c_int32_t) :: density_prof_model, temperature_prof_model(RAYS_NSPECIES), pad_
        real(c_double) :: box_rmax, box_zmin, box_zmax, r_LUFS, z_LUFS, Aphi_LUFS
        real(c_double) :: plasma_AphiN_limit, alphan1, alphan2, AphiN0_d, delta_d, d_scrape_off, T_scrape_off
        real(c_double) :: alphat1(RAYS_NSPECIES), alphat2(RAYS_NSPECIES), AphiN0_t(RAYS_NSPECIES), delta_t(RAYS_NSPECIES)
        type(rays_spline2d) :: Br_spline, Bz_spline, Aphi_spline
    end type
    type, bind(C) :: rays_cfg
        real(c_double) :: clight, eps0
        real(c_double) :: omgrf, k0, dispersion_resid_limit
        integer(c_int32_t) :: ray_param, wave_mode, k0_sign
        integer(c_int32_t) :: nspec
        real(c_double) :: qs(RAYS_NSPECIES), ms(RAYS_NSPECIES), n0s(RAYS_NSPECIES), t0s(RAYS_NSPECIES), eta(RAYS_NSPECIES)
        integer(c_int32_t) :: ode_solver, ray_deriv, nv, nstep_max
        real(c_double) :: ds, s_max, rel_err0, abs_err0, SG_error_limit
        integer(c_int32_t) :: damping_model, multi_spec_damping
        real(c_double) :: total_damping_limit
        integer(c_int32_t) :: integrate_eq_gradients, equilib_model
        type(rays_slab_eq) :: slab
        type(rays_solovev_eq) :: solovev
        type(rays_axisym_eq) :: axisym
        type(rays_mirror_eq) :: mirror
        type(rays_spline1d) :: zfun_re
    end type
    type, bind(C) :: rays_fan
        integer(c_int64_t) :: nray
        type(c_ptr) :: rvec0, rindex_vec0, ray_pwr_wt
    end type
    type, bind(C) :: rays_results
        integer(c_int64_t) :: nray
        integer(c_int32_t) :: nv, npoints_alloc
        type(c_ptr) :: ray_vec, residual, npoints, ray_stop_code, ray_stop_flag, initial_ray_power, ray_trace_time
        type(c_ptr) :: end_residuals, max_residuals, end_ray_parameter, start_ray_vec, end_ray_vec
        real(c_double) :: total_trace_time
        integer(c_int64_t) :: total_ray_steps
    end type

    type, bind(C) :: rays_deposition
        integer(c_int32_t) :: n_bins, pad_
        real(c_double) :: grid_min, grid_max
        type(c_ptr) :: profile
        real(c_double) :: Q_sum
    end type

!   coil_type + /coil_data_list/ + /current_data_list/ (mirror_magnetics_lib/mirror_magnetics_m.f90:62-115)
    type, bind(C) :: rays_coil
        real(c_double) :: inner_radius, outer_radius, z_center, z_width, I_coil
        integer(c_int64_t) :: n_turns
        integer(c_int32_t) :: n_r_layers, n_z_slices
    end type
!   type OX_conv + per-ray flags (post_process_lib/OX_conv_analysis_m.f90:32-47)
    type, bind(C) :: rays_ox_conv
        real(c_double) :: x_max(3), k_max(3), alpha_max, x_cut(3), conv_coeff, nvecx_c(3), nvecy_c(3), nvecz_c(3)
        integer(c_int32_t) :: ray_number, step_number, found_max, found_cutoff, converted, iteration
    end type

    interface
!       calculate_B_on_rz_grid (mirror_magnetics_m.f90:324-368); Br, Bz, Aphi(n_r, n_z) as the module holds them
        integer(c_int) function rays_b200_mirror_brz_grid(coils, n_coils, n_r, r_min, r_max, n_z, z_min, z_max, &
                & r_grid, z_grid, Br, Bz, Aphi) bind(C, name='rays_b200_mirror_brz_grid')
            import :: c_int, c_int32_t, c_double, rays_coil
            type(rays_coil), intent(in) :: coils(*)
            integer(c_int32_t), value :: n_coils, n_r, n_z
            real(c_double), value :: r_min, r_max, z_min, z_max
            real(c_double), intent(out) :: r_grid(*), z_grid(*), Br(*), Bz(*), Aphi(*)
        end function
!       analyze_OX_conv (OX_conv_analysis_m.f90:91-198) on the trajectories the last trace left on the device
        integer(c_int) function rays_b200_ox_conv_analysis(out, n_converted) bind(C, name='rays_b200_ox_conv_analysis')
            import :: c_int, c_int64_t, rays_ox_conv
            type(rays_ox_conv), intent(out) :: out(*)
            integer(c_int64_t), intent(out) :: n_converted
        end function
        integer(c_int) function rays_b200_init(device) bind(C, name='rays_b200_init')
            import :: c_int
            integer(c_int), value :: device
        end function
        integer(c_int) function rays_b200_finalize() bind(C, name='rays_b200_finalize')
            import :: c_int
        end function
!       every GPU of the box from this ONE host process (program rays is single-process): one context + stream per GPU,
!       an NCCL communicator over them (ncclCommInitAll); ngpu <= 0 takes all visible GPUs
        integer(c_int) function rays_b200_init_multi(ngpu) bind(C, name='rays_b200_init_multi')
            import :: c_int
            integer(c_int), value :: ngpu
        end function
!       trace_rays over all GPUs of rays_b200_init_multi: rays iray mod ngpu, results land in the caller's arrays in fan order
        integer(c_int) function rays_b200_trace_multi(cfg, fan, res) bind(C, name='rays_b200_trace_multi')
            import :: c_int, rays_cfg, rays_fan, rays_results
            type(rays_cfg), intent(in) :: cfg
            type(rays_fan), intent(in) :: fan
            type(rays_results), intent(inout) :: res
        end function
!       the same without trajectory storage, deposition binned while tracing, ONE ncclReduce of the profile + ONE ncclAllGather
!       of the per-ray summaries (deposition_profiles_m.f90:228-292 without the 64 GB of ray_vec)
        integer(c_int) function rays_b200_trace_multi_binned(cfg, fan, res, dep) bind(C, name='rays_b200_trace_multi_binned')
            import :: c_int, rays_cfg, rays_fan, rays_results, rays_deposition
            type(rays_cfg), intent(in) :: cfg
            type(rays_fan), intent(in) :: fan
            type(rays_results), intent(inout) :: res
            type(rays_deposition), intent(inout) :: dep
        end function
        integer(c_int) function rays_b200_host_free(p) bind(C, name='rays_b200_host_free')
            import :: c_int, c_ptr
            type(c_ptr), value :: p
        end function
        integer(c_int) function rays_b200_trace(cfg, fan, res) bind(C, name='rays_b200_trace')
            import :: c_int, rays_cfg, rays_fan, rays_results
            type(rays_cfg), intent(in) :: cfg
            type(rays_fan), intent(in) :: fan
            type(rays_results), intent(inout) :: res
        end function
        type(c_ptr) function rays_b200_last_error() bind(C, name='rays_b200_last_error')
            import :: c_ptr
        end function
        integer(c_int) function rays_b200_host_alloc(p, bytes) bind(C, name='rays_b200_host_alloc')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), intent(out) :: p
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function rays_b200_struct_sizes(out, n) bind(C, name='rays_b200_struct_sizes')
            import :: c_int, c_int32_t
            integer(c_int32_t), intent(out) :: out(*)
            integer(c_int), value :: n
        end function
    end interface

contains

    integer function prof_code(name)
        character(len=*), intent(in) :: name
        select case (trim(name))
            case ('zero');       prof_code = 0
            case ('constant');   prof_code = 1
            case ('linear');     prof_code = 2
            case ('linear_2');   prof_code = 3
            case ('parabolic');  prof_code = 4
            case ('Gaussian');   prof_code = 5
            case ('hyperbolic'); prof_code = 6
            case ('density_spline_interp', 'temperature_spline_interp'); prof_code = 7
            case default;        prof_code = -1
        end select
    end function prof_code

!   enum rays_slab_b_model (slab_eq_m.f90:172-215)
    integer function slab_b_code(name)
        character(len=*), intent(in) :: name
        select case (trim(name))
            case ('zero');         slab_b_code = 0
            case ('constant');     slab_b_code = 1
            case ('toroid');       slab_b_code = 2
            case ('linear_shear'); slab_b_code = 3
            case ('linear');       slab_b_code = 4
            case ('linear_2');     slab_b_code = 5
            case default;          slab_b_code = -1
        end select
    end function slab_b_code

!   allocate_ray_results: page-locked ray_vec / residual for the streaming copy-out of rays_b200_trace (the path bench.py's
!   `e2e` times).  ray_results_m.f90:44-45 declares   real(KIND=rkind), allocatable :: ray_vec(:,:,:), residual(:,:)   ;
!   with the two arrays declared   pointer, contiguous   instead, initialize_ray_results_m (:132-133) calls this routine in
!   place of its two allocate statements and everything downstream (write_results_NC, the post-processors) is unchanged.
!   An unmodified host (pageable arrays) works too: the library then copies out in batches (bench.py's `e2e_pageable`).
    subroutine allocate_ray_results(nv, max_number_of_points, nray, ray_vec, residual)
        integer, intent(in) :: nv, max_number_of_points, nray
        real(c_double), pointer, contiguous, intent(out) :: ray_vec(:,:,:), residual(:,:)
        type(c_ptr) :: p
        integer(c_size_t) :: bytes
        integer :: rc
        bytes = int(nv, c_size_t) * int(max_number_of_points, c_size_t) * int(nray, c_size_t) * 8_c_size_t
        rc = rays_b200_host_alloc(p, max(bytes, 8_c_size_t))
        if (rc /= 0) stop 1
        call c_f_pointer(p, ray_vec, [nv, max_number_of_points, nray])
        bytes = int(max_number_of_points, c_size_t) * int(nray, c_size_t) * 8_c_size_t
        rc = rays_b200_host_alloc(p, max(bytes, 8_c_size_t))
        if (rc /= 0) stop 1
        call c_f_pointer(p, residual, [max_number_of_points, nray])
        ray_vec = 0.;  residual = 0.
    end subroutine allocate_ray_results

end module rays_b200_m

!*************************************************************************************************

 subroutine trace_rays
!   Drop-in replacement of RAYS_lib/ray_tracing.f90: packs the module state the reference's ray loop
!   reads (SURVEY.md 8b "inputs to marshal"), calls rays_b200_trace, and leaves the results in the
!   ray_results_m module arrays exactly as the OpenMP loop did.

    use, intrinsic :: iso_c_binding
    use rays_b200_m
    use constants_m, only : rkind, clight, eps0
    use diagnostics_m, only : message, text_message, integrate_eq_gradients
    use species_m, only : nspec, qs, ms, n0s, t0s, eta
    use rf_m, only : omgrf, k0, ray_param, wave_mode, k0_sign, dispersion_resid_limit
    use damping_m, only : damping_model, multi_spec_damping, total_damping_limit
    use ode_m, only : ode_solver_name, ray_deriv_name, nv, ds, s_max, nstep_max
    use SG_ode_m, only : rel_err0, abs_err0, SG_error_limit
    use equilibrium_m, only : equilib_model
    use ray_init_m, only : nray, rvec0, rindex_vec0, ray_pwr_wt
    use ray_results_m, only : ray_vec, residual, npoints, ray_stop_flag, initial_ray_power, ray_trace_time, &
        & end_residuals, max_residuals, end_ray_parameter, start_ray_vec, end_ray_vec, total_trace_time
    use zfunctions_m, only : x_grid, fsplRe, initialize_spline_coeffs, zfun_initialized => initialized

    implicit none

    type(rays_cfg), target :: cfg
    type(rays_fan) :: fan
    type(rays_results) :: res
    integer(c_int32_t), allocatable, target :: stop_code(:)
    integer :: rc, ngpu
    character(len=16) :: ngpu_env
    logical, save :: device_ready = .false., multi = .false.

    if (.not. device_ready) then
!       program rays is ONE host process (RAYS_code/RAYS.f90:9-15): it drives every visible GPU of the box.
!       RAYS_B200_NGPU=1 (or a single visible device) keeps the run on one GPU.
        call get_environment_variable('RAYS_B200_NGPU', ngpu_env, status=rc)
        ngpu = 0
        if (rc == 0) read(ngpu_env, *, iostat=rc) ngpu
        if (ngpu == 1) then
            rc = rays_b200_init(0_c_int)
        else
            rc = rays_b200_init_multi(int(ngpu, c_int))
            multi = .true.
        end if
        if (rc /= 0) call fail('rays_b200_init')
        device_ready = .true.
    end if

!   constants_m, rf_m, species_m (values computed by the Fortran host carry its own rounding)
    cfg%clight = clight;  cfg%eps0 = eps0
    cfg%omgrf = omgrf;    cfg%k0 = k0;  cfg%dispersion_resid_limit = dispersion_resid_limit
    cfg%ray_param = merge(1, 2, trim(ray_param) == 'arcl')
    select case (trim(wave_mode))
        case ('plus');  cfg%wave_mode = 1
        case ('minus'); cfg%wave_mode = 2
        case ('fast');  cfg%wave_mode = 3
        case ('slow');  cfg%wave_mode = 4
    end select
    cfg%k0_sign = k0_sign
    cfg%nspec = nspec
    cfg%qs = qs(0:5);  cfg%ms = ms(0:5);  cfg%n0s = n0s(0:5);  cfg%t0s = t0s(0:5);  cfg%eta = eta(0:5)
!   ode_m, SG_ode_m
    cfg%ode_solver = merge(2, 1, trim(ode_solver_name) == 'SG_ODE')
    cfg%ray_deriv = merge(2, 1, trim(ray_deriv_name) == 'numerical')
    cfg%nv = nv;  cfg%nstep_max = nstep_max;  cfg%ds = ds;  cfg%s_max = s_max
    cfg%rel_err0 = rel_err0;  cfg%abs_err0 = abs_err0;  cfg%SG_error_limit = SG_error_limit
!   damping_m, diagnostics_m
    cfg%damping_model = merge(0, 1, trim(damping_model) == 'no_damp')
    cfg%multi_spec_damping = merge(1, 0, multi_spec_damping)
    cfg%total_damping_limit = total_damping_limit
    cfg%integrate_eq_gradients = merge(1, 0, integrate_eq_gradients)
    if (cfg%damping_model /= 0) then
        if (.not. zfun_initialized) call initialize_spline_coeffs
        cfg%zfun_re%nx = size(x_grid);  cfg%zfun_re%x_grid = c_loc(x_grid);  cfg%zfun_re%fspl = c_loc(fsplRe)
    end if
!   equilibrium_m: the selected model's module data (pack_* are one block of assignments per model, e.g.
!   cfg%solovev%rmaj = rmaj ... cfg%solovev%psiB = psiB from solovev_eq_m; for multiple_mirror the three
!   cube_spline_function_2D objects give nx, ny, c_loc(x grid), c_loc(y grid), c_loc(fspl))
    select case (trim(equilib_model))
        case ('slab');            cfg%equilib_model = 1;  call pack_slab_eq(cfg%slab)
        case ('solovev');         cfg%equilib_model = 2;  call pack_solovev_eq(cfg%solovev)
        case ('axisym_toroid');   cfg%equilib_model = 3;  call pack_axisym_toroid_eq(cfg%axisym)
        case ('multiple_mirror'); cfg%equilib_model = 4;  call pack_multiple_mirror_eq(cfg%mirror)
    end select

!   ray_init_m: the launch fan, as the launcher modules left it (rvec0(3,nray) is xyz-contiguous per ray)
    fan%nray = nray
    fan%rvec0 = c_loc(rvec0);  fan%rindex_vec0 = c_loc(rindex_vec0);  fan%ray_pwr_wt = c_loc(ray_pwr_wt)

!   ray_results_m: caller-owned arrays in their Fortran layout ray_vec(nv, nstep_max+1, nray)
    allocate(stop_code(nray))
    res%nray = nray;  res%nv = nv;  res%npoints_alloc = nstep_max + 1
    res%ray_vec = c_loc(ray_vec);  res%residual = c_loc(residual);  res%npoints = c_loc(npoints)
    res%ray_stop_code = c_loc(stop_code);  res%ray_stop_flag = c_loc(ray_stop_flag)
    res%initial_ray_power = c_loc(initial_ray_power);  res%ray_trace_time = c_loc(ray_trace_time)
    res%end_residuals = c_loc(end_residuals);  res%max_residuals = c_loc(max_residuals)
    res%end_ray_parameter = c_loc(end_ray_parameter)
    res%start_ray_vec = c_loc(start_ray_vec);  res%end_ray_vec = c_loc(end_ray_vec)

    if (multi) then
        rc = rays_b200_trace_multi(cfg, fan, res)
    else
        rc = rays_b200_trace(cfg, fan, res)
    end if
    if (rc /= 0) call fail('rays_b200_trace')

    total_trace_time = res%total_trace_time
    call message('Wall time ray tracing', total_trace_time, 0)
    deallocate(stop_code)
    return

 contains

    subroutine pack_solovev_eq(q)
!   solovev_eq_m module data (RAYS_lib/solovev_eq_m.f90:17-45) -> rays_solovev_eq.  pack_slab_eq,
!   pack_axisym_toroid_eq (+ solovev_magnetics_m) and pack_multiple_mirror_eq (+ the three
!   cube_spline_function_2D objects of mirror_magnetics_spline_interp_m) follow the same pattern:
!   one assignment per field of the struct, strings resolved with prof_code (table in INTEGRATION.md).
        use solovev_eq_m, only : rmaj, kappa, bphi0, iota0, outer_bound, psiB, inner_bound, vert_bound, r_Zmax, &
            & box_rmin, box_rmax, box_zmin, box_zmax, dens_prof_model, alphan1, alphan2, t_prof_model, alphat1, alphat2
        type(rays_solovev_eq), intent(out) :: q
        integer :: is
        q%rmaj = rmaj;  q%kappa = kappa;  q%bphi0 = bphi0;  q%iota0 = iota0;  q%outer_bound = outer_bound;  q%psiB = psiB
        q%inner_bound = inner_bound;  q%vert_bound = vert_bound;  q%r_Zmax = r_Zmax
        q%box_rmin = box_rmin;  q%box_rmax = box_rmax;  q%box_zmin = box_zmin;  q%box_zmax = box_zmax
        q%dens_prof_model = prof_code(dens_prof_model)
        q%alphan1 = alphan1;  q%alphan2 = alphan2
        q%t_prof_model = 0;  q%alphat1 = 0.;  q%alphat2 = 0.
        do is = 0, nspec
            q%t_prof_model(is+1) = prof_code(t_prof_model(is))
            q%alphat1(is+1) = alphat1(is);  q%alphat2(is+1) = alphat2(is)
        end do
    end subroutine pack_solovev_eq

    subroutine pack_slab_eq(q)
!   slab_eq_m module data (RAYS_lib/slab_eq_m.f90:35-85) -> rays_slab_eq
        use slab_eq_m, only : xmin, xmax, ymin, ymax, zmin, zmax, rmaj, rmin, x0, bx_prof_model, by_prof_model, bz_prof_model, &
            & bx0, by0, bz0, LBy_shear_scale, LBz_scale, dBzdx, dens_prof_model, Ln_scale, dndx, alphan1, alphan2, n_min, &
            & t_prof_model, LT_scale, dtdx, alphat1, alphat2, T_min
        type(rays_slab_eq), intent(out) :: q
        integer :: is
        q%xmin = xmin;  q%xmax = xmax;  q%ymin = ymin;  q%ymax = ymax;  q%zmin = zmin;  q%zmax = zmax
        q%rmaj = rmaj;  q%rmin = rmin;  q%x0 = x0
        q%bx_prof_model = slab_b_code(bx_prof_model)
        q%by_prof_model = slab_b_code(by_prof_model)
        q%bz_prof_model = slab_b_code(bz_prof_model)
        q%dens_prof_model = prof_code(dens_prof_model)
        q%bx0 = bx0;  q%by0 = by0;  q%bz0 = bz0
        q%LBy_shear_scale = LBy_shear_scale;  q%LBz_scale = LBz_scale;  q%dBzdx = dBzdx
        q%Ln_scale = Ln_scale;  q%dndx = dndx;  q%alphan1 = alphan1;  q%alphan2 = alphan2;  q%n_min = n_min
        q%LT_scale = LT_scale;  q%dtdx = dtdx
        q%t_prof_model = 0;  q%alphat1 = 0.;  q%alphat2 = 0.;  q%T_min = 0.
        do is = 0, nspec
            q%t_prof_model(is+1) = prof_code(t_prof_model(is))
            q%alphat1(is+1) = alphat1(is);  q%alphat2(is+1) = alphat2(is);  q%T_min(is+1) = T_min(is)
        end do
    end subroutine pack_slab_eq

    subroutine pack_axisym_toroid_eq(q)
!   axisym_toroid_eq_m (RAYS_lib/axisym_toroid_eq_m.f90:56-94) + the selected magnetics: solovev_magnetics_m
!   (solovev_magnetics_m.f90:19-35) or eqdsk_magnetics_spline_interp_m (Psi_profile, T_profile, psiB :31-54) + the tabulated
!   profiles of density_spline_interp_m (:43) and temperature_spline_interp_m (:38-39) -> rays_axisym_eq
        use axisym_toroid_eq_m, only : magnetics_model, r_axis, z_axis, box_rmin, box_rmax, box_zmin, box_zmax, &
            & inner_bound, outer_bound, upper_bound, lower_bound, plasma_psi_limit, density_prof_model, alphan1, alphan2, &
            & d_scrape_off, temperature_prof_model, alphat1, alphat2, T_scrape_off
        use solovev_magnetics_m, only : sm_rmaj => rmaj, sm_kappa => kappa, sm_bphi0 => bphi0, sm_iota0 => iota0, sm_psiB => psiB, &
            & sm_box_rmin => box_rmin, sm_box_rmax => box_rmax, sm_box_zmin => box_zmin, sm_box_zmax => box_zmax
        use eqdsk_magnetics_spline_interp_m, only : Psi_profile, T_profile, eq_psiB => psiB
        use density_spline_interp_m, only : ne_profile_N
        use temperature_spline_interp_m, only : Te_profileN, Ti_profileN
        type(rays_axisym_eq), intent(out) :: q
        integer :: is
        select case (trim(magnetics_model))
            case ('solovev_magnetics');             q%magnetics_model = 1
            case ('eqdsk_magnetics_spline_interp'); q%magnetics_model = 2
            case default;                           q%magnetics_model = -1      ! rejected by rays_b200_set_config
        end select
        q%density_prof_model = prof_code(density_prof_model)
        q%r_axis = r_axis;  q%z_axis = z_axis
        q%box_rmin = box_rmin;  q%box_rmax = box_rmax;  q%box_zmin = box_zmin;  q%box_zmax = box_zmax
        q%inner_bound = inner_bound;  q%outer_bound = outer_bound;  q%upper_bound = upper_bound;  q%lower_bound = lower_bound
        q%plasma_psi_limit = plasma_psi_limit
        q%alphan1 = alphan1;  q%alphan2 = alphan2;  q%d_scrape_off = d_scrape_off;  q%T_scrape_off = T_scrape_off
        q%temperature_prof_model = 0;  q%alphat1 = 0.;  q%alphat2 = 0.
        do is = 0, nspec
            q%temperature_prof_model(is+1) = prof_code(temperature_prof_model(is))
            q%alphat1(is+1) = alphat1(is);  q%alphat2(is+1) = alphat2(is)
        end do
        q%sm_rmaj = 0.;  q%sm_kappa = 0.;  q%sm_bphi0 = 0.;  q%sm_iota0 = 0.;  q%sm_psiB = 0.
        q%sm_box_rmin = 0.;  q%sm_box_rmax = 0.;  q%sm_box_zmin = 0.;  q%sm_box_zmax = 0.
        call no_spline1d(q%ne_spline);  call no_spline1d(q%Te_spline);  call no_spline1d(q%Ti_spline)
        call no_spline1d(q%T_spline);   call no_spline2d(q%Psi_spline)
        q%eq_psibound = 0.
        if (q%magnetics_model == 1) then
            q%sm_rmaj = sm_rmaj;  q%sm_kappa = sm_kappa;  q%sm_bphi0 = sm_bphi0;  q%sm_iota0 = sm_iota0;  q%sm_psiB = sm_psiB
            q%sm_box_rmin = sm_box_rmin;  q%sm_box_rmax = sm_box_rmax;  q%sm_box_zmin = sm_box_zmin;  q%sm_box_zmax = sm_box_zmax
        else if (q%magnetics_model == 2) then
            call pack_spline2d(Psi_profile%nx, Psi_profile%ny, Psi_profile%x_grid, Psi_profile%y_grid, Psi_profile%fspl, q%Psi_spline)
!           T_profile is splined on the g-file's R grid (eqdsk_magnetics_spline_interp_m.f90:184): it shares Psi_profile%x_grid
            q%T_spline%nx = T_profile%nx;  q%T_spline%pad_ = 0
            q%T_spline%x_grid = q%Psi_spline%x_grid;  q%T_spline%fspl = c_loc(T_profile%fspl)
            q%eq_psibound = eq_psiB
        end if
        if (q%density_prof_model == 7) call pack_spline1d(ne_profile_N%nx, ne_profile_N%x_grid, ne_profile_N%fspl, q%ne_spline)
        if (q%temperature_prof_model(1) == 7) call pack_spline1d(Te_profileN%nx, Te_profileN%x_grid, Te_profileN%fspl, q%Te_spline)
        if (nspec >= 1) then
            if (any(q%temperature_prof_model(2:nspec+1) == 7)) &
                & call pack_spline1d(Ti_profileN%nx, Ti_profileN%x_grid, Ti_profileN%fspl, q%Ti_spline)
        end if
    end subroutine pack_axisym_toroid_eq

    subroutine pack_multiple_mirror_eq(q)
!   multiple_mirror_eq_m (RAYS_lib/multiple_mirror_eq_m.f90:63-106) + the three cube_spline_function_2D objects of
!   mirror_magnetics_spline_interp_m (:32-41; all three on the (r, z) grid of the Brz file) -> rays_mirror_eq
        use multiple_mirror_eq_m, only : box_rmax, box_zmin, box_zmax, r_LUFS, z_LUFS, Aphi_LUFS, plasma_AphiN_limit, &
            & density_prof_model, alphan1, alphan2, AphiN0_d, delta_d, d_scrape_off, temperature_prof_model, alphat1, alphat2, &
            & AphiN0_t, delta_t, T_scrape_off
        use mirror_magnetics_spline_interp_m, only : Br_spline, Bz_spline, Aphi_spline
        type(rays_mirror_eq), intent(out) :: q
        integer :: is
        q%density_prof_model = prof_code(density_prof_model);  q%pad_ = 0
        q%box_rmax = box_rmax;  q%box_zmin = box_zmin;  q%box_zmax = box_zmax
        q%r_LUFS = r_LUFS;  q%z_LUFS = z_LUFS;  q%Aphi_LUFS = Aphi_LUFS
        q%plasma_AphiN_limit = plasma_AphiN_limit
        q%alphan1 = alphan1;  q%alphan2 = alphan2;  q%AphiN0_d = AphiN0_d;  q%delta_d = delta_d
        q%d_scrape_off = d_scrape_off;  q%T_scrape_off = T_scrape_off
        q%temperature_prof_model = 0;  q%alphat1 = 0.;  q%alphat2 = 0.;  q%AphiN0_t = 0.;  q%delta_t = 0.
        do is = 0, nspec
            q%temperature_prof_model(is+1) = prof_code(temperature_prof_model(is))
            q%alphat1(is+1) = alphat1(is);  q%alphat2(is+1) = alphat2(is)
            q%AphiN0_t(is+1) = AphiN0_t(is);  q%delta_t(is+1) = delta_t(is)
        end do
        call pack_spline2d(Br_spline%nx, Br_spline%ny, Br_spline%x_grid, Br_spline%y_grid, Br_spline%fspl, q%Br_spline)
        call pack_spline2d(Bz_spline%nx, Bz_spline%ny, Bz_spline%x_grid, Bz_spline%y_grid, Bz_spline%fspl, q%Bz_spline)
        call pack_spline2d(Aphi_spline%nx, Aphi_spline%ny, Aphi_spline%x_grid, Aphi_spline%y_grid, Aphi_spline%fspl, q%Aphi_spline)
!       the library wants Br, Bz, Aphi on ONE (r, z) grid (they are: one Brz file) and compares the grid pointers
        q%Bz_spline%x_grid = q%Br_spline%x_grid;    q%Bz_spline%y_grid = q%Br_spline%y_grid
        q%Aphi_spline%x_grid = q%Br_spline%x_grid;  q%Aphi_spline%y_grid = q%Br_spline%y_grid
    end subroutine pack_multiple_mirror_eq

!   cube_spline_function_1D / _2D (splines_lib/quick_cube_splines_m.f90:36-56) -> rays_spline1d / rays_spline2d:
!   fspl(4,nx) and fspl(4,4,nx,ny) are passed as they lie in memory (column-major = the C index given in rays_b200.h)
    subroutine pack_spline1d(nx, x_grid, fspl, q)
        integer, intent(in) :: nx
        real(KIND=rkind), intent(in), target :: x_grid(:), fspl(:,:)
        type(rays_spline1d), intent(out) :: q
        q%nx = nx;  q%pad_ = 0;  q%x_grid = c_loc(x_grid);  q%fspl = c_loc(fspl)
    end subroutine pack_spline1d
    subroutine pack_spline2d(nx, ny, x_grid, y_grid, fspl, q)
        integer, intent(in) :: nx, ny
        real(KIND=rkind), intent(in), target :: x_grid(:), y_grid(:), fspl(:,:,:,:)
        type(rays_spline2d), intent(out) :: q
        q%nx = nx;  q%ny = ny;  q%x_grid = c_loc(x_grid);  q%y_grid = c_loc(y_grid);  q%fspl = c_loc(fspl)
    end subroutine pack_spline2d
    subroutine no_spline1d(q)
        type(rays_spline1d), intent(out) :: q
        q%nx = 0;  q%pad_ = 0;  q%x_grid = c_null_ptr;  q%fspl = c_null_ptr
    end subroutine no_spline1d
    subroutine no_spline2d(q)
        type(rays_spline2d), intent(out) :: q
        q%nx = 0;  q%ny = 0;  q%x_grid = c_null_ptr;  q%y_grid = c_null_ptr;  q%fspl = c_null_ptr
    end subroutine no_spline2d

    subroutine fail(where)
        character(len=*), intent(in) :: where
        character(kind=c_char), pointer :: msg(:)
        call c_f_pointer(rays_b200_last_error(), msg, [256])
        write(0,*) where, ' failed: ', msg(1:index(transfer(msg, repeat(' ', 256)), c_null_char) - 1)
        stop 1
    end subroutine fail

 end subroutine trace_rays
