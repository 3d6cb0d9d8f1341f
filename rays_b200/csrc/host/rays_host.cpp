// rays_host.cpp — host-side mirror of the reference's Fortran host program above the C ABI.
//
// The reference's toolchain (Fortran) is absent from this image, so the host side that would stay
// Fortran (namelist input, module initialisation, launch-fan selection, result arrays, netCDF
// output) is written here in C++ with the reference's own names and error behaviour:
//   initialize(read_input)   RAYS_lib/intialize.f90:1-94      -> rays_host_initialize
//   trace_rays               RAYS_lib/ray_tracing.f90:1-290   -> rays_host_trace_rays (-> rays_b200_trace)
//   finalize_run             RAYS_lib/finalize_run.f90:1-51   -> rays_host_finalize_run
// Module data is kept in one struct per Fortran module (constants_m, species_m, rf_m, ...).
// Fatal configuration errors, which `stop 1` in the reference, return a nonzero status and set
// rays_host_last_error().  Numerical habits that shape results are reproduced: default-kind real
// literals are single precision widened to double.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/rays_b200.h"
#include "namelist.hpp"
#include "netcdf3.hpp"
#include "splines_setup.hpp"

namespace rays_host {

static inline double f32(double x) { return (double)(float)x; }

struct constants_m {  // RAYS_lib/constants_m.f90:36-60
    double pi, sqrt_pi, clight, mu0, eps0, me, mp, e;
    void initialize() {
        pi = f32(3.1415926535897932385);
        sqrt_pi = std::sqrt(pi);
        clight = f32(2.997930e8);
        mu0 = pi * f32(4.e-7);
        eps0 = 1. / (mu0 * (clight * clight));
        me = f32(9.1094e-31);
        mp = f32(1.6726e-27);
        e = f32(1.6022e-19);
    }
};

struct diagnostics_m {  // RAYS_lib/diagnostics_m.f90:90-167
    int verbosity = 0;
    bool messages_to_stdout = false, write_formatted_ray_files = false, integrate_eq_gradients = false;
    std::string run_description, run_label;
    int date_v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct species_m {  // RAYS_lib/species_m.f90:21-168
    int nspec = 0;
    double n0 = 0.0, neutrality = f32(1.e-10);
    double n0s[6] = {0}, t0s[6] = {0}, qs[6] = {0}, ms[6] = {0}, eta[6] = {0}, nseps[6] = {0}, t0s_eV[6] = {0}, tseps_eV[6] = {0};
    std::string spec_name[6], spec_model[6];
};

struct rf_m {  // RAYS_lib/rf_m.f90:18-95
    double omgrf = 0, k0 = 0, frf = 0, dispersion_resid_limit = 0;
    std::string ray_dispersion_model, wave_mode, ray_param = "arcl";
    int k0_sign = 0;
};

struct damping_m {  // RAYS_lib/damping_m.f90:30-70
    std::string damping_model;
    bool multi_spec_damping = false;
    double total_damping_limit = f32(0.99);
};

struct ode_m {  // RAYS_lib/ode_m.f90:20,89-173 + SG_ode_m.f90:26-70
    int nv = 0, nstep_max = 0;
    std::string ode_solver_name, ray_deriv_name;
    double s_max = 0, ds = 0;
    double rel_err0 = 0, abs_err0 = 0, SG_error_limit = f32(0.1);
};

struct ray_init_m {  // RAYS_lib/ray_init_m.f90:47-127 + launcher modules' namelist data
    std::string ray_init_model;
    int nray_max = 0;
    int64_t nray = 0;
    std::vector<double> rvec0, rindex_vec0, ray_pwr_wt;
    rays_slab_launch slab{};
    rays_solovev_launch solovev{};
    rays_axisym_launch axisym{};
    // one_ray_init_XYZ_k_direction_m / file_input_ray_init_m
    int n_rays_in = 0;
    bool use_this_n_vec = false;
    std::vector<double> rvec_in, rindex_vec_in, ray_pwr_wt_in;
};

struct ray_results_m {  // RAYS_lib/ray_results_m.f90:44-165
    bool write_results_list_directed = false, write_results_netCDF = true;
    int nv = 0, max_number_of_points = 0;
    int64_t number_of_rays = 0;
    std::vector<double> ray_vec, residual, initial_ray_power, ray_trace_time, end_ray_parameter, end_residuals,
        max_residuals, start_ray_vec, end_ray_vec;
    std::vector<int32_t> npoints, ray_stop_code;
    std::vector<char> ray_stop_flag;
    double total_trace_time = 0;
    int64_t total_ray_steps = 0;
    bool allocated = false;
};

struct State {
    constants_m constants;
    diagnostics_m diagnostics;
    species_m species;
    rf_m rf;
    damping_m damping;
    ode_m ode;
    ray_init_m ray_init;
    ray_results_m results;
    std::string equilib_model;
    rays_cfg cfg{};
    // table storage referenced by cfg
    std::vector<double> z_x, z_re, z_im;
    std::vector<double> r_grid, z_grid, Br_fspl, Bz_fspl, Aphi_fspl;
    std::vector<double> ne_grid, ne_fspl, T_grid, Te_fspl, Ti_fspl;   // density/temperature_spline_interp_m
    std::vector<double> eq_r_grid, eq_z_grid, eq_psi_fspl, eq_T_fspl;   // eqdsk_magnetics_spline_interp_m
    std::string workdir;
    std::string namelist_path;
    bool initialized = false;
};

static State *g = nullptr;
static std::string g_err;
static int fail(const std::string &m) { g_err = m; return RAYS_ERR_INVALID_CONFIG; }

static int prof_code(const std::string &s) {
    if (s == "zero") return RAYS_PROF_ZERO;
    if (s == "constant") return RAYS_PROF_CONSTANT;
    if (s == "linear") return RAYS_PROF_LINEAR;
    if (s == "linear_2") return RAYS_PROF_LINEAR_2;
    if (s == "parabolic") return RAYS_PROF_PARABOLIC;
    if (s == "Gaussian") return RAYS_PROF_GAUSSIAN;
    if (s == "hyperbolic") return RAYS_PROF_HYPERBOLIC;
    if (s == "density_spline_interp" || s == "temperature_spline_interp") return RAYS_PROF_SPLINE;
    return -1;
}
static std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// ---------------- module initialisers -----------------------------------------------------------
static int initialize_diagnostics(State &S, const NamelistFile &nml) {
    NamelistGroup G("diagnostics_list");
    G.add("verbosity", &S.diagnostics.verbosity);
    G.add("messages_to_stdout", &S.diagnostics.messages_to_stdout);
    G.add("write_formatted_ray_files", &S.diagnostics.write_formatted_ray_files);
    G.add("run_description", &S.diagnostics.run_description);
    G.add("run_label", &S.diagnostics.run_label);
    G.add("integrate_eq_gradients", &S.diagnostics.integrate_eq_gradients);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    std::time_t t = std::time(nullptr);
    std::tm *lt = std::localtime(&t);
    int *d = S.diagnostics.date_v;
    d[0] = lt->tm_year + 1900; d[1] = lt->tm_mon + 1; d[2] = lt->tm_mday; d[3] = 0; d[4] = lt->tm_hour; d[5] = lt->tm_min; d[6] = lt->tm_sec; d[7] = 0;
    return 0;
}

static int initialize_species_m(State &S, const NamelistFile &nml) {
    static const char *spec_name0[6] = {"electron", "hydrogen", "deuterium", "tritium", "3He", "alpha"};
    static const double qs0[6] = {-1., 1., 1., 1., 2., 2.};
    static const double ms0[6] = {1., 1836., 3670., 5497., 5496., 7294.};
    species_m &sp = S.species;
    NamelistGroup G("species_list");
    G.add("n0", &sp.n0);
    G.add_arr("nseps", sp.nseps, 0, 6);
    G.add_arr("spec_name", sp.spec_name, 0, 6);
    G.add_arr("spec_model", sp.spec_model, 0, 6);
    G.add_arr("qs", sp.qs, 0, 6);
    G.add_arr("ms", sp.ms, 0, 6);
    G.add_arr("t0s_ev", sp.t0s_eV, 0, 6);
    G.add_arr("tseps_ev", sp.tseps_eV, 0, 6);
    G.add_arr("eta", sp.eta, 0, 6);
    G.add("neutrality", &sp.neutrality);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    sp.spec_name[0] = "electron";
    sp.ms[0] = 1.; sp.qs[0] = -1.; sp.eta[0] = 1.;
    sp.nspec = 0;
    for (int is = 1; is <= 5; ++is)
        if (sp.eta[is] > 0.) {
            sp.nspec = sp.nspec + 1;
            for (int j = 1; j <= 5; ++j)
                if (trim(sp.spec_name[sp.nspec]) == spec_name0[j]) { sp.ms[sp.nspec] = ms0[j]; sp.qs[sp.nspec] = qs0[j]; }
        }
    double charge = 0.0;
    for (int is = 0; is <= sp.nspec; ++is) charge += sp.qs[is] * sp.eta[is];
    if (std::fabs(charge) > sp.neutrality) return fail("initialize_species: charge neutrality violated");
    for (int is = 0; is < 6; ++is) {
        sp.ms[is] = S.constants.me * sp.ms[is];
        sp.qs[is] = S.constants.e * sp.qs[is];
        sp.n0s[is] = sp.eta[is] * sp.n0;
        sp.t0s[is] = S.constants.e * sp.t0s_eV[is];
    }
    return 0;
}

static int initialize_rf_m(State &S, const NamelistFile &nml) {
    rf_m &rf = S.rf;
    NamelistGroup G("rf_list");
    G.add("ray_dispersion_model", &rf.ray_dispersion_model);
    G.add("frf", &rf.frf);
    G.add("wave_mode", &rf.wave_mode);
    G.add("k0_sign", &rf.k0_sign);
    G.add("ray_param", &rf.ray_param);
    G.add("dispersion_resid_limit", &rf.dispersion_resid_limit);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    if (rf.frf <= 0.) return fail("initialize_rf: frf <= 0");
    rf.omgrf = 2. * S.constants.pi * rf.frf;
    rf.k0 = rf.omgrf / S.constants.clight;
    return 0;
}

static int initialize_damping_m(State &S, const NamelistFile &nml) {
    NamelistGroup G("damping_list");
    G.add("damping_model", &S.damping.damping_model);
    G.add("multi_spec_damping", &S.damping.multi_spec_damping);
    G.add("total_damping_limit", &S.damping.total_damping_limit);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    // initialize_spline_coeffs (zfunctions_m.f90:436-466)
    if (S.z_x.empty()) {
        S.z_x.resize(ZFUN_NX); S.z_re.resize(4 * ZFUN_NX); S.z_im.resize(4 * ZFUN_NX);
        if (zfun_table(S.z_x.data(), S.z_re.data(), S.z_im.data())) return fail("zfunctions_m: spline setup failed");
    }
    return 0;
}

static int initialize_slab_eq_m(State &S, const NamelistFile &nml) {
    rays_slab_eq &p = S.cfg.slab;
    std::string bx, by, bz, dens, tprof[6];
    NamelistGroup G("slab_eq_list");
    G.add("bx_prof_model", &bx); G.add("by_prof_model", &by); G.add("bz_prof_model", &bz);
    G.add("bx0", &p.bx0); G.add("by0", &p.by0); G.add("bz0", &p.bz0);
    G.add("rmaj", &p.rmaj); G.add("rmin", &p.rmin); G.add("dens_prof_model", &dens);
    G.add("alphan1", &p.alphan1); G.add("alphan2", &p.alphan2); G.add("n_min", &p.n_min);
    G.add_arr("t_prof_model", tprof, 0, S.species.nspec + 1);
    G.add_arr("alphat1", p.alphat1, 0, S.species.nspec + 1);
    G.add_arr("alphat2", p.alphat2, 0, S.species.nspec + 1);
    G.add_arr("t_min", p.T_min, 0, S.species.nspec + 1);
    G.add("ln_scale", &p.Ln_scale); G.add("lt_scale", &p.LT_scale); G.add("lby_shear_scale", &p.LBy_shear_scale);
    G.add("lbz_scale", &p.LBz_scale); G.add("dbzdx", &p.dBzdx); G.add("dndx", &p.dndx); G.add("dtdx", &p.dtdx);
    G.add("x0", &p.x0); G.add("xmin", &p.xmin); G.add("xmax", &p.xmax); G.add("ymin", &p.ymin); G.add("ymax", &p.ymax);
    G.add("zmin", &p.zmin); G.add("zmax", &p.zmax);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    bx = trim(bx); by = trim(by); bz = trim(bz); dens = trim(dens);
    if (bx != "zero") return fail("SLAB: invalid bx_prof_model = " + bx);
    p.bx_prof_model = RAYS_SLAB_B_ZERO;
    if (by == "zero") p.by_prof_model = RAYS_SLAB_B_ZERO;
    else if (by == "constant") p.by_prof_model = RAYS_SLAB_B_CONSTANT;
    else if (by == "toroid") p.by_prof_model = RAYS_SLAB_B_TOROID;
    else if (by == "linear_shear") p.by_prof_model = RAYS_SLAB_B_LINEAR_SHEAR;
    else return fail("SLAB: invalid by_prof_model = " + by);
    if (bz == "constant") p.bz_prof_model = RAYS_SLAB_B_CONSTANT;
    else if (bz == "toroid") p.bz_prof_model = RAYS_SLAB_B_TOROID;
    else if (bz == "linear") p.bz_prof_model = RAYS_SLAB_B_LINEAR;
    else if (bz == "linear_2") p.bz_prof_model = RAYS_SLAB_B_LINEAR_2;
    else return fail("SLAB: invalid bz_prof_model = " + bz);
    int dc = prof_code(dens);
    if (dc < RAYS_PROF_CONSTANT || dc > RAYS_PROF_GAUSSIAN) return fail("SLAB: invalid dens_prof_model =" + dens);
    p.dens_prof_model = dc;
    for (int is = 0; is <= S.species.nspec; ++is) {
        int tc = prof_code(trim(tprof[is]));
        if (tc < RAYS_PROF_ZERO || tc > RAYS_PROF_PARABOLIC) return fail("SLAB: invalid t_prof_model = " + tprof[is]);
        p.t_prof_model[is] = tc;
    }
    return 0;
}

static void solovev_derived(double rmaj, double kappa, double bphi0, double iota0, double outer_bound, double &psiB,
                            double &inner_bound, double &r_Zmax, double &vert_bound) {
    // solovev_eq_m.f90:83-90 == solovev_magnetics_m.f90:100-107
    double bp0 = bphi0 * iota0;
    double d = outer_bound * outer_bound - rmaj * rmaj;
    psiB = .5 * bp0 * (d * d) / (rmaj * rmaj) / 4.;
    inner_bound = std::sqrt(2. * (rmaj * rmaj) - outer_bound * outer_bound);
    double ob2 = outer_bound * outer_bound, ob4 = ob2 * ob2;
    r_Zmax = std::pow(2. * ob2 * (rmaj * rmaj) - ob4, 0.25);
    double rz2 = r_Zmax * r_Zmax, rz4 = rz2 * rz2;
    vert_bound = kappa / (2. * r_Zmax) * std::sqrt(ob4 + 2. * (rz2 - ob2) * (rmaj * rmaj) - rz4);
}

static int initialize_solovev_eq_m(State &S, const NamelistFile &nml) {
    rays_solovev_eq &p = S.cfg.solovev;
    std::string dens, tprof[6];
    NamelistGroup G("solovev_eq_list");
    G.add("rmaj", &p.rmaj); G.add("outer_bound", &p.outer_bound); G.add("kappa", &p.kappa); G.add("bphi0", &p.bphi0);
    G.add("iota0", &p.iota0); G.add("dens_prof_model", &dens); G.add("alphan1", &p.alphan1); G.add("alphan2", &p.alphan2);
    G.add_arr("t_prof_model", tprof, 0, S.species.nspec + 1);
    G.add_arr("alphat1", p.alphat1, 0, S.species.nspec + 1);
    G.add_arr("alphat2", p.alphat2, 0, S.species.nspec + 1);
    G.add("box_rmin", &p.box_rmin); G.add("box_rmax", &p.box_rmax); G.add("box_zmin", &p.box_zmin); G.add("box_zmax", &p.box_zmax);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    solovev_derived(p.rmaj, p.kappa, p.bphi0, p.iota0, p.outer_bound, p.psiB, p.inner_bound, p.r_Zmax, p.vert_bound);
    dens = trim(dens);
    if (dens == "constant") p.dens_prof_model = RAYS_PROF_CONSTANT;
    else if (dens == "parabolic") p.dens_prof_model = RAYS_PROF_PARABOLIC;
    else return fail("solovev_eq invalid dens_prof_model =" + dens);
    for (int is = 0; is <= S.species.nspec; ++is) {
        std::string t = trim(tprof[is]);
        if (t == "zero") p.t_prof_model[is] = RAYS_PROF_ZERO;
        else if (t == "constant") p.t_prof_model[is] = RAYS_PROF_CONSTANT;
        else if (t == "parabolic") p.t_prof_model[is] = RAYS_PROF_PARABOLIC;
        else return fail("SOLOVEV: Unknown t_prof_model: " + t);
    }
    return 0;
}

static int spline2d_init(const std::vector<double> &rg, const std::vector<double> &zg, const std::vector<double> &f_c_order, std::vector<double> &fspl);

// ReadgFile (eqdsk_utilities_m.f90:52-108): formatted reads '(a48,3i4)', '(5e16.9)', '(2i5)' -- fixed-width fields, a new
// record per read statement, missing trailing fields read as blanks (= 0)
struct GFile {
    int NRBOX = 0, NZBOX = 0, NBOUND = 0, NLIM = 0;
    double RBOXLEN = 0, ZBOXLEN = 0, R0 = 0, RBOXLFT = 0, ZOFF = 0, RAXIS = 0, ZAXIS = 0, PSIAXIS = 0, PSIBOUND = 0, B0 = 0, CURRENT = 0;
    std::vector<double> T, P, TTp, Pp, Q, Psi, RBOUND, ZBOUND;
};
static int read_gfile(const std::string &path, GFile &g) {
    FILE *fp = std::fopen(path.c_str(), "r");
    if (!fp) return fail("ReadgFile: cannot open " + path);
    std::vector<std::string> lines;
    {
        std::string cur;
        int ch;
        while ((ch = std::fgetc(fp)) != EOF) {
            if (ch == '\n') { lines.push_back(cur); cur.clear(); } else if (ch != '\r') cur.push_back((char)ch);
        }
        if (!cur.empty()) lines.push_back(cur);
        std::fclose(fp);
    }
    size_t ln = 0;
    bool bad = false;
    auto field = [&](const std::string &L, size_t pos, size_t w) { return pos < L.size() ? L.substr(pos, w) : std::string(); };
    auto to_int = [&](std::string f) { f = trim(f); if (f.empty()) return 0; char *e; long v = std::strtol(f.c_str(), &e, 10); if (*e) bad = true; return (int)v; };
    auto to_real = [&](std::string f) {
        f = trim(f);
        if (f.empty()) return 0.0;
        for (char &c : f) if (c == 'D' || c == 'd') c = 'E';
        char *e; double v = std::strtod(f.c_str(), &e); if (*e) bad = true; return v;
    };
    auto reals = [&](double *out, size_t n) {   // one read statement: '(5e16.9)' over as many records as the list needs
        size_t got = 0;
        while (got < n) {
            if (ln >= lines.size()) { bad = true; return; }
            const std::string &L = lines[ln++];
            for (int k = 0; k < 5 && got < n; ++k) out[got++] = to_real(field(L, 16 * (size_t)k, 16));
        }
    };
    if (lines.empty()) return fail("ReadgFile: empty file " + path);
    {
        const std::string &L = lines[ln++];
        (void)to_int(field(L, 48, 4)); g.NRBOX = to_int(field(L, 52, 4)); g.NZBOX = to_int(field(L, 56, 4));
    }
    if (bad || g.NRBOX < 4 || g.NZBOX < 4) return fail("ReadgFile: bad header (NRBOX, NZBOX) in " + path);
    double h[5];
    reals(h, 5); g.RBOXLEN = h[0]; g.ZBOXLEN = h[1]; g.R0 = h[2]; g.RBOXLFT = h[3]; g.ZOFF = h[4];
    reals(h, 5); g.RAXIS = h[0]; g.ZAXIS = h[1]; g.PSIAXIS = h[2]; g.PSIBOUND = h[3]; g.B0 = h[4];
    reals(h, 5); g.CURRENT = h[0];
    reals(h, 5);
    const size_t nr = (size_t)g.NRBOX, nz = (size_t)g.NZBOX;
    g.T.resize(nr); g.P.resize(nr); g.TTp.resize(nr); g.Pp.resize(nr); g.Q.resize(nr); g.Psi.resize(nr * nz);
    reals(g.T.data(), nr); reals(g.P.data(), nr); reals(g.TTp.data(), nr); reals(g.Pp.data(), nr);
    reals(g.Psi.data(), nr * nz);   // ((Psi(i,j), i = 1, NRBOX), j = 1, NZBOX): R fastest
    reals(g.Q.data(), nr);
    if (bad || ln >= lines.size()) return fail("ReadgFile: truncated or malformed file " + path);
    {
        const std::string &L = lines[ln++];
        g.NBOUND = to_int(field(L, 0, 5)); g.NLIM = to_int(field(L, 5, 5));
    }
    if (bad || g.NBOUND < 1) return fail("ReadgFile: bad NBOUND in " + path);
    std::vector<double> rz(2 * (size_t)g.NBOUND);
    reals(rz.data(), rz.size());
    if (bad) return fail("ReadgFile: truncated boundary in " + path);
    g.RBOUND.resize((size_t)g.NBOUND); g.ZBOUND.resize((size_t)g.NBOUND);
    for (int i = 0; i < g.NBOUND; ++i) { g.RBOUND[(size_t)i] = rz[2 * (size_t)i]; g.ZBOUND[(size_t)i] = rz[2 * (size_t)i + 1]; }
    return 0;   // the limiter contour is not used on the path
}

// initialize_eqdsk_magnetics_spline_interp (eqdsk_magnetics_spline_interp_m.f90:68-200): geometry from the g-file,
// psi shifted to zero on axis, Psi_profile (bicubic) and T_profile (R*Bphi; splined on R_grid as the reference does).
// The q / rho / toroidal-flux profiles built there serve post-processing only and are not on the ray path.
static int initialize_eqdsk_magnetics(State &S, const NamelistFile &nml) {
    rays_axisym_eq &p = S.cfg.axisym;
    std::string file, err;
    NamelistGroup G("eqdsk_magnetics_spline_interp_list");
    G.add("eqdsk_file_name", &file);
    if (!G.read(nml, err)) return fail(err);
    file = trim(file);
    if (!file.empty() && file[0] != '/' && !S.workdir.empty()) file = S.workdir + "/" + file;
    GFile g;
    int rc = read_gfile(file, g);
    if (rc) return rc;
    p.r_axis = g.RAXIS; p.z_axis = g.ZAXIS;
    p.box_rmin = g.RBOXLFT; p.box_rmax = p.box_rmin + g.RBOXLEN;
    p.box_zmin = g.ZOFF - g.ZBOXLEN / 2.; p.box_zmax = g.ZOFF + g.ZBOXLEN / 2.;
    p.inner_bound = *std::min_element(g.RBOUND.begin(), g.RBOUND.end());
    p.outer_bound = *std::max_element(g.RBOUND.begin(), g.RBOUND.end());
    p.lower_bound = *std::min_element(g.ZBOUND.begin(), g.ZBOUND.end());
    p.upper_bound = *std::max_element(g.ZBOUND.begin(), g.ZBOUND.end());
    const int nr = g.NRBOX, nz = g.NZBOX;
    S.eq_r_grid.resize((size_t)nr); S.eq_z_grid.resize((size_t)nz);
    for (int i = 1; i <= nr; ++i) S.eq_r_grid[(size_t)i - 1] = p.box_rmin + (p.box_rmax - p.box_rmin) * (i - 1) / (nr - 1);
    for (int i = 1; i <= nz; ++i) S.eq_z_grid[(size_t)i - 1] = p.box_zmin + (p.box_zmax - p.box_zmin) * (i - 1) / (nz - 1);
    for (double &v : g.Psi) v = v - g.PSIAXIS;
    p.eq_psibound = g.PSIBOUND - g.PSIAXIS;
    if (p.eq_psibound == 0.0) return fail("eqdsk_magnetics: PSIBOUND equals PSIAXIS");
    if (spline2d_init(S.eq_r_grid, S.eq_z_grid, g.Psi, S.eq_psi_fspl)) return fail("cube_spline_2D_init: Psi_profile, grid not evenly spaced or error");
    S.eq_T_fspl.assign((size_t)4 * nr, 0.0);
    for (int i = 0; i < nr; ++i) S.eq_T_fspl[4 * (size_t)i] = g.T[(size_t)i];
    int ilinx = 0;
    if (cspline(S.eq_r_grid.data(), nr, S.eq_T_fspl.data(), &ilinx) || ilinx != 1) return fail("cube_spline_1D_init: T_profile failed");
    p.Psi_spline.nx = nr; p.Psi_spline.ny = nz; p.Psi_spline.x_grid = S.eq_r_grid.data(); p.Psi_spline.y_grid = S.eq_z_grid.data();
    p.Psi_spline.fspl = S.eq_psi_fspl.data();
    p.T_spline.nx = nr; p.T_spline.pad_ = 0; p.T_spline.x_grid = S.eq_r_grid.data(); p.T_spline.fspl = S.eq_T_fspl.data();
    // solovev_magnetics module data is not used by this model
    p.sm_rmaj = p.sm_kappa = p.sm_bphi0 = p.sm_iota0 = p.sm_psiB = 0.0;
    p.sm_box_rmin = p.box_rmin; p.sm_box_rmax = p.box_rmax; p.sm_box_zmin = p.box_zmin; p.sm_box_zmax = p.box_zmax;
    return 0;
}

static int initialize_axisym_toroid_eq_m(State &S, const NamelistFile &nml) {
    rays_axisym_eq &p = S.cfg.axisym;
    p.plasma_psi_limit = 1.0;
    std::string magnetics, dens, tprof[6];
    NamelistGroup G("axisym_toroid_eq_list");
    G.add("magnetics_model", &magnetics); G.add("plasma_psi_limit", &p.plasma_psi_limit);
    G.add("density_prof_model", &dens); G.add("d_scrape_off", &p.d_scrape_off);
    G.add("alphan1", &p.alphan1); G.add("alphan2", &p.alphan2);
    G.add_arr("temperature_prof_model", tprof, 0, S.species.nspec + 1);
    G.add_arr("alphat1", p.alphat1, 0, S.species.nspec + 1);
    G.add_arr("alphat2", p.alphat2, 0, S.species.nspec + 1);
    G.add("t_scrape_off", &p.T_scrape_off);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    magnetics = trim(magnetics);
    p.Psi_spline = rays_spline2d{}; p.T_spline = rays_spline1d{}; p.eq_psibound = 0.0;
    if (magnetics == "eqdsk_magnetics_spline_interp") {
        p.magnetics_model = RAYS_MAG_EQDSK_SPLINE;
        int rc = initialize_eqdsk_magnetics(S, nml);
        if (rc) return rc;
    } else if (magnetics != "solovev_magnetics")
        return fail("initialize_axisym_toroid_eq: unknown magnetics model =" + magnetics + " (eqdsk_magnetics_lin_interp is not provided)");
    else {   // initialize_solovev_magnetics
        p.magnetics_model = RAYS_MAG_SOLOVEV;   // (solovev_magnetics_m.f90:47-120)
        double outer_boundary = 0.0;
        NamelistGroup M("solovev_magnetics_list");
        M.add("rmaj", &p.sm_rmaj); M.add("outer_boundary", &outer_boundary); M.add("kappa", &p.sm_kappa);
        M.add("bphi0", &p.sm_bphi0); M.add("iota0", &p.sm_iota0);
        M.add("box_rmin", &p.sm_box_rmin); M.add("box_rmax", &p.sm_box_rmax); M.add("box_zmin", &p.sm_box_zmin); M.add("box_zmax", &p.sm_box_zmax);
        if (!M.read(nml, err)) return fail(err);
        p.box_rmin = p.sm_box_rmin; p.box_rmax = p.sm_box_rmax; p.box_zmin = p.sm_box_zmin; p.box_zmax = p.sm_box_zmax;
        p.outer_bound = outer_boundary;
        if (p.outer_bound < p.sm_rmaj || p.outer_bound >= (double)std::sqrt(2.f) * p.sm_rmaj)
            return fail("Inner boundary complex, outer_bound >=  sqrt2*rmaj");
        double r_Zmax, vert_bound;
        solovev_derived(p.sm_rmaj, p.sm_kappa, p.sm_bphi0, p.sm_iota0, p.outer_bound, p.sm_psiB, p.inner_bound, r_Zmax, vert_bound);
        p.r_axis = p.sm_rmaj; p.z_axis = 0.;
        p.upper_bound = vert_bound; p.lower_bound = -vert_bound;
    }
    dens = trim(dens);
    p.ne_spline = rays_spline1d{}; p.Te_spline = rays_spline1d{}; p.Ti_spline = rays_spline1d{};
    // profile on a uniform psi_N grid, normalised to its first value, not-a-knot cubic spline
    // (initialize_density_spline_interp, density_spline_interp_m.f90:29-75; cube_spline_1D_init)
    auto build_profile = [&](int ngrid, const std::vector<double> &in, std::vector<double> &grid, std::vector<double> &fspl,
                             rays_spline1d &out) -> int {
        if (ngrid < 4 || ngrid > 200) return fail("spline profile: ngrid must be in 4..200");
        if (in[0] == 0.0) return fail("spline profile: first value (on axis) is zero");
        grid.resize((size_t)ngrid);
        fspl.assign((size_t)4 * ngrid, 0.0);
        for (int i = 1; i <= ngrid; ++i) {
            grid[(size_t)i - 1] = 1.0 * (i - 1) / (ngrid - 1);
            fspl[4 * ((size_t)i - 1)] = in[(size_t)i - 1] / in[0];
        }
        int ilinx = 0;
        if (cspline(grid.data(), ngrid, fspl.data(), &ilinx) || ilinx != 1) return fail("spline profile: cspline failed");
        out.nx = ngrid; out.pad_ = 0; out.x_grid = grid.data(); out.fspl = fspl.data();
        return 0;
    };
    if (dens == "constant") p.density_prof_model = RAYS_PROF_CONSTANT;
    else if (dens == "parabolic") p.density_prof_model = RAYS_PROF_PARABOLIC;
    else if (dens == "density_spline_interp") {
        p.density_prof_model = RAYS_PROF_SPLINE;
        int ngrid = 0;
        std::vector<double> ne_in(200, 0.0);
        NamelistGroup D("density_spline_interp_list");
        D.add("ngrid", &ngrid); D.add_arr("ne_in", ne_in.data(), 1, 200);
        if (!D.read(nml, err)) return fail(err);
        int rc = build_profile(ngrid, ne_in, S.ne_grid, S.ne_fspl, p.ne_spline);
        if (rc) return rc;
    } else return fail("axisym_toroid_eq: Unknown density_prof_model: " + dens);
    int n_T_spline = 0;
    for (int is = 0; is <= S.species.nspec; ++is) {
        std::string t = trim(tprof[is]);
        if (t == "zero") p.temperature_prof_model[is] = RAYS_PROF_ZERO;
        else if (t == "constant") p.temperature_prof_model[is] = RAYS_PROF_CONSTANT;
        else if (t == "parabolic") p.temperature_prof_model[is] = RAYS_PROF_PARABOLIC;
        else if (t == "temperature_spline_interp") { p.temperature_prof_model[is] = RAYS_PROF_SPLINE; ++n_T_spline; }
        else return fail("axisym_toroid_eq: Unknown temperature_prof_model: " + t);
    }
    if (n_T_spline > 0) {   // initialize_temperature_spline_interp (temperature_spline_interp_m.f90:30-78)
        int ngrid = 0;
        std::vector<double> Te_in(200, 0.0), Ti_in(200, 0.0);
        NamelistGroup T("temperature_spline_interp_list");
        T.add("ngrid", &ngrid); T.add_arr("te_in", Te_in.data(), 1, 200); T.add_arr("ti_in", Ti_in.data(), 1, 200);
        if (!T.read(nml, err)) return fail(err);
        std::vector<double> grid2;
        int rc = build_profile(ngrid, Te_in, S.T_grid, S.Te_fspl, p.Te_spline);
        if (rc) return rc;
        if ((rc = build_profile(ngrid, Ti_in, grid2, S.Ti_fspl, p.Ti_spline))) return rc;
        p.Ti_spline.x_grid = S.T_grid.data();
    }
    return 0;
}

// bicubic setup of one field on (r_grid, z_grid): cube_spline_2D_init (quick_cube_splines_m.f90:180-250)
static int spline2d_init(const std::vector<double> &rg, const std::vector<double> &zg, const std::vector<double> &f_c_order,
                         std::vector<double> &fspl) {
    const int nx = (int)rg.size(), ny = (int)zg.size();
    fspl.assign((size_t)16 * nx * ny, 0.0);
    // file variable is C-order (n_z, n_r) == Fortran f(n_r, n_z): f(i,j) = data[j*nx + i]
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) fspl[((size_t)(j * nx + i) * 4 + 0) * 4 + 0] = f_c_order[(size_t)j * nx + i];
    int ilinx, iliny;
    int rc = bcspline(rg.data(), nx, zg.data(), ny, fspl.data(), &ilinx, &iliny);
    if (rc) return rc;
    if (ilinx != 1 || iliny != 1) return 99;  // "grid not evenly spaced"
    return 0;
}
// eval_2D_f at one point on the host (needed once: Aphi_LUFS, mirror_magnetics_spline_interp_m.f90:109)
static double spline2d_eval_f(const rays_spline2d &s, double xg, double yg) {
    const int nx = s.nx, ny = s.ny;
    const double *x = s.x_grid, *y = s.y_grid;
    int nxm = nx - 1, nym = ny - 1;
    int i = std::min(nxm, 1 + (int)(nxm * (xg - x[0]) / (x[nx - 1] - x[0])));
    if (xg < x[i - 1]) i--; else if (xg > x[i]) i++;
    int j = std::min(nym, 1 + (int)(nym * (yg - y[0]) / (y[ny - 1] - y[0])));
    if (yg < y[j - 1]) j--; else if (yg > y[j]) j++;
    double dx = xg - x[i - 1], dy = yg - y[j - 1];
    const double *c = s.fspl + (size_t)((j - 1) * nx + (i - 1)) * 16;
#define F(cx, cy) c[((cy)-1) * 4 + ((cx)-1)]
    return F(1, 1) + dy * (F(1, 2) + dy * (F(1, 3) + dy * F(1, 4))) +
           dx * (F(2, 1) + dy * (F(2, 2) + dy * (F(2, 3) + dy * F(2, 4))) +
                 dx * (F(3, 1) + dy * (F(3, 2) + dy * (F(3, 3) + dy * F(3, 4))) +
                       dx * (F(4, 1) + dy * (F(4, 2) + dy * (F(4, 3) + dy * F(4, 4))))));
#undef F
}

static int initialize_multiple_mirror_eq_m(State &S, const NamelistFile &nml) {
    rays_mirror_eq &p = S.cfg.mirror;
    p.plasma_AphiN_limit = 1.0;
    std::string magnetics, dens, tprof[6];
    NamelistGroup G("multiple_mirror_eq_list");
    G.add("magnetics_model", &magnetics); G.add("plasma_aphin_limit", &p.plasma_AphiN_limit);
    G.add("density_prof_model", &dens); G.add("d_scrape_off", &p.d_scrape_off);
    G.add("alphan1", &p.alphan1); G.add("alphan2", &p.alphan2); G.add("aphin0_d", &p.AphiN0_d); G.add("delta_d", &p.delta_d);
    G.add_arr("temperature_prof_model", tprof, 0, S.species.nspec + 1);
    G.add_arr("alphat1", p.alphat1, 0, S.species.nspec + 1);
    G.add_arr("alphat2", p.alphat2, 0, S.species.nspec + 1);
    G.add_arr("aphin0_t", p.AphiN0_t, 0, S.species.nspec + 1);
    G.add_arr("delta_t", p.delta_t, 0, S.species.nspec + 1);
    G.add("t_scrape_off", &p.T_scrape_off);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    if (trim(magnetics) != "mirror_magnetics_spline_interp") return fail("initialize_multiple_mirror_eq: unknown magnetics model =" + magnetics);
    {   // initialize_mirror_magnetics_spline_interp (mirror_magnetics_spline_interp_m.f90:49-113)
        std::string file;
        NamelistGroup M("mirror_magnetics_spline_interp_list");
        M.add("mirror_field_nc_file", &file);
        if (!M.read(nml, err)) return fail(err);
        std::string path = trim(file);
        if (!path.empty() && path[0] != '/' && !S.workdir.empty()) path = S.workdir + "/" + path;
        NcReader nc;   // read_mirror_fields_Brz_NC (mirror_magnetics_m.f90:452-531)
        if (!nc.open(path)) { g_err = nc.error(); return RAYS_ERR_IO; }
        std::vector<double> v, Br, Bz, Aphi;
        double r_min, r_max, z_min, z_max, r_LUFS, z_LUFS;
        auto scalar = [&](const char *n, double &out) { if (!nc.get_var_double(n, v) || v.empty()) return false; out = v[0]; return true; };
        if (!scalar("r_min", r_min) || !scalar("r_max", r_max) || !scalar("z_min", z_min) || !scalar("z_max", z_max) ||
            !scalar("r_LUFS", r_LUFS) || !scalar("z_LUFS", z_LUFS) || !nc.get_var_double("r_grid", S.r_grid) ||
            !nc.get_var_double("z_grid", S.z_grid) || !nc.get_var_double("Br", Br) || !nc.get_var_double("Bz", Bz) ||
            !nc.get_var_double("Aphi", Aphi)) { g_err = nc.error(); return RAYS_ERR_IO; }
        if (r_min != 0.0) return fail("initialize_mirror_magnetics_spline_interp: non-zero r_min");
        p.box_rmax = r_max; p.box_zmin = z_min; p.box_zmax = z_max;
        p.r_LUFS = r_LUFS; p.z_LUFS = z_LUFS;
        if (spline2d_init(S.r_grid, S.z_grid, Br, S.Br_fspl) || spline2d_init(S.r_grid, S.z_grid, Bz, S.Bz_fspl) ||
            spline2d_init(S.r_grid, S.z_grid, Aphi, S.Aphi_fspl)) return fail("cube_spline_2D_init: init, grid not evenly spaced or error");
        rays_spline2d base; base.nx = (int)S.r_grid.size(); base.ny = (int)S.z_grid.size(); base.x_grid = S.r_grid.data(); base.y_grid = S.z_grid.data();
        p.Br_spline = base; p.Br_spline.fspl = S.Br_fspl.data();
        p.Bz_spline = base; p.Bz_spline.fspl = S.Bz_fspl.data();
        p.Aphi_spline = base; p.Aphi_spline.fspl = S.Aphi_fspl.data();
        p.Aphi_LUFS = spline2d_eval_f(p.Aphi_spline, r_LUFS, z_LUFS);
    }
    dens = trim(dens);
    int dc = prof_code(dens);
    if (dc != RAYS_PROF_CONSTANT && dc != RAYS_PROF_PARABOLIC && dc != RAYS_PROF_HYPERBOLIC)
        return fail("multiple_mirror_eq: Unknown density_prof_model: " + dens);
    p.density_prof_model = dc;
    for (int is = 0; is <= S.species.nspec; ++is) {
        std::string t = trim(tprof[is]);
        int tc = prof_code(t);
        // 'hyperbolic' is evaluated by multiple_mirror_eq but rejected at init (multiple_mirror_eq_m.f90:200-213)
        if (tc != RAYS_PROF_ZERO && tc != RAYS_PROF_CONSTANT && tc != RAYS_PROF_PARABOLIC)
            return fail("multiple_mirror_eq: Unknown temperature_prof_model: " + t);
        p.temperature_prof_model[is] = tc;
    }
    return 0;
}

static int initialize_equilibrium_m(State &S, const NamelistFile &nml) {
    NamelistGroup G("equilibrium_list");
    G.add("equilib_model", &S.equilib_model);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    std::string m = trim(S.equilib_model);
    if (m == "slab") { S.cfg.equilib_model = RAYS_EQ_SLAB; return initialize_slab_eq_m(S, nml); }
    if (m == "solovev") { S.cfg.equilib_model = RAYS_EQ_SOLOVEV; return initialize_solovev_eq_m(S, nml); }
    if (m == "axisym_toroid") { S.cfg.equilib_model = RAYS_EQ_AXISYM_TOROID; return initialize_axisym_toroid_eq_m(S, nml); }
    if (m == "multiple_mirror") { S.cfg.equilib_model = RAYS_EQ_MULTIPLE_MIRROR; return initialize_multiple_mirror_eq_m(S, nml); }
    return fail("initialize_equilibrium: improper equilib_model =" + m);
}

static int initialize_ode_solver_m(State &S, const NamelistFile &nml) {
    ode_m &o = S.ode;
    NamelistGroup G("ode_list");
    G.add("ode_solver_name", &o.ode_solver_name); G.add("ray_deriv_name", &o.ray_deriv_name);
    G.add("nstep_max", &o.nstep_max); G.add("s_max", &o.s_max); G.add("ds", &o.ds);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    std::string solver = trim(o.ode_solver_name);
    if (solver == "SG_ODE") {
        NamelistGroup H("SG_ode_list");
        H.add("rel_err0", &o.rel_err0); H.add("abs_err0", &o.abs_err0); H.add("sg_error_limit", &o.SG_error_limit);
        if (!H.read(nml, err)) return fail(err);
        if (o.rel_err0 < f32(1.e-10) || o.abs_err0 < f32(1.e-10)) return fail("initialize_SG_ode: rel_err0, abs_err0 too small");
    } else if (solver != "RK4_ODE") return fail("read_ode_namelists, invalid ode solver = " + solver);
    o.nv = 7;
    if (trim(S.damping.damping_model) != "no_damp") o.nv = o.nv + 1;
    if (S.damping.multi_spec_damping) o.nv = o.nv + 1 + S.species.nspec;
    if (S.diagnostics.integrate_eq_gradients) o.nv = o.nv + 5;
    return 0;
}

// fills the marshalled POD from the module state (what the ISO_C_BINDING trace_rays would pack)
static int pack_cfg(State &S) {
    rays_cfg &c = S.cfg;
    c.clight = S.constants.clight; c.eps0 = S.constants.eps0;
    c.omgrf = S.rf.omgrf; c.k0 = S.rf.k0; c.dispersion_resid_limit = S.rf.dispersion_resid_limit;
    std::string rp = trim(S.rf.ray_param), wm = trim(S.rf.wave_mode);
    if (rp == "arcl") c.ray_param = RAYS_PARAM_ARCL; else if (rp == "time") c.ray_param = RAYS_PARAM_TIME;
    else return fail("EQN_RAY: invalid ray parameter = " + rp);
    if (wm == "plus") c.wave_mode = RAYS_MODE_PLUS; else if (wm == "minus") c.wave_mode = RAYS_MODE_MINUS;
    else if (wm == "fast") c.wave_mode = RAYS_MODE_FAST; else if (wm == "slow") c.wave_mode = RAYS_MODE_SLOW;
    else return fail("solve_disp: improper wave_mode = " + wm);
    if (trim(S.rf.ray_dispersion_model) != "cold") return fail("check_save: unimplemented ray_dispersion_model");
    c.k0_sign = S.rf.k0_sign;
    c.nspec = S.species.nspec;
    for (int i = 0; i < 6; ++i) { c.qs[i] = S.species.qs[i]; c.ms[i] = S.species.ms[i]; c.n0s[i] = S.species.n0s[i]; c.t0s[i] = S.species.t0s[i]; c.eta[i] = S.species.eta[i]; }
    std::string dm = trim(S.damping.damping_model);
    if (dm == "no_damp") c.damping_model = RAYS_DAMP_NONE; else if (dm == "damp_fund_ECH") c.damping_model = RAYS_DAMP_FUND_ECH;
    else return fail("damping: Unimplemented damping model " + dm);
    c.multi_spec_damping = S.damping.multi_spec_damping ? 1 : 0;
    c.total_damping_limit = S.damping.total_damping_limit;
    c.integrate_eq_gradients = S.diagnostics.integrate_eq_gradients ? 1 : 0;
    c.zfun_re.nx = ZFUN_NX; c.zfun_re.x_grid = S.z_x.data(); c.zfun_re.fspl = S.z_re.data();
    return 0;
}
static int pack_ode(State &S) {
    rays_cfg &c = S.cfg;
    std::string solver = trim(S.ode.ode_solver_name), deriv = trim(S.ode.ray_deriv_name);
    c.ode_solver = solver == "SG_ODE" ? RAYS_ODE_SG : RAYS_ODE_RK4;
    if (deriv == "cold") c.ray_deriv = RAYS_DERIV_COLD; else if (deriv == "numerical") c.ray_deriv = RAYS_DERIV_NUM;
    else return fail("EQN_RAY: invalid value, ray_deriv_name = " + deriv);
    c.nv = S.ode.nv; c.nstep_max = S.ode.nstep_max; c.ds = S.ode.ds; c.s_max = S.ode.s_max;
    c.rel_err0 = S.ode.rel_err0; c.abs_err0 = S.ode.abs_err0; c.SG_error_limit = S.ode.SG_error_limit;
    return 0;
}

// initialize_ray_init_m (ray_init_m.f90:72-127): reads the launcher's namelist, then builds the
// fan ON THE DEVICE (rays_b200_launch_fan_*), and mirrors it into the module arrays.
static int initialize_ray_init_m(State &S, const NamelistFile &nml, bool do_launch) {
    ray_init_m &ri = S.ray_init;
    NamelistGroup G("ray_init_list");
    G.add("ray_init_model", &ri.ray_init_model); G.add("nray_max", &ri.nray_max);
    std::string err;
    if (!G.read(nml, err)) return fail(err);
    std::string m = trim(ri.ray_init_model);
    int64_t nray = 0, ncand = 0;
    int rc = 0;
    if (m == "simple_slab") {
        rays_slab_launch &p = ri.slab;
        p.n_x_launch = p.n_y_launch = p.n_z_launch = 1;
        NamelistGroup L("simple_slab_ray_init_list");
        L.add("n_x_launch", &p.n_x_launch); L.add("x_launch0", &p.x_launch0); L.add("dx_launch", &p.dx_launch);
        L.add("n_y_launch", &p.n_y_launch); L.add("y_launch0", &p.y_launch0); L.add("dy_launch", &p.dy_launch);
        L.add("n_z_launch", &p.n_z_launch); L.add("z_launch0", &p.z_launch0); L.add("dz_launch", &p.dz_launch);
        L.add("n_ky_launch", &p.n_ky_launch); L.add("rindex_y0", &p.rindex_y0); L.add("delta_rindex_y0", &p.delta_rindex_y0);
        L.add("n_kz_launch", &p.n_kz_launch); L.add("rindex_z0", &p.rindex_z0); L.add("delta_rindex_z0", &p.delta_rindex_z0);
        if (!L.read(nml, err)) return fail(err);
        ncand = (int64_t)p.n_x_launch * p.n_ky_launch * p.n_kz_launch;  // (R) ignores n_y, n_z (simple_slab_ray_init_m.f90:108)
        if (ncand <= 0 || ncand > ri.nray_max) return fail("simple slab ray init: improper number of rays");
        if (do_launch) rc = rays_b200_launch_fan_slab(&p, &nray);
    } else if (m == "solovev") {
        rays_solovev_launch &p = ri.solovev;
        p.n_r_launch = p.n_theta_launch = p.n_rindex_theta = p.n_rindex_phi = 1;
        NamelistGroup L("solovev_ray_init_nphi_ktheta_list");
        L.add("n_r_launch", &p.n_r_launch); L.add("r_launch0", &p.r_launch0); L.add("dr_launch", &p.dr_launch);
        L.add("n_theta_launch", &p.n_theta_launch); L.add("theta_launch0", &p.theta_launch0); L.add("dtheta_launch", &p.dtheta_launch);
        L.add("n_rindex_theta", &p.n_rindex_theta); L.add("rindex_theta0", &p.rindex_theta0); L.add("delta_rindex_theta", &p.delta_rindex_theta);
        L.add("n_rindex_phi", &p.n_rindex_phi); L.add("rindex_phi0", &p.rindex_phi0); L.add("delta_rindex_phi", &p.delta_rindex_phi);
        if (!L.read(nml, err)) return fail(err);
        ncand = (int64_t)p.n_r_launch * p.n_theta_launch * p.n_rindex_theta * p.n_rindex_phi;
        if (ncand <= 0 || ncand > ri.nray_max) return fail("solovev ray init: improper number of rays");
        if (S.cfg.equilib_model != RAYS_EQ_SOLOVEV) return fail("ray_init_model 'solovev' needs equilib_model 'solovev'");
        if (do_launch) rc = rays_b200_launch_fan_solovev(&p, &nray);
    } else if (m == "axisym_toroid_ray_init_R_Z_nphi_ntheta") {
        rays_axisym_launch &p = ri.axisym;
        p.n_R_launch = p.n_Z_launch = p.n_rindex_theta = p.n_rindex_phi = 1;
        NamelistGroup L("axisym_toroid_ray_init_R_Z_nphi_ntheta_list");
        L.add("n_r_launch", &p.n_R_launch); L.add("r_launch0", &p.R_launch0);
        L.add("n_z_launch", &p.n_Z_launch); L.add("z_launch0", &p.Z_launch0);
        L.add("n_rindex_theta", &p.n_rindex_theta); L.add("rindex_theta0", &p.rindex_theta0); L.add("delta_rindex_theta", &p.delta_rindex_theta);
        L.add("n_rindex_phi", &p.n_rindex_phi); L.add("rindex_phi0", &p.rindex_phi0); L.add("delta_rindex_phi", &p.delta_rindex_phi);
        if (!L.read(nml, err)) return fail(err);
        ncand = (int64_t)p.n_R_launch * p.n_Z_launch * p.n_rindex_theta * p.n_rindex_phi;
        if (ncand <= 0 || ncand > ri.nray_max) return fail("axisym_toroid ray init: improper number of rays");
        if (S.cfg.equilib_model != RAYS_EQ_AXISYM_TOROID) return fail("axisym_toroid ray init needs equilib_model 'axisym_toroid'");
        if (do_launch) rc = rays_b200_launch_fan_axisym(&p, &nray);
    } else if (m == "one_ray_init_XYZ_n_direction" || m == "file_input_ray_init") {
        const bool from_file = (m == "file_input_ray_init");
        NamelistFile f2;
        const NamelistFile *src = &nml;
        if (from_file) {
            ri.rvec_in.assign((size_t)3 * ri.nray_max, 0.0); ri.rindex_vec_in.assign((size_t)3 * ri.nray_max, 0.0);
            ri.ray_pwr_wt_in.assign((size_t)ri.nray_max, 1.0);
            NamelistGroup L("file_input_ray_init_list");
            L.add("n_rays_in", &ri.n_rays_in);
            L.add_arr("rvec_in", ri.rvec_in.data(), 1, 3 * ri.nray_max);
            L.add_arr("rindex_vec_in", ri.rindex_vec_in.data(), 1, 3 * ri.nray_max);
            L.add_arr("ray_pwr_wt_in", ri.ray_pwr_wt_in.data(), 1, ri.nray_max);
            // file 'ray_init_<run_label>.in' (file_input_ray_init_m.f90:118-125); fall back to rays.in
            std::string p2 = (S.workdir.empty() ? std::string() : S.workdir + "/") + "ray_init_" + trim(S.diagnostics.run_label) + ".in";
            if (f2.load(p2)) src = &f2;
            if (!L.read(*src, err)) return fail(err);
            if (ri.n_rays_in < 1 || ri.n_rays_in > ri.nray_max) return fail("file_input_ray_init: improper number of rays");
        } else {
            double X = 0, Y = 0, Z = 0, nX = 0, nY = 0, nZ = 0;
            NamelistGroup L("one_ray_init_XYZ_k_direction_list");
            L.add("x", &X); L.add("y", &Y); L.add("z", &Z); L.add("nx", &nX); L.add("ny", &nY); L.add("nz", &nZ);
            L.add("use_this_n_vec", &ri.use_this_n_vec);
            if (!L.read(nml, err)) return fail(err);
            ri.n_rays_in = 1;
            ri.rvec_in = {X, Y, Z}; ri.rindex_vec_in = {nX, nY, nZ}; ri.ray_pwr_wt_in = {1.0};
        }
        if (!from_file && ri.use_this_n_vec) {  // one_ray_init_XYZ_k_direction_m.f90:100-105: no dispersion solve
            ri.nray = 1; ri.rvec0 = ri.rvec_in; ri.rindex_vec0 = ri.rindex_vec_in; ri.ray_pwr_wt = {1.0};
            return 0;
        }
        if (do_launch) rc = rays_b200_launch_fan_directions(ri.n_rays_in, ri.rvec_in.data(), ri.rindex_vec_in.data(), &nray);
    } else
        return fail("initialize_ray_init: invalid ray_init_model = " + m);
    if (!do_launch) { ri.nray = 0; return 0; }
    if (rc) { g_err = rays_b200_last_error(); return rc; }
    if (nray == 0) return fail("No successful ray initializations");
    ri.nray = nray;
    ri.rvec0.assign((size_t)3 * nray, 0.0); ri.rindex_vec0.assign((size_t)3 * nray, 0.0); ri.ray_pwr_wt.assign((size_t)nray, 0.0);
    rc = rays_b200_fan_download(ri.rvec0.data(), ri.rindex_vec0.data(), ri.ray_pwr_wt.data());
    if (rc) { g_err = rays_b200_last_error(); return rc; }
    // ray weights as the reference's launchers leave them
    if (m == "simple_slab") for (auto &w : ri.ray_pwr_wt) w = 1.0 / (double)nray / (double)nray;       // (R) divided twice
    else if (m == "file_input_ray_init") {
        double mx = 0.0; for (int i = 0; i < ri.nray_max; ++i) mx = std::max(mx, ri.ray_pwr_wt_in[i]);
        // (R) maxval(ray_pwr_wt_in) == 0 -> 1/nray ; else ray_pwr_wt_temp (never filled: 0) * n_rays_in/nray
        for (auto &w : ri.ray_pwr_wt) w = (mx == 0.0) ? 1.0 / (double)nray : 0.0;
    } else if (m == "one_ray_init_XYZ_n_direction") for (auto &w : ri.ray_pwr_wt) w = 1.0;
    else for (auto &w : ri.ray_pwr_wt) w = 1.0 / (double)nray;
    return 0;
}

// initialize_ray_results_m (ray_results_m.f90:107-165): the dense, zero-filled result arrays.  They are
// allocated on first use (trace_rays / results / finalize_run) rather than inside initialize(): a host that only
// wants the on-device deposition profile of a multi-million-ray fan never needs ray_vec(nv, nstep_max+1, nray).
static void initialize_ray_results_m(State &S) { S.results.allocated = false; }
static void allocate_ray_results(State &S) {
    ray_results_m &r = S.results;
    if (r.allocated) return;
    r.allocated = true;
    const int nv = S.ode.nv;
    const int64_t nray = S.ray_init.nray;
    const int np = S.ode.nstep_max + 1;
    r.nv = nv; r.number_of_rays = nray; r.max_number_of_points = np;
    r.ray_vec.assign((size_t)nv * np * nray, 0.0);
    r.residual.assign((size_t)np * nray, 0.0);
    r.npoints.assign(nray, 0); r.ray_stop_code.assign(nray, 0);
    r.initial_ray_power.assign(nray, 0.0); r.ray_trace_time.assign(nray, 0.0); r.end_ray_parameter.assign(nray, 0.0);
    r.end_residuals.assign(nray, 0.0); r.max_residuals.assign(nray, 0.0);
    r.start_ray_vec.assign((size_t)nv * nray, 0.0); r.end_ray_vec.assign((size_t)nv * nray, 0.0);
    r.ray_stop_flag.assign((size_t)RAYS_FLAG_LEN * nray, ' ');
}

}  // namespace rays_host

using namespace rays_host;

extern "C" {

const char *rays_host_last_error(void) { return g_err.c_str(); }

// initialize(read_input) (RAYS_lib/intialize.f90:1-94).  `namelist_path` plays the role of rays.in
// (the reference copies argv[1] to rays.in, diagnostics_m.f90:132-140); relative data files (the
// Brz netCDF file, ray_init_<label>.in) are looked up next to it.  do_ray_init = 0 skips the
// device launch-fan step (the caller then provides a fan with rays_host_set_fan).
int rays_host_initialize(const char *namelist_path, int do_ray_init) {
    delete g;
    g = new State();
    State &S = *g;
    g_err.clear();
    S.namelist_path = namelist_path ? namelist_path : "rays.in";
    size_t slash = S.namelist_path.find_last_of('/');
    S.workdir = slash == std::string::npos ? std::string() : S.namelist_path.substr(0, slash);
    NamelistFile nml;
    if (!nml.load(S.namelist_path)) return fail(nml.error());
    int rc;
    if ((rc = initialize_diagnostics(S, nml))) return rc;
    S.constants.initialize();
    if ((rc = initialize_species_m(S, nml))) return rc;
    if ((rc = initialize_rf_m(S, nml))) return rc;
    if ((rc = initialize_damping_m(S, nml))) return rc;
    if ((rc = initialize_equilibrium_m(S, nml))) return rc;
    if ((rc = pack_cfg(S))) return rc;
    // ode first here only to know nv for cfg validation on the device; order-independent otherwise
    if ((rc = initialize_ode_solver_m(S, nml))) return rc;
    if ((rc = pack_ode(S))) return rc;
    if (do_ray_init) {
        if ((rc = rays_b200_set_config(&S.cfg))) { g_err = rays_b200_last_error(); return rc; }
    }
    if ((rc = initialize_ray_init_m(S, nml, do_ray_init != 0))) return rc;
    if (nml.has_group("ray_results_list")) {
        NamelistGroup R("ray_results_list");
        R.add("write_results_list_directed", &S.results.write_results_list_directed);
        R.add("write_results_netcdf", &S.results.write_results_netCDF);
        std::string err;
        if (!R.read(nml, err)) return fail(err);
    }
    initialize_ray_results_m(S);
    S.initialized = true;
    return 0;
}

const rays_cfg *rays_host_cfg(void) { return g ? &g->cfg : nullptr; }
int rays_host_nspec(void) { return g ? g->species.nspec : -1; }
const char *rays_host_run_label(void) { return g ? g->diagnostics.run_label.c_str() : ""; }
const char *rays_host_ray_init_model(void) { return g ? g->ray_init.ray_init_model.c_str() : ""; }
// launch-fan namelist data (for tests that run the launcher on the oracle)
int rays_host_launch_params(rays_slab_launch *slab, rays_solovev_launch *sol, rays_axisym_launch *axi) {
    if (!g) return RAYS_ERR_NOT_INITIALIZED;
    if (slab) *slab = g->ray_init.slab;
    if (sol) *sol = g->ray_init.solovev;
    if (axi) *axi = g->ray_init.axisym;
    return 0;
}
int64_t rays_host_directions_in(const double **rvec_in, const double **nvec_in) {
    if (!g) return 0;
    if (rvec_in) *rvec_in = g->ray_init.rvec_in.data();
    if (nvec_in) *nvec_in = g->ray_init.rindex_vec_in.data();
    return g->ray_init.n_rays_in;
}

// the host "pokes module variables between initialize(.false.) calls" (ray_scan.f90:33-49)
int rays_host_set_ode(const char *ode_solver_name, const char *ray_deriv_name, int nstep_max, double s_max, double ds,
                      double rel_err0, double abs_err0, double SG_error_limit) {
    if (!g) return RAYS_ERR_NOT_INITIALIZED;
    State &S = *g;
    if (ode_solver_name && *ode_solver_name) S.ode.ode_solver_name = ode_solver_name;
    if (ray_deriv_name && *ray_deriv_name) S.ode.ray_deriv_name = ray_deriv_name;
    if (nstep_max > 0) S.ode.nstep_max = nstep_max;
    if (s_max > 0) S.ode.s_max = s_max;
    if (ds > 0) S.ode.ds = ds;
    if (rel_err0 > 0) S.ode.rel_err0 = rel_err0;
    if (abs_err0 > 0) S.ode.abs_err0 = abs_err0;
    if (SG_error_limit > 0) S.ode.SG_error_limit = SG_error_limit;
    std::string solver = trim(S.ode.ode_solver_name);
    if (solver != "SG_ODE" && solver != "RK4_ODE") return fail("read_ode_namelists, invalid ode solver = " + solver);
    if (solver == "SG_ODE" && (S.ode.rel_err0 < f32(1.e-10) || S.ode.abs_err0 < f32(1.e-10)))
        return fail("initialize_SG_ode: rel_err0, abs_err0 too small");
    int rc = pack_ode(S);
    if (rc) return rc;
    initialize_ray_results_m(S);
    return 0;
}

int rays_host_set_fan(int64_t nray, const double *rvec0, const double *rindex_vec0, const double *ray_pwr_wt) {
    if (!g) return RAYS_ERR_NOT_INITIALIZED;
    ray_init_m &ri = g->ray_init;
    ri.nray = nray;
    ri.rvec0.assign(rvec0, rvec0 + 3 * nray);
    ri.rindex_vec0.assign(rindex_vec0, rindex_vec0 + 3 * nray);
    if (ray_pwr_wt) ri.ray_pwr_wt.assign(ray_pwr_wt, ray_pwr_wt + nray);
    else ri.ray_pwr_wt.assign(nray, nray ? 1.0 / (double)nray : 0.0);
    initialize_ray_results_m(*g);
    return 0;
}
int64_t rays_host_get_fan(const double **rvec0, const double **rindex_vec0, const double **ray_pwr_wt) {
    if (!g) return 0;
    if (rvec0) *rvec0 = g->ray_init.rvec0.data();
    if (rindex_vec0) *rindex_vec0 = g->ray_init.rindex_vec0.data();
    if (ray_pwr_wt) *ray_pwr_wt = g->ray_init.ray_pwr_wt.data();
    return g->ray_init.nray;
}

// views of the module arrays of ray_results_m, in the layout rays_results documents
int rays_host_results(rays_results *out) {
    if (!g) return RAYS_ERR_NOT_INITIALIZED;
    allocate_ray_results(*g);
    ray_results_m &r = g->results;
    out->nray = r.number_of_rays; out->nv = r.nv; out->npoints_alloc = r.max_number_of_points;
    out->ray_vec = r.ray_vec.data(); out->residual = r.residual.data(); out->npoints = r.npoints.data();
    out->ray_stop_code = r.ray_stop_code.data(); out->ray_stop_flag = r.ray_stop_flag.data();
    out->initial_ray_power = r.initial_ray_power.data(); out->ray_trace_time = r.ray_trace_time.data();
    out->end_residuals = r.end_residuals.data(); out->max_residuals = r.max_residuals.data();
    out->end_ray_parameter = r.end_ray_parameter.data(); out->start_ray_vec = r.start_ray_vec.data();
    out->end_ray_vec = r.end_ray_vec.data(); out->total_trace_time = r.total_trace_time; out->total_ray_steps = r.total_ray_steps;
    return 0;
}

// trace_rays (RAYS_lib/ray_tracing.f90:1-290): the call the C ABI replaces
int rays_host_trace_rays(void) {
    if (!g || !g->initialized) { g_err = "trace_rays called before initialize"; return RAYS_ERR_NOT_INITIALIZED; }
    State &S = *g;
    rays_fan fan;
    fan.nray = S.ray_init.nray; fan.rvec0 = S.ray_init.rvec0.data(); fan.rindex_vec0 = S.ray_init.rindex_vec0.data();
    fan.ray_pwr_wt = S.ray_init.ray_pwr_wt.data();
    rays_results res;
    rays_host_results(&res);
    int rc = rays_b200_trace(&S.cfg, &fan, &res);
    if (rc) { g_err = rays_b200_last_error(); return rc; }
    S.results.total_trace_time = res.total_trace_time;
    S.results.total_ray_steps = res.total_ray_steps;
    return 0;
}

// write_deposition_profiles_NC (post_process_lib/deposition_profiles_m.f90:336-420): deposition_profiles.<run_label>.nc with
// the reference's structure -- record (unlimited) dimension n_profiles, n_bins, n_bins_p1, d20; per-profile Q_sum, n_bins,
// grid_min, grid_max, profile_name, grid_name, grid (edges: grid_min + delta*(i-1), :151-159), profile; global attributes
// RAYS_run_label and date_vector.  names: n_profiles x 20 characters, blank padded; profile: [n_profiles][n_bins].
int rays_host_write_deposition_profiles(const char *outdir, int n_profiles, const char *profile_names, const char *grid_names, int n_bins,
                                        const double *grid_min, const double *grid_max, const double *profile, const double *Q_sum) {
    if (!g || !g->initialized) { g_err = "write_deposition_profiles called before initialize"; return RAYS_ERR_NOT_INITIALIZED; }
    if (n_profiles < 1 || n_bins < 1 || !profile_names || !grid_names || !grid_min || !grid_max || !profile || !Q_sum)
        return fail("write_deposition_profiles_NC: bad arguments");
    State &S = *g;
    NcWriter w;
    const int d_prof = w.def_unlimited_dim("n_profiles");   // NF90_UNLIMITED
    const int d_bins = w.def_dim("n_bins", n_bins);
    const int d_bp1 = w.def_dim("n_bins_p1", n_bins + 1);
    const int d20 = w.def_dim("d20", 20);
    const int v_q = w.def_var("Q_sum", NC_DOUBLE, {d_prof});
    const int v_nb = w.def_var("n_bins", NC_INT, {d_prof});
    const int v_gmin = w.def_var("grid_min", NC_DOUBLE, {d_prof});
    const int v_gmax = w.def_var("grid_max", NC_DOUBLE, {d_prof});
    const int v_pn = w.def_var("profile_name", NC_CHAR, {d_prof, d20});      // Fortran [d20, n_profiles]
    const int v_gn = w.def_var("grid_name", NC_CHAR, {d_prof, d20});
    const int v_grid = w.def_var("grid", NC_DOUBLE, {d_prof, d_bp1});
    const int v_p = w.def_var("profile", NC_DOUBLE, {d_prof, d_bins});
    w.put_att_text("RAYS_run_label", S.diagnostics.run_label);
    w.put_att_int("date_vector", std::vector<int32_t>(S.diagnostics.date_v, S.diagnostics.date_v + 8));
    w.set_numrecs(n_profiles);
    std::vector<int32_t> nb((size_t)n_profiles, n_bins);
    std::vector<double> grid((size_t)n_profiles * (n_bins + 1));
    for (int p = 0; p < n_profiles; ++p) {
        const double delta = (grid_max[p] - grid_min[p]) / (double)(float)n_bins;     // real(n_bins): default real
        for (int i = 1; i <= n_bins + 1; ++i) grid[(size_t)p * (n_bins + 1) + i - 1] = grid_min[p] + delta * (i - 1);
    }
    w.put_double(v_q, Q_sum, (size_t)n_profiles);
    w.put_int(v_nb, nb.data(), nb.size());
    w.put_double(v_gmin, grid_min, (size_t)n_profiles);
    w.put_double(v_gmax, grid_max, (size_t)n_profiles);
    w.put_char(v_pn, profile_names, (size_t)n_profiles * 20);
    w.put_char(v_gn, grid_names, (size_t)n_profiles * 20);
    w.put_double(v_grid, grid.data(), grid.size());
    w.put_double(v_p, profile, (size_t)n_profiles * n_bins);
    const std::string full = (outdir && outdir[0] ? std::string(outdir) + "/" : std::string()) + "deposition_profiles." + trim(S.diagnostics.run_label) + ".nc";
    std::string err;
    if (!w.close(full, err)) { g_err = err; return RAYS_ERR_IO; }
    return 0;
}

// program mirror_magnetics (mirror_magnetics_lib/mirror_magnetics.f90:1-137): initialize_mirror_magnetics
// (mirror_magnetics_m.f90:134-243: /mirror_magnetics_list/ from `namelist_path`, /coil_data_list/ and
// /current_data_list/ from the files it names), calculate_B_on_rz_grid on the GPU (rays_b200_mirror_brz_grid),
// write_mirror_fields_Brz_NC (:372-450) -> <outdir>/Brz_fields.<coil_set>_<current_set>_<case>.nc, the file
// mirror_magnetics_spline_interp reads.  out_path (optional) receives the file name.
int rays_host_mirror_magnetics(const char *namelist_path, const char *outdir, char *out_path, int out_len) {
    std::string path = namelist_path ? namelist_path : "mirror_magnetics.nml";
    size_t slash = path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? std::string() : path.substr(0, slash);
    auto resolve = [&](std::string f) { f = trim(f); return (!f.empty() && f[0] != '/' && !dir.empty()) ? dir + "/" + f : f; };
    NamelistFile nml;
    if (!nml.load(path)) return fail(nml.error());
    std::string coil_data_file, current_data_file, case_name, err;
    int n_r = 0, n_z = 0, n_coils = 0;
    double r_min = 0.0, r_max = 0.0, z_min = 0.0, z_max = 0.0, r_LUFS = 0.0, z_LUFS = 0.0;
    NamelistGroup G("mirror_magnetics_list");
    G.add("case_name", &case_name); G.add("coil_data_file", &coil_data_file); G.add("current_data_file", &current_data_file);
    G.add("n_coils", &n_coils); G.add("n_r", &n_r); G.add("n_z", &n_z); G.add("r_min", &r_min); G.add("r_max", &r_max);
    G.add("z_min", &z_min); G.add("z_max", &z_max); G.add("r_lufs", &r_LUFS); G.add("z_lufs", &z_LUFS);
    if (!G.read(nml, err)) return fail(err);
    if (n_coils < 1 || n_r < 1 || n_z < 1) return fail("mirror_magnetics_list: n_coils, n_r, n_z must be >= 1");
    const size_t nc = (size_t)n_coils;
    std::vector<double> inner_radius(nc, 0.0), outer_radius(nc, 0.0), z_width(nc, 0.0), z_center(nc, 0.0), I_coil(nc, 0.0);
    std::vector<int> n_turns(nc, 0), n_r_layers(nc, 0), n_z_slices(nc, 0);
    std::string coil_set_name, current_set_name;
    NamelistFile cf, uf;
    if (!cf.load(resolve(coil_data_file))) return fail(cf.error());
    NamelistGroup Cg("coil_data_list");
    Cg.add("coil_set_name", &coil_set_name);
    Cg.add_arr("inner_radius", inner_radius.data(), 1, n_coils); Cg.add_arr("outer_radius", outer_radius.data(), 1, n_coils);
    Cg.add_arr("z_width", z_width.data(), 1, n_coils); Cg.add_arr("z_center", z_center.data(), 1, n_coils);
    Cg.add_arr("n_turns", n_turns.data(), 1, n_coils); Cg.add_arr("n_r_layers", n_r_layers.data(), 1, n_coils);
    Cg.add_arr("n_z_slices", n_z_slices.data(), 1, n_coils);
    if (!Cg.read(cf, err)) return fail(err);
    if (!uf.load(resolve(current_data_file))) return fail(uf.error());
    NamelistGroup Ug("current_data_list");
    Ug.add("current_set_name", &current_set_name); Ug.add_arr("i_coil", I_coil.data(), 1, n_coils);
    if (!Ug.read(uf, err)) return fail(err);
    std::vector<rays_coil> coils(nc);
    for (size_t i = 0; i < nc; ++i) {
        rays_coil &c = coils[i];
        c.inner_radius = inner_radius[i]; c.outer_radius = outer_radius[i]; c.z_center = z_center[i]; c.z_width = z_width[i];
        c.I_coil = I_coil[i]; c.n_turns = n_turns[i]; c.n_r_layers = n_r_layers[i]; c.n_z_slices = n_z_slices[i];
    }
    const size_t n = (size_t)n_r * n_z;
    std::vector<double> r_grid((size_t)n_r), z_grid((size_t)n_z), Br(n), Bz(n), Aphi(n);
    int rc = rays_b200_mirror_brz_grid(coils.data(), n_coils, n_r, r_min, r_max, n_z, z_min, z_max, r_grid.data(), z_grid.data(), Br.data(),
                                       Bz.data(), Aphi.data());
    if (rc) { g_err = rays_b200_last_error(); return rc; }
    const std::string base_file_name = trim(coil_set_name) + "_" + trim(current_set_name) + "_" + trim(case_name);
    const std::string nc_name = "Brz_fields." + base_file_name + ".nc";
    NcWriter w;
    int d_r = w.def_dim("n_r", n_r), d_z = w.def_dim("n_z", n_z);
    int v0 = w.def_var("r_min", NC_DOUBLE, {}), v1 = w.def_var("r_max", NC_DOUBLE, {}), v2 = w.def_var("z_min", NC_DOUBLE, {});
    int v3 = w.def_var("z_max", NC_DOUBLE, {}), v4 = w.def_var("r_LUFS", NC_DOUBLE, {}), v5 = w.def_var("z_LUFS", NC_DOUBLE, {});
    int v6 = w.def_var("r_grid", NC_DOUBLE, {d_r}), v7 = w.def_var("z_grid", NC_DOUBLE, {d_z});
    // Fortran dims [n_r_id, n_z_id] are stored reversed: C order (n_z, n_r)
    int v8 = w.def_var("Br", NC_DOUBLE, {d_z, d_r}), v9 = w.def_var("Bz", NC_DOUBLE, {d_z, d_r}), v10 = w.def_var("Aphi", NC_DOUBLE, {d_z, d_r});
    w.put_att_text("NC_file_name", nc_name);
    w.put_double(v0, &r_min, 1); w.put_double(v1, &r_max, 1); w.put_double(v2, &z_min, 1); w.put_double(v3, &z_max, 1);
    w.put_double(v4, &r_LUFS, 1); w.put_double(v5, &z_LUFS, 1);
    w.put_double(v6, r_grid.data(), r_grid.size()); w.put_double(v7, z_grid.data(), z_grid.size());
    w.put_double(v8, Br.data(), n); w.put_double(v9, Bz.data(), n); w.put_double(v10, Aphi.data(), n);
    const std::string full = (outdir && outdir[0] ? std::string(outdir) + "/" : std::string()) + nc_name;
    if (!w.close(full, err)) { g_err = err; return RAYS_ERR_IO; }
    if (out_path && out_len > 0) { std::strncpy(out_path, full.c_str(), (size_t)out_len - 1); out_path[out_len - 1] = 0; }
    return 0;
}

// finalize_run (RAYS_lib/finalize_run.f90:1-51) -> write_results_NC (ray_results_m.f90:171-249):
// writes <outdir>/run_results.<run_label>.nc in netCDF classic format with the reference's
// dimensions, variable names and types so post_process_RAYS / graphics_RAYS read it unchanged.
int rays_host_finalize_run(const char *outdir) {
    if (!g || !g->initialized) { g_err = "finalize_run called before initialize"; return RAYS_ERR_NOT_INITIALIZED; }
    State &S = *g;
    allocate_ray_results(S);
    ray_results_m &r = S.results;
    if (!r.write_results_netCDF) return 0;
    const int64_t nray = r.number_of_rays;
    const int nv = r.nv, np = r.max_number_of_points;
    int amax = 0;
    for (int64_t i = 0; i < nray; ++i) amax = std::max(amax, (int)r.npoints[i]);  // actual_max_npoints
    if (nray < 1 || amax < 1) return fail("write_results_NC: no rays to write (number_of_rays = 0): the classic format has no zero-length fixed dimension");
    NcWriter w;
    int d_rays = w.def_dim("number_of_rays", nray);
    int d_pts = w.def_dim("max_number_of_points", amax);
    int d_v = w.def_dim("dim_v_vector", nv);
    int d8 = w.def_dim("d8", 8);
    int d60 = w.def_dim("d60", 60);
    // Fortran dims (a,b,c) are stored reversed: C order (c,b,a)
    int v_date = w.def_var("date_vector", NC_INT, {d8});
    int v_rv = w.def_var("ray_vec", NC_DOUBLE, {d_rays, d_pts, d_v});
    int v_res = w.def_var("residual", NC_DOUBLE, {d_rays, d_pts});
    int v_np = w.def_var("npoints", NC_INT, {d_rays});
    int v_pw = w.def_var("initial_ray_power", NC_FLOAT, {d_rays});
    int v_tt = w.def_var("ray_trace_time", NC_FLOAT, {d_rays});
    int v_er = w.def_var("end_residuals", NC_FLOAT, {d_rays});
    int v_mr = w.def_var("max_residuals", NC_FLOAT, {d_rays});
    int v_ep = w.def_var("end_ray_parameter", NC_FLOAT, {d_rays});
    int v_sv = w.def_var("start_ray_vec", NC_FLOAT, {d_rays, d_v});
    int v_ev = w.def_var("end_ray_vec", NC_FLOAT, {d_rays, d_v});
    int v_fl = w.def_var("ray_stop_flag", NC_CHAR, {d_rays, d60});
    int v_tot = w.def_var("total_trace_time", NC_FLOAT, {});
    w.put_att_text("RAYS_run_label", S.diagnostics.run_label);
    std::vector<int32_t> date(S.diagnostics.date_v, S.diagnostics.date_v + 8);
    w.put_int(v_date, date.data(), 8);
    std::vector<double> rv((size_t)nray * amax * nv), rs((size_t)nray * amax);
    for (int64_t i = 0; i < nray; ++i) {
        std::memcpy(&rv[(size_t)i * amax * nv], &r.ray_vec[(size_t)i * np * nv], sizeof(double) * (size_t)amax * nv);
        std::memcpy(&rs[(size_t)i * amax], &r.residual[(size_t)i * np], sizeof(double) * (size_t)amax);
    }
    w.put_double(v_rv, rv.data(), rv.size());
    w.put_double(v_res, rs.data(), rs.size());
    w.put_int(v_np, r.npoints.data(), nray);
    w.put_double(v_pw, r.initial_ray_power.data(), nray);
    w.put_double(v_tt, r.ray_trace_time.data(), nray);
    w.put_double(v_er, r.end_residuals.data(), nray);
    w.put_double(v_mr, r.max_residuals.data(), nray);
    w.put_double(v_ep, r.end_ray_parameter.data(), nray);
    w.put_double(v_sv, r.start_ray_vec.data(), (size_t)nray * nv);
    w.put_double(v_ev, r.end_ray_vec.data(), (size_t)nray * nv);
    w.put_char(v_fl, r.ray_stop_flag.data(), (size_t)nray * 60);
    w.put_double(v_tot, &r.total_trace_time, 1);
    std::string path = std::string(outdir && *outdir ? outdir : ".") + "/run_results." + trim(S.diagnostics.run_label) + ".nc";
    std::string err;
    if (!w.close(path, err)) { g_err = err; return RAYS_ERR_IO; }
    return 0;
}

int rays_host_deallocate(void) { delete g; g = nullptr; return 0; }

// small probes used by CPU tests of the host logic
int rays_host_zfun(double x, double y, double *re, double *im) { zfun_D(x, y, re, im); return 0; }
int rays_host_cspline(const double *x, int nx, double *fspl) { int il; return cspline(x, nx, fspl, &il); }
int rays_host_bcspline(const double *x, int nx, const double *y, int ny, double *fspl) { int a, b; return bcspline(x, nx, y, ny, fspl, &a, &b); }

}  // extern "C"
