// siafd_test.cc -- TEST: the shape of the reference's src/stressbalance/sia/siafd_test.cc (verification test F
// through the SIAFD class), driving the C++ host class SIAFD_B200 over the PETSc-free mirrors.
//   usage: siafd_test [-Mx N] [-My N] [-Mz N]
// Prints the reference's report line ("surf vels : maxUvec avUvec", m/year) and, for the pytest wrapper,
// D_max and a few checksums.  Exit code 0 = ok, 1 = pism::RuntimeError (message on stderr).
//
// The exact solution is the reference's own exactTestsFG.cc, compiled unmodified into
// oracle/_ref/libpism_exact.so (oracle/exact_wrap.cc); this driver is test code and may link it.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pism_b200/host/SIAFD_B200.hh"

extern "C" int ref_exactFG(double t, double r, int Mz, const double *z, double Cp, double *H, double *M, double *T,
                           double *U, double *w, double *Sig, double *Sigc);

using namespace pism;

namespace {

struct FG {
  double H, M;
  std::vector<double> T, U, w, Sig, Sigc;
};
FG exactFG(double t, double r, const std::vector<double> &z, double Cp) {
  FG f;
  const int n = (int)z.size();
  f.T.resize(n), f.U.resize(n), f.w.resize(n), f.Sig.resize(n), f.Sigc.resize(n);
  if (ref_exactFG(t, r, n, z.data(), Cp, &f.H, &f.M, f.T.data(), f.U.data(), f.w.data(), f.Sig.data(), f.Sigc.data())) {
    throw RuntimeError("exactFG failed");
  }
  return f;
}

// ColdEnthalpyConverter::enthalpy_permissive(T, 0, p) = c_i (T - T_0) (EnthalpyConverter.cc:298-313)
double enthalpy_cold(const Config &c, double T) {
  return c.get_number("constants.ice.specific_heat_capacity") * (T - c.get_number("enthalpy_converter.T_reference"));
}

// siafd_test.cc:181-224
void setInitStateF(const IceGrid &grid, IceModelVec2S &bed, IceModelVec2CellType &mask, IceModelVec2S &surface,
                   IceModelVec2S &thickness, IceModelVec3 &enthalpy) {
  const double ST = 1.67e-5, Tmin = 223.15, LforFG = 750000;
  bed.set(0.0);
  mask.set(MASK_GROUNDED);
  std::vector<double> T(grid.Mz());
  for (int j = grid.ys(); j < grid.ys() + grid.ym(); ++j) {
    for (int i = grid.xs(); i < grid.xs() + grid.xm(); ++i) {
      const double r = std::max(radius(grid, i, j), 1.0), Ts = Tmin + ST * r;
      if (r > LforFG - 1.0) {
        thickness(i, j) = 0.0;
        enthalpy.set_column(i, j, enthalpy_cold(*grid.config(), Ts));
      } else {
        FG F = exactFG(0.0, r, grid.z(), 0.0);
        thickness(i, j) = F.H;
        for (unsigned int k = 0; k < grid.Mz(); ++k) T[k] = enthalpy_cold(*grid.config(), F.T[k]);
        enthalpy.set_column(i, j, T.data());
      }
    }
  }
  thickness.update_ghosts();
  surface.copy_from(thickness);
  enthalpy.update_ghosts();
}

// siafd_test.cc:105-151
void computeSurfaceVelocityErrors(const IceGrid &grid, const IceModelVec2S &ice_thickness, const IceModelVec3 &u3,
                                  const IceModelVec3 &v3, const IceModelVec3 &w3, double &gmaxUerr, double &gavUerr,
                                  double &gmaxWerr, double &gavWerr) {
  double maxUerr = 0.0, avUerr = 0.0, maxWerr = 0.0, avWerr = 0.0;
  const double LforFG = 750000;
  for (int j = grid.ys(); j < grid.ys() + grid.ym(); ++j) {
    for (int i = grid.xs(); i < grid.xs() + grid.xm(); ++i) {
      const double xx = grid.x(i), yy = grid.y(j), r = sqrt(xx * xx + yy * yy);
      if ((r >= 1.0) && (r <= LforFG - 1.0)) {
        const double H = ice_thickness(i, j);
        FG F = exactFG(0.0, r, std::vector<double>(1, H), 0.0);
        const double uex = (xx / r) * F.U[0], vex = (yy / r) * F.U[0];
        const double du = u3.getValZ(i, j, H) - uex, dv = v3.getValZ(i, j, H) - vex;
        const double Uerr = sqrt(du * du + dv * dv);
        maxUerr = std::max(maxUerr, Uerr);
        avUerr += Uerr;
        const double Werr = fabs(w3.getValZ(i, j, H) - F.w[0]);
        maxWerr = std::max(maxWerr, Werr);
        avWerr += Werr;
      }
    }
  }
  gmaxUerr = maxUerr;
  gavUerr = avUerr / (grid.Mx() * grid.My());
  gmaxWerr = maxWerr;
  gavWerr = avWerr / (grid.Mx() * grid.My());
}

} // namespace

int main(int argc, char *argv[]) {
  using namespace pism::stressbalance;
  int Mx = 61, My = 61, Mz = 61;
  for (int a = 1; a + 1 < argc; a += 2) {
    if (!strcmp(argv[a], "-Mx")) Mx = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-My")) My = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-Mz")) Mz = atoi(argv[a + 1]);
  }
  try {
    Config::Ptr config(new Config());
    config->set_flag("stress_balance.sia.grain_size_age_coupling", false); // siafd_test.cc:283-284
    config->set_string("stress_balance.sia.flow_law", "arr");
    config->set_flag("enthalpy_converter.cold_mode", true);               // ColdEnthalpyConverter, :317
    config->set_number("stress_balance.sia.bed_smoother.range", 0.0);     // flat bed: nothing to smooth

    const double Lz = 4000.0;
    IceGrid::Ptr grid(new IceGrid(config, Mx, My, 900e3, 900e3, IceGrid::compute_vertical_levels(Lz, Mz, EQUAL)));

    const int WIDE_STENCIL = (int)config->get_number("grid.max_stencil_width");
    IceModelVec3 enthalpy(grid, "enthalpy", WITH_GHOSTS, WIDE_STENCIL);
    Geometry geometry(grid);
    geometry.sea_level_elevation.set(0.0);

    // stress_balance de-allocates sia (siafd_test.cc:340-346: StressBalance(grid, ZeroSliding, SIAFD))
    SIAFD_B200 *sia_p = new SIAFD_B200(grid);
    StressBalance_B200 stress_balance(grid, sia_p);
    SIAFD_B200 &sia = *sia_p;
    IceModelVec2V no_sliding(grid, "velbar", WITH_GHOSTS, 1); // ZeroSliding: identically zero
    no_sliding.set(0.0);

    setInitStateF(*grid, geometry.bed_elevation, geometry.cell_type, geometry.ice_surface_elevation,
                  geometry.ice_thickness, enthalpy);
    // Geometry::ensure_consistency (Geometry.cc:121-187) with bed = 0, sea level 0, all cells grounded: ice-free
    // cells become ice-free bedrock, the surface is bed + thickness
    const double H_min = config->get_number("geometry.ice_free_thickness_standard");
    for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
      for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i)
        geometry.cell_type(i, j) = geometry.ice_thickness(i, j) > H_min ? MASK_GROUNDED : MASK_ICE_FREE_BEDROCK;
    geometry.cell_type.update_ghosts();

    stress_balance.init();
    Inputs inputs;
    inputs.geometry = &geometry;
    inputs.enthalpy = &enthalpy;
    inputs.age = NULL;

    stress_balance.update(inputs, true);

    double maxUerr, avUerr, maxWerr, avWerr;
    computeSurfaceVelocityErrors(*grid, geometry.ice_thickness, stress_balance.velocity_u(), stress_balance.velocity_v(),
                                 stress_balance.velocity_w(), maxUerr, avUerr, maxWerr, avWerr);
    const double secpera = 365.242198781 * 86400.0;
    printf("surf vels :     maxUvec      avUvec        maxW         avW\n");
    printf("           %12.6f%12.6f%12.6f%12.6f\n", maxUerr * secpera, avUerr * secpera, maxWerr * secpera, avWerr * secpera);
    double sumD = 0.0, sumQ = 0.0, sumU = 0.0;
    for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
      for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) {
        sumD += sia.diffusivity()(i, j, 0) + sia.diffusivity()(i, j, 1);
        sumQ += fabs(sia.diffusive_flux()(i, j, 0)) + fabs(sia.diffusive_flux()(i, j, 1));
        sumU += fabs(sia.velocity_u().get_column(i, j)[Mz / 2]);
      }
    printf("D_max %.17g\nsum_D %.17g\nsum_absQ %.17g\nsum_absU_mid %.17g\n", sia.max_diffusivity(), sumD, sumQ, sumU);

    // full_update = false must leave u, v untouched (SIAFD.cc:149-154) and reproduce D, Q
    const double u_probe = sia.velocity_u().get_column(Mx / 3, My / 2)[Mz / 3];
    sia.update(no_sliding, inputs, false);
    printf("flux_only_ok %d\n", (int)(sia.velocity_u().get_column(Mx / 3, My / 2)[Mz / 3] == u_probe));

    // error path: thickness above the top of the grid -> the reference's RuntimeError (IceGrid.cc:434-437)
    geometry.ice_thickness(Mx / 2, My / 2) = 2.0 * Lz;
    geometry.ice_surface_elevation(Mx / 2, My / 2) = 2.0 * Lz;
    try {
      sia.update(no_sliding, inputs, true);
      printf("error_path none\n");
    } catch (RuntimeError &e) {
      printf("error_path status %d: %s\n", e.status(), e.what());
    }
  } catch (RuntimeError &e) {
    fprintf(stderr, "PISM ERROR: %s\n", e.what());
    return 1;
  }
  return 0;
}
