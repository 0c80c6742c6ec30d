// pismv_test_C.cc -- `pismv -test C` (isothermal SIA with a time-dependent surface mass balance, BASELINE configs[0])
// through the C++ host classes: the time step of IceModel::step (src/icemodel/IceModel.cc:388-640) with
// StressBalance_B200 (SIAFD_B200 + vertical velocity + strain heating + CFL) and GeometryEvolution_B200 over the C ABI.
// Prints the reference's report (src/verification/iceCompModel.cc:661-667), which test/regression/test_15.sh diffs
// against its golden rows.  TEST CODE: the exact solution comes from the reference's own exactTestsABCD.c compiled
// into oracle/_ref/libpism_exact.so.
//
// Several ranks: one PROCESS per rank (`-rank R -size S -prefix P`, P a path prefix every rank can write to), each with
// its patch of PISM's decomposition; SIAFD_B200 forms the library's communicator, the host arrays' ghost updates
// (DMLocalToLocal over MPI in PISM) go through siafd_b200_comm_exchange, and the report's sums / maxima through
// siafd_b200_comm_allreduce.  `-poison_rank R` makes the thickness negative at one point owned by rank R before the
// first update: every rank must fail with the reference's message (ParallelSection, util/error_handling.cc:189-214).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../pism_b200/host/GeometryEvolution_B200.hh"

extern "C" int ref_exactC(double t, double r, double *H, double *M);

using namespace pism;

int main(int argc, char *argv[]) {
  using namespace pism::stressbalance;
  int Mx = 31, My = 31, Mz = 31;
  double run_length_years = 5000.0;
  int rank = 0, size = 1, poison_rank = -1, trace = 0;
  std::string prefix;
  for (int a = 1; a + 1 < argc; a += 2) {
    if (!strcmp(argv[a], "-rank")) rank = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-size")) size = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-prefix")) prefix = argv[a + 1];
    if (!strcmp(argv[a], "-poison_rank")) poison_rank = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-trace")) trace = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-Mx")) Mx = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-My")) My = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-Mz")) Mz = atoi(argv[a + 1]);
    if (!strcmp(argv[a], "-y")) run_length_years = atof(argv[a + 1]);
  }
  try {
    const double year = StressBalance_B200::seconds_per_year_udunits(); // calendar "none", pismv.cc:51
    Config::Ptr config(new Config());
    // IceCompModel::IceCompModel for test C, iceCompModel.cc:60-135
    config->set_number("stress_balance.sia.enhancement_factor", 1.0);
    config->set_number("stress_balance.sia.bed_smoother.range", 0.0);
    config->set_flag("geometry.update.use_basal_melt_rate", false);
    config->set_string("stress_balance.sia.flow_law", "isothermal_glen");
    config->set_number("flow_law.isothermal_Glen.ice_softness", 1.0e-16 / year);
    config->set_flag("ocean.always_grounded", true);
    config->set_flag("enthalpy_converter.cold_mode", true); // pismv.cc:61
    // pismv_grid_defaults + vertical_grid_from_options: 2000 km x 2000 km x 4000 m, QUADRATIC levels (pismv.cc:96-102, :155)
    IceGrid::Ptr grid(new IceGrid(config, Mx, My, 1000e3, 1000e3, IceGrid::compute_vertical_levels(4000.0, Mz, QUADRATIC), rank, size));
    const int WIDE_STENCIL = (int)config->get_number("grid.max_stencil_width");
    const double ice_density = config->get_number("constants.ice.density");
    const double ice_free_thickness = config->get_number("geometry.ice_free_thickness_standard");

    Geometry geometry(grid);
    IceModelVec3 enthalpy(grid, "enthalpy", WITH_GHOSTS, WIDE_STENCIL);
    enthalpy.set(config->get_number("constants.ice.specific_heat_capacity") * (263.15 - config->get_number("enthalpy_converter.T_reference")));
    IceModelVec2S mass_flux(grid, "climatic_mass_balance", WITHOUT_GHOSTS);
    geometry.bed_elevation.set(0.0);
    geometry.sea_level_elevation.set(0.0);

    double time = 0.0;
    const double run_end = run_length_years * year;
    // initTestABCDH, iceCompModel.cc:301-356
    for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
      for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) {
        double H, M;
        ref_exactC(time, radius(*grid, i, j), &H, &M);
        geometry.ice_thickness(i, j) = H;
      }
    int ndev = 1;
    if (const char *e = getenv("PISMV_NDEV")) ndev = std::max(1, atoi(e)); // (GPUs to spread the ranks over; 1: share one)
    StressBalance_B200 stress_balance(grid, new SIAFD_B200(grid, rank % ndev, size > 1 ? prefix.c_str() : NULL));
    stress_balance.init();
    siafd_b200_handle *handle = stress_balance.modifier()->handle();
    GeometryEvolution_B200 geometry_evolution(grid, handle);
    if (size > 1) {
      // IceModelVec::update_ghosts of the HOST arrays between processes: through the device and the communicator
      grid->set_ghost_exchanger([handle](const std::string &name, double *a, int width, unsigned int) {
        const int f = name == "thk" ? SIAFD_B200_F_THICKNESS : (name == "mask" ? SIAFD_B200_F_MASK : (name == "usurf" ? SIAFD_B200_F_SURFACE : -1));
        if (f < 0) throw RuntimeError::formatted(-1, "no ghost exchange for '%s' in this driver", name.c_str());
        int st = siafd_b200_upload(handle, f, a);
        if (st == SIAFD_B200_OK) st = siafd_b200_comm_exchange(handle, 1, &f, &width);
        if (st == SIAFD_B200_OK) st = siafd_b200_download(handle, f, a);
        if (st != SIAFD_B200_OK) throw RuntimeError::formatted(st, "%s", siafd_b200_last_error(handle));
      });
    }
    geometry.ice_thickness.update_ghosts();

    Inputs inputs;
    inputs.geometry = &geometry;
    inputs.enthalpy = &enthalpy;
    inputs.new_bed_elevation = false;

    geometry.ensure_consistency(ice_free_thickness); // IceModel::run, IceModel.cc:763
    if (poison_rank == rank) { // an owned point away from the patch edge: only this rank sees it
      geometry.ice_thickness(grid->xs() + grid->xm() / 2, grid->ys() + grid->ym() / 2) = -1.0;
    }
    const double max_dt = config->get_number("time_stepping.maximum_time_step") * year;
    int steps = 0;
    while (time < run_end) { // IceModel::run :790 / IceModel::step
      grid->set_current_time(time);
      stress_balance.update(inputs, true);
      // IceModel::max_timestep, icemodel/timestepping.cc:111-232 (hit_multiples = 0, skip off)
      std::vector<std::pair<double, std::string> > restrictions;
      restrictions.push_back(std::make_pair(stress_balance.max_timestep_cfl_3d().dt_max, "energy")); // EnergyModel.cc:314-324
      restrictions.push_back(std::make_pair(max_dt, "max"));
      if (run_end - time > 0.0) restrictions.push_back(std::make_pair(run_end - time, "end of the run"));
      restrictions.push_back(std::make_pair(stress_balance.max_timestep_cfl_2d().dt_max, "2D CFL"));
      const double D_max = stress_balance.max_diffusivity();
      if (D_max > 0.0) { // :52-68
        const double dx = grid->dx(), dy = grid->dy(), grid_factor = 1.0 / (dx * dx) + 1.0 / (dy * dy);
        restrictions.push_back(std::make_pair(config->get_number("time_stepping.adaptive_ratio") * 2.0 / (D_max * grid_factor), "diffusivity"));
      } else {
        restrictions.push_back(std::make_pair(max_dt, "max time step"));
      }
      std::sort(restrictions.begin(), restrictions.end());
      const double dt = restrictions[0].first;

      geometry_evolution.flow_step(geometry, dt, stress_balance.advective_velocity(), stress_balance.diffusive_flux(), NULL, NULL);
      geometry_evolution.apply_flux_divergence(geometry);
      geometry.ensure_consistency(ice_free_thickness);
      // surface::Verification::update_ABCDH at the start-of-step time, PSVerification.cc:165-224
      for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
        for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) {
          double H, M;
          ref_exactC(time, radius(*grid, i, j), &H, &M);
          mass_flux(i, j) = M * ice_density;
        }
      geometry_evolution.source_term_step(geometry, dt, NULL, mass_flux, NULL);
      geometry_evolution.apply_mass_fluxes(geometry);
      geometry.ensure_consistency(ice_free_thickness);
      if (trace) { // per step: dt, D_max and the ice volume (decomposition-independence debugging)
        double sH[1] = {0.0};
        for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
          for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) sH[0] += geometry.ice_thickness(i, j);
        if (size > 1) siafd_b200_comm_allreduce(handle, 2, 1, sH);
        if (rank == 0) {
          printf("trace step %d dt %.17g D_max %.17g cfl3 %.17g cfl2 %.17g sumH %.17g\n", steps, dt, D_max,
                 stress_balance.max_timestep_cfl_3d().dt_max, stress_balance.max_timestep_cfl_2d().dt_max, sH[0]);
        }
      }
      time += dt; // Time::step, Time.cc:206-215
      if (run_end > time && run_end - time < 1e-3) time = run_end;
      steps += 1;
    }

    // IceCompModel::computeGeometryErrors + reportErrors, iceCompModel.cc:442-583, :661-667
    const double a = grid->dx() * grid->dy() * 1e-3 * 1e-3, m = (2.0 * 3.0 + 2.0) / 3.0;
    double vol = 0, volexact = 0, Herr = 0, avHerr = 0, etaerr = 0, domeHexact = 0;
    for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j)
      for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) {
        double Hexact, M;
        ref_exactC(time, radius(*grid, i, j), &Hexact, &M);
        const double H = geometry.ice_thickness(i, j);
        if (H > 0) vol += a * H * 1e-3;
        if (Hexact > 0) volexact += a * Hexact * 1e-3;
        if (i == ((int)grid->Mx() - 1) / 2 and j == ((int)grid->My() - 1) / 2) domeHexact = Hexact;
        Herr = std::max(Herr, fabs(H - Hexact));
        etaerr = std::max(etaerr, fabs(pow(H, m) - pow(Hexact, m)));
        avHerr += fabs(H - Hexact);
      }
    if (size > 1) { // GlobalSum / GlobalMax, iceCompModel.cc:540-560
      double sums[3] = {vol, volexact, avHerr}, maxs[3] = {Herr, etaerr, domeHexact};
      int st = siafd_b200_comm_allreduce(handle, 2, 3, sums);
      if (st == SIAFD_B200_OK) st = siafd_b200_comm_allreduce(handle, 0, 3, maxs);
      if (st != SIAFD_B200_OK) throw RuntimeError::formatted(st, "%s", siafd_b200_last_error(handle));
      vol = sums[0], volexact = sums[1], avHerr = sums[2], Herr = maxs[0], etaerr = maxs[1], domeHexact = maxs[2];
    }
    if (rank != 0) {
      printf("rank %d done, steps %d\n", rank, steps);
      return 0;
    }
    printf("NUMERICAL ERRORS evaluated at final time (relative to exact solution):\n");
    printf("geometry  :    prcntVOL        maxH         avH   relmaxETA\n");
    printf("           %12.6f%12.6f%12.6f%12.6f\n", 100 * fabs(vol - volexact) / volexact, Herr, avHerr / (grid->Mx() * grid->My()),
           etaerr / pow(domeHexact, m));
    printf("NUM ERRORS DONE\n");
    printf("steps %d\n", steps);
  } catch (RuntimeError &e) {
    fprintf(stderr, "PISM ERROR (rank %d): %s\n", rank, e.what());
    return 1;
  }
  return 0;
}
