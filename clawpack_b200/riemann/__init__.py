"""clawpack.riemann-compatible namespace: solver objects by the reference's module names
(riemann/riemann/__init__.py:56-111) plus ``static`` and the ``*_constants`` field indices."""
import sys as _sys

from . import static
from ._make import make as _make

# name -> (rp_id (include/clawcuda.h), COMMON /cparam/ members in declaration order, num_dim)
_SOLVERS = {
    'advection_2D': (1, ('u', 'v'), 2),                       # riemann/src/rpn2_advection.f90
    'acoustics_2D': (2, ('rho', 'bulk', 'cc', 'zz'), 2),      # rpn2_acoustics.f90:52
    'euler_4wave_2D': (3, ('gamma',), 2),                     # rpn2_euler_4wave.f90:71
    'shallow_roe_with_efix_2D': (4, ('grav',), 2),            # rpn2_shallow_roe_with_efix.f90:59
    'vc_acoustics_3D': (5, (), 3),                            # rpn3_vc_acoustics.f90 (no common block)
    'euler_3D': (6, ('gamma',), 3),                           # rpn3_euler.f90 / rpt3_euler.f90 / rptt3_euler.f90
    'vc_acoustics_2D': (7, (), 2),                            # rpn2_vc_acoustics.f90 (coefficients in aux)
    'euler_5wave_2D': (9, ('gamma',), 2),                     # rpn2_euler_5wave.f90:64
    'vc_advection_2D': (8, (), 2),                            # rpn2_vc_advection.f90 (edge velocities in aux)
    'euler_hlle_2D': (11, ('gamma',), 2),                     # rpn2_euler_hlle.f90:32 (+ rpt2_dummy.f90)
    'shallow_hlle_2D': (12, ('grav',), 2),                    # rpn2_shallow_hlle.f90:31 (+ rpt2_dummy.f90)
    'acoustics_mapped_2D': (13, (), 2),                       # rpn2_acoustics_mapped.f90 (geometry + coefficients in aux)
    'shallow_bathymetry_fwave_2D': (14, ('grav', 'dry_tolerance', 'sea_level'), 2),   # rpn2_shallow_bathymetry_fwave.f90
    'psystem_2D': (10, (), 2),                                # rpn2_psystem.f90 (f-wave; coefficients + strain in aux)
    'burgers_2D': (16, (), 2),                                # rpn2_burgers.f90, rpt2_burgers.f90
    'vc_elasticity_2D': (19, (), 2),                          # rpn2_vc_elasticity.f90 (material in aux)
    'vc_advection_vector_2D': (20, (), 2),                    # examples/advection_reaction_2d/rpn2_vc_advection_vector.f90
    'kpp_2D': (18, (), 2),                                    # rpn2_kpp.f90 (+ rpt2_dummy.f90)
    'burgers_3D': (15, ('coeff',), 3),
    'vc_advection_3D': (17, (), 3),                           # rpn3_vc_advection.f90 (edge velocities in aux)                        # rpn3_burgers.f90:45: common /cparam/ coeff(3)
    # 1-D (claw_step1_host): riemann/src/rp1_*.f90
    'advection_1D': (21, ('u',), 1),
    'acoustics_1D': (22, ('rho', 'bulk', 'cc', 'zz'), 1),
    'burgers_1D': (23, (), 1),
    'shallow_roe_with_efix_1D': (24, ('grav',), 1),
    'shallow_hlle_1D': (25, ('grav',), 1),
    'traffic_1D': (26, ('umax',), 1),                         # rp1_traffic.f90:39
    'advection_color_1D': (27, (), 1),                        # rp1_advection_color.f90 (velocity in aux)
    'acoustics_variable_1D': (28, (), 1),                     # rp1_acoustics_variable.f90 (Z, c in aux)
    'shallow_bathymetry_fwave_1D': (29, ('grav', 'dry_tolerance', 'sea_level'), 1),   # rp1_shallow_bathymetry_fwave.f90:42
    'euler_with_efix_1D': (30, ('gamma',), 1),                # rp1_euler_with_efix.f90:66
    'euler_hlle_1D': (31, ('gamma',), 1),                     # rp1_euler_hlle.f90:47
    'nonlinear_elasticity_fwave_1D': (32, (), 1),             # rp1_nonlinear_elasticity_fwave.f90 (rho, coefA, coefB in aux)
    'cubic_1D': (33, (), 1),                                  # rp1_cubic.f90
    'traffic_vc_1D': (34, (), 1),                             # rp1_traffic_vc.f90 (speed limit in aux)
    'traffic_vc_fwave_1D': (35, (), 1),                       # rp1_traffic_vc_fwave.f90
    'traffic_vc_tracer_1D': (36, (), 1),                      # rp1_traffic_vc_tracer.f90
}

for _name, (_id, _cp, _nd) in _SOLVERS.items():
    _m = _make(_name, _id, _cp, _nd)
    globals()[_name] = _m
    _sys.modules[__name__ + '.' + _name] = _m

from . import acoustics_2D_constants, advection_2D_constants, euler_4wave_2D_constants  # noqa: E402,F401
from . import shallow_roe_with_efix_2D_constants, vc_acoustics_3D_constants             # noqa: E402,F401
from . import euler_3D_constants, vc_acoustics_2D_constants, vc_advection_2D_constants   # noqa: E402,F401
from . import euler_5wave_2D_constants                                                   # noqa: E402,F401
from .plugin import build_cuda_solver, compile_cuda_solver, load_cuda_solver      # noqa: E402,F401
