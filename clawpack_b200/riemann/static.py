"""Solver metadata by name -- same three tables as riemann/riemann/static.py:1-166 (only the
multi-dimensional solvers this build accelerates are executable; the other names are listed so
that ``Solver.__init__`` resolves num_eqn / num_waves for them as the reference does)."""

num_eqn = {
    'advection_2D': 1, 'acoustics_2D': 3, 'acoustics_mapped_2D': 3, 'burgers_2D': 1, 'euler_mapgrid_2D': 4,
    'euler_5wave_2D': 5, 'euler_4wave_2D': 4, 'euler_hlle_2D': 4, 'kpp_2D': 1, 'psystem_2D': 3,
    'shallow_roe_with_efix_2D': 3, 'shallow_hlle_2D': 3, 'shallow_sphere_2D': 4, 'vc_advection_2D': 1,
    'vc_acoustics_2D': 3, 'vc_elasticity_2D': 5, 'vc_acoustics_3D': 4, 'euler_3D': 5,
    'advection_1D': 1, 'acoustics_1D': 2, 'burgers_1D': 1, 'shallow_roe_with_efix_1D': 2, 'shallow_hlle_1D': 2,
    'shallow_bathymetry_fwave_2D': 3, 'burgers_3D': 1, 'vc_advection_3D': 1,
    'traffic_1D': 1, 'advection_color_1D': 1, 'acoustics_variable_1D': 2, 'shallow_bathymetry_fwave_1D': 2,
    'euler_with_efix_1D': 3, 'euler_hlle_1D': 3, 'nonlinear_elasticity_fwave_1D': 2, 'cubic_1D': 1,
    'traffic_vc_1D': 1, 'traffic_vc_fwave_1D': 1, 'traffic_vc_tracer_1D': 2, 'vc_advection_vector_2D': 2,
}

num_waves = {
    'advection_2D': 1, 'acoustics_2D': 2, 'acoustics_mapped_2D': 2, 'burgers_2D': 1, 'euler_mapgrid_2D': 4,
    'euler_5wave_2D': 5, 'euler_4wave_2D': 4, 'euler_hlle_2D': 2, 'kpp_2D': 1, 'psystem_2D': 2,
    'shallow_roe_with_efix_2D': 3, 'shallow_hlle_2D': 2, 'shallow_sphere_2D': 3, 'vc_advection_2D': 1,
    'vc_acoustics_2D': 2, 'vc_elasticity_2D': 4, 'vc_acoustics_3D': 2, 'euler_3D': 3,
    'advection_1D': 1, 'acoustics_1D': 2, 'burgers_1D': 1, 'shallow_roe_with_efix_1D': 2, 'shallow_hlle_1D': 2,
    'shallow_bathymetry_fwave_2D': 3, 'burgers_3D': 1, 'vc_advection_3D': 1,
    'traffic_1D': 1, 'advection_color_1D': 1, 'acoustics_variable_1D': 2, 'shallow_bathymetry_fwave_1D': 2,
    'euler_with_efix_1D': 3, 'euler_hlle_1D': 2, 'nonlinear_elasticity_fwave_1D': 2, 'cubic_1D': 1,
    'traffic_vc_1D': 1, 'traffic_vc_fwave_1D': 1, 'traffic_vc_tracer_1D': 2, 'vc_advection_vector_2D': 1,
}

fwave = {name: False for name in num_eqn}
fwave.update({'psystem_2D': True, 'shallow_bathymetry_fwave_2D': True,
              'shallow_bathymetry_fwave_1D': True, 'nonlinear_elasticity_fwave_1D': True,
              'traffic_vc_fwave_1D': True})      # riemann/riemann/static.py:155,158
