"""Field indices of vc_advection_2D (same names and values as clawpack.riemann.vc_advection_2D_constants)."""
num_waves = 1
num_eqn = 1
num_aux = 2

q = 0

x_velocity = 0
y_velocity = 1
