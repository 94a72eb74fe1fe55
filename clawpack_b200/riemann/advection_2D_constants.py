"""Field indices of advection_2D (same names and values as clawpack.riemann.advection_2D_constants)."""
num_waves = 1
num_eqn = 1

q = 0
