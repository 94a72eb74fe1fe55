"""Field indices of acoustics_2D (same names and values as clawpack.riemann.acoustics_2D_constants)."""
num_waves = 2
num_eqn = 3

pressure = 0
x_velocity = 1
y_velocity = 2
