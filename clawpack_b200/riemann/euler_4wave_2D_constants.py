"""Field indices of euler_4wave_2D (same names and values as clawpack.riemann.euler_4wave_2D_constants)."""
num_waves = 4
num_eqn = 4

density = 0
x_momentum = 1
y_momentum = 2
energy = 3
