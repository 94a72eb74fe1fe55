"""Field indices of euler_3D (same names and values as clawpack.riemann.euler_3D_constants)."""
num_waves = 3
num_eqn = 5

density = 0
x_momentum = 1
y_momentum = 2
z_momentum = 3
energy = 4
