"""Field indices of shallow_roe_with_efix_2D (same names and values as clawpack.riemann.shallow_roe_with_efix_2D_constants)."""
num_waves = 3
num_eqn = 3

depth = 0
x_momentum = 1
y_momentum = 2
