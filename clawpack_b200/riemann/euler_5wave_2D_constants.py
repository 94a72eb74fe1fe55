"""Field indices of euler_5wave_2D (same names and values as clawpack.riemann.euler_5wave_2D_constants;
num_waves follows riemann/static.py and the solver body, which produce 5 waves)."""
num_waves = 5
num_eqn = 5

density = 0
x_momentum = 1
y_momentum = 2
energy = 3
tracer = 4
