"""Field indices of vc_acoustics_2D (same names and values as clawpack.riemann.vc_acoustics_2D_constants)."""
num_waves = 2
num_eqn = 3
num_aux = 2

pressure = 0
x_momentum = 1
y_momentum = 2

density = 0
sound_speed = 1
