"""Field indices of vc_acoustics_3D (same names and values as clawpack.riemann.vc_acoustics_3D_constants)."""
num_waves = 2
num_eqn = 4
num_aux = 2

pressure = 0
x_velocity = 1
y_velocity = 2
z_velocity = 3

impedance = 0
sound_speed = 1
