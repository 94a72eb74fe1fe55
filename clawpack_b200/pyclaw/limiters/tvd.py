"""TVD limiter ids understood by the classic kernels (philim.f90:16-35).

Only the ids are provided: the numpy limiter functions of the reference's tvd.py serve the
pure-Python 1-D kernel, which is outside this build's scope."""
minmod = 1
superbee = 2
vanleer = 3
MC = 4
# id 5 (Beam-Warming) is accepted by the kernels as in philim.f90:31
