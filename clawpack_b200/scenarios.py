"""Problem definitions (initial data + solver settings) for the configurations named in
BASELINE.json / SURVEY.md section 8(d), restated from the reference examples:

  advection_2d      pyclaw/examples/advection_2d/advection_2d.py:21-80
  acoustics_2d      pyclaw/examples/acoustics_2d_homogeneous/acoustics_2d.py:31-99
  quadrants         pyclaw/examples/euler_2d/quadrants.py:73-115
  radial_dam_break  pyclaw/examples/shallow_2d/radial_dam_break.py:23-72
  shock_bubble      synthetic 4-equation Euler shock/bubble (state formulas from
                    pyclaw/examples/euler_2d/shock_bubble_interaction.py:57-59), point-sampled
  acoustics_3d      pyclaw/examples/acoustics_3d_variable/acoustics_3d_interface.py:25-135
  euler_3d_*        pyclaw/examples/euler_3d/shocktube.py, shock_bubble.py

Each builder returns a plain ``dict`` (no PyClaw objects), so the same definition feeds the
CUDA solver (via ``make_claw``), the oracle in tests, and bench.py.  Cell centres follow
``Dimension.centers`` (pyclaw/src/pyclaw/geometry.py:546-553): ``lower + (i+0.5)*delta``.
"""
import numpy as np

EXTRAP, PERIODIC, WALL = 1, 2, 3
MINMOD, SUPERBEE, VANLEER, MC = 1, 2, 3, 4


def centers(lower, upper, n):
    delta = (upper - lower) / float(n)
    return lower + (np.arange(n) + 0.5) * delta


def _grid(lower, upper, n, slab=None):
    """Cell-centre meshgrid; ``slab = (s0, s1)`` keeps only cells [s0, s1) of the LAST dimension (one rank's share of a
    slab-decomposed run: the centres are still those of the global grid, bit for bit)."""
    cs = [centers(lo, up, k) for lo, up, k in zip(lower, upper, n)]
    if slab is not None:
        cs[-1] = cs[-1][slab[0]:slab[1]]
    return np.meshgrid(*cs, indexing='ij')


def _base(rp, lower, upper, n, q, **kw):
    d = dict(rp=rp, lower=list(map(float, lower)), upper=list(map(float, upper)), num_cells=list(map(int, n)),
             q=np.asfortranarray(q), aux=None, problem_data={}, limiters=MINMOD, order=2,
             dimensional_split=True, transverse_waves=None, cfl_max=1.0, cfl_desired=0.9,
             dt_initial=0.1, bc_lower=None, bc_upper=None, aux_bc_lower=None, aux_bc_upper=None,
             tfinal=1.0, num_output_times=10, index_capa=-1, custom_aux_bc_lower=None, custom_aux_bc_upper=None,
             slab=None, device_callbacks=None, mapc2p=None, step_source=None, source_split=1, before_step=None, fwave=False, gauges=None, custom_bc_lower_full=None,
             compute_gauge_values=None)
    d.update(kw)
    return d


def advection_2d(mx=50, my=50, dimensional_split=True, transverse_waves=0):
    X, Y = _grid((0., 0.), (1., 1.), (mx, my))
    q = np.empty((1, mx, my), order='F')
    q[0] = 0.9 * (0.1 < X) * (X < 0.6) * (0.1 < Y) * (Y < 0.6) + 0.1
    no_trans = (not dimensional_split) and (not transverse_waves)
    return _base('advection_2D', (0., 0.), (1., 1.), (mx, my), q,
                 problem_data={'u': 0.5, 'v': 1.0}, limiters=VANLEER,
                 dimensional_split=bool(dimensional_split), transverse_waves=int(transverse_waves),
                 cfl_max=0.5 if no_trans else 1.0, cfl_desired=0.45 if no_trans else 0.9,
                 bc_lower=[PERIODIC, PERIODIC], bc_upper=[PERIODIC, PERIODIC], tfinal=2.0)


def acoustics_2d(mx=100, my=100, tfinal=0.12, num_output_times=10):
    X, Y = _grid((-1., -1.), (1., 1.), (mx, my))
    r = np.sqrt(X ** 2 + Y ** 2)
    width = 0.2
    q = np.zeros((3, mx, my), order='F')
    q[0] = (np.abs(r - 0.5) <= width) * (1. + np.cos(np.pi * (r - 0.5) / width))
    rho, bulk = 1.0, 4.0
    cc = np.sqrt(bulk / rho)
    zz = rho * cc
    delta = [2.0 / mx, 2.0 / my]
    return _base('acoustics_2D', (-1., -1.), (1., 1.), (mx, my), q,
                 problem_data={'rho': rho, 'bulk': bulk, 'zz': zz, 'cc': cc}, limiters=MC,
                 dimensional_split=True, cfl_max=0.5, cfl_desired=0.45,
                 dt_initial=np.min(delta) / cc * 0.45,
                 bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal,
                 num_output_times=num_output_times)


def acoustics_2d_variable(mx=50, my=50, tfinal=0.6, num_output_times=20, dimensional_split=False,
                          transverse_waves=1):
    """pyclaw/examples/acoustics_2d_variable/acoustics_2d_interface.py:22-101 (classic branch): two materials,
    aux = (density, sound speed), wall / extrap, unsplit with the default transverse_waves = trans_inc
    (classic/solver.py:446), MC."""
    X, Y = _grid((-1., -1.), (1., 1.), (mx, my))
    rho_left, rho_right, bulk_left, bulk_right = 4.0, 1.0, 4.0, 4.0
    c_left, c_right = np.sqrt(bulk_left / rho_left), np.sqrt(bulk_right / rho_right)
    aux = np.empty((2, mx, my), order='F')
    aux[0] = rho_left * (X < 0.) + rho_right * (X >= 0.)
    aux[1] = c_left * (X < 0.) + c_right * (X >= 0.)
    r = np.sqrt((X + 0.5) ** 2 + Y ** 2)
    width, rad = 0.1, 0.25
    q = np.zeros((3, mx, my), order='F')
    q[0] = (np.abs(r - rad) <= width) * (1. + np.cos(np.pi * (r - rad) / width))
    return _base('vc_acoustics_2D', (-1., -1.), (1., 1.), (mx, my), q, aux=aux, limiters=MC,
                 dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves,
                 bc_lower=[WALL, WALL], bc_upper=[EXTRAP, EXTRAP], aux_bc_lower=[WALL, WALL],
                 aux_bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal, num_output_times=num_output_times)


def _nodes_with_ghost(lower, upper, n, g=2):
    # Dimension.nodes / nodes_with_ghost (pyclaw/src/pyclaw/geometry.py:538-545, 591-597)
    delta = (upper - lower) / float(n)
    nodes = np.array([lower + i * delta for i in range(n + 1)])
    pre = np.linspace(lower - g * delta, lower - delta, g)
    post = np.linspace(upper + delta, upper + g * delta, g)
    return np.hstack((pre, nodes, post))


def advection_annulus(mr=40, mtheta=120, tfinal=1.0, num_output_times=10, dimensional_split=False,
                      transverse_waves=2, order=2):
    """pyclaw/examples/advection_2d_annulus/advection_annulus.py:22-205: colour equation on an annulus in polar
    computational coordinates (r, theta); aux = (u at left edges, v at bottom edges, capacity = cell area ratio)
    from the stream function pi*(x^2+y^2) (one rotation per unit time); index_capa = 2; extrap in r with
    custom aux ghost cells (edge velocities evaluated on the extended grid), periodic in theta; van Leer."""
    lower, upper = (0.2, 0.0), (1.0, np.pi * 2.0)
    dx, dy = (upper[0] - lower[0]) / mr, (upper[1] - lower[1]) / mtheta
    rn = _nodes_with_ghost(lower[0], upper[0], mr)
    tn = _nodes_with_ghost(lower[1], upper[1], mtheta)
    RN, TN = np.meshgrid(rn, tn, indexing='ij')          # computational nodes incl. ghost cells
    XN, YN = RN * np.cos(TN), RN * np.sin(TN)            # mapc2p_annulus (:22-38)

    def edge_data(X, Y):                                  # edge_velocities_and_area (:94-134)
        def stream(xp, yp):
            return np.pi * (xp ** 2 + yp ** 2)
        x0, y0 = X[:-1, :-1], Y[:-1, :-1]
        x1, y1 = X[:-1, 1:], Y[:-1, 1:]
        x2, y2 = X[1:, 1:], Y[1:, 1:]
        x3, y3 = X[1:, :-1], Y[1:, :-1]
        a = np.empty((3,) + x0.shape, order='F')
        a[0] = (stream(x1, y1) - stream(x0, y0)) / dy
        a[1] = -(stream(x3, y3) - stream(x0, y0)) / dx
        area = 1. / 2. * ((y0 + y1) * (x1 - x0) + (y1 + y2) * (x2 - x1) + (y2 + y3) * (x3 - x2) + (y3 + y0) * (x0 - x3))
        a[2] = area / (dx * dy)
        return a

    aux = edge_data(XN[2:-2, 2:-2], YN[2:-2, 2:-2])

    def aux_lower(auxbc, g):                              # ghost_velocities_lower (:78-91): all theta incl. ghosts
        auxbc[:, 0:g, :] = edge_data(XN[0:g + 1, :], YN[0:g + 1, :])

    def aux_upper(auxbc, g):                              # ghost_velocities_upper (:62-75)
        auxbc[:, -g:, :] = edge_data(XN[-g - 1:, :], YN[-g - 1:, :])

    RC, TC = _grid(lower, upper, (mr, mtheta))
    R, Theta = RC * np.cos(TC), RC * np.sin(TC)           # p_centers (qinit uses the PHYSICAL centres, :58-60)
    q = np.empty((1, mr, mtheta), order='F')
    q[0] = (1. * np.exp(-40. * (np.square(R + 0.5) + np.square(Theta - 0.)))
            + -1. * np.exp(-40. * (np.square(R - 0.5) + np.square(Theta - 0.))))
    return _base('vc_advection_2D', lower, upper, (mr, mtheta), q, aux=aux, limiters=VANLEER, order=order,
                 dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves,
                 cfl_max=0.5, cfl_desired=0.4, bc_lower=[EXTRAP, PERIODIC], bc_upper=[EXTRAP, PERIODIC],
                 aux_bc_lower=[0, PERIODIC], aux_bc_upper=[0, PERIODIC], custom_aux_bc_lower=aux_lower,
                 custom_aux_bc_upper=aux_upper, index_capa=2, tfinal=tfinal, num_output_times=num_output_times,
                 mapc2p=lambda xc, yc: [xc * np.cos(yc), xc * np.sin(yc)])


def quadrants_tracer(mx=60, my=50, **kw):
    """The four-quadrant Riemann problem of `quadrants` with a fifth, passively advected component, for
    euler_5wave_2D without host callbacks (device-resident path)."""
    sc = quadrants(mx=mx, my=my)
    X, Y = _grid((0., 0.), (1., 1.), (mx, my))
    q = np.empty((5, mx, my), order='F')
    q[:4] = sc['q']
    q[4] = 1.0 * ((X - 0.5) ** 2 + (Y - 0.5) ** 2 < 0.09)
    sc.update(rp='euler_5wave_2D', q=q, limiters=[MC, MC, MC, MC, SUPERBEE])
    sc.update(kw)
    return sc


def shock_bubble_interaction(q0, tfinal=0.2, num_output_times=1, dimensional_split=True, transverse_waves=None,
                             device_callbacks=False):
    """pyclaw/examples/euler_2d/shock_bubble_interaction.py:93-215 (classic branch): axisymmetric Euler + tracer
    (euler_5wave_2D), Mach-2.95 shock hitting a light bubble; aux(1) = radial coordinate of the cell centre;
    Godunov-split geometric source term advanced on the HOST by a 2-stage Runge-Kutta step; inflow (BC.custom) at
    x-lower, wall at the axis.  `q0` is the example's initial state (cells cut by the bubble edge carry sub-cell
    volume fractions from scipy.integrate.quad): tests/golden/reference_goldens2.npz
    ['euler_2d__shock_bubble_interaction_q0'], produced by running the reference's own qinit."""
    gamma = 1.4
    gamma1 = gamma - 1.
    mx, my = q0.shape[1:]
    lower, upper = (0.0, 0.0), (2.0, 0.5)
    pinf = 5.
    rinf = (gamma1 + pinf * (gamma + 1.)) / ((gamma + 1.) + gamma1 * pinf)
    vinf = 1. / np.sqrt(gamma) * (pinf - 1.) / np.sqrt(0.5 * ((gamma + 1.) / gamma) * pinf + 0.5 * gamma1 / gamma)
    einf = 0.5 * rinf * vinf ** 2 + pinf / gamma1
    aux = np.empty((1, mx, my), order='F')
    aux[0, :, :] = centers(lower[1], upper[1], my)[None, :]

    def inflow(qbc, g):                                  # incoming_shock (:93-109)
        for i in range(g):
            qbc[0, i, ...] = rinf
            qbc[1, i, ...] = rinf * vinf
            qbc[2, i, ...] = 0.
            qbc[3, i, ...] = einf
            qbc[4, i, ...] = 0.

    def radial_source(q, aux_, dt):                      # step_Euler_radial (:112-143): RK2 on the geometric terms
        dt2 = dt / 2.
        rad = aux_[0, :, :]
        rho = q[0, :, :]
        u = q[1, :, :] / rho
        v = q[2, :, :] / rho
        press = (gamma - 1.) * (q[3, :, :] - 0.5 * rho * (u ** 2 + v ** 2))
        qs = np.empty(q.shape)
        qs[0, :, :] = q[0, :, :] - dt2 / rad * q[2, :, :]
        qs[1, :, :] = q[1, :, :] - dt2 / rad * rho * u * v
        qs[2, :, :] = q[2, :, :] - dt2 / rad * rho * v * v
        qs[3, :, :] = q[3, :, :] - dt2 / rad * v * (q[3, :, :] + press)
        rho = qs[0, :, :]
        u = qs[1, :, :] / rho
        v = qs[2, :, :] / rho
        press = (gamma - 1.) * (qs[3, :, :] - 0.5 * rho * (u ** 2 + v ** 2))
        q[0, :, :] = q[0, :, :] - dt / rad * qs[2, :, :]
        q[1, :, :] = q[1, :, :] - dt / rad * rho * u * v
        q[2, :, :] = q[2, :, :] - dt / rad * rho * v * v
        q[3, :, :] = q[3, :, :] - dt / rad * v * (qs[3, :, :] + press)

    dc = None
    if device_callbacks:                                 # source term and inflow BC as device functors: no host round trip
        from .pyclaw import device_callbacks as D
        dc = {'step_source': D.EulerRadialSource(gamma), 'user_bc_lower': D.FixedStateBC([rinf, rinf * vinf, 0., einf, 0.])}
    return _base('euler_5wave_2D', lower, upper, (mx, my), q0, aux=aux, problem_data={'gamma': gamma},
                 device_callbacks=dc,
                 limiters=[MC, MC, MC, MC, SUPERBEE], dimensional_split=bool(dimensional_split),
                 transverse_waves=transverse_waves, cfl_max=0.5, cfl_desired=0.45,
                 bc_lower=[0, WALL], bc_upper=[EXTRAP, EXTRAP], aux_bc_lower=[EXTRAP, EXTRAP],
                 aux_bc_upper=[EXTRAP, EXTRAP], custom_bc_lower=inflow, step_source=radial_source, source_split=1,
                 tfinal=tfinal, num_output_times=num_output_times)


def acoustics_2d_mapped(num_cells=100, tfinal=0.9, num_output_times=1, transverse_waves=None):
    """pyclaw/examples/acoustics_2d_mapped/acoustics_2d_inclusions.py:150-235 (classic branch): acoustics on a mapped
    grid with two circular inclusions (impedance 12 / 10, sound speed 0.3 / 1.5), solver acoustics_mapped_2D, unsplit,
    MC; aux = (left-edge normal x, y, length ratio, bottom-edge normal x, y, length ratio, cell area ratio [capacity],
    impedance, sound speed); an incoming square wave enters through a BC.custom fill at x-lower that reads auxbc."""
    circles = (((0.15, 0.204), (0.45, 0.3)), ((0.15, 0.204), (0.7, 0.75)))
    impedance, sound_speed = (12., 10.), (0.3, 1.5)

    def inclusion_mapping(xc, yc):                          # :30-82: square -> circle inside every inclusion box
        xp, yp = xc + 0., yc + 0.
        for (r1, r2), (x0, y0) in circles:
            xdist, ydist = np.abs(xc - x0), np.abs(yc - y0)
            insquare = np.where(np.maximum(xdist, ydist) <= r2)
            xc0, yc0 = (xc[insquare] - x0) / r2, (yc[insquare] - y0) / r2
            xc1, yc1 = np.abs(xc0), np.abs(yc0)
            d = np.maximum(xc1, yc1)
            d = np.maximum(d, 1.e-10)
            d = np.minimum(d, 0.99999)
            d1 = d * r2 / np.sqrt(2)
            R = r1 ** 2 / (r2 * d)
            ij = np.where(d > (r1 / r2))                    # morph back to the square outside the circle
            d1[ij] = r1 / np.sqrt(2) + (d[ij] - r1 / r2) * (r2 - r1 / np.sqrt(2)) / (1. - r1 / r2)
            R[ij] = r1 * ((1. - r1 / r2) / (1. - d[ij])) ** (r2 / r1 + 0.5)
            xp2, yp2 = d1 / d * xc1, d1 / d * yc1
            center = d1 - np.sqrt(R ** 2 - d1 ** 2)
            ij = np.where(xc1 >= yc1)
            xp2[ij] = center[ij] + np.sqrt(R[ij] ** 2 - yp2[ij] ** 2)
            ij = np.where(yc1 >= xc1)
            yp2[ij] = center[ij] + np.sqrt(R[ij] ** 2 - xp2[ij] ** 2)
            xp[insquare] = x0 + np.sign(xc0) * xp2
            yp[insquare] = y0 + np.sign(yc0) * yp2
        return xp, yp

    n = num_cells
    lower, upper = (0., 0.), (1., 1.)
    dx = dy = 1.0 / n
    xn = lower[0] + np.arange(n + 1) * dx                   # Dimension.nodes
    index = np.indices((n + 1, n + 1))
    xcor, ycor = inclusion_mapping(xn[index[0]], xn[index[1]])          # grid.p_nodes
    llx, lly = xcor[:-1, :-1], ycor[:-1, :-1]               # compute_geometry (:84-127)
    ulx, uly = xcor[:-1, 1:], ycor[:-1, 1:]
    lrx, lry = xcor[1:, :-1], ycor[1:, :-1]
    urx, ury = xcor[1:, 1:], ycor[1:, 1:]
    a_x, a_y = uly - lly, -(ulx - llx)
    anorm = np.sqrt(a_x ** 2 + a_y ** 2)
    a_x, a_y = a_x / anorm, a_y / anorm
    b_x, b_y = -(lry - lly), lrx - llx
    bnorm = np.sqrt(b_x ** 2 + b_y ** 2)
    b_x, b_y = b_x / bnorm, b_y / bnorm
    area = 0 * a_x
    area += 0.5 * (lly + uly) * (ulx - llx)
    area += 0.5 * (uly + ury) * (urx - ulx)
    area += 0.5 * (ury + lry) * (lrx - urx)
    area += 0.5 * (lry + lly) * (llx - lrx)
    area = area / (dx * dy)
    aux = np.empty((9, n, n), order='F')
    aux[0], aux[1], aux[2] = a_x, a_y, anorm / dy
    aux[3], aux[4], aux[5] = b_x, b_y, bnorm / dx
    aux[6] = area
    aux[7] = 1.0
    aux[8] = 1.0
    X, Y = _grid(lower, upper, (n, n))
    xp, yp = inclusion_mapping(X, Y)                        # grid.p_centers
    for i, ((r1, r2), (x0, y0)) in enumerate(circles):
        in_circle = np.where(np.sqrt((xp - x0) ** 2 + (yp - y0) ** 2) <= r1)
        aux[7][in_circle] = impedance[i]
        aux[8][in_circle] = sound_speed[i]
    q = np.zeros((3, n, n), order='F')

    def incoming_square_wave(qbc, auxbc, t, g):             # :129-147
        s_ = 1.0 if t < 0.05 else 0.0
        for i in range(g):
            ri = 2 * g - i - 1
            alpha, beta = auxbc[0, i, :], auxbc[1, i, :]
            u_normal = alpha * qbc[1, ri, :] + beta * qbc[2, ri, :]
            u_tangential = -beta * qbc[1, ri, :] + alpha * qbc[2, ri, :]
            u_normal = 2.0 * s_ - u_normal
            qbc[0, i, :] = qbc[0, ri, :]
            qbc[1, i, :] = alpha * u_normal - beta * u_tangential
            qbc[2, i, :] = beta * u_normal + alpha * u_tangential

    return _base('acoustics_mapped_2D', lower, upper, (n, n), q, aux=aux, limiters=MC, dimensional_split=False,
                 transverse_waves=transverse_waves, bc_lower=[0, EXTRAP], bc_upper=[EXTRAP, EXTRAP],
                 aux_bc_lower=[EXTRAP, EXTRAP], aux_bc_upper=[EXTRAP, EXTRAP], custom_bc_lower_full=incoming_square_wave,
                 index_capa=6, tfinal=tfinal, num_output_times=num_output_times, mapc2p=inclusion_mapping)


def sill(my=20, tfinal=10.0, num_output_times=40):
    """pyclaw/examples/shallow_2d/sill.py:27-86: shallow water over a Gaussian hump, f-wave solver
    shallow_bathymetry_fwave_2D (bathymetry in aux), dimensionally split, minmod; a wave maker (BC.custom) at x-lower
    sets only part of the ghost cells -- the rest keep their previous contents, as in the reference."""
    mx = 4 * my
    lower, upper = (0., 0.), (4., 1.)
    X, Y = _grid(lower, upper, (mx, my))
    aux = np.empty((1, mx, my), order='F')
    aux[0] = 0.8 * np.exp(-10 * ((X - 1.) ** 2 + (Y - 0.5) ** 2))
    q = np.zeros((3, mx, my), order='F')
    q[0] = 1. - aux[0]
    amplitude, t_bdy = 0.1, 5.

    def wave_maker_bc(qbc, auxbc, t, g):                    # :31-42
        qbc[0, :g, :] = qbc[0, g, :]
        vwall = amplitude * (np.sin(t * np.pi / 1.5)) if t <= t_bdy else 0.
        for ibc in range(g - 1):
            qbc[1, g - ibc - 1, :] = 2 * vwall - qbc[1, g + ibc, :]

    return _base('shallow_bathymetry_fwave_2D', lower, upper, (mx, my), q, aux=aux,
                 problem_data={'grav': 1.0, 'dry_tolerance': 1.e-3, 'sea_level': 0.}, limiters=MINMOD,
                 dimensional_split=True, bc_lower=[0, WALL], bc_upper=[EXTRAP, WALL], aux_bc_lower=[EXTRAP, EXTRAP],
                 aux_bc_upper=[EXTRAP, EXTRAP], custom_bc_lower_full=wave_maker_bc, fwave=True, tfinal=tfinal,
                 num_output_times=num_output_times)


def psystem_2d(cells_per_layer=6, tfinal=1.0, num_output_times=10, stress_rel=2, dimensional_split=False,
               transverse_waves=None, size=20.0):
    """pyclaw/examples/psystem_2d/psystem_2d.py:126-226 (classic branch): nonlinear elasticity in a checkerboard
    medium, f-wave solver psystem_2D, superbee, unsplit; aux = (density, modulus, stress-strain relation id, strain);
    `before_step` copies q[0] into aux[3] for the transverse solver (:90-95); four gauges record the stress
    (:121-123, :221-223).  `stress_rel`: 1 linear, 2 exponential (the example), 3 cubic."""
    KA, KB, rhoA, rhoB = 1., 4., 1., 4.
    lower, upper = (0.25, 0.25), (0.25 + size, 0.25 + size)
    mx = int((upper[0] - lower[0]) * cells_per_layer)
    my = int((upper[1] - lower[1]) * cells_per_layer)
    x, y = centers(lower[0], upper[0], mx), centers(lower[1], upper[1], my)
    alphax = alphay = 0.5
    deltax = deltay = 1.
    xfrac = x - np.floor(x / deltax) * deltax                  # setaux, checkerboard medium (:38-66)
    yfrac = y - np.floor(y / deltay) * deltay
    yf, xf = np.meshgrid(yfrac, xfrac)
    aux = np.empty((4, mx, my), order='F')
    a_ = (xf <= alphax * deltax) * (yf <= alphay * deltay)
    b_ = (xf > alphax * deltax) * (yf > alphay * deltay)
    c_ = (xf > alphax * deltax) * (yf <= alphay * deltay)
    d_ = (xf <= alphax * deltax) * (yf > alphay * deltay)
    aux[0] = rhoA * a_ + rhoA * b_ + rhoB * c_ + rhoB * d_
    aux[1] = KA * a_ + KA * b_ + KB * c_ + KB * d_
    aux[2] = stress_rel
    aux[3] = 0.0
    X, Y = _grid(lower, upper, (mx, my))
    A, x0, y0, varx, vary = 10., 0.25, 0.25, 0.5, 0.5          # qinit (:27-36)
    stress = A * np.exp(-(X - x0) ** 2 / (2 * varx) - (Y - y0) ** 2 / (2 * vary))
    q = np.zeros((3, mx, my), order='F')
    K = aux[1]
    if stress_rel == 1:
        q[0] = stress / K
    elif stress_rel == 2:
        q[0] = np.log(stress + 1) / K
    else:                                                      # cubic law has no closed-form inverse: small-strain data
        q[0] = 0.05 * stress / K

    def b4step(q_, aux_):                                      # :90-95
        aux_[3, :, :] = q_[0, :, :]

    def gauge_stress(q_, aux_):                                # :121-123
        p = np.exp(q_[0] * aux_[1]) - 1
        return [p, 10 * p]

    return _base('psystem_2D', lower, upper, (mx, my), q, aux=aux, limiters=SUPERBEE,
                 dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves,
                 cfl_max=0.9, cfl_desired=0.8, bc_lower=[WALL, WALL], bc_upper=[EXTRAP, EXTRAP],
                 aux_bc_lower=[WALL, WALL], aux_bc_upper=[EXTRAP, EXTRAP], fwave=True, before_step=b4step,
                 tfinal=tfinal, num_output_times=num_output_times,
                 gauges=[[0.25, 0.25], [17.85, 1.25], [3.25, 18.75], [11.75, 11.75]], compute_gauge_values=gauge_stress)


def quadrants(mx=100, my=100, tfinal=0.8, riemann_solver='roe'):
    xx, yy = _grid((0., 0.), (1., 1.), (mx, my))
    gamma = 1.4
    l = xx < 0.8
    r = xx >= 0.8
    b = yy < 0.8
    t = yy >= 0.8
    q = np.empty((4, mx, my), order='F')
    q[0] = 1.5 * r * t + 0.532258064516129 * l * t + 0.137992831541219 * l * b + 0.532258064516129 * r * b
    u = 0.0 * r * t + 1.206045378311055 * l * t + 1.206045378311055 * l * b + 0.0 * r * b
    v = 0.0 * r * t + 0.0 * l * t + 1.206045378311055 * l * b + 1.206045378311055 * r * b
    p = 1.5 * r * t + 0.3 * l * t + 0.029032258064516 * l * b + 0.3 * r * b
    q[1] = q[0] * u
    q[2] = q[0] * v
    q[3] = 0.5 * q[0] * (u ** 2 + v ** 2) + p / (gamma - 1.0)
    if riemann_solver == 'hlle':                      # quadrants.py:82-86
        return _base('euler_hlle_2D', (0., 0.), (1., 1.), (mx, my), q, problem_data={'gamma': gamma},
                     limiters=MINMOD, dimensional_split=True, transverse_waves=0, cfl_desired=0.4, cfl_max=0.5,
                     bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal)
    # quadrants.py sets transverse_waves=2 but leaves dimensional_split at its default True
    return _base('euler_4wave_2D', (0., 0.), (1., 1.), (mx, my), q, problem_data={'gamma': gamma},
                 limiters=MINMOD, dimensional_split=True, transverse_waves=2,
                 bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal)


def radial_dam_break(mx=150, my=150, tfinal=2.5, riemann_solver='roe', slab=None):
    X, Y = _grid((-2.5, -2.5), (2.5, 2.5), (mx, my), slab)
    r = np.sqrt((X - 0.) ** 2 + (Y - 0.) ** 2)
    q = np.zeros((3,) + X.shape, order='F')
    q[0] = 2. * (r <= 0.5) + 1. * (r > 0.5)
    rp = 'shallow_hlle_2D' if riemann_solver == 'hlle' else 'shallow_roe_with_efix_2D'     # radial_dam_break.py:41-44
    return _base(rp, (-2.5, -2.5), (2.5, 2.5), (mx, my), q,
                 problem_data={'grav': 1.0}, limiters=MC, dimensional_split=True,
                 bc_lower=[EXTRAP, EXTRAP], bc_upper=[WALL, WALL], tfinal=tfinal, slab=slab)


def shock_bubble(mx=8192, my=8192, ly=1.0, y_offset=0.0, ly_total=None, dimensional_split=False,
                 transverse_waves=2, tfinal=0.02, slab=None):
    """BASELINE configs 2/5: 4-equation Euler, planar Mach shock at x=0.2 hitting a low-density
    bubble; unsplit with transverse corrections, MC limiter, extrapolation BCs, dt_initial=0.1 so
    the first step is a deliberate reject.  ``ly``/``y_offset``/``ly_total`` describe a y-slab of a
    taller weak-scaling domain (bubble centred in the full domain)."""
    gamma = 1.4
    gamma1 = gamma - 1.
    if ly_total is None:
        ly_total = ly
    x = centers(0., 1., mx)
    y = centers(y_offset, y_offset + ly, my)
    if slab is not None:                       # one rank's rows of the global grid (same centres, bit for bit)
        y = y[slab[0]:slab[1]]
    X, Y = np.meshgrid(x, y, indexing='ij')
    pinf = 5.
    rinf = (gamma1 + pinf * (gamma + 1.)) / ((gamma + 1.) + gamma1 * pinf)
    vinf = 1. / np.sqrt(gamma) * (pinf - 1.) / np.sqrt(0.5 * ((gamma + 1.) / gamma) * pinf + 0.5 * gamma1 / gamma)
    einf = 0.5 * rinf * vinf ** 2 + pinf / gamma1
    rr = np.sqrt((X - 0.5) ** 2 + (Y - 0.5 * ly_total) ** 2)
    shocked = X < 0.2
    bubble = rr <= 0.2
    rho = np.where(shocked, rinf, np.where(bubble, 0.1, 1.0))
    q = np.empty((4,) + X.shape, order='F')
    q[0] = rho
    q[1] = np.where(shocked, rinf * vinf, 0.0)
    q[2] = 0.0
    q[3] = np.where(shocked, einf, 1.0 / gamma1)
    return _base('euler_4wave_2D', (0., y_offset), (1., y_offset + ly), (mx, my), q,
                 problem_data={'gamma': gamma}, limiters=MC, dimensional_split=bool(dimensional_split),
                 transverse_waves=int(transverse_waves), cfl_max=1.0, cfl_desired=0.9, dt_initial=0.1,
                 bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal, slab=slab)


def acoustics_3d(mx=30, my=30, mz=30, heterogeneous=True, dimensional_split=None, transverse_waves=22,
                 periodic=False, tfinal=2.0, num_output_times=10, slab=None):
    """pyclaw/examples/acoustics_3d_variable/acoustics_3d_interface.py:25-135: problem='heterogeneous'
    (impedance/sound speed 1 for x<0, 2 for x>=0; unsplit, lower wall / upper periodic) or
    'homogeneous' (dimension-split, all periodic).  ``periodic=True`` makes every face periodic
    (sharded runs).  mx=my=mz=15 with the defaults reproduces the stored vector
    verify_classic_heterogeneous.txt."""
    X, Y, Z = _grid((-1., -1., -1.), (1., 1., 1.), (mx, my, mz), slab)
    zl, cl = 1.0, 1.0
    zr, cr = (2.0, 2.0) if heterogeneous else (1.0, 1.0)
    aux = np.empty((2,) + X.shape, order='F')
    aux[0] = zl * (X < 0.) + zr * (X >= 0.)
    aux[1] = cl * (X < 0.) + cr * (X >= 0.)
    x0, y0, z0 = -0.5, 0., 0.
    q = np.zeros((4,) + X.shape, order='F')
    if heterogeneous:
        r = np.sqrt((X - x0) ** 2 + (Y - y0) ** 2 + (Z - z0) ** 2)
        width = 0.1
        q[0] = (np.abs(r - 0.3) <= width) * (1. + np.cos(np.pi * (r - 0.3) / width))
    else:
        r = np.sqrt((X - x0) ** 2)
        width = 0.2
        q[0] = (np.abs(r) <= width) * (1. + np.cos(np.pi * (r) / width))
    if dimensional_split is None:
        dimensional_split = not heterogeneous
    lo = [WALL] * 3 if (heterogeneous and not periodic) else [PERIODIC] * 3
    up = [PERIODIC] * 3
    return _base('vc_acoustics_3D', (-1., -1., -1.), (1., 1., 1.), (mx, my, mz), q, aux=np.asfortranarray(aux),
                 limiters=MC, dimensional_split=bool(dimensional_split), transverse_waves=int(transverse_waves),
                 cfl_max=1.0, cfl_desired=0.9, dt_initial=0.1,
                 bc_lower=lo, bc_upper=up, aux_bc_lower=lo, aux_bc_upper=up,
                 tfinal=tfinal, num_output_times=num_output_times, slab=slab)


def acoustics_3d_general(mx=18, my=15, mz=13, **kw):
    """vc_acoustics_3D stress variant: impedance / sound speed varying smoothly in all three directions and a
    state with non-zero pressure and velocities everywhere (every aux cell of the 3x3 neighbourhood and every
    limiter branch matters); all faces periodic."""
    sc = acoustics_3d(mx=mx, my=my, mz=mz, periodic=True)
    X, Y, Z = _grid((-1., -1., -1.), (1., 1., 1.), (mx, my, mz))
    aux = np.empty((2, mx, my, mz), order='F')
    aux[0] = 1.5 + 0.5 * np.sin(3 * X + 1) * np.cos(2 * Y) + 0.3 * np.sin(4 * Z + X * Y)
    aux[1] = 1.2 + 0.4 * np.cos(2 * X - Y + 0.5 * Z) + 0.2 * np.sin(5 * Y * Z)
    q = np.empty((4, mx, my, mz), order='F')
    q[0] = np.exp(-3 * ((X + 0.2) ** 2 + Y ** 2 + (Z - 0.1) ** 2)) + 0.1 * np.sin(7 * X * Y * Z)
    q[1] = 0.3 * np.sin(2 * X + Y)
    q[2] = -0.2 * np.cos(3 * Y - Z)
    q[3] = 0.25 * np.sin(Z + 2 * X)
    sc = dict(sc, q=q, aux=aux)
    sc.update(kw)
    return sc


def acoustics_3d_capa(mx=18, my=15, mz=13, **kw):
    """vc_acoustics_3D on a grid with a capacity function (step3.f90:132-136, 229-273; no reference example uses one in
    3-D): the stress case above with a third aux plane kappa(x, y, z) in [0.6, 1.5] and ``index_capa = 2``."""
    sc = acoustics_3d_general(mx=mx, my=my, mz=mz)
    X, Y, Z = _grid((-1., -1., -1.), (1., 1., 1.), (mx, my, mz))
    aux = np.empty((3, mx, my, mz), order='F')
    aux[:2] = sc['aux']
    aux[2] = 1.05 + 0.3 * np.sin(2 * X - Y) * np.cos(3 * Z + 0.5) + 0.15 * np.cos(4 * X * Y + Z)
    sc = dict(sc, aux=aux, index_capa=2)
    sc.update(kw)
    return sc


def burgers_2d(mx=60, my=50, dimensional_split=False, transverse_waves=2, tfinal=0.3, num_output_times=2):
    """2-D Burgers equation (riemann burgers_2D; no reference example): data of both signs, periodic."""
    X, Y = _grid((0., 0.), (1., 1.), (mx, my))
    q = np.empty((1, mx, my), order='F')
    q[0] = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.3
    return _base('burgers_2D', (0., 0.), (1., 1.), (mx, my), q, limiters=VANLEER,
                 dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves, dt_initial=0.01,
                 bc_lower=[PERIODIC, PERIODIC], bc_upper=[PERIODIC, PERIODIC], tfinal=tfinal,
                 num_output_times=num_output_times)


def advection_3d_variable(mx=20, my=16, mz=18, dimensional_split=False, transverse_waves=22, tfinal=0.3,
                          num_output_times=2):
    """Colour-equation advection in a divergence-free 3-D velocity field given by edge velocities (riemann
    vc_advection_3D; the reference ships the solver but no example): u = sin(2 pi y) + 0.3 at x-faces,
    v = cos(2 pi z) - 0.2 at y-faces, w = sin(2 pi x) at z-faces (each independent of its own direction); periodic."""
    lower, upper, n = (0., 0., 0.), (1., 1., 1.), (mx, my, mz)
    X, Y, Z = _grid(lower, upper, n)
    aux = np.empty((3, mx, my, mz), order='F')
    aux[0] = np.sin(2 * np.pi * Y) + 0.3
    aux[1] = np.cos(2 * np.pi * Z) - 0.2
    aux[2] = np.sin(2 * np.pi * X)
    q = np.empty((1, mx, my, mz), order='F')
    q[0] = np.exp(-40 * ((X - 0.5) ** 2 + (Y - 0.4) ** 2 + (Z - 0.6) ** 2)) + 0.1 * (X > 0.7)
    return _base('vc_advection_3D', lower, upper, n, q, aux=aux, limiters=VANLEER,
                 dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves, cfl_max=1.0,
                 cfl_desired=0.8, dt_initial=0.01, bc_lower=[PERIODIC] * 3, bc_upper=[PERIODIC] * 3,
                 aux_bc_lower=[PERIODIC] * 3, aux_bc_upper=[PERIODIC] * 3, tfinal=tfinal, num_output_times=num_output_times)


def burgers_3d(mx=24, my=20, mz=18, coeff=(1.0, 0.7, -0.5), dimensional_split=False, transverse_waves=22, tfinal=0.2,
               num_output_times=2):
    """3-D Burgers equation (riemann burgers_3D; the reference ships the solver but no example): smooth data with both
    signs, so that transonic rarefactions (the entropy fix) and all three coefficient signs occur; periodic."""
    X, Y, Z = _grid((0., 0., 0.), (1., 1., 1.), (mx, my, mz))
    q = np.empty((1, mx, my, mz), order='F')
    q[0] = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.4 * np.sin(2 * np.pi * (Z + 0.3 * X)) + 0.1
    return _base('burgers_3D', (0., 0., 0.), (1., 1., 1.), (mx, my, mz), q, problem_data={'coeff': list(coeff)},
                 limiters=VANLEER, dimensional_split=bool(dimensional_split), transverse_waves=transverse_waves,
                 cfl_max=1.0, cfl_desired=0.8, dt_initial=0.01, bc_lower=[PERIODIC] * 3, bc_upper=[PERIODIC] * 3,
                 tfinal=tfinal, num_output_times=num_output_times)


def euler_3d_shocktube(mx=10, my=10, mz=128, tfinal=0.2, num_output_times=10):
    """pyclaw/examples/euler_3d/shocktube.py:17-66 as run by test_euler3d.py:27 (dimension-split, MC)."""
    gamma = 1.4
    X, Y, Z = _grid((-1., -1., -1.), (1., 1., 1.), (mx, my, mz))
    q = np.zeros((5, mx, my, mz), order='F')
    pressure = 3. * (Z <= 0) + 1. * (Z > 0)
    q[0] = 3. * (Z <= 0) + 1. * (Z > 0)
    q[4] = pressure / (gamma - 1.)
    return _base('euler_3D', (-1., -1., -1.), (1., 1., 1.), (mx, my, mz), q, problem_data={'gamma': gamma},
                 limiters=MC, dimensional_split=True, transverse_waves=22, cfl_max=1.0, cfl_desired=0.8,
                 bc_lower=[EXTRAP] * 3, bc_upper=[EXTRAP] * 3, tfinal=tfinal, num_output_times=num_output_times)


def euler_3d_shock_bubble(q0, consts, tfinal=0.1, num_output_times=10, dimensional_split=False):
    """pyclaw/examples/euler_3d/shock_bubble.py:75-156 as run by test_euler3d.py:31 (64x16x16, unsplit,
    transverse_waves=22, default minmod limiter, custom inflow BC at x-lower, walls at y/z-lower).
    ``q0`` is the initial state (scipy sub-cell quadrature; stored fixture, tests/golden/make_golden.py)
    and ``consts`` = (gamma, rho_shock, v_shock, e_shock)."""
    gamma, rho_shock, v_shock, e_shock = [float(c) for c in consts]
    n = q0.shape[1:]

    def incoming_shock(qbc, num_ghost):                 # shock_bubble.py:63-72
        for i in range(num_ghost):
            qbc[0, i, ...] = rho_shock
            qbc[1, i, ...] = rho_shock * v_shock
            qbc[2, i, ...] = 0.
            qbc[3, i, ...] = 0.
            qbc[4, i, ...] = e_shock

    return _base('euler_3D', (0., 0., 0.), (2., 0.5, 0.5), n, np.array(q0, order='F'), problem_data={'gamma': gamma},
                 limiters=MINMOD, dimensional_split=bool(dimensional_split), transverse_waves=22,
                 bc_lower=[0, WALL, WALL], bc_upper=[EXTRAP] * 3, custom_bc_lower=incoming_shock,
                 tfinal=tfinal, num_output_times=num_output_times)


def elasticity_2d(mx=80, my=80, tfinal=0.25, num_output_times=5, dimensional_split=False, transverse_waves=2,
                  inclusion=True):
    """vc_elasticity_2D (riemann/src/rpn2_vc_elasticity.f90, rpt2_vc_elasticity.f90; the reference ships no PyClaw
    example for it): a compressional pulse sigma11 = sigma22 = exp(-150 r^2) in a unit square of (rho, lambda, mu) =
    (1, 2, 1), optionally with a stiff, dense disc (4, 8, 4) off-centre that converts part of it into shear waves.
    q = (sigma11, sigma22, sigma12, u, v); aux = (rho, lambda, mu, cp, cs); MC limiter, extrapolation."""
    X, Y = _grid((0., 0.), (1., 1.), (mx, my))
    q = np.zeros((5, mx, my), order='F')
    r2 = (X - 0.5) ** 2 + (Y - 0.5) ** 2
    q[0] = np.exp(-150.0 * r2)
    q[1] = q[0]
    inside = ((X - 0.72) ** 2 + (Y - 0.6) ** 2 <= 0.15 ** 2) if inclusion else np.zeros_like(X, dtype=bool)
    rho = np.where(inside, 4.0, 1.0)
    lam = np.where(inside, 8.0, 2.0)
    mu = np.where(inside, 4.0, 1.0)
    aux = np.empty((5, mx, my), order='F')
    aux[0], aux[1], aux[2] = rho, lam, mu
    aux[3] = np.sqrt((lam + 2.0 * mu) / rho)
    aux[4] = np.sqrt(mu / rho)
    return _base('vc_elasticity_2D', (0., 0.), (1., 1.), (mx, my), q, aux=aux, limiters=MC,
                 dimensional_split=bool(dimensional_split), transverse_waves=int(transverse_waves),
                 cfl_max=1.0, cfl_desired=0.9, dt_initial=0.5 / max(mx, my) / 2.0,
                 bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], aux_bc_lower=[EXTRAP, EXTRAP],
                 aux_bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal, num_output_times=num_output_times)


def advection_reaction_2d(mx=100, my=100, tfinal=4.0, num_output_times=10):
    """pyclaw/examples/advection_reaction_2d/advection_reaction.py:36-110 (classic, dimensionally split): two unknowns
    transported by the time-periodic swirl velocity field psi = sin^2(pi x) sin^2(pi y) / pi (edge velocities in aux,
    rewritten by ``before_step`` from t + dt/2 every step) and coupled by the source term p_t = eps q, q_t = eps p
    (``step_source``, Godunov splitting); the example's own Riemann solver rpn2_vc_advection_vector + rpt2_dummy;
    minmod (default), extrapolation."""
    t_period, epsilon = 4.0, 0.1
    dx, dy = 1.0 / mx, 1.0 / my
    xe = np.array([0.0 + i * dx for i in range(mx + 1)])           # Dimension.nodes (geometry.py:541-545)
    ye = np.array([0.0 + j * dy for j in range(my + 1)])
    X, Y = np.meshgrid(xe, ye, indexing='ij')

    def psi(x, y):
        return (np.sin(np.pi * x)) ** 2 * (np.sin(np.pi * y)) ** 2 / np.pi

    def velocities():
        a = np.empty((2, mx, my), order='F')
        a[0] = -(psi(X[:-1, 1:], Y[:-1, 1:]) - psi(X[:-1, :-1], Y[:-1, :-1])) / dy
        a[1] = (psi(X[1:, :-1], Y[1:, :-1]) - psi(X[:-1, :-1], Y[:-1, :-1])) / dx
        return a

    def set_velocities(q, aux, t, dt):
        v_t = np.cos(2 * np.pi * (t + dt / 2) / t_period)
        aux[...] = velocities()
        aux[:] = v_t * aux[:]

    def source_step(q, aux, dt):
        qq = q.copy()
        q[0] = q[0] + epsilon * dt * qq[1]
        q[1] = q[1] + epsilon * dt * qq[0]

    Xc, Yc = _grid((0., 0.), (1., 1.), (mx, my))
    q = np.empty((2, mx, my), order='F')
    q[0] = (Xc <= 0.5)
    q[1] = (Yc <= 0.5)
    return _base('vc_advection_vector_2D', (0., 0.), (1., 1.), (mx, my), q, aux=velocities(), limiters=MINMOD,
                 dimensional_split=True, bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP],
                 aux_bc_lower=[EXTRAP, EXTRAP], aux_bc_upper=[EXTRAP, EXTRAP], before_step_full=set_velocities,
                 step_source=source_step, source_split=1, tfinal=tfinal, num_output_times=num_output_times)


def kpp_2d(mx=200, my=200, tfinal=1.0, num_output_times=10):
    """pyclaw/examples/kpp/kpp.py:18-55 (classic): q_t + sin(q)_x + cos(q)_y = 0, disc of 3.5 pi in a sea of pi/4,
    dimensional splitting, minmod, extrapolation."""
    X, Y = _grid((-2., -2.), (2., 2.), (mx, my))
    r = np.sqrt(X ** 2 + Y ** 2)
    q = np.empty((1, mx, my), order='F')
    q[0] = 0.25 * np.pi + 3.25 * np.pi * (r <= 1.0)
    return _base('kpp_2D', (-2., -2.), (2., 2.), (mx, my), q, limiters=MINMOD, dimensional_split=True,
                 cfl_max=1.0, cfl_desired=0.9, bc_lower=[EXTRAP, EXTRAP], bc_upper=[EXTRAP, EXTRAP], tfinal=tfinal,
                 num_output_times=num_output_times)


# ---- 1-D (ClawSolver1D / step1.f90) ------------------------------------------------------------------
def advection_1d(nx=400, tfinal=1.0):
    """pyclaw/examples/advection_1d/advection_1d.py:30-78 (Fortran classic): Gaussian pulse, periodic."""
    xc = centers(0.0, 1.0, nx)
    q = np.empty((1, nx), order='F')
    beta, gamma, x0 = 100, 0, 0.75
    q[0] = np.exp(-beta * (xc - x0) ** 2) * np.cos(gamma * (xc - x0))
    return _base('advection_1D', (0.0,), (1.0,), (nx,), q, problem_data={'u': 1.0}, limiters=MINMOD,
                 bc_lower=[PERIODIC], bc_upper=[PERIODIC], tfinal=tfinal)


def acoustics_1d(num_cells=100, tfinal=1.0):
    """pyclaw/examples/acoustics_1d_homogeneous/acoustics_1d.py:47-114 (classic): MC limiter, periodic."""
    xc = centers(0.0, 1.0, num_cells)
    q = np.zeros((2, num_cells), order='F')
    q[0] = np.exp(-200 * (xc - 0.5) ** 2) * np.cos(0 * (xc - 0.5))
    rho, bulk = 1.0, 1.0
    zz, cc = np.sqrt(rho * bulk), np.sqrt(bulk / rho)
    return _base('acoustics_1D', (0.0,), (1.0,), (num_cells,), q,
                 problem_data={'rho': rho, 'bulk': bulk, 'zz': zz, 'cc': cc}, limiters=MC,
                 bc_lower=[PERIODIC], bc_upper=[PERIODIC], dt_initial=(1.0 / num_cells) / cc * 0.1, tfinal=tfinal)


def burgers_1d(mx=500, tfinal=0.5):
    """pyclaw/examples/burgers_1d/burgers_1d.py:22-64 (classic): van Leer, periodic, sine + 0.5."""
    xc = centers(0.0, 1.0, mx)
    q = np.empty((1, mx), order='F')
    q[0] = np.sin(np.pi * 2 * xc) + 0.50
    return _base('burgers_1D', (0.0,), (1.0,), (mx,), q, limiters=VANLEER, bc_lower=[PERIODIC], bc_upper=[PERIODIC],
                 tfinal=tfinal)


def dam_break_1d(mx=500, riemann_solver='roe', tfinal=2.0):
    """pyclaw/examples/shallow_1d/dam_break.py:22-99 (classic): van Leer, extrapolation, h = 3 | 1."""
    xc = centers(-5.0, 5.0, mx)
    q = np.zeros((2, mx), order='F')
    q[0] = 3. * (xc <= 0.) + 1. * (xc > 0.)
    rp = 'shallow_hlle_1D' if riemann_solver == 'hlle' else 'shallow_roe_with_efix_1D'
    return _base(rp, (-5.0,), (5.0,), (mx,), q, problem_data={'grav': 1.0}, limiters=VANLEER,
                 bc_lower=[EXTRAP], bc_upper=[EXTRAP], tfinal=tfinal)


def traffic_1d(mx=500, tfinal=2.0):
    """pyclaw/examples/traffic/traffic.py:14-50 (classic): LWR traffic flow, minmod (default), extrapolation."""
    xc = centers(-1.0, 1.0, mx)
    q = np.empty((1, mx), order='F')
    q[0] = 0.75 * (xc < 0) + 0.1 * (xc > 0.)
    return _base('traffic_1D', (-1.0,), (1.0,), (mx,), q, problem_data={'efix': True, 'umax': 1.}, limiters=MINMOD,
                 bc_lower=[EXTRAP], bc_upper=[EXTRAP], tfinal=tfinal)


def advection_1d_variable(mx=100, tfinal=1.0):
    """pyclaw/examples/advection_1d_variable/variable_coefficient_advection.py:22-107 (Fortran classic): colour
    equation q_t + u(x) q_x = 0 with u = 2 + sin(2 pi x) in aux, MC limiter, periodic."""
    x = centers(0.0, 1.0, mx)
    beta, gamma, x0, x1, x2 = 100., 0., 0.3, 0.7, 0.9
    qg = np.exp(-beta * (x - x0) ** 2) * np.cos(gamma * (x - x0))
    qs = (x > x1) * 1.0 - (x > x2) * 1.0
    q = np.empty((1, mx), order='F')
    q[0] = qg + qs
    aux = np.empty((1, mx), order='F')
    aux[0] = np.sin(2. * np.pi * x) + 2
    return _base('advection_color_1D', (0.0,), (1.0,), (mx,), q, aux=aux, limiters=MC, bc_lower=[PERIODIC],
                 bc_upper=[PERIODIC], aux_bc_lower=[PERIODIC], aux_bc_upper=[PERIODIC], tfinal=tfinal)


def sill_1d(mx=500, tfinal=1.0):
    """pyclaw/examples/shallow_1d/sill.py:22-70 (Fortran): flow over a sill, f-wave solver with bathymetry in aux,
    van Leer, extrapolation."""
    xc = centers(-1.0, 1.0, mx)
    aux = np.empty((1, mx), order='F')
    aux[0] = 0.8 * np.exp(-xc ** 2 / 0.2 ** 2) - 1.0
    q = np.zeros((2, mx), order='F')
    q[0] = 0.1 * np.exp(-(xc + 0.4) ** 2 / 0.2 ** 2) - aux[0]
    return _base('shallow_bathymetry_fwave_1D', (-1.0,), (1.0,), (mx,), q, aux=aux,
                 problem_data={'grav': 9.8, 'dry_tolerance': 1e-3, 'sea_level': 0.0}, limiters=VANLEER, fwave=True,
                 bc_lower=[EXTRAP], bc_upper=[EXTRAP], aux_bc_lower=[EXTRAP], aux_bc_upper=[EXTRAP], tfinal=tfinal)


def stegoton_1d(xupper=300.0, cells_per_layer=12, tfinal=50.0, num_output_times=20, exponential=True):
    """pyclaw/examples/stegoton_1d/stegoton.py:33-170 (Fortran classic): nonlinear elasticity in a layered medium,
    f-waves, minmod (default); the pulse enters through a moving-wall ``BC.custom`` on the left (:62-76) that reads
    the time and the interior aux and leaves the outermost momentum ghost cell untouched; extrapolation on the
    right.  ``exponential=False`` switches to the linear stress relation (coefA = 0, coefB = K), which involves no
    ``exp`` and is therefore bit-reproducible across math libraries."""
    mx = int(round(xupper - 0.0)) * cells_per_layer
    x = centers(0.0, xupper, mx)
    alpha, KA, KB, rhoA, rhoB = 0.5, 1.0, 4.0, 1.0, 4.0
    xfrac = x - np.floor(x)
    aux = np.empty((3, mx), order='F')
    aux[0] = rhoA * (xfrac < alpha) + rhoB * (xfrac >= alpha)
    aux[1] = KA * (xfrac < alpha) + KB * (xfrac >= alpha)
    aux[2] = 0.
    if not exponential:
        aux[2] = aux[1]
        aux[1] = 0.
    q = np.zeros((2, mx), order='F')
    t1, tw1, a1 = 10.0, 10.0, 0.1
    kwall = 1 if exponential else 2

    def moving_wall(qbc, auxbc, t, g):
        qbc[0, :g] = qbc[0, g]
        t0 = (t - t1) / tw1
        vwall = -a1 * (1. + np.cos(t0 * np.pi)) if abs(t0) <= 1. else 0.
        for ibc in range(g - 1):
            qbc[1, g - ibc - 1] = 2 * vwall * auxbc[kwall, g + ibc] - qbc[1, g + ibc]

    return _base('nonlinear_elasticity_fwave_1D', (0.0,), (xupper,), (mx,), q, aux=aux, limiters=MINMOD, fwave=True,
                 bc_lower=[0], bc_upper=[EXTRAP], aux_bc_lower=[EXTRAP], aux_bc_upper=[EXTRAP],
                 custom_bc_lower_full=moving_wall, tfinal=tfinal, num_output_times=num_output_times)


def shocktube_1d(mx=800, tfinal=0.4, riemann_solver='hlle'):
    """pyclaw/examples/euler_1d/shocktube.py:24-68 (classic, Fortran branch = euler_hlle_1D): Sod shock tube,
    minmod (default), extrapolation.  ``riemann_solver='roe'`` runs the same data through euler_with_efix_1D."""
    gamma = 1.4
    x = centers(-1.0, 1.0, mx)
    q = np.zeros((3, mx), order='F')
    q[0] = (x < 0.) * 1. + (x >= 0.) * (1. / 8)
    velocity = q[1] / q[0]
    pressure = (x < 0.) * 1. + (x >= 0.) * 0.1
    q[2] = pressure / (gamma - 1.) + 0.5 * q[0] * velocity ** 2
    rp = 'euler_hlle_1D' if riemann_solver == 'hlle' else 'euler_with_efix_1D'
    return _base(rp, (-1.0,), (1.0,), (mx,), q, problem_data={'gamma': gamma, 'gamma1': gamma - 1.},
                 limiters=MINMOD, bc_lower=[EXTRAP], bc_upper=[EXTRAP], tfinal=tfinal)


def blast_1d(mx=800, tfinal=0.038):
    """pyclaw/examples/euler_1d/woodward_colella_blast.py:27-82 with solver_type='classic': euler_with_efix_1D, MC
    limiter, solid walls."""
    gamma = 1.4
    x = centers(0.0, 1.0, mx)
    q = np.zeros((3, mx), order='F')
    q[0] = 1.
    q[2] = ((x < 0.1) * 1.e3 + (0.1 <= x) * (x < 0.9) * 1.e-2 + (0.9 <= x) * 1.e2) / (gamma - 1.)
    return _base('euler_with_efix_1D', (0.0,), (1.0,), (mx,), q, problem_data={'gamma': gamma}, limiters=MC,
                 bc_lower=[WALL], bc_upper=[WALL], tfinal=tfinal)


def traffic_vc_1d(mx=400, tfinal=1.0, kind='vc'):
    """LWR traffic with a variable speed limit (riemann traffic_vc_1D / traffic_vc_fwave_1D / traffic_vc_tracer_1D; the
    reference ships no PyClaw example for them): periodic road, speed limit 1 with a slow stretch (0.5) between
    x = 0.4 and 0.7, a platoon of dense traffic behind it; the tracer variant carries a second, passively advected
    component.  MC limiter."""
    x = centers(0.0, 1.0, mx)
    aux = np.empty((1, mx), order='F')
    aux[0] = 1.0 - 0.5 * ((x > 0.4) & (x < 0.7))
    rows = [0.2 + 0.5 * np.exp(-200.0 * (x - 0.2) ** 2)]
    rp = {'vc': 'traffic_vc_1D', 'fwave': 'traffic_vc_fwave_1D', 'tracer': 'traffic_vc_tracer_1D'}[kind]
    if kind == 'tracer':
        rows.append(np.sin(2 * np.pi * x))
    q = np.asfortranarray(np.stack(rows))
    return _base(rp, (0.0,), (1.0,), (mx,), q, aux=aux, limiters=MC, fwave=(kind == 'fwave'), bc_lower=[PERIODIC],
                 bc_upper=[PERIODIC], aux_bc_lower=[PERIODIC], aux_bc_upper=[PERIODIC], tfinal=tfinal)


def cubic_1d(mx=1000, tfinal=0.2, qL=4.0, qR=-2.0):
    """pyclaw/examples/cubic_1d/cubic.py:14-54 (classic): q_t + (q^3)_x = 0, Riemann data qL | qR at x = -0.5,
    van Leer, cfl_desired 0.5, extrapolation.  test_cubic.py:3-21 checks the shock position x0 = -0.5 + 12 t."""
    x = centers(-1.0, 3.0, mx)
    q = np.empty((1, mx), order='F')
    q[0] = (x < -0.5) * qL + (x >= -0.5) * qR
    return _base('cubic_1D', (-1.0,), (3.0,), (mx,), q, limiters=VANLEER, cfl_max=1.0, cfl_desired=0.5,
                 bc_lower=[EXTRAP], bc_upper=[EXTRAP], tfinal=tfinal)


def acoustics_1d_variable(mx=400, tfinal=0.6):
    """acoustics_variable_1D (riemann/src/rp1_acoustics_variable.f90; the reference ships no PyClaw example for the
    classic solver): a pulse crossing an impedance jump at x = 0.5, aux = (Z, c), MC, extrapolation."""
    x = centers(0.0, 1.0, mx)
    rho = 1.0 + 3.0 * (x > 0.5)
    bulk = 1.0 + 0.0 * x
    aux = np.empty((2, mx), order='F')
    aux[1] = np.sqrt(bulk / rho)
    aux[0] = rho * aux[1]
    q = np.zeros((2, mx), order='F')
    q[0] = np.exp(-400 * (x - 0.3) ** 2)
    q[1] = q[0] / aux[0]
    return _base('acoustics_variable_1D', (0.0,), (1.0,), (mx,), q, aux=aux, limiters=MC, bc_lower=[EXTRAP],
                 bc_upper=[EXTRAP], aux_bc_lower=[EXTRAP], aux_bc_upper=[EXTRAP], tfinal=tfinal)


def make_claw(sc, kernel_language='CUDA', pyclaw=None, **solver_kw):
    """Build a PyClaw-API Controller (solver + solution) from a scenario dict."""
    if pyclaw is None:
        from clawpack_b200 import pyclaw
    from clawpack_b200 import riemann
    ndim = len(sc['num_cells'])
    Solver = {1: pyclaw.ClawSolver1D, 2: pyclaw.ClawSolver2D, 3: pyclaw.ClawSolver3D}[ndim]
    # 'rp_object': a solver object that is not one of clawpack_b200.riemann's (a user-compiled CUDA plug-in)
    solver = Solver(sc['rp_object'] if sc.get('rp_object') is not None else getattr(riemann, sc['rp']))
    solver.kernel_language = kernel_language
    solver.limiters = sc['limiters']
    solver.order = sc['order']
    if ndim > 1:
        solver.dimensional_split = sc['dimensional_split']
        if sc['transverse_waves'] is not None:
            solver.transverse_waves = sc['transverse_waves']
    solver.cfl_max = sc['cfl_max']
    solver.cfl_desired = sc['cfl_desired']
    solver.dt_initial = sc['dt_initial']
    for i in range(ndim):
        solver.bc_lower[i] = sc['bc_lower'][i]
        solver.bc_upper[i] = sc['bc_upper'][i]
        if sc['aux'] is not None:
            solver.aux_bc_lower[i] = sc['aux_bc_lower'][i]
            solver.aux_bc_upper[i] = sc['aux_bc_upper'][i]
    if sc.get('custom_bc_lower') is not None:           # BC.custom (solver.py:393-398): q fill of one face
        f = sc['custom_bc_lower']
        solver.user_bc_lower = lambda state, dim, t, qbc, auxbc, num_ghost: f(qbc, num_ghost)
    if sc.get('custom_bc_lower_full') is not None:      # BC.custom fill that also reads auxbc and t
        ff = sc['custom_bc_lower_full']
        solver.user_bc_lower = lambda state, dim, t, qbc, auxbc, num_ghost: ff(qbc, auxbc, t, num_ghost)
    if sc.get('custom_aux_bc_lower') is not None:
        fl = sc['custom_aux_bc_lower']
        solver.user_aux_bc_lower = lambda state, dim, t, qbc, auxbc, num_ghost: fl(auxbc, num_ghost)
    if sc.get('custom_aux_bc_upper') is not None:
        fu = sc['custom_aux_bc_upper']
        solver.user_aux_bc_upper = lambda state, dim, t, qbc, auxbc, num_ghost: fu(auxbc, num_ghost)
    if sc.get('step_source') is not None:               # host callbacks: f(q, aux, dt) / f(q, aux) on the interior arrays
        fs = sc['step_source']
        solver.step_source = lambda solver_, state, dt: fs(state.q, state.aux, dt)
        solver.source_split = sc.get('source_split', 1)
    if sc.get('device_callbacks') is not None:          # the same hooks as device functors (pyclaw.device_callbacks)
        dc = sc['device_callbacks']
        if 'step_source' in dc:
            solver.step_source = dc['step_source']
        if 'user_bc_lower' in dc:
            solver.user_bc_lower = dc['user_bc_lower']
    if sc.get('before_step') is not None:
        fb = sc['before_step']
        solver.before_step = lambda solver_, state: fb(state.q, state.aux)
    if sc.get('before_step_full') is not None:          # f(q, aux, t, dt): dt is the size of the PREVIOUS step (solver.py:609)
        fbf = sc['before_step_full']
        solver.before_step = lambda solver_, state: fbf(state.q, state.aux, state.t, solver_.dt)
    if sc.get('fwave'):
        solver.fwave = True
    if sc.get('compute_gauge_values') is not None:
        solver.compute_gauge_values = sc['compute_gauge_values']
    for k, v in solver_kw.items():
        setattr(solver, k, v)
    names = ['x', 'y', 'z'][:ndim]
    dims = [pyclaw.Dimension(lo, up, n, name=nm)
            for lo, up, n, nm in zip(sc['lower'], sc['upper'], sc['num_cells'], names)]
    domain = pyclaw.Domain(dims)
    if sc.get('mapc2p') is not None:
        domain.grid.mapc2p = sc['mapc2p']
    num_aux = 0 if sc['aux'] is None else sc['aux'].shape[0]
    state = pyclaw.State(domain, sc['q'].shape[0], num_aux)
    state.q[...] = sc['q']
    if num_aux:
        state.aux[...] = sc['aux']
    state.index_capa = sc.get('index_capa', -1)
    state.problem_data.update(sc['problem_data'])
    if sc.get('gauges'):
        state.grid.add_gauges(sc['gauges'])
        state.keep_gauges = True
    claw = pyclaw.Controller()
    claw.output_format = None          # no frame files unless the caller asks (the reference tests pass disable_output)
    claw.tfinal = sc['tfinal']
    claw.num_output_times = sc['num_output_times']
    claw.solution = pyclaw.Solution(state, domain)
    claw.solver = solver
    claw.output_format = None
    claw.keep_copy = True
    return claw
