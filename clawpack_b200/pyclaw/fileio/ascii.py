"""Classic Clawpack ascii frame files: byte-for-byte the layout of pyclaw/src/pyclaw/fileio/ascii.py:17-126
(write) and readable by its reader (:129-330) and by visclaw.

  <prefix>.tNNNN   time, num_eqn, nstates, num_aux, num_dim, num_ghost, file_format   ("%18.8e     name" / "%5i   ...")
  <prefix>.qNNNN   per patch: header (patch_number, AMR_level, m<dim>.., <dim>low.., d<dim>..), blank line, then one
                   line per cell with num_eqn "%18.8e" fields, x fastest; a blank line closes every x-row (and every
                   xy-plane in 3-D)
  <prefix>.aNNNN   the same for aux (only when requested)
  <prefix>.pklNNNN pickle of {t, num_eqn, nstates, num_aux, num_dim, write_aux, problem_data, mapc2p}

The arrays are formatted row-block by row-block with numpy instead of one Python write per number.
"""
import os
import pickle

import numpy as np


def _frame_name(path, prefix, kind, frame):
    return os.path.join(path, '%s.%s' % (prefix, kind)) + str(frame).zfill(4)


def write(solution, frame, path, file_prefix='fort', write_aux=False, options={}, write_p=False):
    with open(_frame_name(path, file_prefix, 't', frame), 'w') as f:                     # ascii.py:41-49
        f.write('%18.8e     time\n' % solution.t)
        f.write('%5i                  num_eqn\n' % solution.num_eqn)
        f.write('%5i                  nstates\n' % len(solution.states))
        f.write('%5i                  num_aux\n' % solution.num_aux)
        f.write('%5i                  num_dim\n' % solution.domain.num_dim)
        f.write('%5i                  num_ghost\n' % 0)
        f.write('%s                  file_format\n' % 'ascii')
    with open(_frame_name(path, file_prefix, 'q', frame), 'w') as f:                     # :52-60
        for state in solution.states:
            write_patch_header(f, state.patch)
            write_array(f, state.patch, state.p if write_p else state.q)
    if solution.num_aux > 0 and write_aux:                                               # :63-68
        with open(_frame_name(path, file_prefix, 'a', frame), 'w') as f:
            for state in solution.states:
                write_patch_header(f, state.patch)
                write_array(f, state.patch, state.aux)
    sol_dict = {'t': solution.t, 'num_eqn': solution.num_eqn, 'nstates': len(solution.states),   # :71-80
                'num_aux': solution.num_aux, 'num_dim': solution.domain.num_dim, 'write_aux': write_aux,
                'problem_data': solution.problem_data, 'mapc2p': solution.state.grid.mapc2p}
    if write_p:
        sol_dict['num_eqn'] = solution.mp
    try:
        blob = pickle.dumps(sol_dict)
    except Exception:                       # a locally defined mapc2p cannot be pickled: keep the rest
        sol_dict['mapc2p'] = None
        blob = pickle.dumps(sol_dict)
    with open(_frame_name(path, file_prefix, 'pkl', frame), 'wb') as f:
        f.write(blob)


def write_patch_header(f, patch):                                                        # :83-93
    f.write('%5i                  patch_number\n' % patch.patch_index)
    f.write('%5i                  AMR_level\n' % patch.level)
    for dim in patch.dimensions:
        f.write('%5i                  m%s\n' % (dim.num_cells, dim.name))
    for dim in patch.dimensions:
        f.write('%18.8e     %slow\n' % (dim.lower, dim.name))
    for dim in patch.dimensions:
        f.write('%18.8e     d%s\n' % (dim.delta, dim.name))
    f.write('\n')


def _write_rows(f, block):
    """block[k, m] -> one line per k with the m fields back to back."""
    fmt = '%18.8e' * block.shape[1]
    f.write('\n'.join(fmt % tuple(row) for row in block))
    f.write('\n')


def write_array(f, patch, q):                                                            # :96-126
    q = np.asarray(q)
    nd = patch.num_dim
    if nd == 1:
        _write_rows(f, q.T)
    elif nd == 2:
        for j in range(q.shape[2]):
            _write_rows(f, q[:, :, j].T)
            f.write('\n')
    elif nd == 3:
        for l in range(q.shape[3]):
            for j in range(q.shape[2]):
                _write_rows(f, q[:, :, j, l].T)
                f.write('\n')
            f.write('\n')
    else:
        raise Exception('Dimension Exception in writing fort file.')


def _data_line(f, data_type=float):
    """First token of the next line (pyclaw/src/pyclaw/util.py read_data_line, single value)."""
    line = f.readline()
    while line and not line.strip():
        line = f.readline()
    return data_type(line.split()[0])


def read_t(frame, path='./', file_prefix='fort'):                                        # :236-288
    with open(_frame_name(path, file_prefix, 't', frame), 'r') as f:
        t = _data_line(f)
        num_eqn = _data_line(f, int)
        nstates = _data_line(f, int)
        num_aux = _data_line(f, int)
        num_dim = _data_line(f, int)
        try:
            num_ghost = _data_line(f, int)
        except Exception:
            num_ghost = 0
        try:
            file_format = _data_line(f, str)
        except Exception:
            file_format = None
    return [t, num_eqn, nstates, num_aux, num_dim, num_ghost, file_format]


def read_patch_header(f, num_dim, geometry):                                             # :291-330
    patch_index = _data_line(f, int)
    level = _data_line(f, int)
    n = [_data_line(f, int) for _ in range(num_dim)]
    lower = [_data_line(f) for _ in range(num_dim)]
    d = [_data_line(f) for _ in range(num_dim)]
    names = ['x', 'y', 'z']
    dims = [geometry.Dimension(lower[i], lower[i] + n[i] * d[i], n[i], name=names[i]) for i in range(num_dim)]
    patch = geometry.Patch(dims)
    patch.patch_index = patch_index
    patch.level = level
    return patch


def read_array(f, state, num_var):
    """The next prod(num_cells) non-blank lines -> array (num_var, mx[, my[, mz]]) in Fortran order."""
    shape = tuple(state.patch.num_cells_global)
    ncell = int(np.prod(shape))
    vals = np.empty((ncell, num_var))
    k = 0
    while k < ncell:
        line = f.readline()
        if not line:
            raise IOError('unexpected end of file: %d of %d cells read' % (k, ncell))
        if not line.strip():
            continue
        # fields are 18 characters wide and may touch ("-1.00000000e+00-2.00000000e+00" never happens with the
        # leading pad of %18.8e, but split on the fixed width anyway, like a Fortran formatted read)
        line = line.rstrip('\n')
        vals[k] = [float(line[18 * m:18 * (m + 1)]) for m in range(num_var)]
        k += 1
    return np.asfortranarray(vals.T.reshape((num_var,) + shape, order='F'))


def read(solution, frame, path='./', file_prefix='fort', read_aux=False, options={}):   # :129-233
    from .. import geometry
    State = solution.claw_package.State
    problem_data, mapc2p = None, None
    pkl = _frame_name(path, file_prefix, 'pkl', frame)
    if os.path.exists(pkl):
        with open(pkl, 'rb') as f:
            d = pickle.load(f)
        problem_data = d.get('problem_data', None)
        mapc2p = d.get('mapc2p', None)
    t, num_eqn, nstates, num_aux, num_dim, num_ghost, file_format = read_t(frame, path, file_prefix)
    patches = []
    with open(_frame_name(path, file_prefix, 'q', frame), 'r') as f:
        for _ in range(nstates):
            patch = read_patch_header(f, num_dim, geometry)
            state = State(patch, num_eqn, num_aux)
            state.t = t
            if problem_data is not None:
                state.problem_data = problem_data
            if mapc2p is not None:
                state.grid.mapc2p = mapc2p
            if num_aux > 0:
                state.aux[:] = np.nan                  # "uninitialised" marker, as the reference does
            state.q = read_array(f, state, num_eqn)
            solution.states.append(state)
            patches.append(state.patch)
    solution.domain = geometry.Domain(patches)
    if num_aux > 0 and read_aux:
        fname = _frame_name(path, file_prefix, 'a', frame)
        if not os.path.exists(fname):
            fname = _frame_name(path, file_prefix, 'a', 0)       # aux of the initial frame serves every frame
            if not os.path.exists(fname):
                return
        with open(fname, 'r') as f:
            for state in solution.states:
                aux_patch = read_patch_header(f, num_dim, geometry)
                for da, dq in zip(aux_patch.dimensions, state.patch.dimensions):
                    assert da.num_cells == dq.num_cells, 'num_cells in the aux file header does not match the q file'
                    assert abs(da.lower - dq.lower) <= 1e-8 + 1e-15 * abs(dq.lower), 'lower in the aux file does not match'
                    assert abs(da.delta - dq.delta) <= 1e-8 + 1e-15 * abs(dq.delta), 'delta in the aux file does not match'
                state.aux = read_array(f, state, num_aux)
