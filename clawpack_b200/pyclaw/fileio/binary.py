"""Raw binary frame files in the layout the reference's reader understands
(pyclaw/src/pyclaw/fileio/binary.py:19-208; the files are those of AMRClaw's Fortran ``valout``):

  <prefix>.tNNNN   the ascii header of fileio/ascii.py with ``file_format`` = binary64 / binary32 and ``num_ghost`` = 0
  <prefix>.qNNNN   patch headers only (patch_number, AMR_level, m<dim>.., <dim>low.., d<dim>.., blank line)
  <prefix>.bNNNN   q of every patch, (num_eqn, mx[, my[, mz]]) in Fortran order, float64 (or float32), back to back
  <prefix>.aNNNN   aux of every patch in the same way (only when requested)
  <prefix>.pklNNNN as in the ascii format (problem_data, mapc2p)

The reference has no PyClaw-side writer for this format ("the code that writes these files is in AMRClaw"); this one
exists because an 8192^2 frame is 2.1 GB raw and would be 4.8 GB of text: ``claw.output_format = 'binary'``.
``read`` follows the reference reader, including ghost-cell stripping when the .t header says ``num_ghost`` > 0.
"""
import os
import pickle

import numpy as np

from .ascii import _data_line, _frame_name, read_patch_header, read_t, write_patch_header

_DTYPES = {'binary': np.float64, 'binary64': np.float64, 'binary32': np.float32}


def _dtype(fmt):
    try:
        return _DTYPES[fmt]
    except KeyError:
        raise Exception('Unrecognized file_format: %s' % fmt)


def write(solution, frame, path, file_prefix='fort', write_aux=False, options={}, write_p=False):
    fmt = options.get('format', 'binary64')
    if fmt == 'binary':
        fmt = 'binary64'
    dt = _dtype(fmt)
    with open(_frame_name(path, file_prefix, 't', frame), 'w') as f:
        f.write('%18.8e     time\n' % solution.t)
        f.write('%5i                  num_eqn\n' % (solution.mp if write_p else solution.num_eqn))
        f.write('%5i                  nstates\n' % len(solution.states))
        f.write('%5i                  num_aux\n' % solution.num_aux)
        f.write('%5i                  num_dim\n' % solution.domain.num_dim)
        f.write('%5i                  num_ghost\n' % 0)
        f.write('%s                  file_format\n' % fmt)
    with open(_frame_name(path, file_prefix, 'q', frame), 'w') as f:
        for state in solution.states:
            write_patch_header(f, state.patch)
    with open(_frame_name(path, file_prefix, 'b', frame), 'wb') as f:
        for state in solution.states:
            np.asarray(state.p if write_p else state.q).astype(dt, copy=False).ravel(order='F').tofile(f)
    if solution.num_aux > 0 and write_aux:
        with open(_frame_name(path, file_prefix, 'a', frame), 'wb') as f:
            for state in solution.states:
                np.asarray(state.aux).astype(dt, copy=False).ravel(order='F').tofile(f)
    sol_dict = {'t': solution.t, 'num_eqn': solution.mp if write_p else solution.num_eqn,
                'nstates': len(solution.states), 'num_aux': solution.num_aux, 'num_dim': solution.domain.num_dim,
                'write_aux': write_aux, 'problem_data': solution.problem_data, 'mapc2p': solution.state.grid.mapc2p}
    try:
        blob = pickle.dumps(sol_dict)
    except Exception:                       # a locally defined mapc2p cannot be pickled: keep the rest
        sol_dict['mapc2p'] = None
        blob = pickle.dumps(sol_dict)
    with open(_frame_name(path, file_prefix, 'pkl', frame), 'wb') as f:
        f.write(blob)


def _strip(a, g, num_dim):
    if g == 0:
        return a
    return a[(slice(None),) + (slice(g, -g),) * num_dim]


def read(solution, frame, path='./', file_prefix='fort', read_aux=False, options={}):
    from .. import geometry
    State = solution.claw_package.State
    t, num_eqn, nstates, num_aux, num_dim, num_ghost, header_format = read_t(frame, path, file_prefix)
    # the reference reader takes the precision from options['format'] (default binary64, binary.py:56); the .t header
    # carries it too, which settles the case where the option is absent
    fmt = options.get('format', header_format if header_format in _DTYPES else 'binary64')
    dt = _dtype(fmt)
    problem_data, mapc2p = None, None
    pkl = _frame_name(path, file_prefix, 'pkl', frame)
    if os.path.exists(pkl):
        with open(pkl, 'rb') as f:
            d = pickle.load(f)
        problem_data, mapc2p = d.get('problem_data', None), d.get('mapc2p', None)
    qdata = np.fromfile(_frame_name(path, file_prefix, 'b', frame), dtype=dt)
    start = 0
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
                state.aux[:] = np.nan
            shape = [num_eqn] + [n + 2 * num_ghost for n in patch.num_cells_global]
            size = int(np.prod(shape))
            if start + size > qdata.size:
                raise IOError('%s holds %d values, patch %d needs %d more' %
                              (_frame_name(path, file_prefix, 'b', frame), qdata.size, patch.patch_index, size))
            block = np.reshape(qdata[start:start + size], shape, order='F')
            state.q = np.asfortranarray(_strip(block, num_ghost, num_dim), dtype=np.float64)
            start += size
            solution.states.append(state)
            patches.append(state.patch)
    solution.domain = geometry.Domain(patches)
    if num_aux > 0 and read_aux:
        fname = _frame_name(path, file_prefix, 'a', frame)
        if not os.path.exists(fname):
            fname = _frame_name(path, file_prefix, 'a', 0)
            if not os.path.exists(fname):
                return
        auxdata = np.fromfile(fname, dtype=dt)
        start = 0
        for state in solution.states:
            shape = [num_aux] + [n + 2 * num_ghost for n in state.patch.num_cells_global]
            size = int(np.prod(shape))
            block = np.reshape(auxdata[start:start + size], shape, order='F')
            state.aux = np.asfortranarray(_strip(block, num_ghost, num_dim), dtype=np.float64)
            start += size
