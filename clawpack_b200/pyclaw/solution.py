"""Solution = states + domain (pyclaw/src/pyclaw/solution.py:17-253), with ``write`` / ``read`` of frame files in
the classic ascii layout (fileio/ascii.py; the hdf5 / netcdf / petsc formats of the reference need libraries
outside this build)."""
import logging
import os

from .geometry import Dimension, Domain, Patch


class Solution(object):
    _state_attrs = ('t', 'num_eqn', 'mp', 'mF', 'q', 'p', 'F', 'aux', 'capa', 'problem_data', 'num_aux', 'num_dim',
                    'p_centers', 'p_edges', 'c_centers', 'c_edges', 'num_cells', 'lower', 'upper', 'delta',
                    'centers', 'edges', 'gauges', 'grid')
    _start_frame = 0

    def __init__(self, *arg, **kargs):
        import sys
        pkg = kargs.get('claw_package')
        if pkg is not None and pkg in sys.modules:
            self.claw_package = sys.modules[pkg]
        else:
            name = self.__module__.rpartition('.')[0]
            if name not in sys.modules:
                raise NotImplementedError('Unable to determine solver package, please provide one')
            self.claw_package = sys.modules[name]
        State = self.claw_package.State
        self.states = []
        self.domain = None
        if len(arg) == 2:
            geom = arg[1]
            if isinstance(geom, Domain):
                self.domain = geom
            else:
                if not isinstance(geom, (list, tuple)):
                    geom = [geom]
                if isinstance(geom[0], Dimension):
                    self.domain = Domain(Patch(list(geom)))
                elif isinstance(geom[0], Patch):
                    self.domain = Domain(list(geom))
                else:
                    raise Exception('Invalid arguments for Solution initialization.')
            st = arg[0]
            if isinstance(st, State):
                self.states.append(st)
            elif isinstance(st, (list, tuple)):
                if isinstance(st[0], State):
                    self.states = list(st)
                elif isinstance(st[0], int):
                    self.states.append(State(self.domain, st[0], st[1]))
                else:
                    raise Exception('Invalid arguments for Solution initialization.')
            elif isinstance(st, int):
                self.states.append(State(self.domain, st))
            if self.states == [] or self.domain is None:
                raise Exception('Invalid arguments for Solution initialization.')
        elif len(arg) == 1 and isinstance(arg[0], int):        # Solution(frame, path=..., file_format=...): solution.py:150-165
            frame = arg[0]
            if 'count_from_zero' in kargs and kargs['count_from_zero'] is True:
                self._start_frame = 0
            else:
                self._start_frame = frame
            kargs.pop('count_from_zero', None)
            kargs.pop('claw_package', None)
            self.read(frame, **kargs)
        elif len(arg) == 0 and not [k for k in kargs if k != 'claw_package']:
            pass
        else:
            raise Exception('Invalid arguments for Solution initialization.')

    def __getattr__(self, key):
        if key in Solution._state_attrs:
            return getattr(self.states[0], key)
        raise AttributeError("'Solution' object has no attribute '" + key + "'")

    def __setattr__(self, key, value):
        if key in ('t', 'mp', 'mF'):
            for state in self.states:
                setattr(state, key, value)
        else:
            self.__dict__[key] = value

    state = property(lambda self: self.states[0])
    patch = property(lambda self: self.domain.patch)
    start_frame = property(lambda self: self._start_frame)

    def is_valid(self):
        return all(state.is_valid() for state in self.states)

    def set_all_states(self, attr, value, overwrite=True):
        for state in self.states:
            if getattr(state, attr) is None or overwrite:
                setattr(state, attr, value)

    def __deepcopy__(self, memo={}):
        import copy
        result = self.__class__()
        for state in self.states:
            result.states.append(copy.deepcopy(state))
        result.domain = copy.deepcopy(self.domain)
        return result

    # ---- frame files (solution.py:254-378) ------------------------------------------------------
    def get_write_func(self, file_format):
        if file_format == 'ascii':
            from .fileio import ascii
            return ascii.write
        if file_format[:6] == 'binary':                 # 'binary' / 'binary64' / 'binary32' (solution.py:286-293)
            from .fileio import binary
            if file_format == 'binary':
                return binary.write
            return lambda sol, frame, path, **kw: binary.write(
                sol, frame, path, **dict(kw, options=dict(kw.get('options') or {}, format=file_format)))
        raise ValueError("File format '%s' is not part of this build ('ascii', 'binary', 'binary32'; hdf5 / netcdf / "
                         "petsc / vtk need libraries that are not available here)." % file_format)

    def get_read_func(self, file_format):
        if file_format == 'ascii':
            from .fileio import ascii
            return ascii.read
        if file_format[:6] == 'binary':
            from .fileio import binary
            if file_format == 'binary':
                return binary.read
            return lambda sol, frame, path, **kw: binary.read(
                sol, frame, path, **dict(kw, options=dict(kw.get('options') or {}, format=file_format)))
        raise ValueError("File format '%s' cannot be read by this build ('ascii', 'binary', 'binary32')." % file_format)

    def write(self, frame, path='./', file_format='ascii', file_prefix=None, write_aux=False, options={},
              write_p=False):
        """Write frame `frame` in every requested format; state.q of a device-resident run is downloaded on access."""
        path = os.path.expandvars(os.path.expanduser(path))
        os.makedirs(path, exist_ok=True)
        format_list = [file_format] if isinstance(file_format, str) else list(file_format)
        for form in format_list:
            write_func = self.get_write_func(form)
            if file_prefix is None:
                write_func(self, frame, path, write_aux=write_aux, options=options, write_p=write_p)
            else:
                write_func(self, frame, path, file_prefix=file_prefix, write_aux=write_aux, options=options,
                           write_p=write_p)
            logging.getLogger('pyclaw.fileio').info('Wrote out solution in format %s for time t=%s' % (form, self.t))

    def read(self, frame, path='./_output', file_format=None, file_prefix='fort', read_aux=True, options={}, **kargs):
        """Fill this (empty) Solution from the files of frame `frame`; the format is taken from the .t header when
        not given (solution.py:301-378)."""
        path = os.path.expandvars(os.path.expanduser(path))
        if file_format is None:
            from .fileio.ascii import read_t
            file_format = read_t(frame, path, file_prefix)[6] or 'ascii'
        self.get_read_func(file_format)(self, frame, path, file_prefix=file_prefix, read_aux=read_aux, options=options)
        logging.getLogger('pyclaw.fileio').info('Read in solution for time t=%s' % self.t)

    def __str__(self):
        return 'states:\n' + ''.join(str(s) for s in self.states)
