"""Solution = states + domain (pyclaw/src/pyclaw/solution.py:17-253).  File I/O (write/read of
frames) is outside this build's scope: ``write`` supports the
reference's ascii layout only as far as the controller needs it and is otherwise a stub."""
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
        elif len(arg) == 0 and not [k for k in kargs if k != 'claw_package']:
            pass
        else:
            raise NotImplementedError('Reading a Solution from file is outside this build (frame I/O is host-side '
                                      'and out of scope); construct it from a State and a Domain.')

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

    def write(self, frame, path='./', file_format='ascii', file_prefix=None, write_aux=False, options={},
              write_p=False):
        raise NotImplementedError('Solution.write: frame file I/O is outside this build (set '
                                  'controller.output_format = None and use controller.frames / state.q).')

    def __str__(self):
        return 'states:\n' + ''.join(str(s) for s in self.states)
