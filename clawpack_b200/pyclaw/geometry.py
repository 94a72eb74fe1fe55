"""Dimension / Grid / Patch / Domain with the public surface of pyclaw/src/pyclaw/geometry.py
that the classic solvers, the examples and the controller use (uniform grids, optional mapc2p)."""
import numpy as np


class Dimension(object):
    """One coordinate direction (geometry.py:476-645): ``lower, upper, num_cells, name``;
    ``delta``, ``centers`` (lower + (i+0.5)*delta, :546-553), ``nodes`` (lower + i*delta, :538-545)."""

    def __init__(self, lower, upper, num_cells, name='x', on_lower_boundary=None, on_upper_boundary=None,
                 units=None):
        if isinstance(lower, str):
            raise Exception('Passing dimension name as first argument is deprecated. '
                            'Pass it as a keyword argument instead.')
        self._lower = float(lower)
        self._upper = float(upper)
        self._num_cells = int(num_cells)
        self.name = name
        self.on_lower_boundary = on_lower_boundary
        self.on_upper_boundary = on_upper_boundary
        self.units = units
        self._centers = None
        self._nodes = None
        self._check_validity()

    def _check_validity(self):
        assert isinstance(self.num_cells, int), 'Dimension.num_cells must be an integer; got %s' % type(self.num_cells)
        assert isinstance(self.lower, float), 'Dimension.lower must be a float'
        assert isinstance(self.upper, float), 'Dimension.upper must be a float'
        assert self.num_cells > 0, 'Dimension.num_cells must be positive'
        assert self.upper > self.lower, 'Dimension.upper must be greater than lower'

    def _reset(self):
        self._centers = None
        self._nodes = None
        self._check_validity()

    lower = property(lambda self: self._lower)
    upper = property(lambda self: self._upper)
    num_cells = property(lambda self: self._num_cells)

    @lower.setter
    def lower(self, v):
        self._lower = float(v)
        self._reset()

    @upper.setter
    def upper(self, v):
        self._upper = float(v)
        self._reset()

    @num_cells.setter
    def num_cells(self, v):
        self._num_cells = int(v)
        self._reset()

    @property
    def delta(self):
        return (self.upper - self.lower) / float(self.num_cells)

    @property
    def nodes(self):
        if self._nodes is None:
            self._nodes = self.lower + np.arange(self.num_cells + 1) * self.delta
        return self._nodes

    @property
    def centers(self):
        if self._centers is None:
            self._centers = self.lower + (np.arange(self.num_cells) + 0.5) * self.delta
        return self._centers

    edges = nodes

    def centers_with_ghost(self, num_ghost):
        pre = self.lower + (np.arange(-num_ghost, 0) + 0.5) * self.delta
        post = self.upper + self.delta * (np.arange(num_ghost) + 0.5)
        return np.hstack((pre, self.centers, post))

    def nodes_with_ghost(self, num_ghost):
        pre = np.linspace(self.lower - num_ghost * self.delta, self.lower - self.delta, num_ghost)
        post = np.linspace(self.upper + self.delta, self.upper + num_ghost * self.delta, num_ghost)
        return np.hstack((pre, self.nodes, post))

    def __len__(self):
        return self.num_cells

    def __str__(self):
        s = 'Dimension %s' % self.name
        if self.units:
            s += ' (%s)' % self.units
        return s + ':  (num_cells,delta,[lower,upper]) = (%s,%s,[%s,%s])' % (
            self.num_cells, self.delta, self.lower, self.upper)


def _identity_map(*coords):
    return list(coords)


class _DimensionHolder(object):
    """Shared by Grid and Patch: named Dimension attributes + per-dimension attribute lists."""

    def _init_dims(self, dimensions):
        if isinstance(dimensions, Dimension):
            dimensions = [dimensions]
        self._dimensions = []
        for dim in dimensions:
            self.add_dimension(dim)

    def add_dimension(self, dimension):
        if dimension.name in self._dimensions:
            raise Exception('Unable to add dimension. A dimension of the same name: {name}, already exists.'
                            .format(name=dimension.name))
        self._dimensions.append(dimension.name)
        setattr(self, dimension.name, dimension)

    def get_dim_attribute(self, attr):
        return [getattr(getattr(self, name), attr) for name in self._dimensions]

    @property
    def num_dim(self):
        return len(self._dimensions)

    @property
    def dimensions(self):
        return [getattr(self, name) for name in self._dimensions]


class Grid(_DimensionHolder):
    """Cell-centre / node coordinates of a logically rectangular grid (geometry.py:27-474)."""

    _dim_attrs = ('num_cells', 'lower', 'upper', 'delta', 'units', 'centers', 'nodes', 'on_lower_boundary',
                  'on_upper_boundary')

    def __init__(self, dimensions):
        self.gauges = []
        self.gauge_file_names = []
        self.gauge_files = []
        self.gauge_dir_name = '_gauges'
        self._mapc2p = None
        self._init_dims(dimensions)

    def __getattr__(self, key):
        if key in Grid._dim_attrs:
            return self.get_dim_attribute(key)
        raise AttributeError("'Grid' object has no attribute '" + key + "'")

    def add_gauges(self, gauge_coords):
        """Cell indices of the gauges that lie on this grid (geometry.py:392-405)."""
        for gauge in gauge_coords:
            if all(self.lower[n] <= gauge[n] < self.upper[n] for n in range(self.num_dim)):
                self.gauges.append([int((gauge[n] - self.lower[n]) // self.delta[n]) for n in range(self.num_dim)])
                self.gauge_file_names.append('gauge' + '_'.join(str(coord) for coord in gauge) + '.txt')

    def setup_gauge_files(self, outdir):
        """Open one text file per gauge under outdir/_gauges (geometry.py:407-423)."""
        import os
        gauge_path = os.path.join(outdir, self.gauge_dir_name)
        os.makedirs(gauge_path, exist_ok=True)
        for name in self.gauge_file_names:
            path = os.path.join(gauge_path, name)
            if os.path.isfile(path):
                os.remove(path)
            self.gauge_files.append(open(path, 'a'))

    @property
    def mapc2p(self):
        return self._mapc2p if self._mapc2p is not None else _identity_map

    @mapc2p.setter
    def mapc2p(self, f):
        self._mapc2p = f

    @property
    def c_centers(self):
        index = np.indices(self.num_cells)
        return [c[index[i, ...]] for i, c in enumerate(self.get_dim_attribute('centers'))]

    @property
    def c_nodes(self):
        index = np.indices([n + 1 for n in self.num_cells])
        return [c[index[i, ...]] for i, c in enumerate(self.get_dim_attribute('nodes'))]

    @property
    def p_centers(self):
        return self.mapc2p(*self.c_centers)

    @property
    def p_nodes(self):
        return self.mapc2p(*self.c_nodes)

    c_edges = c_nodes
    p_edges = p_nodes

    def c_centers_with_ghost(self, num_ghost):
        cs = [d.centers_with_ghost(num_ghost) for d in self.dimensions]
        index = np.indices([n + 2 * num_ghost for n in self.num_cells])
        return [c[index[i, ...]] for i, c in enumerate(cs)]

    def p_centers_with_ghost(self, num_ghost):
        return self.mapc2p(*self.c_centers_with_ghost(num_ghost))

    def c_nodes_with_ghost(self, num_ghost):
        ns = [d.nodes_with_ghost(num_ghost) for d in self.dimensions]
        index = np.indices([n + 2 * num_ghost + 1 for n in self.num_cells])
        return [c[index[i, ...]] for i, c in enumerate(ns)]

    def p_nodes_with_ghost(self, num_ghost):
        return self.mapc2p(*self.c_nodes_with_ghost(num_ghost))

    def __str__(self):
        return '%s-dimensional domain (%s)\n' % (self.num_dim, ','.join(self._dimensions)) + \
            '\n'.join(str(d) for d in self.dimensions)


class Patch(_DimensionHolder):
    """A rectangular region with one Grid (geometry.py:647-742)."""

    def __init__(self, dimensions):
        self.level = 1
        self.patch_index = 1
        if isinstance(dimensions, Dimension):
            dimensions = [dimensions]
        self._dimensions = []
        for dim in dimensions:
            dim.on_lower_boundary = True
            dim.on_upper_boundary = True
            self.add_dimension(dim)
        self.grid = Grid(dimensions)

    num_cells_global = property(lambda self: self.get_dim_attribute('num_cells'))
    lower_global = property(lambda self: self.get_dim_attribute('lower'))
    upper_global = property(lambda self: self.get_dim_attribute('upper'))
    delta = property(lambda self: self.get_dim_attribute('delta'))
    name = property(lambda self: self._dimensions)

    def __deepcopy__(self, memo={}):
        import copy
        result = self.__class__(copy.deepcopy(self.dimensions))
        result.grid.mapc2p = self.grid._mapc2p
        result.level = self.level
        result.patch_index = self.patch_index
        return result

    def __str__(self):
        return 'Patch %s:\n' % self.patch_index + '\n'.join(str(d) for d in self.dimensions)


class Domain(object):
    """A list of patches; ``Domain(lower, upper, num_cells)``, ``Domain(dims)`` or ``Domain(patches)``
    (geometry.py:744-815)."""

    def __init__(self, *arg):
        if len(arg) > 1:
            lower, upper, n = arg[0], arg[1], arg[2]
            names = ['x', 'y', 'z']
            dims = [Dimension(lo, up, nn, name=nm) for lo, up, nn, nm in zip(lower, upper, n, names)]
            self.patches = [self._patch_class()(dims)]
        else:
            geom = arg[0]
            if not isinstance(geom, (list, tuple)):
                geom = [geom]
            if isinstance(geom[0], Patch):
                self.patches = list(geom)
            elif isinstance(geom[0], Dimension):
                self.patches = [self._patch_class()(list(geom))]
            else:
                raise Exception('Invalid arguments for Domain initialization.')

    @staticmethod
    def _patch_class():
        return Patch

    num_dim = property(lambda self: self.patches[0].num_dim)
    patch = property(lambda self: self.patches[0])
    grid = property(lambda self: self.patches[0].grid)

    def __deepcopy__(self, memo={}):
        import copy
        return self.__class__(copy.deepcopy(self.patches))
