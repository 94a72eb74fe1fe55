"""Frame file I/O of ``Solution.write`` / ``Solution.read`` (pyclaw/src/pyclaw/fileio/): the classic Clawpack ascii
layout (fort.tNNNN, fort.qNNNN, fort.aNNNN, fort.pklNNNN) and the raw binary layout of fileio/binary.py (fort.bNNNN).
The hdf5 / netcdf / petsc / vtk writers of the reference need libraries that are not part of this build."""
from . import ascii  # noqa: F401
from . import binary  # noqa: F401
