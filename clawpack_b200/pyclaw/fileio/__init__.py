"""Frame file I/O of ``Solution.write`` / ``Solution.read`` (pyclaw/src/pyclaw/fileio/).  Only the classic
Clawpack ascii layout is built (fort.tNNNN, fort.qNNNN, fort.aNNNN, fort.pklNNNN); the hdf5 / netcdf / petsc
writers of the reference need libraries that are not part of this build."""
from . import ascii  # noqa: F401
