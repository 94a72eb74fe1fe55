"""``clawpack`` import namespace of the B200 build.

The reference's examples and user scripts start with ``from clawpack import pyclaw, riemann`` or
``import clawpack.petclaw as pyclaw`` (reference ``clawpack/__init__.py:3-24`` stitches its subpackages into one
namespace the same way).  This package makes those imports resolve to ``clawpack_b200.pyclaw`` / ``.riemann`` /
``.petclaw`` -- the SAME module objects, no copies -- so that a reference script runs unmodified on the CUDA path
(``kernel_language`` defaults to ``'CUDA'`` here).  Anything the build does not contain (sharpclaw, visclaw, amrclaw,
geoclaw, forestclaw ...) raises ImportError with a message saying so; nothing falls back to a CPU implementation.

If the real Clawpack is installed it wins over this directory only when it comes first on ``sys.path``.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import clawpack_b200 as _impl

__version__ = '5.14.0+b200.' + _impl.__version__
_SUBPACKAGES = ('pyclaw', 'riemann', 'petclaw')
__all__ = list(_SUBPACKAGES)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        mod = importlib.import_module(self.target)       # the clawpack_b200 module itself
        self.spec = mod.__spec__
        return mod

    def exec_module(self, module):
        module.__spec__ = self.spec                      # already initialised under its own name: keep its own spec


class _AliasFinder(importlib.abc.MetaPathFinder):
    """``clawpack.<sub>[.rest]`` -> ``clawpack_b200.<sub>[.rest]``."""

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split('.')
        if len(parts) < 2 or parts[0] != 'clawpack':
            return None
        if parts[1] not in _SUBPACKAGES:
            raise ImportError("clawpack.%s is not part of the B200 build (it provides clawpack.pyclaw, clawpack.riemann "
                              "and clawpack.petclaw for the classic solvers)" % parts[1])
        real = 'clawpack_b200.' + '.'.join(parts[1:])
        try:
            mod = importlib.import_module(real)
        except ImportError as e:
            raise ImportError('%s is not part of the B200 build (%s)' % (fullname, e))
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=hasattr(mod, '__path__'))
        return spec


def _install():
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())
    for sub in ('pyclaw', 'riemann'):
        mod = importlib.import_module('clawpack_b200.' + sub)
        sys.modules['clawpack.' + sub] = mod
        globals()[sub] = mod


_install()


def __getattr__(name):
    if name == 'petclaw':                               # imports torch.distributed helpers: load on first use
        mod = importlib.import_module('clawpack_b200.petclaw')
        sys.modules['clawpack.petclaw'] = mod
        return mod
    raise AttributeError("module 'clawpack' has no attribute %r" % name)
