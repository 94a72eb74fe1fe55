"""Controller of the distributed runs (pyclaw/src/petclaw/controller.py): same loop on every rank."""
from ..pyclaw.controller import Controller as _Controller
from . import util


class Controller(_Controller):
    def __init__(self):
        super(Controller, self).__init__()
        self.rank = util.rank()
