"""nls_b200 -- B200-native time-stepping engine behind the ``daskol/nls`` API.

``from nls_b200 import Problem`` is the drop-in for ``from nls import Problem``
(``Problem().model(...).solve().report()``, ref ``README.md:47-57``); ``nls_b200.native.nls`` is the
drop-in for the f2py module ``nls.native.nls``; ``nls_b200.engine`` is the device-resident API
(torch tensors on the GPU, ensembles, slabs) on top of the same C ABI.
"""

from .version import version
from .pumping import *        # noqa: F401,F403
from .model import *          # noqa: F401,F403

__all__ = ["model", "native", "pumping", "solver", "version", "engine"]
