"""asyncrl_b200: B200-native A3C worker hot path behind the reference's Python interfaces.

The directory name carries a hyphen, so import it with
    pkg = importlib.import_module("async-rl-tensorflow_b200")
(``tests/conftest.py`` and ``bench.py`` do exactly that).  Sub-modules mirror the
reference: ``config``, ``src.environment``, ``src.history``, ``src.ops``, ``src.network``,
``src.agent``.  All math runs in ``libasyncrl_b200.so`` (csrc/, C-ABI in include/asyncrl_b200.h).
"""
from . import _cabi, config                                   # noqa: F401
from .src import agent, base, environment, history, network, network_nature, ops   # noqa: F401
from .src.agent import Agent                                  # noqa: F401
from .src.environment import GymEnvironment, SimpleGymEnvironment, SyntheticAtari  # noqa: F401
from .src.history import History                              # noqa: F401
from .src.network import Network, make_network                # noqa: F401
from .src.network_nature import NatureNetwork                 # noqa: F401

__version__ = "0.1.0"
