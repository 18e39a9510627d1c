"""sde_mc_b200 -- B200-native drop-in for the hot path of piers-hinds/sde_mc.

Same flat, star-exported namespace as the reference package (/root/reference/sde_mc/__init__.py:1-11, including
the `torch`, `np`, `nn`, `time`, `optim` names its tests and examples rely on), so
`import sde_mc_b200 as sde_mc` or `from sde_mc_b200 import *` replaces `from sde_mc import *`.
"""
from .version import __version__
from .sde import *
from .mc import *
from .varred import *
from .options import *
from .nets import *
from .helpers import *
from .solvers import *
from .levy import *
from .problem import *
from .mlmc import *
