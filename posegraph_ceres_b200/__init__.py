"""Importable alias of the `posegraph-ceres_b200/` package directory (the hyphen is not a legal
Python identifier): extends __path__ so that `posegraph_ceres_b200.api`, `.datasets`, ... resolve
to the files under posegraph-ceres_b200/."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "posegraph-ceres_b200"))

from .api import *  # noqa: E402,F401,F403
from . import datasets  # noqa: E402,F401
