"""CPU oracle (test infrastructure only -- see oracle/oracle.py and oracle/oracle.c)."""
from .oracle import *  # noqa: F401,F403
from . import oracle as _o

build = _o.build
