import types

from . import signal  # noqa: F401

distributions = types.SimpleNamespace()
