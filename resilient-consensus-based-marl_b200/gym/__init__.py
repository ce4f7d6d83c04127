"""`gym` names the reference imports (main.py:3,6; environments/grid_world.py:2-5): only the `Env` base class and
the `spaces` module are touched, never any gym functionality."""
import sys
import types


class Env(object):
    metadata = {}


spaces = types.ModuleType("gym.spaces")
sys.modules.setdefault("gym.spaces", spaces)
