"""Minimal `gym` stand-in (TEST INFRASTRUCTURE): the reference only subclasses
gym.Env and imports gym.spaces (environments/grid_world.py:2-5, main.py:3,6)."""
import types, sys


class Env(object):
    pass


spaces = types.ModuleType("gym.spaces")
sys.modules["gym.spaces"] = spaces
