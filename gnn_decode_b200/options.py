"""Library tunables (csrc/gd_options.cuh): the planners read their switches from ONE table that is filled from the
environment once, at first use; these helpers change it at run time through the C ABI (gd_set_option)."""
import contextlib

from . import _cabi


def set_option(name, value=1):
    _cabi.check(_cabi.lib().gd_set_option(name.encode(), int(value), 0), "gd_set_option")


def unset_option(name):
    _cabi.check(_cabi.lib().gd_set_option(name.encode(), 0, 1), "gd_set_option")


@contextlib.contextmanager
def option(name, value=1):
    """with option("GD_NO_LEAN"): ...   (restores 'unset' on exit: options have no other default state)"""
    set_option(name, value)
    try:
        yield
    finally:
        unset_option(name)
