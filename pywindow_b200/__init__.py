"""pywindow_b200 -- B200 (sm_100a) implementation of pywindow's per-frame pore-geometry path.

Drop-in for the reference's public names on that path (``pywindow/__init__.py:3-19``); see DESIGN.md.
"""
__all__ = ["Engine", "default_engine"]


def __getattr__(name):
    if name in ("Engine", "default_engine"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
