from .simulation2d import Simulation2D, get_closest

__all__ = ["Simulation2D", "get_closest"]
