"""Toy geometries (reference: megastep/toys.py:5-29)."""
import numpy as np

from . import geometry
from .arrdict import arrdict


def _square(width, centre):
    angles = np.arange(np.pi / 4, 2 * np.pi, np.pi / 2)
    return width / 2 ** .5 * np.stack([np.cos(angles), np.sin(angles)], -1) + centre


def box(width=5):
    """One square room of side `width` with a single light in the middle."""
    centre = width / 2 + geometry.MARGIN
    corners = _square(width, centre)
    walls = np.stack(geometry.cyclic_pairs(corners))
    return arrdict(walls=walls, lights=np.full((1, 2), centre), masks=geometry.masks(walls, [corners]), res=geometry.RES)


def column(width=5, column_width=.1):
    """A small square column in the middle of an (unwalled) room, lit from four corners."""
    centre = width / 2 + geometry.MARGIN
    walls = np.stack(geometry.cyclic_pairs(_square(column_width, centre)))
    room = _square(width, centre)
    return arrdict(walls=walls, lights=_square(2., centre), masks=geometry.masks(walls, [room]), res=geometry.RES)
