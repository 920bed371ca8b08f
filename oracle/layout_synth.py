"""Seeded synthetic room layouts for the layout-modality parity tests (TEST INFRASTRUCTURE).

A layout = a star-shaped room polygon around the origin (ZInD-normalised units, as PanoData.room_vertices_local_2d holds them)
and a few windows / doors / openings lying on its walls.  `nodes(...)` returns plain objects shaped like the reference's
PoseGraph2d.nodes[i] (room_vertices_local_2d, doors, windows, openings); the W/D/O items are the reference's own WDO class when
`wdo_cls` is given (scripts/make_golden_layout.py), else a small stand-in with the same fields.
"""

from __future__ import annotations

from types import SimpleNamespace

import numpy as np


class PlainWDO:
    def __init__(self, pt1, pt2, type):
        self.pt1, self.pt2, self.type = tuple(pt1), tuple(pt2), type

    @property
    def vertices_local_2d(self) -> np.ndarray:
        return np.array([self.pt1, self.pt2])


def synth_room(seed: int, radius: float = 1.6):
    """(vertices (n, 2) float64, [(type, pt1, pt2), ...])."""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(5, 12))
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = radius * rng.uniform(0.55, 1.0, n)
    v = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    wdos = []
    for k, typ in enumerate(("doors", "windows", "openings", "windows")):
        e = int(rng.integers(0, n))
        a, b = v[e], v[(e + 1) % n]
        t0 = rng.uniform(0.1, 0.5)
        t1 = t0 + rng.uniform(0.2, 0.4)
        wdos.append((typ, tuple(a + t0 * (b - a)), tuple(a + t1 * (b - a))))
    return v, wdos


def nodes(seeds, wdo_cls=None, sim2_cls=None):
    out = {}
    for i, sd in enumerate(seeds):
        v, wd = synth_room(sd)
        items = {"doors": [], "windows": [], "openings": []}
        for typ, p1, p2 in wd:
            if wdo_cls is None:
                items[typ].append(PlainWDO(p1, p2, typ))
            else:
                items[typ].append(wdo_cls(global_Sim2_local=sim2_cls(np.eye(2), np.zeros(2), 1.0), pt1=p1, pt2=p2, bottom_z=-1.0, top_z=1.0, type=typ))
        out[i] = SimpleNamespace(room_vertices_local_2d=v, **items)
    return SimpleNamespace(nodes=out)
