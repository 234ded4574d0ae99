"""Seeded synthetic scenes for the physics step: procedural heightfield terrain plus
sphere bodies. Host-side numpy only; every array is float32/int32 so the oracle, the compiled
reference and the CUDA path all see the very same bits.

Terrain triangulation follows the reference's index pattern (procedural/src/ProcIsland.cpp:56-68):
vertices row-major with z outer / x inner, and per cell `(b+stride, b+1, b)` then
`(b+stride, b+1+stride, b+1)`, which gives every triangle an upward (y>0) normal
(asserted at ProcIsland.cpp:81-82). Heights are a bounded smooth value noise (our own lattice
noise, same idea as procedural/src/Noise.cpp:7-39) with amplitude <= 0.75*cell.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

WORLD_HALF = 1000.0  # physics/src/Physics.cpp:7-9 -> OSP(2000, 500)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _value_noise(nx: int, nz: int, seed: int, octaves=((16.0, 1.0), (5.0, 0.35), (2.0, 0.12))) -> np.ndarray:
    """Smooth noise in [0,1] sampled on an (nz, nx) vertex lattice."""
    rng = _rng(seed)
    out = np.zeros((nz, nx), dtype=np.float64)
    tot = 0.0
    xs = np.arange(nx, dtype=np.float64)
    zs = np.arange(nz, dtype=np.float64)
    for period, amp in octaves:
        lx = int(nx / period) + 3
        lz = int(nz / period) + 3
        lat = rng.random((lz, lx))
        fx = xs / period
        fz = zs / period
        ix = np.floor(fx).astype(np.int64)
        iz = np.floor(fz).astype(np.int64)
        tx = fx - ix
        tz = fz - iz
        tx = tx * tx * (3.0 - 2.0 * tx)
        tz = tz * tz * (3.0 - 2.0 * tz)
        a = lat[np.ix_(iz, ix)]
        b = lat[np.ix_(iz, ix + 1)]
        c = lat[np.ix_(iz + 1, ix)]
        d = lat[np.ix_(iz + 1, ix + 1)]
        top = a + (b - a) * tx[None, :]
        bot = c + (d - c) * tx[None, :]
        out += amp * (top + (bot - top) * tz[:, None])
        tot += amp
    return out / tot


def terrain_heights(G: int, c: float, seed: int, amp: float | None = None) -> np.ndarray:
    """(G+1, G+1) float32 vertex heights, index [iz, ix]."""
    if amp is None:
        amp = 0.75 * c
    h = _value_noise(G + 1, G + 1, seed) * amp
    return h.astype(np.float32)


def terrain_triangles(G: int, c: float, seed: int, amp: float | None = None):
    """Return (tris[2*G*G, 9] float32, heights[(G+1),(G+1)] float32). Grid centred on the origin."""
    h = terrain_heights(G, c, seed, amp)
    half = np.float32(G * c / 2.0)
    coords = (np.arange(G + 1, dtype=np.float32) * np.float32(c) - half).astype(np.float32)
    X, Z = np.meshgrid(coords, coords)  # [iz, ix]
    V = np.stack([X, h, Z], axis=-1).astype(np.float32)  # (G+1, G+1, 3)
    b = V[:-1, :-1]      # b
    b1 = V[:-1, 1:]      # b + 1
    bs = V[1:, :-1]      # b + stride
    bs1 = V[1:, 1:]      # b + 1 + stride
    t1 = np.concatenate([bs, b1, b], axis=-1)     # (G, G, 9)
    t2 = np.concatenate([bs, bs1, b1], axis=-1)
    tris = np.stack([t1, t2], axis=2).reshape(-1, 9)
    return np.ascontiguousarray(tris, dtype=np.float32), h


def box_triangles(center, size) -> np.ndarray:
    """12 triangles of an axis-aligned box ("house", cf. procedural/include/WorldGenerator.h:28-36)."""
    cx, cy, cz = center
    sx, sy, sz = (s / 2.0 for s in size)
    p = np.array([[cx - sx, cy - sy, cz - sz], [cx + sx, cy - sy, cz - sz], [cx + sx, cy + sy, cz - sz],
                  [cx - sx, cy + sy, cz - sz], [cx - sx, cy - sy, cz + sz], [cx + sx, cy - sy, cz + sz],
                  [cx + sx, cy + sy, cz + sz], [cx - sx, cy + sy, cz + sz]], dtype=np.float32)
    quads = [(0, 1, 2, 3), (5, 4, 7, 6), (4, 0, 3, 7), (1, 5, 6, 2), (3, 2, 6, 7), (4, 5, 1, 0)]
    out = []
    for a, b, c, d in quads:
        out.append(np.concatenate([p[a], p[b], p[c]]))
        out.append(np.concatenate([p[a], p[c], p[d]]))
    return np.stack(out).astype(np.float32)


def _height_upper(h: np.ndarray, G: int, c: float, x: np.ndarray, z: np.ndarray) -> np.ndarray:
    """Upper bound of the terrain height under (x, z): max of the 4 cell-corner heights."""
    half = G * c / 2.0
    ix = np.clip(np.floor((x + half) / c).astype(np.int64), 0, G - 1)
    iz = np.clip(np.floor((z + half) / c).astype(np.int64), 0, G - 1)
    return np.maximum(np.maximum(h[iz, ix], h[iz, ix + 1]), np.maximum(h[iz + 1, ix], h[iz + 1, ix + 1]))


@dataclass
class Scene:
    """A physics scene in the insertion order every implementation uses:
    static geometry first (geom ids 0..), then bodies (body ids 0..)."""
    name: str
    tris: list = field(default_factory=list)          # list of (ntri, 9) float32 arrays, one per static geometry
    body_pos: np.ndarray = None                        # (nb, 3) float32 linear positions
    body_vel: np.ndarray = None                        # (nb, 3)
    body_active: np.ndarray = None                     # (nb,) int32
    body_gravity: np.ndarray = None                    # (nb,) int32
    body_scale: np.ndarray = None                      # (nb,) float32 uniform scale of W (1 = identity)
    sphere_start: np.ndarray = None                    # (nb+1,) int32 CSR into obj_spheres
    obj_spheres: np.ndarray = None                     # (ns, 4) float32 object-space centre + radius
    dt_ms: int = 5
    meta: dict = field(default_factory=dict)

    @property
    def n_bodies(self) -> int:
        return int(self.body_pos.shape[0])

    @property
    def n_spheres(self) -> int:
        return int(self.obj_spheres.shape[0])

    @property
    def n_tris(self) -> int:
        return int(sum(t.shape[0] for t in self.tris))


def _single_sphere_bodies(pos, vel, r, gravity=1, active=1) -> dict:
    n = pos.shape[0]
    obj = np.zeros((n, 4), dtype=np.float32)
    obj[:, 3] = np.float32(r) if np.isscalar(r) else r.astype(np.float32)
    return dict(body_pos=pos.astype(np.float32), body_vel=vel.astype(np.float32),
                body_active=np.full(n, active, dtype=np.int32), body_gravity=np.full(n, gravity, dtype=np.int32),
                body_scale=np.ones(n, dtype=np.float32), sphere_start=np.arange(n + 1, dtype=np.int32),
                obj_spheres=obj)


def falling_spheres(n: int, G: int, c: float, r: float = 0.5, y_range=(5.0, 30.0), above_terrain: bool = False,
                    seed: int = 1234, name: str = "falling") -> Scene:
    """Config C1/C3 family: n single-sphere bodies dropped on a G x G-cell heightfield."""
    tris, h = terrain_triangles(G, c, seed)
    rng = _rng(seed + 1)
    half = G * c / 2.0 - 2.0
    x = rng.uniform(-half, half, n)
    z = rng.uniform(-half, half, n)
    y = rng.uniform(y_range[0], y_range[1], n)
    if above_terrain:
        y = y + _height_upper(h, G, c, x, z)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    vel = np.zeros_like(pos)
    return Scene(name=name, tris=[tris], meta=dict(G=G, c=c, r=r, seed=seed), **_single_sphere_bodies(pos, vel, r))


def config1(seed: int = 1234) -> Scene:
    """C1: 1,000 spheres r=0.5 falling on a 64x64-cell (8,192-triangle) heightfield, c=2."""
    return falling_spheres(1000, 64, 2.0, 0.5, (5.0, 30.0), False, seed, "C1")


def config2(n: int = 100_000, seed: int = 1234, fill: float = 0.03) -> Scene:
    """C2: n spheres r=0.5 uniform in a box at ~3% volume fraction, |v|<=5, gravity off, no triangles."""
    r = 0.5
    edge = (n * (4.0 / 3.0) * np.pi * r ** 3 / fill) ** (1.0 / 3.0)
    rng = _rng(seed)
    pos = rng.uniform(-edge / 2, edge / 2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    vel = (d * rng.uniform(0, 5.0, (n, 1))).astype(np.float32)
    return Scene(name="C2", tris=[], meta=dict(edge=edge, r=r, seed=seed),
                 **_single_sphere_bodies(pos, vel, r, gravity=0))


def config3(n: int = 1_000_000, G: int = 1000, seed: int = 1234, y_range=(2.0, 12.0)) -> Scene:
    """C3: n spheres r=0.5 over a GxG-cell (2*G*G triangles) terrain, c=2, y in [2,12] above local height."""
    return falling_spheres(n, G, 2.0, 0.5, y_range, True, seed, "C3")


def config4(n_bodies: int = 262_144, G: int = 1000, seed: int = 1234, r: float = 0.4) -> Scene:
    """C4: compound bodies, 8 collision spheres each on a 2x2x2 lattice of spacing r."""
    tris, h = terrain_triangles(G, 2.0, seed)
    rng = _rng(seed + 2)
    half = G * 2.0 / 2.0 - 3.0
    x = rng.uniform(-half, half, n_bodies)
    z = rng.uniform(-half, half, n_bodies)
    y = rng.uniform(2.0, 12.0, n_bodies) + _height_upper(h, G, 2.0, x, z)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    offs = np.array([[sx, sy, sz] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)], dtype=np.float32) * r
    obj = np.zeros((n_bodies, 8, 4), dtype=np.float32)
    obj[:, :, :3] = offs[None]
    obj[:, :, 3] = r
    return Scene(name="C4", tris=[tris], body_pos=pos, body_vel=np.zeros_like(pos),
                 body_active=np.ones(n_bodies, np.int32), body_gravity=np.ones(n_bodies, np.int32),
                 body_scale=np.ones(n_bodies, np.float32),
                 sphere_start=(np.arange(n_bodies + 1, dtype=np.int32) * 8), obj_spheres=obj.reshape(-1, 4),
                 meta=dict(G=G, c=2.0, r=r, seed=seed))


def config5(n: int = 16_000_000, G: int = 2000, seed: int = 1234, houses: bool = True, x_range=None) -> Scene:
    """C5: n spheres r=0.25 on a GxG-cell terrain of c=2000/G plus static box "houses" on a jittered
    200-unit tile grid. `x_range=(lo, hi)` keeps only the bodies of one x-slab (multi-GPU sharding)
    while static geometry stays replicated."""
    c = 2000.0 / G
    tris, h = terrain_triangles(G, c, seed)
    geoms = [tris]
    if houses:
        rng_h = _rng(seed + 7)
        boxes = []
        for tx in np.arange(-900.0, 901.0, 200.0):
            for tz in np.arange(-900.0, 901.0, 200.0):
                cx = tx + rng_h.uniform(-60, 60)
                cz = tz + rng_h.uniform(-60, 60)
                s = rng_h.uniform(5.0, 10.0, 3)
                base = float(_height_upper(h, G, c, np.array([cx]), np.array([cz]))[0])
                boxes.append(box_triangles((cx, base + s[1] / 2 - 0.5, cz), s))
        geoms.append(np.concatenate(boxes).astype(np.float32))
    rng = _rng(seed + 3)
    half = G * c / 2.0 - 2.0
    lo, hi = (-half, half) if x_range is None else (max(-half, x_range[0]), min(half, x_range[1]))
    x = rng.uniform(lo, hi, n)
    z = rng.uniform(-half, half, n)
    y = rng.uniform(1.0, 6.0, n) + _height_upper(h, G, c, x, z)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    return Scene(name="C5", tris=geoms, meta=dict(G=G, c=c, r=0.25, seed=seed),
                 **_single_sphere_bodies(pos, np.zeros_like(pos), 0.25))


def pile(n: int = 64, seed: int = 5, r: float = 0.5, gravity: int = 0) -> Scene:
    """Sphere-sphere ordering scene: n overlapping/nearby spheres in a small box, no triangles."""
    rng = _rng(seed)
    edge = (n ** (1.0 / 3.0)) * 2.2 * r
    pos = rng.uniform(-edge / 2, edge / 2, (n, 3)).astype(np.float32)
    vel = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    return Scene(name="pile", tris=[], meta=dict(r=r, seed=seed), **_single_sphere_bodies(pos, vel, r, gravity=gravity))
