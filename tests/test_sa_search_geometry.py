"""CPU property test of the row trimming in the fused SA kernels (csrc/sa_fused.cu, round 2): a numpy float32
restatement of the grid arithmetic (csrc/ball_query.cu grid_build_kernel + cell_coord) and of the per-row chord
trimming, checked for CONSERVATIVENESS -- every point the kernel's own fp32 distance test accepts (d2 < r2, operations
rounded separately as in sn2::dist2) must lie in a row the trimming keeps and inside the kept x-cell range of that row.
The GPU parity tests establish that on the benchmark clouds; this test hammers the margins with adversarial inputs:
points a few ulps inside the radius, centroids on cell faces, coordinate offsets up to 1e4 (ulp 1e-3), flat clouds."""
import numpy as np
import pytest

f32 = np.float32
GRID_MAX, GRID_CELLS = 64, 4096


def _cell(v, mn, inv, g):
    c = np.floor((v.astype(f32) - mn).astype(f32) * inv).astype(np.int64)
    return np.clip(c, 0, g - 1)


def _grid(pts, r):
    """grid_build_kernel's header: origin, 1 / cell edge, dims (xy edge shared, z layers as many as fit)."""
    mn, mx = pts.min(0), pts.max(0)
    ext = max(f32(mx[0] - mn[0]), f32(mx[1] - mn[1]))
    cs = max(f32(r * f32(1.0001)), f32(ext / f32(GRID_MAX - 1)), f32(1e-20))
    inv = f32(1.0) / cs
    gx = min(GRID_MAX, int(np.floor(f32(mx[0] - mn[0]) * inv)) + 1)
    gy = min(GRID_MAX, int(np.floor(f32(mx[1] - mn[1]) * inv)) + 1)
    gz, invz, csz = 1, f32(0.0), f32(np.inf)
    zext = f32(mx[2] - mn[2])
    gzmax = GRID_CELLS // (gx * gy)
    if gzmax > 1 and zext > 0:
        need = int(np.floor(zext / cs)) + 1
        if need <= gzmax:
            gz, csz = need, cs
        else:
            gz, csz = gzmax, f32(zext / f32(gz + 0.0 if False else gzmax - 0.5))
        invz = f32(1.0) / csz
        gz = min(gz, int(np.floor(zext * invz)) + 1)
    return dict(ox=mn[0], oy=mn[1], oz=mn[2], inv=inv, invz=invz, cs=cs, csz=csz, gx=gx, gy=gy, gz=gz)


def _dist2(p, q):
    d = (p - q).astype(f32)
    return ((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32) + (d[:, 2] * d[:, 2]).astype(f32)


def _check(pts, queries, r, shrink=1.0):
    pts, queries = pts.astype(f32), queries.astype(f32)
    g = _grid(pts, f32(r))
    r2 = f32(np.float64(f32(r)) ** 2)
    cx = _cell(pts[:, 0], g["ox"], g["inv"], g["gx"])
    cy = _cell(pts[:, 1], g["oy"], g["inv"], g["gy"])
    cz = _cell(pts[:, 2], g["oz"], g["invz"], g["gz"])
    cs, csz = g["cs"], g["csz"]
    bad = 0
    for q in queries:
        hit = _dist2(pts, q[None, :]) < r2
        if not hit.any():
            continue
        ix = int(_cell(q[0:1], g["ox"], g["inv"], g["gx"])[0])
        iy = int(_cell(q[1:2], g["oy"], g["inv"], g["gy"])[0])
        iz = int(_cell(q[2:3], g["oz"], g["invz"], g["gz"])[0])
        x0, x1 = max(ix - 1, 0), min(ix + 1, g["gx"] - 1)
        ey = f32(1e-4) * cs + f32(1e-6) * (abs(q[1]) + abs(g["oy"]) + cs * g["gy"])
        ez = (f32(1e-4) * csz + f32(1e-6) * (abs(q[2]) + abs(g["oz"]) + csz * g["gz"])) if np.isfinite(csz) else f32(0)
        ex = f32(1e-4) * cs + f32(1e-6) * (abs(q[0]) + abs(g["ox"]) + cs * g["gx"])
        for k in np.nonzero(hit)[0]:
            ry, rz = int(cy[k]), int(cz[k])
            # the 3 x 3 x 3 block itself (the assumption the kernels already made before the trimming)
            assert abs(ry - iy) <= 1 and abs(rz - iz) <= 1 and x0 <= cx[k] <= x1
            dy = (g["oy"] + f32(ry) * cs) - q[1] if ry > iy else (q[1] - (g["oy"] + f32(ry + 1) * cs) if ry < iy else f32(0))
            if np.isfinite(csz):
                dz = (g["oz"] + f32(rz) * csz) - q[2] if rz > iz else (q[2] - (g["oz"] + f32(rz + 1) * csz) if rz < iz else f32(0))
            else:
                dz = f32(0)
            dy, dz = max(f32(dy - ey), f32(0)), max(f32(dz - ez), f32(0))
            rem = f32(r2 * f32(1.00001)) - f32(dy * dy) - f32(dz * dz)
            if not rem > 0:
                bad += 1
                continue
            w = (f32(np.sqrt(rem)) * f32(1.00001) + ex) * f32(shrink)
            xa = max(x0, int(_cell(np.array([q[0] - w], f32), g["ox"], g["inv"], g["gx"])[0]))
            xb = min(x1, int(_cell(np.array([q[0] + w], f32), g["ox"], g["inv"], g["gx"])[0]))
            if not xa <= cx[k] <= xb:
                bad += 1
    return bad


@pytest.mark.parametrize("offset", [0.0, 37.5, 1.0e3, 1.0e4])
@pytest.mark.parametrize("r", [np.sqrt(2.0), np.sqrt(8.0), 0.3])
def test_row_trimming_never_drops_an_in_radius_point(offset, r):
    rng = np.random.default_rng(int(offset) + int(r * 100))
    n = 600
    base = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), np.abs(rng.normal(0, 4, n))], 1)
    q = base[rng.choice(n, 40, replace=False)].copy()
    extra = []
    for c in q:  # points a hair inside / outside the sphere in random directions, and exactly axis-aligned
        d = rng.normal(size=(6, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        for s in (1 - 1e-7, 1 - 1e-6, 1 - 1e-4, 1 + 1e-6):
            extra.append(c + d * r * s)
        for ax in range(3):
            for sgn in (-1, 1):
                e = np.zeros(3)
                e[ax] = sgn * r * (1 - 3e-7)
                extra.append((c + e)[None, :])
    pts = np.concatenate([base] + extra, 0) + offset
    assert _check(pts, q + offset, r) == 0


def test_row_trimming_centroids_on_cell_faces_and_flat_clouds():
    rng = np.random.default_rng(9)
    r = np.sqrt(2.0)
    # a flat cloud (z extent 0 -> one layer, invz = 0) and centroids snapped onto the cell faces of the grid
    pts = np.stack([rng.uniform(0, 20, 1500), rng.uniform(0, 20, 1500), np.zeros(1500)], 1).astype(f32)
    g = _grid(pts, f32(r))
    k = rng.integers(1, 10, (50, 2))
    q = np.stack([g["ox"] + k[:, 0] * g["cs"], g["oy"] + k[:, 1] * g["cs"], np.zeros(50)], 1)
    assert _check(np.concatenate([pts, q.astype(f32)], 0), q, r) == 0
    # tall thin cloud: z layers thicker than the xy cell (cell budget exhausted)
    pts = np.stack([rng.uniform(0, 80, 3000), rng.uniform(0, 80, 3000), rng.uniform(0, 60, 3000)], 1)
    q = pts[rng.choice(3000, 60, replace=False)]
    assert _check(pts, q, r) == 0


def test_the_property_test_has_teeth():
    """A trimming whose chord is 30 % too short must be caught (points just inside the radius fall out of the kept cells)."""
    rng = np.random.default_rng(1)
    r = np.sqrt(2.0)
    pts = np.stack([rng.uniform(-10, 10, 4000), rng.uniform(-10, 10, 4000), np.abs(rng.normal(0, 4, 4000))], 1)
    q = pts[rng.choice(4000, 300, replace=False)]
    assert _check(pts, q, r) == 0
    assert _check(pts, q, r, shrink=0.7) > 0
