"""CPU oracle: restated parcel tiling, plot preparation and band finalisation.  TEST INFRASTRUCTURE ONLY.

Follows the reference step by step, with the reference's own library calls where they are installed here
(scipy cKDTree.query_ball_point, sklearn NearestNeighbors.radius_neighbors):
  extract_cloud / extract_cloud_data     /root/reference/inference/prepare_utils.py:47-81  (+ prepare.py:91-94: > 50 points)
  normalize_z_with_minz_in_a_radius      /root/reference/utils/load_data.py:237-249
  center_cloud, add_fake_empty_ground_points, rescale_cloud, sample_cloud
                                         /root/reference/data_loader/loader.py:73-105, 127-158, 233-255
  insert_hard_med_veg_raster_band, finalize_merged_raster (without admissibility)
                                         /root/reference/inference/geotiff_raster.py:121-146, 273-291
dtype rules of the pinned NumPy 1.x (float32 array op float64 scalar -> float32) are written out explicitly so that
the result does not depend on the NumPy installed here.  np.random in sample_cloud is replaced by the canonical
counter-based hash shared with csrc/parcel.cu (S smallest hashes, order kept; up-sampling picks hash % n); points of a
plot are taken in ascending parcel index (the kd-tree order is arbitrary).  Parity with the real pipeline on real LAS
data is unpinned (laspy / shapefiles absent)."""
import numpy as np

MIN_N_POINTS_FOR_INFERENCE = 50


def sample_hash(seed, j):
    """uint32 hash of csrc/parcel.cu::sample_hash, vectorised over j."""
    with np.errstate(over="ignore"):
        h = (np.uint32(seed) * np.uint32(0x9E3779B1) + np.asarray(j, dtype=np.uint32) * np.uint32(0x85EBCA77) + np.uint32(0x165667B1)).astype(np.uint32)
        h ^= h >> np.uint32(15)
        h = (h * np.uint32(0x2C1B3C6D)).astype(np.uint32)
        h ^= h >> np.uint32(12)
        h = (h * np.uint32(0x297A2D39)).astype(np.uint32)
        h ^= h >> np.uint32(15)
    return h


def fake_ground_points(diam_meters, n_feats=10):
    """data_loader/loader.py:90-105 -> float32 [n_feats, 316]."""
    c = np.arange(-diam_meters // 2, diam_meters // 2, 1) + 0.5
    xx, yy = np.meshgrid(c, c, sparse=True)
    x, y = (xx + 0 * yy).flatten(), (yy + 0 * xx).flatten()
    keep = np.sqrt(x ** 2 + y ** 2) < diam_meters // 2
    pts = np.zeros((n_feats, int(keep.sum())), dtype=np.float32)
    pts[0], pts[1] = x[keep], y[keep]
    return pts


def prepare_plot(parcel_cloud, tree, center, seed, args):
    """One plot: -> None (too few points) or dict(xyz (3,S) f32, cloud (10,S) f32, n, src (S,) int)."""
    from sklearn.neighbors import NearestNeighbors

    f32 = np.float32
    idx = np.sort(np.asarray(tree.query_ball_point(center, r=args.diam_meters // 2), dtype=np.int64))  # canonical: ascending index
    n = idx.size
    if n <= MIN_N_POINTS_FOR_INFERENCE:  # < 50 -> None (prepare_utils.py:67-69); prepare.py:93 keeps only > 50
        return None, n
    cloud = parcel_cloud[:, idx].astype(f32).copy()
    # pre_transform: z - min z of the neighbours within the radius (load_data.py:237-249)
    xy = cloud[:2].T
    nn = NearestNeighbors(n_neighbors=500, algorithm="kd_tree").fit(xy)
    _, neigh = nn.radius_neighbors(xy, args.znorm_radius_in_meters)
    z = cloud[2]
    zmin = np.array([z[nb].min() for nb in neigh], dtype=f32)
    cloud[2] = z - zmin
    # load_cloud (loader.py:73-87): centre (float32 array - scalar -> float32 in NumPy 1.x), fake points, xyz, rescale, sample
    cloud[0] = cloud[0] - f32(center[0])
    cloud[1] = cloud[1] - f32(center[1])
    fake = fake_ground_points(args.diam_meters, cloud.shape[0])
    src = np.concatenate([idx, -1 - np.arange(fake.shape[1])])
    cloud = np.concatenate([cloud, fake], axis=1)
    xyz = cloud[:3].copy()
    cloud[0] = cloud[0] / f32(10)
    cloud[1] = cloud[1] / f32(10)
    cloud[2] = cloud[2] / f32(args.z_max)
    for r in (3, 4, 5, 6):
        cloud[r] = cloud[r] / f32(65536)
    cloud[7] = cloud[7] / f32(32768)
    for r in (8, 9):
        cloud[r] = (cloud[r] - f32(1)) / f32(6)
    # sample_cloud (loader.py:233-247) with the canonical hash in place of np.random
    ntot, S = cloud.shape[1], args.subsample_size
    if ntot > S:
        h = sample_hash(seed, np.arange(ntot))
        order = np.lexsort((np.arange(ntot), h))  # by hash, ties by position
        pick = np.sort(order[:S])
    else:
        extra = (sample_hash(np.uint32(seed) ^ np.uint32(0x5bd1e995), np.arange(S - ntot)) % np.uint32(ntot)).astype(np.int64)
        pick = np.concatenate([np.arange(ntot), extra])
    return dict(xyz=xyz[:, pick], cloud=cloud[:, pick], src=src[pick]), n


def prepare_plots(parcel_cloud, centers, args, seeds=None):
    """All plots of a parcel (prepare.py:71-98 + the DataLoader's load_cloud).  -> list of (dict | None), counts."""
    from scipy.spatial import cKDTree

    tree = cKDTree(parcel_cloud[:2].transpose().astype(np.float64), 50)
    out, counts = [], []
    for i, c in enumerate(np.asarray(centers, dtype=np.float64)):
        d, n = prepare_plot(parcel_cloud, tree, c, i if seeds is None else int(seeds[i]), args)
        out.append(d)
        counts.append(n)
    return out, np.asarray(counts)


def insert_hard_med_veg_raster_band(mosaic):
    """geotiff_raster.py:121-146, verbatim semantics (10 001 thresholds, first arg-min)."""
    image = mosaic[1]
    mask = np.isnan(image)
    with np.errstate(invalid="ignore"):
        target = np.nanmean(image)
        lin = np.linspace(0, 1, 10001)
        delta = np.ones_like(lin)
        for i, thr in enumerate(lin):
            hard = 1.0 * (image > thr)
            hard[mask] = np.nan
            delta[i] = abs(target - np.nanmean(hard))
        thr = lin[np.argmin(delta)]
        hard = 1.0 * (image > thr)
    hard[mask] = np.nan
    return np.insert(mosaic, 3, hard, axis=0), thr, target


def finalize_merged_raster(mosaic):
    """geotiff_raster.py:273-291 without insert_admissibility_raster (rasterio / shapely)."""
    mosaic = mosaic[:4]
    mosaic, thr, target = insert_hard_med_veg_raster_band(mosaic)
    none = np.sum(np.isnan(mosaic[:3]), axis=0) == 3
    mosaic = np.nan_to_num(mosaic, nan=0.0, posinf=None, neginf=None)
    mosaic[:, none] = np.nan
    return mosaic, thr, target
