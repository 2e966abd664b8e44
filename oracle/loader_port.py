"""CPU oracle: restated train-time loader transforms.  TEST INFRASTRUCTURE ONLY.

augment + rescale_cloud of /root/reference/data_loader/loader.py:135-214 with the random draws passed in (the
reference draws them with np.random at the same places); numpy float32 / float64 exactly as the reference mixes them."""
import numpy as np


def rotate_around_z(cloud, angle):
    """loader.py:210-214."""
    c, s = np.cos(angle), np.sin(angle)
    M = np.array(((c, -s), (s, c)))
    cloud[:2] = np.dot(cloud[:2].T, M).T
    return cloud


def augment(cloud, xyz, angle, flip_x, flip_y, noise):
    """loader.py:161-207.  cloud (10,N), xyz (3,N) float32 (modified in place); noise (6,N) float64 standard normal."""
    cloud = rotate_around_z(cloud, angle)
    xyz = rotate_around_z(xyz, angle)
    if flip_x:
        cloud[0] = -cloud[0]
        xyz[0] = -xyz[0]
    if flip_y:
        cloud[1] = -cloud[1]
        xyz[1] = -xyz[1]
    sigma, clip = 0.01 * 10, 0.03 * 10
    cloud[:2] = cloud[:2] + np.clip(sigma * noise[:2], a_min=-clip, a_max=clip).astype(np.float32)
    clip = 0.03 * 65536
    for k, idx in enumerate((3, 4, 5, 6)):  # red, green, blue, near_infrared; the reference multiplies by the xy `sigma` here too
        cloud[idx] = cloud[idx] + np.clip(sigma * noise[2 + k], a_min=-clip, a_max=clip).astype(np.float32)
    return cloud, xyz


def rescale_cloud(cloud, z_max):
    """loader.py:135-158 (float32 array / python number -> float32)."""
    f32 = np.float32
    cloud[0] = cloud[0] / f32(10)
    cloud[1] = cloud[1] / f32(10)
    cloud[2] = cloud[2] / f32(z_max)
    for idx in (3, 4, 5, 6):
        cloud[idx] = cloud[idx] / f32(65536)
    cloud[7] = cloud[7] / f32(32768)
    for idx in (8, 9):
        cloud[idx] = (cloud[idx] - f32(1)) / f32(7 - 1)
    return cloud


def load_transform(raw, z_max, angle=None, flip=None, noise=None):
    """raw (10,N) float32 centred plot with fake points -> (xyz (3,N), cloud (10,N)) as load_cloud(train=...) leaves them
    before sampling (loader.py:77-85)."""
    cloud = raw.astype(np.float32).copy()
    xyz = cloud[:3].copy()
    if angle is not None:
        cloud, xyz = augment(cloud, xyz, angle, bool(flip[0]), bool(flip[1]), noise)
    return xyz, rescale_cloud(cloud, z_max)
