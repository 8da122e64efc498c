"""Pupil sample generators with the reference's semantics
(/root/reference/src/rayoptics/raytr/sampler.py:15-54,105-122): square grid with
accumulated stepping, and the concentric square->disk map used by ``RayList``.
Host-side, O(num^2) Python/numpy -- they produce *coordinates*, the rays are
traced in one launch."""
import math

import numpy as np


def grid_ray_generator(grid_rng):
    """sampler.py:15-39 -- x outer, y inner, ``sample_pt += step`` accumulation."""
    start, stop, num = grid_rng
    sample_pt = np.array(start, dtype=float)
    step = np.array((np.asarray(stop) - np.asarray(start))/(num - 1))
    for i in range(num):
        for j in range(num):
            yield np.array(sample_pt)
            sample_pt[1] += step[1]
        sample_pt[0] += step[0]
        sample_pt[1] = start[1]


def concentric_sample_disk(u, offset=True):
    """sampler.py:105-122 -- map a 2d unit-square sample to the unit disk."""
    uOffset = 2*u - np.array([1, 1]) if offset else u
    if uOffset[0] == 0 and uOffset[1] == 0:
        return np.array([0, 0])
    if abs(uOffset[0]) > abs(uOffset[1]):
        r = uOffset[0]
        theta = np.pi/4*(uOffset[1]/uOffset[0])
    else:
        r = uOffset[1]
        theta = np.pi/2 - np.pi/4*(uOffset[0]/uOffset[1])
    return r*np.array([math.cos(theta), math.sin(theta)])


def csd_grid_ray_generator(grid_rng):
    """sampler.py:42-54 -- square grid pushed through concentric_sample_disk."""
    start = np.array(grid_rng[0], dtype=float)
    stop = grid_rng[1]
    num = grid_rng[2]
    step = np.array((np.asarray(stop) - start)/(num - 1))
    for i in range(num):
        for j in range(num):
            yield concentric_sample_disk(start, offset=False)
            start[1] += step[1]
        start[0] += step[0]
        start[1] = grid_rng[0][1]


def polar_grid_ray_generator(grid_rng):
    """sampler.py:55-66 -- the same traversal as the square grid (the reference does not map
    to polar coordinates here either)"""
    return grid_ray_generator(grid_rng)


def phi(d):
    """generalised golden ratio by the nested radical, sampler.py:73-77"""
    x = 2.0000
    for i in range(10):
        x = pow(1 + x, 1/(d + 1))
    return x


def R_2_quasi_random_generator(n):
    """2-D R2 low-discrepancy sequence, sampler.py:80-102"""
    d = 2
    g = phi(d)
    alpha = np.zeros(d)
    for j in range(d):
        alpha[j] = pow(1/g, j + 1) % 1
    seed = 0.5
    z = np.zeros((n, d))
    for i in range(n):
        z[i] = (seed + alpha*(i + 1)) % 1
        yield z[i]


def create_generator(sampler, *sampler_args, mapper=None, **kwargs):
    """sampler.py:125-132 -- chain a sampler with an optional mapping function"""
    def gen():
        for xy in sampler(*sampler_args):
            yield mapper(xy, **kwargs) if mapper else xy
    return gen()
