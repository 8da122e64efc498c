"""Pupil sample sets as coordinate ARRAYS.

The reference produces pupil samples one ``yield`` at a time
(/root/reference/src/rayoptics/raytr/sampler.py:15-132) because it traces one ray
per iteration.  Here a whole sample set is one ``[n, 2]`` float64 array that goes
to the device as the ``pupil_x`` / ``pupil_y`` tables of a grid launch, so the
producers below are whole-array computations:

    square_grid_axes(grid_rng)    -> (xs[num], ys[num])   the two axes only: the grid
                                     kernel forms the product on the device
    square_grid_points(grid_rng)  -> [num*num, 2]         x outer, y inner
    disk_grid_points(grid_rng)    -> [num*num, 2]         square grid -> unit disk
    concentric_disk(uv)           -> [n, 2]               Shirley's concentric map
    r2_sequence(n)                -> [n, 2]               R2 low-discrepancy points

Values are the reference's doubles bit for bit (``tests/test_host.py``): the axes are
running sums (``start``, ``start+step``, ``(start+step)+step`` ...; ``np.add.accumulate``
adds strictly left to right), and sine / cosine go through ``math`` (libm) element by
element because numpy's SIMD kernels may round differently.

The reference's generator names are kept as thin iterators over the arrays, for callers
written against ``rayoptics.raytr.sampler``.
"""
import math

import numpy as np


def _running_sum(start, step, num):
    """start, start+step, (start+step)+step, ...: sampler.py:30-39 adds the step to the
    running value, it does not multiply (the last sample is not exactly ``stop``)."""
    terms = np.full(num, step, dtype=np.float64)
    terms[0] = start
    return np.add.accumulate(terms)


def square_grid_axes(grid_rng):
    """The x and y sample positions of the reference's square grid."""
    start, stop, num = grid_rng
    lo = np.asarray(start, dtype=np.float64)
    step = (np.asarray(stop, dtype=np.float64) - lo)/(num - 1)
    return _running_sum(lo[0], step[0], num), _running_sum(lo[1], step[1], num)


def square_grid_points(grid_rng):
    """``[num*num, 2]``: x outer, y inner -- the traversal of sampler.py:15-39."""
    xs, ys = square_grid_axes(grid_rng)
    n = len(xs)
    pts = np.empty((n, n, 2))
    pts[:, :, 0] = xs[:, None]
    pts[:, :, 1] = ys[None, :]
    return pts.reshape(n*n, 2)


def _libm(fn, a):
    return np.fromiter(map(fn, a.tolist()), dtype=np.float64, count=a.size)


def concentric_disk(uv, offset=True):
    """Concentric square -> disk map of sampler.py:105-122 for ``[n, 2]`` samples.
    ``offset``: the samples are in [0, 1]^2 and are first moved to [-1, 1]^2."""
    uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
    if offset:
        uv = 2*uv - np.array([1, 1])
    a, b = uv[:, 0], uv[:, 1]
    wide = np.abs(a) > np.abs(b)
    origin = (a == 0) & (b == 0)
    with np.errstate(divide='ignore', invalid='ignore'):
        theta = np.where(wide, np.pi/4*(b/a), np.pi/2 - np.pi/4*(a/b))
    theta[origin] = 0.0
    r = np.where(wide, a, b)
    out = np.stack([r*_libm(math.cos, theta), r*_libm(math.sin, theta)], axis=1)
    out[origin] = 0.0
    return out


def disk_grid_points(grid_rng):
    """The square grid pushed through the concentric map (sampler.py:42-54): RayList's
    default pupil sampling."""
    return concentric_disk(square_grid_points(grid_rng), offset=False)


def phi(d):
    """Positive root of x**(d+1) = x + 1 by ten rounds of the fixed-point iteration the
    reference uses (sampler.py:73-77); d = 2 gives the plastic number."""
    x = 2.0
    for _ in range(10):
        x = pow(1 + x, 1/(d + 1))
    return x


def r2_sequence(n, d=2):
    """First ``n`` points of the R_d additive-recurrence sequence (sampler.py:80-102):
    frac(1/2 + k*alpha), alpha_j = phi^-(j+1)."""
    g = phi(d)
    alpha = np.array([pow(1/g, j + 1) % 1 for j in range(d)])
    k = np.arange(1, n + 1, dtype=np.int64)[:, None]
    return (0.5 + alpha[None, :]*k) % 1


# ---- the reference's generator names: iterators over the arrays above ---------------

def grid_ray_generator(grid_rng):
    return iter(square_grid_points(grid_rng))


def csd_grid_ray_generator(grid_rng):
    return iter(disk_grid_points(grid_rng))


def polar_grid_ray_generator(grid_rng):
    """The reference's version does not map to polar coordinates either (sampler.py:55-66)."""
    return iter(square_grid_points(grid_rng))


def R_2_quasi_random_generator(n):
    return iter(r2_sequence(n))


def concentric_sample_disk(u, offset=True):
    """One sample through ``concentric_disk``."""
    return concentric_disk(np.asarray(u, dtype=np.float64).reshape(1, 2), offset=offset)[0]


_ARRAY_FORM = {grid_ray_generator: square_grid_points, csd_grid_ray_generator: disk_grid_points,
               polar_grid_ray_generator: square_grid_points, R_2_quasi_random_generator: r2_sequence}


def sample_points(sampler, *sampler_args, mapper=None, **kwargs):
    """``[n, 2]`` array of a sampler (one of the names above, or any iterable-returning
    callable) optionally pushed through ``mapper``."""
    make = _ARRAY_FORM.get(sampler)
    pts = (make(*sampler_args) if make is not None
           else np.array([np.asarray(p, dtype=np.float64) for p in sampler(*sampler_args)]).reshape(-1, 2))
    if mapper is None:
        return pts
    if mapper is concentric_sample_disk:
        return concentric_disk(pts, **kwargs)
    return np.array([mapper(p, **kwargs) for p in pts]).reshape(-1, 2)


def create_generator(sampler, *sampler_args, mapper=None, **kwargs):
    """sampler.py:125-132: a sampler chained with an optional mapping function."""
    return iter(sample_points(sampler, *sampler_args, mapper=mapper, **kwargs))
