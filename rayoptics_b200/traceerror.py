"""Per-ray failure types of the trace, by the reference's names
(/root/reference/src/rayoptics/raytr/traceerror.py:11-52): what a kernel ``status`` code turns into
on the host (``raytrace.package_ray``).  Used when the reference package itself is not importable;
``raytrace.py`` prefers the reference's own classes so that ``except TraceError`` in reference code
keeps working after ``install()``.

Each type only records what was known where the ray stopped.  The constructor arguments (order
and defaults) and attribute names are the reference's, declared here as data: ``_args`` are the
positional parameters, ``_unset`` the attributes that exist from the start but are filled in later
by the tracer (``surf`` / ``ray_pkg`` of the base type, ``ifc`` / ``int_pt`` of a TIR failure).
"""

_REQUIRED = object()


class TraceError(Exception):
    """a ray did not make it through the model; ``surf``: interface index, ``ray_pkg``: the partial ray"""
    _args = (('surf', None), ('ray_pkg', None))
    _unset = ()

    def __init__(self, *values, **named):
        spec = type(self)._args
        if len(values) > len(spec):
            raise TypeError(f'{type(self).__name__}() takes at most {len(spec)} arguments')
        given = dict(zip((k for k, _ in spec), values))
        for k in named:
            if k in given or k not in dict(spec):
                raise TypeError(f'{type(self).__name__}(): unexpected argument {k!r}')
        given.update(named)
        for k in type(self)._unset:
            setattr(self, k, None)
        for k, default in spec:
            if k not in given and default is _REQUIRED:
                raise TypeError(f'{type(self).__name__}() missing argument {k!r}')
            setattr(self, k, given.get(k, default))


def _failure(name, doc, args, unset=()):
    return type(name, (TraceError,), {'__doc__': doc, '_args': tuple(args), '_unset': tuple(unset),
                                      '__module__': __name__})


_R = _REQUIRED
TraceMissedSurfaceError = _failure(
    'TraceMissedSurfaceError', 'status 1: no intersection with the interface (square root of a negative)',
    [('ifc', None), ('prev_seg', None)])
TraceTIRError = _failure(
    'TraceTIRError', 'status 2: total internal reflection at a refracting interface',
    [('inc_dir', _R), ('normal', _R), ('prev_indx', _R), ('follow_indx', _R)], unset=('ifc', 'int_pt'))
TraceRayBlockedError = _failure(
    'TraceRayBlockedError', 'status 3: stopped by an aperture of the interface',
    [('ifc', _R), ('int_pt', _R)])
TraceEvanescentRayError = _failure(
    'TraceEvanescentRayError', 'status 4: evanescent diffraction order at a phase element',
    [('ifc', _R), ('int_pt', _R), ('inc_dir', _R), ('normal', _R), ('prev_indx', _R), ('follow_indx', _R)])
