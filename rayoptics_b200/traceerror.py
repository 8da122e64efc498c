"""Ray-trace exceptions with the reference's names and attributes
(/root/reference/src/rayoptics/raytr/traceerror.py:11-52).  Used when the
reference package itself is not importable; ``raytrace.py`` prefers the
reference's own classes so that ``except TraceError`` in reference code keeps
working after ``install()``."""


class TraceError(Exception):
    """Exception raised when ray tracing a model"""

    def __init__(self, surf=None, ray_pkg=None):
        self.surf = surf
        self.ray_pkg = ray_pkg


class TraceMissedSurfaceError(TraceError):
    """Exception raised when ray misses an interface"""

    def __init__(self, ifc=None, prev_seg=None):
        self.ifc = ifc
        self.prev_seg = prev_seg


class TraceTIRError(TraceError):
    """Exception raised when ray TIRs at an interface"""

    def __init__(self, inc_dir, normal, prev_indx, follow_indx):
        self.ifc = None
        self.int_pt = None
        self.inc_dir = inc_dir
        self.normal = normal
        self.prev_indx = prev_indx
        self.follow_indx = follow_indx


class TraceEvanescentRayError(TraceError):
    """Exception raised when ray diffracts evanescently at an interface"""

    def __init__(self, ifc, int_pt, inc_dir, normal, prev_indx, follow_indx):
        self.ifc = ifc
        self.int_pt = int_pt
        self.inc_dir = inc_dir
        self.normal = normal
        self.prev_indx = prev_indx
        self.follow_indx = follow_indx


class TraceRayBlockedError(TraceError):
    """Exception raised when ray is blocked by an aperture on an interface"""

    def __init__(self, ifc, int_pt):
        self.ifc = ifc
        self.int_pt = int_pt
