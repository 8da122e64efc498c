"""Run the REFERENCE's own analysis layer (raytr/trace.py, raytr/analyses.py) on a mirror model.

TEST INFRASTRUCTURE like ref_harness.py, build container only.  ``rayoptics.raytr.trace`` and
``rayoptics.raytr.analyses`` import here (their opticalglass-dependent siblings do not), so a
"hybrid" optical model -- the reference's Surface / ThinLens objects on the mirror's path,
the mirror's optical specification (whose start-ray code is pinned to the reference's,
tests/test_startrays_vs_reference.py) and first-order data (pinned too) -- lets the reference's
``trace_base`` / ``trace_fan`` / ``trace_grid`` / ``RayFan`` / ``RayList`` / ``RayGrid`` run
unmodified on the reference's own ``trace_raw``.  Used to generate
tests/golden/vectors/*_analyses.npz and by CPU tests that skip without /root/reference.
"""
from __future__ import annotations

import importlib

from . import ref_harness as rh


class HybridSeq:
    """what the reference's trace / analyses read from a SequentialModel"""

    def __init__(self, sm):
        self.sm = sm
        self._paths = {}
        self.gaps, self.z_dir, self.stop_surface = sm.gaps, sm.z_dir, sm.stop_surface
        self.lcl_tfrms, self.wvlns = sm.lcl_tfrms, sm.wvlns
        self.ifcs = [seg[0] for seg in rh.ref_path(sm, sm.central_wavelength())]

    def path(self, wl=None, start=None, stop=None, step=1):
        wl = self.sm.central_wavelength() if wl is None else wl
        if wl not in self._paths:
            p = rh.ref_path(self.sm, wl)
            for seg, ifc in zip(p, self.ifcs):
                seg[0] = ifc
            self._paths[wl] = p
        return iter(self._paths[wl])

    def reverse_path(self, start=None, stop=None, step=-1, wl=None):
        """the mirror's image-to-object path with the reference's interface objects in it"""
        mine = {id(seg[0]): ifc for seg, ifc in zip(self.sm.path(), self.ifcs)}
        return iter([[mine[id(seg[0])]] + list(seg[1:])
                     for seg in self.sm.reverse_path(start, stop, step, wl)])

    def get_num_surfaces(self):
        return self.sm.get_num_surfaces()

    def central_wavelength(self):
        return self.sm.central_wavelength()

    def index_for_wavelength(self, wvl):
        return self.sm.index_for_wavelength(wvl)


class HybridModel:
    def __init__(self, opm):
        self.opm = opm
        self.seq_model = HybridSeq(opm.seq_model)
        self.optical_spec = opm.optical_spec
        self.analysis_results = opm.analysis_results

    def __getitem__(self, key):
        return {'seq_model': self.seq_model, 'sm': self.seq_model,
                'optical_spec': self.optical_spec, 'osp': self.optical_spec,
                'analysis_results': self.analysis_results, 'ar': self.analysis_results}[key]

    def nm_to_sys_units(self, nm):
        return self.opm.nm_to_sys_units(nm)


def modules():
    """(rayoptics.raytr.trace, rayoptics.raytr.analyses) of the reference"""
    rh.ref()
    return (importlib.import_module('rayoptics.raytr.trace'),
            importlib.import_module('rayoptics.raytr.analyses'))
