#!/usr/bin/env python3
"""Collect the catalog glasses whose dispersion coefficients the reference's bundled .roa lens
files carry (manufacturer datasheet numbers stored by opticalglass when those files were saved)
into rayoptics_b200/glass_table.json, the fallback of the .seq / .zmx readers for glass names
that the caller's glass_map does not resolve.  Build container only (reads /root/reference);
numbers only, no code.  17 glasses: not a catalog, a convenience."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayoptics_b200.roa import _val      # noqa: E402

SERIES = {'SchottGlass': ('sellmeier', 'Schott'), 'OharaGlass': ('sellmeier', 'Ohara'),
          'CDGMGlass': ('sellmeier', 'CDGM'), 'SumitaGlass': ('sellmeier', 'Sumita'),
          'HoyaGlass': ('power_series', 'Hoya'), 'HikariGlass': ('power_series', 'Hikari')}
table = {}


def walk(o, src):
    if isinstance(o, dict):
        it = o.get('__instance_type__')
        if isinstance(it, list) and len(it) == 2 and it[1] in SERIES and 'attributes' in o:
            a = o['attributes']
            if a.get('gname') and a.get('coefs') is not None:
                form, maker = SERIES[it[1]]
                table.setdefault(a['gname'].upper(), {'name': a['gname'], 'catalog': maker, 'form': form,
                                                      'coefs': [float(c) for c in _val(a['coefs'])],
                                                      'source': src})
        for v in o.values():
            walk(v, src)
    elif isinstance(o, list):
        for v in o:
            walk(v, src)


for f in sorted(glob.glob('/root/reference/src/rayoptics/**/*.roa', recursive=True)):
    try:
        walk(json.load(open(f)), f.split('/rayoptics/', 1)[1])
    except Exception:
        pass
out = os.path.join(ROOT, 'rayoptics_b200', 'glass_table.json')
json.dump(dict(sorted(table.items())), open(out, 'w'), indent=1)
print(len(table), 'glasses ->', out)
