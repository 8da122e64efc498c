"""bench.py's reference arm runs here (no GPU): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--model', 'rc', '--num', '64'], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
              'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['value'] > 0 and line['unit'] == 'rays/s'
    assert line['e2e'] == {'value': line['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and 'sample' in cb
    if cb['kind'] == 'reference':               # baseline/_ref installed: the unmodified Python reference
        assert cb['port']['kind'] == 'port' and cb['port']['value'] > cb['value']
        assert 'trace_grid' in cb['sample']
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ''
