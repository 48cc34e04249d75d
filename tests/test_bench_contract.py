"""bench.py's reference arm on the CPU (the one leg that runs without a GPU): exactly one JSON line on stdout with the
keys the driver reads.  A 2-geometry sample keeps it to a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--batch', '2', '--steps', '1',
                          '--warmup', '1', '--no-configs'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'points/s' and d['higher_is_better'] is True
    for k in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
              'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['value'] > 0 and d['gpu_launches'] == 0 and d['vs_baseline'] is None
    assert 'workload' in d['config'] and '2 geometries per GPU' in d['config']['workload']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    e = d['e2e']
    assert e['value'] == d['value'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'], env=env,
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''
