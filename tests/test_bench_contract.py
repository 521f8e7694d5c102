"""The CPU-runnable parts of bench.py's contract: the reference arm's JSON line, and that the GPU arm refuses to run
without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          env=env, timeout=900, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('x4 SR images/sec') and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['value'] > 0 and abs(d['ms_per_step'] * d['value'] / 1e3 - 2) < 1e-6          # two images per step
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] == (os.cpu_count() or 1) and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_gpu_arm_fails_loudly_without_a_device():
    r = _run('--steps', '1')
    assert r.returncode != 0
    assert 'no CUDA device' in (r.stderr + r.stdout) and 'no CPU fallback' in (r.stderr + r.stdout)
