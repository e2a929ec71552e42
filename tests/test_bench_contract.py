"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) prints one JSON
line with the keys the driver reads; the GPU arm is covered by the driver itself on the B200 box."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--planes', '1k', '--cpu-sample-images', '2']
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=300,
                       cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'hypotheses/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and 'sample' in d['cpu_baseline']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and 'workload' in d['config'] and d['dtype'] == 'f32'
