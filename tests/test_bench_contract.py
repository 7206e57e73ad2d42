"""bench.py contract on the CPU: the reference arm (`--impl reference`, oracle ports on the host cores) prints ONE JSON
line with the keys the driver reads; our arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'graphs/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('graphs/sec preprocess+forward') and d['value'] > 0 and d['steps'] == 2
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert 'workload' in d['config']


def test_our_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0 and 'no CPU path' in (p.stderr + p.stdout)
