"""bench.py's contract pieces that do not need a GPU: the reference arm's JSON line, rank handling under
torchrun's environment, and the stdout guard that keeps library banners (NCCL prints its version with a plain
printf to fd 1) away from the one JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, 'bench.py')


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    e['SSW_REF_BUDGET_S'] = '5'
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(['--impl', 'reference', '--workload', 'c3', '--steps', '1', '--warmup', '0'])   # 1080p frames: seconds, not a minute
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'Mpix/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] >= 1 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['frame'] == [1920, 1080] and d['config']['mark_len'] == 1000


def test_reference_arm_other_ranks_exit_silently():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_stdout_guard_keeps_foreign_output_off_the_json_line():
    code = (
        "import importlib.util, os, sys\n"
        "sys.argv = ['bench.py']\n"
        "spec = importlib.util.spec_from_file_location('bench', %r)\n"
        "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
        "m.protect_stdout(); os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('noise from python'); m.emit({'ok': 1})\n"
    ) % BENCH
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert 'NCCL version' in r.stderr and 'noise from python' in r.stderr


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('GPU present')
    r = _run(['--steps', '1', '--warmup', '0'])
    assert r.returncode != 0 and r.stdout.strip() == ''
    assert 'no CPU fallback' in r.stderr or 'CUDA' in r.stderr
