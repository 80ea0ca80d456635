"""The drop-in boundary exercised from plain C: tests/cabi/cabi_flow.c links libssw.so and replays the reference's doc
examples (/root/reference/src/lib.rs:22-66) and tests/single_simple.rs on the reference's fixture (tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CABI = os.path.join(ROOT, 'tests', 'cabi')


def build():
    subprocess.check_call(['make', '-C', CABI, 'cabi_flow'], stdout=subprocess.DEVNULL)
    return os.path.join(CABI, 'cabi_flow')


def test_c_caller_builds_against_the_header_and_library():
    """gcc compiles the C caller against include/ssw.h and links libssw.so (no GPU needed for that); without a device the
    program must fail loudly at ssw_ctx_create -- there is no CPU path to fall back to"""
    exe = build()
    assert os.path.exists(exe)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        r = subprocess.run([exe, '/dev/null', '0', '0', '/dev/null', '0', '/dev/null'], capture_output=True, text=True)
        assert r.returncode == 2 and 'ssw_ctx_create' in r.stderr and 'no CUDA device' in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_caller_replays_the_reference_examples(tmp_path):
    exe = build()
    g = os.path.join(ROOT, 'tests', 'golden')
    cat = np.load(os.path.join(g, 'cat_rgb8.npz'))['rgb']
    marks = np.load(os.path.join(g, 'marks.npz'))
    golden = np.load(os.path.join(g, 'watermarked_with_1.npz'))['rgb']
    h, w = cat.shape[:2]
    paths = {}
    for name, arr in (('cat', cat), ('mark', marks['seed_1']), ('golden', golden), ('rnd', marks['seed_baaaaaad'])):
        paths[name] = str(tmp_path / (name + '.bin'))
        np.ascontiguousarray(arr).tofile(paths[name])
    r = subprocess.run([exe, paths['cat'], str(w), str(h), paths['mark'], '1000', paths['golden'], paths['rnd']],
                       capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert 'all checks passed' in r.stdout
