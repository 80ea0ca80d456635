"""Mark storage files of the reference CLI (examples/main.rs:110-131, 321-344): host-side parsing only."""
import json

import numpy as np
import pytest


def test_version1_roundtrip(wm, tmp_path):
    from spread_spectrum_watermarking_b200 import storage
    rng = np.random.default_rng(0)
    marks = rng.standard_normal((3, 50)).astype(np.float32)
    cfg = {'method': 2, 'alpha': 0.1, 'ordering': 0}
    p = tmp_path / 'cat_wm.json'
    storage.save(str(p), cfg, marks, ['a', 'b', 'c'])
    doc = json.loads(p.read_text())
    assert doc['Version1']['config'] == {'insert_extract': {'alpha': pytest.approx(0.1), 'method': 'Option2'}, 'ordering': 'Energy'}
    cfg2, marks2, desc = storage.load(str(p))
    assert cfg2 == {'method': 2, 'alpha': pytest.approx(0.1), 'ordering': 0} and desc == ['a', 'b', 'c']
    assert (marks2 == marks).all()            # shortest round-trip f32 text


def test_legacy_wm_file(wm):
    from spread_spectrum_watermarking_b200 import storage
    cfg, marks, desc = storage.loads(json.dumps({'alpha': 0.25, 'length': 4, 'version': '1', 'wm': [0.5, -1.0, 2.0, 0.125]}))
    assert cfg == {'method': 2, 'alpha': 0.25, 'ordering': 2} and marks.shape == (1, 4) and desc == ['']
    with pytest.raises(wm.SswError):
        storage.loads('{"something": 1}')
