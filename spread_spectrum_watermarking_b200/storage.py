"""Mark storage files of the reference's CLI (host I/O glue).

`WatermarkStorage::Version1 { config, watermarks: [{values, description}] }` as serde_json writes it
(/root/reference/examples/main.rs:110-131, written at :285-304) and the legacy `.wm` JSON
`{alpha, length, version, wm}` (:321-344, read as Option2 + Legacy ordering).  Loading gives the
configuration and the marks as one [M][n] float32 array, ready for `Bank` (device-resident bank for
`Tester::similarity` against many marks, README.md:62).  No arithmetic happens here.
"""
import json

import numpy as np

from ._lib import SswError, SSW_ERR_INVALID

METHODS = {'Option1': 1, 'Option2': 2, 'Option3': 3}
ORDERINGS = {'Energy': 0, 'EnergyOrthogonal': 1, 'Legacy': 2}


def _names(table, value):
    for k, v in table.items():
        if v == value:
            return k
    raise SswError(SSW_ERR_INVALID, 'not serialisable: %r' % (value,))


def loads(text):
    """-> (config dict {'method', 'alpha', 'ordering'}, marks float32 [M][n], descriptions [M])"""
    doc = json.loads(text)
    if isinstance(doc, dict) and 'Version1' in doc:
        v1 = doc['Version1']
        ie = v1['config']['insert_extract']
        cfg = {'method': METHODS[ie['method']], 'alpha': float(ie['alpha']), 'ordering': ORDERINGS[v1['config']['ordering']]}
        wms = v1['watermarks']
    elif isinstance(doc, dict) and 'wm' in doc and 'alpha' in doc:   # legacy .wm (examples/main.rs:321-344)
        cfg = {'method': 2, 'alpha': float(doc['alpha']), 'ordering': ORDERINGS['Legacy']}
        wms = [{'values': doc['wm'], 'description': ''}]
    else:
        raise SswError(SSW_ERR_INVALID, 'not a watermark storage file')
    lens = {len(w['values']) for w in wms}
    if len(lens) > 1:
        raise SswError(SSW_ERR_INVALID, 'marks of different lengths in one file')
    marks = np.array([w['values'] for w in wms], dtype=np.float32).reshape(len(wms), lens.pop() if lens else 0)
    return cfg, marks, [w.get('description', '') for w in wms]


def load(path):
    with open(path) as f:
        return loads(f.read())


def dumps(config, marks, descriptions=None):
    """the Version1 form the reference writes (examples/main.rs:285-304)"""
    marks = np.asarray(marks, dtype=np.float32)
    if marks.ndim == 1:
        marks = marks[None, :]
    descriptions = descriptions or [''] * len(marks)
    # f32 values are written with the shortest round-trip repr, like serde_json does for f32
    vals = [[float(np.format_float_positional(v, unique=True, trim='0')) if np.isfinite(v) else None for v in m] for m in marks]
    return json.dumps({'Version1': {
        'config': {'insert_extract': {'alpha': float(np.float32(config['alpha'])), 'method': _names(METHODS, config['method'])},
                   'ordering': _names(ORDERINGS, config['ordering'])},
        'watermarks': [{'values': v, 'description': d} for v, d in zip(vals, descriptions)]}})


def save(path, config, marks, descriptions=None):
    with open(path, 'w') as f:
        f.write(dumps(config, marks, descriptions))


def bank_from_file(path, ctx=None):
    """device-resident Bank of every mark in the file + its configuration"""
    from . import Bank
    cfg, marks, desc = load(path)
    return Bank(marks, ctx=ctx), cfg, desc
