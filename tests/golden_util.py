import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_algebra_golden():
    with open(os.path.join(GOLDEN_DIR, 'algebra_golden.json')) as fh:
        meta = json.load(fh)
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, 'algebra_golden.npz')))
    return meta, arrays
