import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_sessionfinish(session, exitstatus):
    """Measured parity errors (max normalised error per test and column, collected by tests/test_parity_gpu.compare)
    go to gpurun_out/parity_errors.json so that a GPU run leaves the numbers behind, not just pass/fail."""
    import json
    mod = sys.modules.get('test_parity_gpu')
    log = getattr(mod, 'ERROR_LOG', None)
    if not log:
        return
    worst = {}
    for test, col, err in log:
        key = test.split('::')[-1]
        worst.setdefault(key, {})
        worst[key][col] = max(worst[key].get(col, 0.), err)
    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_errors.json'), 'w') as f:
            json.dump(worst, f, indent=1, sort_keys=True)
    except OSError:
        pass
