import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line(
        'markers', 'needs_reference: executes /root/reference functions (authoring container only)')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_extract
    have_ref = ref_extract.available()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    for item in items:
        if 'needs_reference' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='/root/reference not present'))
        if 'gpu' in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
