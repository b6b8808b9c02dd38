"""Run the reference's own hot-path functions, unmodified, on NumPy.  TEST INFRASTRUCTURE.

The reference scripts cannot be imported here (chainer / cupy / chainercv / skimage are
not installed), but the functions on the hot path are written against ``xp`` and run on
NumPy unchanged.  This module parses a reference file with ``ast``, keeps only the wanted
``FunctionDef`` nodes and ``exec``s them in a namespace with a three-symbol shim
(SURVEY.md section 8c).  Nothing is copied into the repo: the source is read from
``/root/reference`` at call time, so this only works in the authoring container.  It is
used by ``oracle/gen_golden.py`` (to freeze golden vectors) and by the
``needs_reference`` tests (skipped where ``/root/reference`` is absent, e.g. the GPU box).
"""
from __future__ import annotations

import ast
import os
import random
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get('SPALIGN_REFERENCE_ROOT', '/root/reference')

_WANTED = {
    'batch_spalign_kmeans.py': {
        'create_prior', 'weighted_average', 'kmeans', 'weighted_kmeans', 'superpixel_align',
        'batch_superpixel_align', 'batch_create_prior',
    },
    'direct_clustering.py': {'create_prior', 'weighted_average', 'kmeans'},
    'superpixel_overlaps.py': {'create_prior', 'weighted_average', 'kmeans'},
}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'batch_spalign_kmeans.py'))


class _Cuda:
    """chainer.cuda shim: everything is NumPy, transfers are the identity."""

    @staticmethod
    def get_array_module(*_):
        return np

    @staticmethod
    def to_cpu(x):
        return x

    @staticmethod
    def to_gpu(x, *_a, **_k):
        return x


class _Chainer:
    class Variable:  # isinstance(feature_map, chainer.Variable) must be False
        pass


def load(script: str = 'batch_spalign_kmeans.py', seed: int | None = 1111) -> types.SimpleNamespace:
    """Return a namespace holding the reference functions of ``script``.

    ``seed`` re-creates the import-time seeding of the scripts (random / np.random seeded
    with 1111, batch_spalign_kmeans.py:33-35); pass None to leave the streams alone.
    """
    if not available():
        raise FileNotFoundError('reference tree not found at %s' % REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, script)
    with open(path) as fp:
        tree = ast.parse(fp.read(), filename=path)
    wanted = _WANTED[script]
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    mod = ast.Module(body=body, type_ignores=[])
    if not hasattr(np, 'float'):  # alias removed in numpy 1.24, used at :233
        np.float = float  # type: ignore[attr-defined]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from scipy.ndimage import measurements  # deprecated alias, still importable
    ns = {'np': np, 'random': random, 'measurements': measurements, 'cuda': _Cuda,
          'chainer': _Chainer}
    exec(compile(mod, path, 'exec'), ns)
    if seed is not None:
        random.seed(seed)
        np.random.seed(seed)
    return types.SimpleNamespace(**{k: ns[k] for k in wanted if k in ns})
