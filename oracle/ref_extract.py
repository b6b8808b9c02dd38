"""Run the reference's own hot-path functions, unmodified, on NumPy.  TEST INFRASTRUCTURE.

The reference scripts cannot be imported here (chainer / cupy / chainercv / skimage are
not installed), but the functions on the hot path are written against ``xp`` and run on
NumPy unchanged.  This module parses a reference file with ``ast``, keeps only the wanted
``FunctionDef`` nodes and ``exec``s them in a namespace with a three-symbol shim
(SURVEY.md section 8c).  Nothing is copied into the repo: the source is read from
``/root/reference`` at call time (authoring container) or, on the GPU box, from the copies
``oracle/make_ref.py`` leaves under the git-ignored ``oracle/_ref/``.  It is used by
``oracle/gen_golden.py`` (to freeze golden vectors), by the ``needs_reference`` tests and by
``bench.py``'s CPU baseline of kind "reference".
"""
from __future__ import annotations

import ast
import os
import random
import types
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _root() -> str:
    """/root/reference where it exists (authoring container), else the copies that
    oracle/make_ref.py left under oracle/_ref/ (they travel to the GPU box with gpurun)."""
    env = os.environ.get('SPALIGN_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isfile('/root/reference/batch_spalign_kmeans.py'):
        return '/root/reference'
    return os.path.join(_HERE, '_ref')


REFERENCE_ROOT = _root()

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


class _Concat:
    """chainer.functions.concat shim: ``F.concat(xs, axis).array``."""

    def __init__(self, xs, axis=1):
        self.array = np.concatenate([np.asarray(getattr(x, 'array', x)) for x in xs], axis=axis)


class _F:
    concat = _Concat


def _statements_in(tree, func: str, first: int, last: int):
    """The statement nodes of ``func`` whose source lines lie inside [first, last]: the
    shallowest body list that holds statements entirely inside the range."""
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func)

    def search(body):
        hit = [st for st in body if st.lineno >= first and st.end_lineno <= last]
        if hit:
            return hit
        for st in body:
            if st.lineno <= first and st.end_lineno >= last:
                for field in ('body', 'orelse', 'finalbody'):
                    sub = getattr(st, field, None)
                    if sub:
                        got = search(sub)
                        if got:
                            return got
        return []
    return search(fn.body)


def load_inline(script: str, func: str, first: int, last: int, name: str, args, returns):
    """Wrap the INLINE statements ``first..last`` of the reference function ``func`` into a
    function ``name(*args) -> returns`` and return it: the reference code runs unmodified (the
    statement nodes are taken from the parsed source, nothing is retyped), only the ``def``
    line and the ``return`` are added.  Used for the overlap-refine loop
    (superpixel_overlaps.py:360-369) and the direct feature build / prior tiling
    (direct_clustering.py:298-308), which the reference does not factor into functions."""
    if not available():
        raise FileNotFoundError('reference tree not found at %s' % REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, script)
    with open(path) as fp:
        tree = ast.parse(fp.read(), filename=path)
    body = _statements_in(tree, func, first, last)
    if not body or body[0].lineno != first or body[-1].end_lineno != last:
        raise ValueError('%s:%d-%d does not delimit whole statements of %s (got %s)'
                         % (script, first, last, func,
                            [(b.lineno, b.end_lineno) for b in body]))
    ret = ast.Return(value=ast.Tuple(elts=[ast.Name(id=r, ctx=ast.Load()) for r in returns],
                                     ctx=ast.Load()))
    fdef = ast.FunctionDef(
        name=name,
        args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=a) for a in args], kwonlyargs=[],
                           kw_defaults=[], defaults=[]),
        body=list(body) + [ret], decorator_list=[], type_params=[])
    mod = ast.fix_missing_locations(ast.Module(body=[fdef], type_ignores=[]))
    import cv2 as cv
    # helper functions the statements call (create_prior, ...)
    helpers = {}
    if script in _WANTED:
        helpers = vars(load(script, seed=None))
    ns = {'np': np, 'cv': cv, 'F': _F, 'cuda': _Cuda, 'chainer': _Chainer, **helpers}
    exec(compile(mod, path, 'exec'), ns)
    return ns[name]


def refine_loop():
    """superpixel_overlaps.py:360-369 as ``f(road_mask, superpixel, args) -> (refined_roadmap,)``
    (road_mask: bool/uint8 cell or pixel mask of one image, superpixel: its label map)."""
    return load_inline('superpixel_overlaps.py', 'estimate_road_mask', 360, 369,
                       'ref_refine', ['road_mask', 'superpixel', 'args'], ['refined_roadmap'])


def direct_feature_build():
    """direct_clustering.py:298-303 as ``f(use_maps, xp) -> (feature_maps, n, h, w)``."""
    return load_inline('direct_clustering.py', 'estimate_road_mask', 298, 303,
                       'ref_direct_features', ['use_maps', 'xp'], ['feature_maps', 'n', 'h', 'w'])


def direct_prior_tiling():
    """direct_clustering.py:307-308 as ``f(h, w, n, args) -> (prior,)``."""
    return load_inline('direct_clustering.py', 'estimate_road_mask', 307, 308,
                       'ref_direct_prior', ['h', 'w', 'n', 'args'], ['prior'])


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
