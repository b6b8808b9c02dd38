"""Recipe for oracle/_ref/: copies of the three reference scripts whose hot-path functions the
CPU baseline of kind "reference" executes (bench.py, through oracle/ref_extract.py: AST
extraction, NumPy shim, nothing modified).  TEST / BASELINE INFRASTRUCTURE.

    python oracle/make_ref.py        # needs /root/reference (authoring container)

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so
the copies travel to the GPU box next to the built .so files; /root/reference itself does not
exist there.  __graft_entry__.build() runs this when /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('SPALIGN_REFERENCE_SRC', '/root/reference')
SCRIPTS = ('batch_spalign_kmeans.py', 'direct_clustering.py', 'superpixel_overlaps.py')


def make(verbose: bool = False) -> bool:
    if not os.path.isfile(os.path.join(SRC, SCRIPTS[0])):
        return False
    dst = os.path.join(HERE, '_ref')
    os.makedirs(dst, exist_ok=True)
    for name in SCRIPTS:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(dst, name))
        if verbose:
            print('oracle/_ref/%s' % name)
    return True


if __name__ == '__main__':
    sys.exit(0 if make(verbose=True) else 1)
