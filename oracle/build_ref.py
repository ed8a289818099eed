"""ORACLE (test infrastructure): make the UNMODIFIED reference travel to the GPU box.

    python -m oracle.build_ref          # in the build container, where /root/reference exists

copies the reference's Python package -- opendrift/ as it lies under /root/reference, byte for byte, nothing edited -- into
oracle/_ref/opendrift (git-ignored, NOT gpurun-ignored: it ships to the GPU box with the snapshot like the built .so files).
oracle/refrun.py then imports it from there when /root/reference is absent, with the same stub modules for the plotting / IO
packages that this image lacks and the same fake pyproj (oracle/geod_karney.py, oracle/proj_stere.py).  bench.py's reference arm
(`--impl reference`) and cpu_baseline time OceanDrift.run() of that package: kind = "reference".

No reference SOURCE enters the repository's history: oracle/_ref/ is listed in .gitignore and this recipe is what is committed.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('OPENDRIFT_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')


def tree_digest(root):
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith('.pyc'):
                continue
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, root).encode())
            h.update(open(p, 'rb').read())
    return h.hexdigest()


def build(verbose=True):
    src = os.path.join(SRC, 'opendrift')
    if not os.path.isdir(src):
        if verbose:
            print('oracle/build_ref: no reference tree at %s (GPU box?): using what is in oracle/_ref' % SRC)
        return os.path.isdir(os.path.join(DST, 'opendrift'))
    dst = os.path.join(DST, 'opendrift')
    want = tree_digest(src)
    stamp = os.path.join(DST, 'SHA256')
    if os.path.isdir(dst) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return True
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for extra in ('LICENSE', 'history.md'):
        p = os.path.join(SRC, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, extra))
    assert tree_digest(dst) == want, 'copy differs from the reference tree'
    open(stamp, 'w').write(want + '\n')
    if verbose:
        print('oracle/build_ref: %s -> %s (sha256 %s)' % (src, dst, want[:16]))
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
