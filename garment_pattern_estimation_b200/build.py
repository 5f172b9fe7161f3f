"""Build recipe for libnt_b200.so -- the C-ABI CUDA library of the hot path (include/nt_b200.h).

nvcc cross-compiles for sm_100a without a GPU; the .so is written IN-TREE (garment_pattern_estimation_b200/lib/) so it
travels to the GPU box with the repository snapshot.  No JIT, no torch extension machinery: the library has no torch
types in its interface and is loaded with ctypes (see _lib.py).
"""
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_DIR = os.path.join(_HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libnt_b200.so')
INCLUDE = os.path.join(os.path.dirname(_HERE), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-DNT_BUILT_ARCH=100', '--shared',
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _fingerprint():
    h = hashlib.sha256()
    for path in sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith('.cuh')] + \
            [os.path.join(INCLUDE, 'nt_b200.h')]:
        with open(path, 'rb') as f:
            h.update(os.path.basename(path).encode() + b'\0' + f.read())      # path-independent: the tree is copied to the GPU box
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return None


def _compile_one(nvcc, src, obj, flags, verbose):
    cmd = [nvcc] + flags + ['-I', INCLUDE, '-c', src, '-o', obj]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)


def is_current():
    """True if lib/libnt_b200.so exists and its stamp matches the sources (csrc/*.cu, *.cuh, include/nt_b200.h, flags)."""
    stamp = os.path.join(LIB_DIR, 'libnt_b200.stamp')
    return os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == _fingerprint()


def build(force=False, verbose=False, extra_flags=()):
    """Compile every .cu under csrc/ (in parallel, one object per source, re-using up-to-date objects) and link them into
    lib/libnt_b200.so.  Returns the library path."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'libnt_b200.stamp')
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build libnt_b200.so (the product has no CPU fallback)')
    obj_dir = os.path.join(LIB_DIR, 'obj')
    os.makedirs(obj_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != '--shared'] + list(extra_flags)
    headers = hashlib.sha256()
    for path in [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith('.cuh')] + \
            [os.path.join(INCLUDE, 'nt_b200.h')]:
        with open(path, 'rb') as f:
            headers.update(f.read())
    headers.update(' '.join(flags).encode())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        with open(src, 'rb') as f:
            key = hashlib.sha256(headers.digest() + f.read()).hexdigest()
        keyfile = obj + '.key'
        objs.append(obj)
        if force or not (os.path.exists(obj) and os.path.exists(keyfile) and open(keyfile).read() == key):
            jobs.append((src, obj, keyfile, key))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as pool:
        futures = [pool.submit(_compile_one, nvcc, src, obj, flags, verbose) for src, obj, _, _ in jobs]
        for fut in futures:
            fut.result()
    for _, obj, keyfile, key in jobs:
        with open(keyfile, 'w') as f:
            f.write(key)
    link = [nvcc, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs
    if verbose:
        print(' '.join(link))
    subprocess.check_call(link)
    with open(stamp, 'w') as f:
        f.write(fp)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose=True,
                extra_flags=['-Xptxas', '-v'] if '--ptxas' in sys.argv else []))
