"""In-tree build of libsynthsr_b200.so (sm_100a only) with nvcc.  No JIT cache: the .so lives next to the sources
so it travels to the GPU box with the repository snapshot.

    python -m synthsr_b200.build          # build if stale
    python -m synthsr_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libsynthsr_b200.so')

# (source, extra flags).  generator.cu must not contract a*b+c into FMA: the label-resampling coordinates have
# to round like the reference's op-by-op float32 graph (bit-exact nearest-neighbour output).
SOURCES = [
    ('capi.cu', []),
    ('generator.cu', ['-fmad=false']),
    ('unet_kernels.cu', []),
    ('conv_tc.cu', []),
    ('seg_loss.cu', []),
]

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
COMMON = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
          '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    objs = []
    logs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + COMMON + extra + ['-c', s, '-o', o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append((src, r.stderr))
            if verbose:
                print(' '.join(cmd))
                print(r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    if force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'shared']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(os.path.join(LIBDIR, 'ptxas.log'), 'a' if not force else 'w') as f:
        for src, log in logs:
            f.write('==== %s ====\n%s\n' % (src, log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
