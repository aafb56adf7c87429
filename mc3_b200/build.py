"""Build libmc3b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mc3_b200.build [--force]

The library has no torch / Python dependency: plain CUDA runtime + a C ABI
(include/mc3b200.h).  Objects and the .so are git-ignored but travel to the GPU
box with the working-tree snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmc3b200.so')
SOURCES = ['runtime.cu', 'chisq.cu', 'chisq_grid.cu', 'sampler.cu', 'small.cu', 'dwt.cu', 'timeavg.cu', 'poststat.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '--fmad=true',
    '-Xptxas', '-v',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
               if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'mc3b200.h'))
    objs = []
    logs = []
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for src in srcs:
        path = os.path.join(CSRC, src)
        obj = os.path.join(CSRC, src[:-3] + '.o')
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ['-c', path, '-o', obj]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                               stderr=subprocess.PIPE, text=True)))
    for src, pr in jobs:                       # the translation units compile side by side
        out, err = pr.communicate()
        logs.append(err)
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError(f'nvcc failed on {src}')
        if verbose:
            sys.stderr.write(err)
    if force or _stale(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    if logs:
        with open(os.path.join(CSRC, 'ptxas.log'), 'w') as f:
            f.write('\n'.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
