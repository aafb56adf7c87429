"""Scratch tool: build libmc3b200 variants that differ in compile-time constants of
one translation unit (for A/B runs on the GPU box; select one with MC3B_LIBPATH).

    python profiles/build_variants.py [--src chisq_grid.cu] name:-DMC3B_GRID_RESTART=4 name2:-DA=1,-DB=2
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc3_b200 import build as b

args = sys.argv[1:]
src = 'chisq_grid.cu'
if args and args[0] == '--src':
    src = args[1]
    args = args[2:]
b.build()
out = os.path.join(os.path.dirname(b.HERE), 'variants')
os.makedirs(out, exist_ok=True)
others = [os.path.join(b.CSRC, s[:-3] + '.o') for s in b.SOURCES if s != src]
for spec in args:
    name, flags = spec.split(':', 1)
    obj = os.path.join(out, f'{src[:-3]}_{name}.o')
    r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + [f for f in flags.split(',') if f] +
                       ['-c', os.path.join(b.CSRC, src), '-o', obj],
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr); raise SystemExit(1)
    grab = 0
    for ln in r.stderr.splitlines():
        if ('sinegrid' in ln or 'SineGridModeldLi32ELi1' in ln) and 'Compiling' in ln:
            grab = 3
        if grab > 0 and ('Used' in ln or 'spill' in ln):
            print(name, ln.strip()); grab -= 1
    lib = os.path.join(out, f'libmc3b200_{name}.so')
    subprocess.run([b._nvcc(), '-shared', '-o', lib, obj] + others + ['-lcudart'], check=True)
    print('built', lib)
