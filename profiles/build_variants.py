"""Scratch tool: build libmc3b200 variants that differ in compile-time constants of
chisq.cu (for A/B runs on the GPU box; select one with MC3B_LIBPATH).

    python profiles/build_variants.py name:-DMC3B_RESIDENT=8 name2:-DMC3B_TILE_F64=256,-DMC3B_WARPS=8
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc3_b200 import build as b

b.build()
out = os.path.join(os.path.dirname(b.HERE), 'variants')
os.makedirs(out, exist_ok=True)
others = [os.path.join(b.CSRC, s[:-3] + '.o') for s in b.SOURCES if s != 'chisq.cu']
for spec in sys.argv[1:]:
    name, flags = spec.split(':', 1)
    obj = os.path.join(out, f'chisq_{name}.o')
    r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + flags.split(',') +
                       ['-c', os.path.join(b.CSRC, 'chisq.cu'), '-o', obj],
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr); raise SystemExit(1)
    for ln in r.stderr.splitlines():
        if 'SineGridModeldLi32ELi1' in ln and 'Compiling' in ln:
            grab = 3
        if 'grab' in dir() and grab > 0 and ('Used' in ln or 'spill' in ln):
            print(name, ln.strip()); grab -= 1
    lib = os.path.join(out, f'libmc3b200_{name}.so')
    subprocess.run([b._nvcc(), '-shared', '-o', lib, obj] + others + ['-lcudart'], check=True)
    print('built', lib)
