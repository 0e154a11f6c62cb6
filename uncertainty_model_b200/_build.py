"""Build libusl.so (sm_100a only) in-tree with nvcc.

    python -m uncertainty_model_b200._build

The shared library is git-ignored but travels to the GPU box with the
repository snapshot.  `__graft_entry__.build()` calls `build()`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libusl.so')
SOURCES = ['pyramid.cu', 'warp.cu', 'misc.cu', 'glue.cu', 'ssim.cu', 'loss_kernels.cu',
           'col_kernels.cu', 'col_inst_512.cu', 'col_inst_256.cu',
           'col_inst_128.cu', 'col_inst_64.cu', 'cons_kernels.cu', 'cons_rows.cu',
           'spars.cu']
HEADERS = ['usl_math.cuh', 'usl_common.cuh', 'loss_core.cuh', 'cons_core.cuh',
           'col_core.cuh',
           'col_launch.cuh', 'col_kernel_impl.cuh', 'cons_launch.cuh',
           os.path.join('..', '..', 'include', 'usl.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    obj_dir = os.path.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT,
                                            text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    with open(os.path.join(obj_dir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    cmd = [nvcc, '-shared', '-o', LIB] + objs
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if out.returncode != 0:
        raise RuntimeError(f'link failed:\n{out.stdout}')
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
