"""
Builds libirsgmcmc.so (hand-written CUDA for sm_100a, C ABI in include/irsgmcmc.h) in-tree with nvcc.
nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repository snapshot.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libirsgmcmc.so')

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
              '-diag-suppress', '177']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(nvcc):
        raise RuntimeError('nvcc not found: libirsgmcmc.so cannot be built (there is no non-CUDA build)')
    return nvcc


def _stale(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile_object(src, obj, deps, defines, verbose, force):
    if not force and not _stale(obj, [src] + deps):
        return None
    cmd = [_nvcc(), '-c'] + NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else []) + [src, '-o', obj]
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def build_library(force=False, verbose=False, defines=(), out=None):
    """one object per translation unit (compiled in parallel, rebuilt only when stale), then one link.
    `defines` / `out`: development variants (e.g. defines=['IRS_X1=1'], out='libirsgmcmc_x1.so'; picked up at run time
    through the IRSGMCMC_LIB environment variable, see _lib.py)."""
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    headers = glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(ROOT, 'include', 'irsgmcmc.h')]
    lib = LIB if out is None else os.path.join(HERE, out)
    tag = '' if not defines else '_' + '_'.join(d.replace('=', '-') for d in defines)
    objdir = os.path.join(HERE, 'build', 'obj' + tag)
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        p = _compile_object(src, obj, headers, list(defines), verbose, force)
        if p is not None:
            procs.append((src, p))
    log = ''
    for src, p in procs:
        text, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{text}')
        log += text
    if verbose:
        print(log)
    if procs or not os.path.isfile(lib) or _stale(lib, objs):
        res = subprocess.run([_nvcc(), '-shared', '-o', lib] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return lib


def build_host_emulation(force=False):
    """test infrastructure: the library's __host__ __device__ arithmetic in CPU loops (tests/host_emul)"""
    src = os.path.join(ROOT, 'tests', 'host_emul', 'host_emul.cu')
    out = os.path.join(ROOT, 'tests', 'host_emul', 'libirs_host_emul.so')
    deps = [src] + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(ROOT, 'include', 'irsgmcmc.h')]
    if not force and not _stale(out, deps):
        return out
    cmd = [_nvcc(), '-shared', '-O2', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
           '-diag-suppress', '20013', src, '-o', out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc (host emulation) failed:\n' + res.stdout + res.stderr)
    return out


def build_oracle_c(force=False):
    """test infrastructure: the plain-C restatement of the bit-exact nearest-neighbour warp (oracle/nearest_warp.c)"""
    src = os.path.join(ROOT, 'oracle', 'nearest_warp.c')
    out = os.path.join(ROOT, 'oracle', 'libnearest_warp.so')
    if not os.path.isfile(src):
        return None
    if not force and not _stale(out, [src]):
        return out
    res = subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-ffp-contract=off', src, '-o', out, '-lm'],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc (oracle) failed:\n' + res.stdout + res.stderr)
    return out


if __name__ == '__main__':
    print(build_library(force=True, verbose=False))
    print(build_host_emulation(force=True))
    print(build_oracle_c(force=True))
